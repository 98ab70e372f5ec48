"""Coarse-to-fine lifted variational inference: drop-in for the reference's
``C2FVarInference.py``.

Continuous evidence starts lumped into one class per domain, integrated as a fixed Gaussian
``(mean, variance)`` (``C2FVarInference.py:110-113,266-267``); every ``update_obs_its``
iterations the evidence classes whose spread exceeds a shrinking threshold are split by 1-D
k-means, colour passing re-runs, children inherit their parent's parameters and Adam moments
(``:33-61``), the new compressed graph is lowered and uploaded, and the kernels continue
(``:301-352``).  The Adam step counter ``t`` runs across rounds.
"""
from __future__ import annotations

from math import sqrt

from . import lowering
from ._vi_base import VIBase
from .CompressedGraphWithObs import CompressedGraph


class VarInference(VIBase):
    k_mean_k = 2
    k_mean_its = 10
    update_obs_its = 10
    output_its = 0
    min_obs_var = 0
    gaussian_obs = True

    def __init__(self, g, num_mixtures=5, num_quadrature_points=3, *, dtype="float64", device=None):
        self.g = CompressedGraph(g)
        self._init_common(num_mixtures, num_quadrature_points, dtype, device)

    def _handles(self):
        return sorted(self.g.rvs)

    def _lower(self):
        return lowering.lower_compressed(self.g, self.K, self.T, gaussian_obs=self.gaussian_obs,
                                         min_obs_var=self.min_obs_var)

    def _ground_graph(self):
        return self.g.g

    def _handle_of(self, rv):
        return rv.cluster if hasattr(rv, "cluster") and not hasattr(rv, "rvs") else rv

    def _gauss_evidence(self, h):
        if self.gaussian_obs and h.value is not None and h.variance > self.min_obs_var:
            return (h.value, h.variance)
        return None

    # ---- refinement -----------------------------------------------------------------------
    def split_evidence(self, epsilon):
        n = -1
        while n != len(self.g.rvs):
            n = len(self.g.rvs)
            self.g.split_evidence(self.k_mean_k, self.k_mean_its, epsilon)

    def split_rvs(self):
        """Structure split in which the pieces of a hidden class inherit its parameters and
        Adam moments (``C2FVarInference.py:39-61``)."""
        for rv in sorted(self.g.rvs):
            pieces = rv.split_by_structure()
            if rv.value is not None:
                self.g.note_evidence_split(rv, pieces)
            else:
                for piece in pieces:
                    if piece is rv:
                        continue
                    self.eta[piece] = self.eta[rv]
                    if rv.domain.continuous:
                        self.eta_g[0][piece] = self.eta_g[0][rv]
                        self.eta_g[1][piece] = self.eta_g[1][rv]
                    else:
                        self.eta_tau[piece] = self.eta_tau[rv]
                        self.eta_tau_g[0][piece] = self.eta_tau_g[0][rv]
                        self.eta_tau_g[1][piece] = self.eta_tau_g[1][rv]
            self.g.rvs |= pieces

    def cp_run(self):
        n = -1
        while n != len(self.g.rvs):
            n = len(self.g.rvs)
            self.g.split_factors()
            self.split_rvs()
        self._drop_engine()      # graph changed: lower + upload again before the next pass

    def run(self, iteration=100, lr=0.1, is_log=True, log_fe=True):
        self._start_run(lr, is_log, log_fe)
        self.g.init_cluster(is_split_cont_evidence=False)
        self._drop_engine()
        self.init_param()
        self._zero_moments()
        self.cp_run()

        epsilon = 0
        for rv in self.g.rvs:
            if rv.value is not None:
                epsilon = max(sqrt(rv.variance), epsilon)
        d = epsilon * self.update_obs_its / (iteration - self.output_its)
        epsilon -= d

        if self.is_log:
            self.time_log = list()
            self.total_time = 0

        for _ in range(int(iteration / self.update_obs_its)):       # remainder dropped (H10)
            self.split_evidence(epsilon)
            self.cp_run()
            epsilon = max(epsilon - d, self.min_obs_var)
            print('split, num of rvs:', len(self.g.rvs))
            self.ADAM_update(self.update_obs_its)
