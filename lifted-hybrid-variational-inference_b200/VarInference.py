"""Ground variational inference: drop-in for the reference's ``VarInference.py``.

Bethe-free-energy minimisation with K-component mixture beliefs (fully factorised Gaussian x
categorical components) over a ground ``Graph``; every expectation, gradient and optimiser
step runs in the sm_100a kernels of ``liblhvi.so`` (see ``_vi_base.py`` / ``engine.py``).
"""
from __future__ import annotations

import numpy as np

from . import lowering
from ._vi_base import VIBase, mixture_mode, norm_pdf


class VarInference(VIBase):
    def __init__(self, g, num_mixtures=5, num_quadrature_points=3, *, dtype="float64", device=None, compat=None):
        self.g = g
        self._init_common(num_mixtures, num_quadrature_points, dtype, device, compat)

    def _handles(self):
        return sorted(self.g.rvs, key=lambda rv: rv.id)

    def _lower(self):
        return lowering.lower_ground(self.g, self.K, self.T)

    def _ground_graph(self):
        return self.g

    def rvs_map(self, rvs):
        """Joint MAP of several variables under the mixture belief by coordinate ascent
        (reference ``VarInference.py:378-456``): start every variable at its best marginal
        candidate, then 10 sweeps, each maximising one variable with the others' component
        responsibilities held fixed."""
        rvs = list(rvs)
        res = {}
        for rv in rvs:
            if rv.value is not None:
                res[rv] = rv.value
                continue
            eta = self.eta[rv]
            if rv.domain.continuous:
                cand = list(eta[:, 0])
                score = [(self.w * norm_pdf(x, eta[:, 0], eta[:, 1])).sum() for x in cand]
            else:
                cand = list(rv.domain.values)
                score = [(self.w * eta[:, d]).sum() for d in range(len(cand))]
            res[rv] = cand[int(np.argmax(score))]

        def comp(rv, x):
            eta = self.eta[rv]
            if rv.domain.continuous:
                return norm_pdf(x, eta[:, 0], eta[:, 1])
            return eta[:, rv.domain.values.index(x)]

        hidden = [rv for rv in rvs if rv.value is None]
        for _ in range(10):
            for rv in hidden:
                resp = np.array(self.w, dtype=float, copy=True)
                for other in hidden:
                    if other is not rv:
                        resp = resp * comp(other, res[other])
                eta = self.eta[rv]
                if rv.domain.continuous:
                    res[rv] = float(mixture_mode(resp, eta[None, :, 0], eta[None, :, 1],
                                                 np.array([float(res[rv])]))[0])
                else:
                    res[rv] = rv.domain.values[int(np.argmax((resp[:, None] * eta).sum(axis=0)))]
        return res
