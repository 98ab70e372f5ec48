"""Lifted variational inference: drop-in for the reference's ``LiftedVarInference.py``.

The ground graph is compressed once by colour passing (``CompressedGraphWithObs``); the
kernels then process one record per *class* of factors, weighted by class sizes and
neighbour counts (``LiftedVarInference.py:74,90,111-112,131-132,162``).  ``belief`` and
``map`` take ground variables and dereference ``rv.cluster`` (``:356-380``).
"""
from __future__ import annotations

from . import lowering
from ._vi_base import VIBase
from .CompressedGraphWithObs import CompressedGraph


class VarInference(VIBase):
    def __init__(self, g, num_mixtures=5, num_quadrature_points=3, *, dtype="float64", device=None, compat=None):
        self.g = CompressedGraph(g)
        self.g.run()
        self._init_common(num_mixtures, num_quadrature_points, dtype, device, compat)

    def _handles(self):
        return sorted(self.g.rvs)

    def _lower(self):
        return lowering.lower_compressed(self.g, self.K, self.T)

    def _ground_graph(self):
        return self.g.g

    def _handle_of(self, rv):
        return rv.cluster if hasattr(rv, "cluster") and not hasattr(rv, "rvs") else rv
