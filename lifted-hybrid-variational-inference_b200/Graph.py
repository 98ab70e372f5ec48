"""Factor-graph data model accepted by the B200 VI engines.

Mirrors the attribute contract of the reference's ``Graph.py`` (Domain ``:11-19``,
Potential ``:32-51``, RV ``:54-90``, F ``:93-129``, Graph ``:137-172``) so graphs
built for the reference load unchanged: the engines only ever touch
``g.rvs / g.factors``, ``rv.domain / rv.value / rv.nb / rv.N``, ``f.potential / f.nb``
and ``domain.values / domain.continuous``.  The reference's own objects are
accepted too (duck typing) -- nothing here is required to be *this* class.
"""
from __future__ import annotations

import itertools
from abc import ABC, abstractmethod

import numpy as np


class Domain:
    """Value set of a random variable.

    Discrete: ``values`` enumerates the states.  Continuous: ``values`` is the
    ``(low, high)`` range, ``integral_points`` an optional evaluation grid
    (only used by plotting / KL utilities, never by the kernels).
    """

    __slots__ = ("values", "continuous", "integral_points")

    def __init__(self, values, continuous=False, integral_points=None):
        self.values = tuple(values)
        self.continuous = bool(continuous)
        self.integral_points = None
        if self.continuous:
            self.integral_points = (
                np.linspace(self.values[0], self.values[1], 30)
                if integral_points is None else integral_points
            )

    def __repr__(self):
        tag = "cont" if self.continuous else "disc"
        return f"Domain<{tag} {self.values}>"


class Potential(ABC):
    """Plugin base class: a factor potential is anything with ``get(x) -> psi(x) >= 0``.

    ``symmetric`` tells colour passing that argument order is irrelevant
    (reference ``Graph.py:33-35``).  Subclasses may also provide
    ``get_quadratic_params() -> (A, b, c)`` with ``log psi = x'Ax + b'x + c``;
    the lowering layer uses it when present and otherwise probes ``get``.
    """

    def __init__(self, symmetric=False):
        self.symmetric = symmetric
        self.alpha = 0.001

    @abstractmethod
    def get(self, parameters):
        ...

    def gradient(self, parameters, wrt):
        x = np.asarray(parameters, dtype=float)
        x_step = x + np.asarray(wrt, dtype=float) * self.alpha
        return (self.get(x_step) - self.get(x)) / self.alpha

    def log_gradient(self, parameters, wrt):
        x = np.asarray(parameters, dtype=float)
        x_step = x + np.asarray(wrt, dtype=float) * self.alpha
        return (np.log(self.get(x_step)) - np.log(self.get(x))) / self.alpha


class _Node:
    """Shared id / ordering behaviour of RV and F."""

    __slots__ = ()

    def __lt__(self, other):
        return self.id < other.id


class RV(_Node):
    """Random variable; ``value is None`` means hidden, otherwise point evidence."""

    _ids = itertools.count()

    def __init__(self, domain, value=None):
        self.domain = domain
        self.value = value
        self.id = next(RV._ids)
        self.nb = []
        self.N = 0
        self.cluster = None  # set by colour passing (CompressedGraphWithObs)

    @property
    def dstates(self):
        return None if self.domain.continuous else len(self.domain.values)

    @property
    def domain_type(self):
        return "c-g" if self.domain.continuous else f"d-{self.dstates}"

    def __repr__(self):
        return f"{self.domain_type} rv #{self.id}"


class F(_Node):
    """Factor: a potential applied to an ordered neighbour list ``nb``."""

    _ids = itertools.count()

    def __init__(self, potential=None, nb=None):
        self.potential = potential
        self.nb = [] if nb is None else nb
        self.id = next(F._ids)
        self.cluster = None

    def __repr__(self):
        return f"factor #{self.id}"


class Graph:
    """Ground factor graph.  ``rvs`` / ``factors`` may be sets or lists."""

    def __init__(self, rvs=None, factors=None):
        self.rvs = set() if rvs is None else rvs
        self.factors = set() if factors is None else factors
        if rvs is not None and factors is not None:
            self.init_nb()

    def init_nb(self):
        """Rebuild ``rv.nb`` (one entry per occurrence) and ``rv.N`` (ground degree)."""
        for rv in self.rvs:
            rv.nb = []
        for f in sorted(self.factors):
            for rv in f.nb:
                rv.nb.append(f)
        for rv in self.rvs:
            rv.N = len(rv.nb)

    @property
    def rvs_list(self):
        return sorted(self.rvs)

    @property
    def factors_list(self):
        return sorted(self.factors)
