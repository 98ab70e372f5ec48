"""Colour-passing compression with evidence clustering (host side of the lifted engines).

Produces the same partition and the same products as the reference's
``CompressedGraphWithObs.py`` -- ``SuperRV{rvs, domain, value, variance, nb, N, count}``
(``:8-45``), ``SuperF{factors, potential, nb}`` (``:133-150``), ``rv.cluster`` /
``f.cluster`` back-pointers (``:17-18,138-139``) and the entry points ``init_cluster``
(``:187-234``), ``split_factors`` (``:260-262``), ``split_rvs`` (``:249-258``),
``split_evidence`` (``:236-247``), ``run`` (``:264-271``) -- but is deterministic:
members are kept in id order and every "pick a representative / first k values" choice
the reference makes by set-iteration order is made by smallest id here.  The refinement
fixed point (coarsest equitable partition) does not depend on that order.
"""
from __future__ import annotations

import itertools
from collections import Counter

import numpy as np


def _by_id(items):
    return sorted(items, key=lambda o: o.id)


def _group(members, key_fn):
    """Stable partition of ``members`` by ``key_fn`` (first-seen order of keys)."""
    groups = {}
    for m in members:
        groups.setdefault(key_fn(m), []).append(m)
    return list(groups.values())


class SuperRV:
    """A colour class of ground variables.  For evidence classes ``value`` is the mean of
    the members' values and ``variance`` their population variance (``:9-13,38-39``)."""

    _uids = itertools.count()

    def __init__(self, rvs, domain=None, value=None):
        self.uid = next(SuperRV._uids)
        self.rvs = _by_id(rvs)
        first = self.rvs[0]
        self.domain = first.domain if domain is None else domain
        if value is None and first.value is not None:
            value = self.mean_value(self.rvs)
        self.value = value
        self.variance = None if self.value is None else self.member_variance()
        self.nb = None
        self.N = 0
        self.count = None
        for rv in self.rvs:
            rv.cluster = self

    def __lt__(self, other):
        return self.uid < other.uid

    def __repr__(self):
        return f"SuperRV#{self.uid}(n={len(self.rvs)}, value={self.value})"

    @staticmethod
    def mean_value(rvs):
        acc = 0
        for rv in rvs:
            acc += rv.value
        return acc / len(rvs)

    def member_variance(self):
        return np.var(tuple(rv.value for rv in self.rvs))

    def update_nb(self):
        """Lifted neighbourhood of the class, read off its representative: how many
        ground factors of each factor class touch one member (``:41-45``)."""
        rep = self.rvs[0]
        self.count = Counter(f.cluster for f in rep.nb)
        self.nb = tuple(self.count)
        self.N = rep.N

    def split_by_structure(self):
        """Split by the multiset of neighbouring factor classes (``:47-76``).  The first
        group stays in this object so parameters keyed by it remain valid."""
        parts = _group(self.rvs, lambda rv: tuple(sorted(f.cluster.uid for f in rv.nb)))
        self.rvs = parts[0]
        if self.value is not None:
            self.value = self.mean_value(self.rvs)
            self.variance = self.member_variance()
        out = [self] + [SuperRV(p, self.domain, None) for p in parts[1:]]
        for c in out:
            c.update_nb()
        return set(out)

    def split_by_evidence(self, k=2, iteration=10):
        """1-D k-means over the members' evidence values (``:78-130``): centroids start at
        the first ``k`` distinct values, ``iteration`` Lloyd sweeps on the value histogram,
        then members go to their nearest centroid; class values become the centroids."""
        if len(self.rvs) <= 1:
            return {self}
        hist = Counter(rv.value for rv in self.rvs)
        k = min(k, len(hist))
        if k <= 1:
            return {self}
        centroids = np.array(list(itertools.islice(hist, k)), dtype=float)
        vals = np.fromiter(hist.keys(), dtype=float, count=len(hist))
        cnts = np.fromiter(hist.values(), dtype=float, count=len(hist))
        for _ in range(iteration):
            owner = np.abs(vals[:, None] - centroids[None, :]).argmin(axis=1)
            mass = np.bincount(owner, weights=cnts, minlength=k)
            tot = np.bincount(owner, weights=vals * cnts, minlength=k)
            with np.errstate(invalid="ignore", divide="ignore"):
                centroids = tot / mass
        parts = [[] for _ in range(k)]
        for rv in self.rvs:
            parts[int(np.abs(centroids - rv.value).argmin())].append(rv)
        self.rvs = parts[0]
        self.value = centroids[0]
        self.variance = self.member_variance()
        out = [self] + [SuperRV(parts[j], self.domain, centroids[j]) for j in range(1, k)]
        for c in out:
            c.update_nb()
        return set(out)


class SuperF:
    """A colour class of ground factors sharing one potential."""

    _uids = itertools.count()

    def __init__(self, factors):
        self.uid = next(SuperF._uids)
        self.factors = _by_id(factors)
        self.potential = self.factors[0].potential
        self.nb = None
        for f in self.factors:
            f.cluster = self

    def __lt__(self, other):
        return self.uid < other.uid

    def __repr__(self):
        return f"SuperF#{self.uid}(n={len(self.factors)})"

    def update_nb(self):
        self.nb = tuple(rv.cluster for rv in self.factors[0].nb)

    def split_by_structure(self):
        """Split by the (ordered, or sorted if the potential is symmetric) tuple of
        neighbouring variable classes (``:152-175``)."""
        def signature(f):
            sig = [rv.cluster.uid for rv in f.nb]
            return tuple(sorted(sig)) if f.potential.symmetric else tuple(sig)

        parts = _group(self.factors, signature)
        self.factors = parts[0]
        out = [self] + [SuperF(p) for p in parts[1:]]
        for c in out:
            c.update_nb()
        return set(out)


class CompressedGraph:
    """Colour passing over a ground ``Graph`` (reference ``:178-271``)."""

    def __init__(self, graph):
        self.g = graph
        self.rvs = set()
        self.factors = set()
        self.clustered_evidence = set()

    def init_cluster(self, is_split_cont_evidence=True):
        self.rvs.clear()
        self.factors.clear()
        self.clustered_evidence.clear()

        for same_domain in _group(_by_id(self.g.rvs), lambda rv: id(rv.domain)):
            hidden = [rv for rv in same_domain if rv.value is None]
            evidence = [rv for rv in same_domain if rv.value is not None]
            if hidden:
                self.rvs.add(SuperRV(hidden))
            if not evidence:
                continue
            if not is_split_cont_evidence and same_domain[0].domain.continuous:
                # coarse start of C2F: one class for all continuous observations
                lumped = SuperRV(evidence)
                self.rvs.add(lumped)
                self.clustered_evidence.add(lumped)
            else:
                for same_value in _group(evidence, lambda rv: rv.value):
                    self.rvs.add(SuperRV(same_value))

        # potentials colour factors through their own __hash__/__eq__
        for same_pot in _group(_by_id(self.g.factors), lambda f: f.potential):
            self.factors.add(SuperF(same_pot))

    def split_evidence(self, k=2, iteration=10, epsilon=0):
        for rv in sorted(self.clustered_evidence):
            if np.sqrt(rv.variance) > epsilon:
                pieces = rv.split_by_evidence(k, iteration)
                if len(pieces) > 1:
                    for piece in pieces:
                        if piece.variance > epsilon:
                            self.clustered_evidence.add(piece)
                elif len(next(iter(pieces)).rvs) == 1:
                    self.clustered_evidence -= pieces
                self.rvs |= pieces

    def note_evidence_split(self, rv, pieces):
        """Book-keeping shared by ``split_rvs`` here and in the C2F engine (``:253-257``)."""
        if rv.value is not None:
            if len(pieces) > 1:
                self.clustered_evidence |= pieces
            elif len(next(iter(pieces)).rvs) == 1:
                self.clustered_evidence -= pieces

    def split_rvs(self):
        for rv in sorted(self.rvs):
            pieces = rv.split_by_structure()
            self.note_evidence_split(rv, pieces)
            self.rvs |= pieces

    def split_factors(self):
        for f in sorted(self.factors):
            self.factors |= f.split_by_structure()

    def run(self):
        self.init_cluster(is_split_cont_evidence=True)
        n_before = -1
        while n_before != len(self.rvs):
            n_before = len(self.rvs)
            self.split_factors()
            self.split_rvs()
