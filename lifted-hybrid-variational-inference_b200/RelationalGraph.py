"""Relational model builder: drop-in for the reference's ``RelationalGraph.py`` (``LV``, ``Atom``,
``ParamF``, ``RelationalGraph``), with an array-native grounding next to the
object one.

A parametric factor lists atom expressions such as ``'loss(c,b)'`` or ``'recession($all)'``
(``$x`` is the constant instance ``x``, a bare token is a logical variable ranging over the
instances of the atom's logical variable at that position, ``RelationalGraph.py:44-81``);
grounding substitutes every combination of the tokens' instances -- filtered by the optional
``constrain(substitution)`` -- and creates one factor per combination over the ground atoms
``(name, instance, ...)`` (``:104-132``).  Ground atoms are created on first use, so only the
ones some factor touches exist.

* ``ground_graph()`` / ``add_evidence(data)``: the reference's object route (a ``Graph`` of ``RV`` /
  ``F`` and the ``rvs_dict`` keyed by ground atom), for small models and for the reference's demos.
* ``ground_arrays(data=None)``: the same grounding as ``lifting.GroundArrays`` -- cross products
  and ground-atom ids are computed on index arrays, no Python object per ground atom or factor --
  which ``lifting.ArrayVI`` / ``lifting.C2FArrayVI`` take directly.
"""
from __future__ import annotations

import re
from itertools import product

import numpy as np

from .Graph import *  # noqa: F401,F403  (the reference's scripts rely on `from RelationalGraph import *` for Domain, RV, F)
from .Graph import F, RV, Graph

_TOKEN = re.compile(r"\$?\w+")


class LV:
    """Logical variable: a tuple / list of instance names."""

    def __init__(self, instances):
        self.instances = instances


class Atom:
    """Relational atom ``name(lv_1, ..., lv_n)`` whose groundings take values in ``domain``."""

    def __init__(self, domain, logical_variables, name=None):
        self.domain = domain
        self.lvs = logical_variables
        self.name = name


class ParamF:
    """Parametric factor: a potential over atom expressions, optionally constrained."""

    def __init__(self, potential, nb=None, constrain=None):
        self.potential = potential
        self.constrain = constrain
        self.nb = [] if nb is None else nb


def _parts(expression):
    return _TOKEN.findall(expression) if isinstance(expression, str) else list(expression)


class GroundIndex:
    """Ground atoms of an array grounding: ``index_of(key)`` / ``key_of(i)`` translate between
    the reference's keys ``(atom name, instance, ...)`` and variable indices."""

    def __init__(self, atoms, codes, offsets):
        self._atoms = atoms              # name -> Atom
        self._codes = codes              # name -> sorted int64 codes of the groundings in use
        self._offsets = offsets          # name -> first variable index of that atom
        self._order = sorted(offsets, key=offsets.get)
        self._pos = {}

    def _position(self, atom, i):
        key = (atom.name, i)
        if key not in self._pos:
            self._pos[key] = {inst: j for j, inst in enumerate(atom.lvs[i].instances)}
        return self._pos[key]

    def _shape(self, atom):
        return tuple(len(lv.instances) for lv in atom.lvs)

    def index_of(self, key):
        """Variable index of ground atom ``key``, or -1 if no factor touches it."""
        atom = self._atoms[key[0]]
        try:
            idx = tuple(self._position(atom, i)[inst] for i, inst in enumerate(key[1:]))
        except KeyError:
            return -1
        code = int(np.ravel_multi_index(idx, self._shape(atom)))
        codes = self._codes[atom.name]
        j = int(np.searchsorted(codes, code))
        if j >= codes.size or codes[j] != code:
            return -1
        return self._offsets[atom.name] + j

    def key_of(self, i):
        name = self._order[int(np.searchsorted([self._offsets[n] for n in self._order], i, side="right")) - 1]
        atom = self._atoms[name]
        code = int(self._codes[name][i - self._offsets[name]])
        idx = np.unravel_index(code, self._shape(atom))
        return (name,) + tuple(lv.instances[int(j)] for lv, j in zip(atom.lvs, idx))

    def __len__(self):
        return sum(c.size for c in self._codes.values())


class RelationalGraph:
    def __init__(self, atoms, parametric_factors):
        self.atoms = atoms
        self.param_factors = parametric_factors
        self.atoms_dict = {atom.name: atom for atom in atoms}
        self.rvs_dict = dict()
        self.grounding = None

    # ---- shared: tokens of a parametric factor and their instance tables -----------------------
    def _tokens(self, param_f):
        """``{token: instances}`` in order of first appearance (``extract_lvs``, ``:62-75``)."""
        lvs = dict()
        for expression in param_f.nb:
            parts = _parts(expression)
            atom = self.atoms_dict[parts[0]]
            for i in range(len(atom.lvs)):
                s = parts[i + 1]
                if s[0] != "$":
                    lvs[s] = atom.lvs[i].instances
        return lvs

    # ---- object route ------------------------------------------------------------------------
    def ground_graph(self):
        """``(Graph, rvs_dict)``: one ``F`` per admissible substitution of every parametric factor
        (``:104-132``); ground atoms are created on first use."""
        factors = []
        for param_f in self.param_factors:
            lvs = self._tokens(param_f)
            tokens = list(lvs)
            expressions = [_parts(e) for e in param_f.nb]
            for combination in product(*[lvs[t] for t in tokens]):
                substitution = dict(zip(tokens, combination))
                if param_f.constrain is not None and not param_f.constrain(substitution):
                    continue
                nb = []
                for parts in expressions:
                    key = (parts[0],) + tuple(s[1:] if s[0] == "$" else substitution[s] for s in parts[1:])
                    if key not in self.rvs_dict:
                        self.rvs_dict[key] = RV(self.atoms_dict[parts[0]].domain)
                    nb.append(self.rvs_dict[key])
                factors.append(F(potential=param_f.potential, nb=nb))
        g = Graph()
        g.rvs = set(self.rvs_dict.values())
        g.factors = set(factors)
        g.init_nb()
        self.grounding = g
        return self.grounding, self.rvs_dict

    def add_evidence(self, data):
        """``data``: ``{(atom name, instance, ...): value}``; every other ground atom becomes
        hidden (``:93-102``)."""
        for key, rv in self.rvs_dict.items():
            rv.value = data[key] if key in data else None
        return self.grounding, self.rvs_dict

    # ---- array route -------------------------------------------------------------------------
    def ground_arrays(self, data=None):
        """The same grounding as ``(lifting.GroundArrays, GroundIndex)``.  Variables are numbered
        atom by atom (in the order the atoms were given), groundings of one atom in row-major
        order of their instances; one ``FactorBlock`` per parametric factor, factors in the
        reference's substitution order."""
        from .lifting import FactorBlock, GroundArrays
        position = {}

        def pos(atom, i):
            if (atom.name, i) not in position:
                position[(atom.name, i)] = {inst: j for j, inst in enumerate(atom.lvs[i].instances)}
            return position[(atom.name, i)]

        raw_blocks = []                 # (param_f, [(atom name, codes [n])] per expression)
        used = {atom.name: [] for atom in self.atoms}
        for param_f in self.param_factors:
            lvs = self._tokens(param_f)
            tokens = list(lvs)
            sizes = [len(lvs[t]) for t in tokens]
            n = int(np.prod(sizes)) if sizes else 1
            grids = np.indices(sizes).reshape(len(sizes), n) if sizes else np.zeros((0, 1), dtype=np.int64)
            keep = np.ones(n, dtype=bool)
            if param_f.constrain is not None:
                for j in range(n):
                    substitution = {t: lvs[t][int(grids[a, j])] for a, t in enumerate(tokens)}
                    keep[j] = bool(param_f.constrain(substitution))
            grids = grids[:, keep]
            per_expr = []
            for expression in param_f.nb:
                parts = _parts(expression)
                atom = self.atoms_dict[parts[0]]
                shape = tuple(len(lv.instances) for lv in atom.lvs)
                idx = []
                for i, s in enumerate(parts[1:]):
                    if s[0] == "$":
                        idx.append(np.full(grids.shape[1], pos(atom, i)[s[1:]], dtype=np.int64))
                    else:
                        idx.append(grids[tokens.index(s)].astype(np.int64))
                codes = np.ravel_multi_index(tuple(idx), shape) if idx else np.zeros(grids.shape[1], np.int64)
                per_expr.append((atom.name, codes))
                used[atom.name].append(codes)
            raw_blocks.append((param_f, per_expr))

        codes_of, offset_of, cursor = {}, {}, 0
        domains, dom_index, var_dom = [], {}, []
        for atom in self.atoms:
            codes = np.unique(np.concatenate(used[atom.name])) if used[atom.name] else np.zeros(0, np.int64)
            codes_of[atom.name], offset_of[atom.name] = codes, cursor
            cursor += codes.size
            if id(atom.domain) not in dom_index:
                dom_index[id(atom.domain)] = len(domains)
                domains.append(atom.domain)
            var_dom.append(np.full(codes.size, dom_index[id(atom.domain)], dtype=np.int32))
        blocks = []
        for param_f, per_expr in raw_blocks:
            cols = [offset_of[name] + np.searchsorted(codes_of[name], codes) for name, codes in per_expr]
            blocks.append(FactorBlock(param_f.potential, np.stack(cols, axis=1).astype(np.int64)))
        index = GroundIndex(self.atoms_dict, codes_of, offset_of)
        var_value = np.full(cursor, np.nan)
        if data:
            for key, value in data.items():
                i = index.index_of(key)
                if i >= 0:
                    var_value[i] = float(value)
        ga = GroundArrays(domains, np.concatenate(var_dom) if var_dom else np.zeros(0, np.int32), var_value, blocks)
        return ga, index
