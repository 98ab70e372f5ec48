"""Gaussian belief propagation on the device: drop-in for the reference's ``GaBP`` (``GaBP.py:7-216``),
the cross-check of the Gaussian configurations (BASELINE config 4).

    bp = GaBP(g); bp.run(20); mu, var = bp.get_belief_params(rv); bp.map(rv); bp.belief(x, rv)

The reference sweeps a dict of ``(mu, sig)`` messages between Python objects: all variable -> factor
messages, then all factor -> variable messages (``run`` ``:139-165``), for the four pairwise
potential classes it knows (``message_f_to_rv`` ``:37-136``).  Here the model is lowered once to its
information form -- ``log psi_f = -1/2 x' J_f x + h_f' x`` read off the same coefficient table the
variational kernels use, evidence substituted, so every exp-quadratic potential is covered -- and the
sweeps run as CUDA kernels over the directed edge list (``csrc/lhvi_gabp.cu``, ``lhvi_gabp_sweeps``):
same flooding schedule, same start (every message at ``(0, 1)``), hence the same numbers after the same
number of iterations (``tests/golden/gabp_grid.json`` holds the reference's).  There is no CPU path.

Hazard kept out of the parity fixtures: for a ``GaussianPotential`` with an observed neighbour the
reference returns ``-u2 - a2 (value - u1) / a3`` (``:72``; the conditional mean is ``u2 - ...``), which
is only right for ``u2 = 0``.  This module computes the conditional from the quadratic form.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _cabi, lowering


@dataclass
class GabpArrays:
    """Information form of a pairwise Gaussian model (``include/lhvi.h``, ``lhvi_gabp``)."""
    n_vars: int
    jd: np.ndarray          # f64 [V]   precision of the unary / evidence-reduced factors of each variable
    hd: np.ndarray          # f64 [V]   their potential
    nfac: np.ndarray        # i64 [V]   factors on each variable (the reference's initial messages are (0, 1) each)
    src: np.ndarray         # i32 [E]
    dst: np.ndarray         # i32 [E]
    rev: np.ndarray         # i32 [E]
    coef: np.ndarray        # f64 [5, E]  J_ss, J_dd, J_sd, h_s, h_d


def reduce_quadratic(coef, nh, ev):
    """Substitute the point evidence ``ev`` [ne, n] into packed quadratics ``coef`` [n, ncoef] over
    ``nh + ne`` arguments (layout of ``lowering.PotentialTable``: c, b, upper-triangular A row-major):
    returns ``(lin [nh, n], quad {(i, j): [n]})`` over the first ``nh`` arguments."""
    ne = ev.shape[0]
    nct = nh + ne
    lin = [coef[:, 1 + i].copy() for i in range(nh)]
    quad = {}
    p = 1 + nct
    for i in range(nct):
        for j in range(i, nct):
            a = coef[:, p]
            p += 1
            if j < nh:
                quad[(i, j)] = a.copy()
            elif i < nh:
                lin[i] += a * ev[j - nh]
    return lin, quad


def gabp_from_model(model: lowering.LoweredModel) -> GabpArrays:
    """Information form of a lowered all-continuous pairwise model (any exp-quadratic potentials; point
    evidence already substituted record by record).  Variables are the model's parameter slots, in slot
    order."""
    if np.any(model.var_kind != 0):
        raise NotImplementedError("GaBP: discrete hidden variables (the reference's GaBP is Gaussian only)")
    V = model.n_vars
    lookup = np.full(int(model.var_off.max()) + 1 if V else 1, -1, dtype=np.int64)      # slot offset -> variable
    lookup[model.var_off.astype(np.int64)] = np.arange(V)
    jd, hd, nfac = np.zeros(V), np.zeros(V), np.zeros(V, dtype=np.int64)
    blocks = []                           # per pairwise group: (v0, v1, [J00, J11, J01, h0, h1])
    ptab = np.asarray(model.ptab, dtype=np.float64)
    for g in model.groups:
        if g.node or g.n == 0:
            continue
        if g.nd or g.ng or g.kind != lowering.POT_QUADRATIC:
            raise NotImplementedError("GaBP: only exp-quadratic potentials over continuous variables")
        if g.nc == 0:
            continue                       # all arguments observed: a constant
        if g.nc > 2:
            raise NotImplementedError("GaBP: factors over more than two hidden variables (GaBP.py:38 'only for pairwise potential')")
        ncoef = lowering.ncoef_for(g.nc + g.ne)
        coef = ptab[g.pot.astype(np.int64)[:, None] + np.arange(ncoef)[None, :]]
        lin, quad = reduce_quadratic(coef, g.nc, g.ecval)
        v0 = lookup[g.poff[0]]
        np.add.at(nfac, v0, 1)
        if g.nc == 1:                      # unary or evidence-reduced factor: a constant message of its variable
            np.add.at(jd, v0, -2.0 * quad[(0, 0)])
            np.add.at(hd, v0, lin[0])
            continue
        v1 = lookup[g.poff[1]]
        if np.any(v0 == v1):
            raise NotImplementedError("GaBP: a factor that takes the same variable twice")
        np.add.at(nfac, v1, 1)
        blocks.append((v0, v1, np.stack([-2.0 * quad[(0, 0)], -2.0 * quad[(1, 1)], -quad[(0, 1)], lin[0], lin[1]])))
    # directed slots: per group the messages 0 -> 1 of all its factors, then the messages 1 -> 0; rev pairs them
    src, dst, rev, cols, at = [], [], [], [], 0
    for v0, v1, c in blocks:
        n = v0.size
        src += [v0, v1]
        dst += [v1, v0]
        cols += [c, c[[1, 0, 2, 4, 3]]]
        rev += [np.arange(at + n, at + 2 * n), np.arange(at, at + n)]
        at += 2 * n
    if blocks:
        s, d, rev = (np.concatenate(x).astype(np.int32) for x in (src, dst, rev))
        c = np.concatenate(cols, axis=1)
    else:
        s = d = rev = np.zeros(0, dtype=np.int32)
        c = np.zeros((5, 0))
    return GabpArrays(V, jd, hd, nfac, s, d, rev, np.ascontiguousarray(c))


class DeviceGaBP:
    """The sweeps on the device over a ``GabpArrays`` model (``lhvi_gabp_sweeps`` / ``lhvi_gabp_marginals``)."""

    def __init__(self, arrays: GabpArrays, dtype="float64", device=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("lhvi: no CUDA device visible. GaBP runs only on the GPU (csrc/lhvi_gabp.cu); "
                               "there is no CPU fallback.")
        self.lib = _cabi.load()
        self.torch = torch
        self.a = arrays
        self.device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.tdtype = torch.float64 if str(dtype) in ("float64", "torch.float64", "f64") else torch.float32
        dev = lambda x, dt: torch.as_tensor(np.ascontiguousarray(x)).to(dt).to(self.device).contiguous()
        V, E = arrays.n_vars, int(arrays.src.size)
        self.src, self.dst, self.rev = (dev(x, torch.int32) for x in (arrays.src, arrays.dst, arrays.rev))
        self.coef = dev(arrays.coef, self.tdtype)
        self.jd, self.hd = dev(arrays.jd, self.tdtype), dev(arrays.hd, self.tdtype)
        self.P = torch.ones(2 * max(E, 1), dtype=self.tdtype, device=self.device)        # messages start at (mu, sig) = (0, 1)
        self.H = torch.zeros(2 * max(E, 1), dtype=self.tdtype, device=self.device)
        self.SP = torch.zeros(2 * max(V, 1), dtype=self.tdtype, device=self.device)
        self.SH = torch.zeros(2 * max(V, 1), dtype=self.tdtype, device=self.device)
        self.SP[:V] = dev(arrays.nfac, self.tdtype)
        d = self.desc = _cabi.LhviGabp()
        d.dtype = _cabi.LHVI_F64 if self.tdtype == torch.float64 else _cabi.LHVI_F32
        d.parity, d.n_vars, d.n_edges = 0, V, E
        d.src, d.dst, d.rev, d.coef = self.src.data_ptr(), self.dst.data_ptr(), self.rev.data_ptr(), self.coef.data_ptr()
        d.jd, d.hd = self.jd.data_ptr(), self.hd.data_ptr()
        d.P, d.H, d.SP, d.SH = self.P.data_ptr(), self.H.data_ptr(), self.SP.data_ptr(), self.SH.data_ptr()
        self.sweeps_done = 0

    def sweeps(self, n):
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        with self.torch.cuda.device(self.device):
            _cabi.check(self.lib.lhvi_gabp_sweeps(C.byref(self.desc), int(n), C.c_void_p(stream)), self.lib)
        self.desc.parity = (self.desc.parity + int(n)) & 1
        self.sweeps_done += int(n)

    def marginals(self):
        """``(mean [V], variance [V])`` as device tensors."""
        V = self.a.n_vars
        mean = self.torch.empty(max(V, 1), dtype=self.tdtype, device=self.device)
        var = self.torch.empty(max(V, 1), dtype=self.tdtype, device=self.device)
        stream = self.torch.cuda.current_stream(self.device).cuda_stream
        with self.torch.cuda.device(self.device):
            _cabi.check(self.lib.lhvi_gabp_marginals(C.byref(self.desc), mean.data_ptr(), var.data_ptr(), C.c_void_p(stream)), self.lib)
        return mean[:V], var[:V]


class GaBP:
    """Same surface as the reference class (``GaBP.py:7-216``)."""

    def __init__(self, g=None, dtype="float64", device=None):
        self.g = g
        self.dtype, self.device = dtype, device
        self.message = dict()          # (kept for attribute compatibility; the messages live on the device)
        self._mean = self._var = None

    @staticmethod
    def norm_pdf(x, mu, sig):
        u = x - mu
        return np.exp(-u * u * 0.5 / sig) / (2.506628274631 * sig)

    def run(self, iteration=10, log_enable=False):
        """``iteration`` variable -> factor passes and ``iteration - 1`` factor -> variable passes, as the
        reference's loop (``:147-165``): the beliefs read the factor -> variable messages, i.e. the state
        after ``iteration - 1`` sweeps."""
        self.model = lowering.lower_ground(self.g, 1, 3)
        self.arrays = gabp_from_model(self.model)
        self.engine = DeviceGaBP(self.arrays, self.dtype, self.device)
        self.engine.sweeps(max(int(iteration) - 1, 0))
        mean, var = self.engine.marginals()
        self._mean, self._var = mean.double().cpu().numpy(), var.double().cpu().numpy()
        return self

    def get_belief_params(self, rv):
        assert rv.value is None
        i = self.model.index[rv]
        return float(self._mean[i]), float(self._var[i])

    def belief(self, x, rv):
        if rv.value is None:
            mu, var = self.get_belief_params(rv)
            return float(self.norm_pdf(x, mu, var))
        return 1 if x == rv.value else 0

    def map(self, rv):
        return self.get_belief_params(rv)[0] if rv.value is None else rv.value
