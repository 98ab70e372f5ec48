"""ORACLE (test infrastructure) -- ctypes wrapper of ``oracle/vi_port.c`` (the plain C / OpenMP
restatement of one pass of the reference's VI loop; see the header of that file for the
reference lines it follows).  Used by ``tests/test_c_port.py`` and, through
``oracle/cpu_port.py``, by bench.py's CPU-baseline / reference arm.  Never imported by the
product path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from oracle.vi_numpy import NumpyVI, quadrature, tau_gradients

LIB = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libvi_oracle.so")
MAX_AXES = 6


class VoGroup(C.Structure):
    _fields_ = [("nd", C.c_int), ("nc", C.c_int), ("ng", C.c_int), ("ne", C.c_int),
                ("dims", C.c_int * MAX_AXES), ("node", C.c_int), ("weighted", C.c_int), ("pure", C.c_int),
                ("kind", C.c_int), ("n", C.c_longlong), ("pot", C.c_void_p), ("poff", C.c_void_p),
                ("egval", C.c_void_p), ("egvar", C.c_void_p), ("ecval", C.c_void_p),
                ("wf", C.c_void_p), ("gam", C.c_void_p), ("nscale", C.c_void_p)]


_lib = None


def available() -> bool:
    return os.path.exists(LIB)


def load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB)
        lib.vo_grad_pass.restype = C.c_int
        lib.vo_grad_pass.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_longlong, C.POINTER(VoGroup), C.c_int, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_int]
        lib.vo_max_threads.restype = C.c_int
        _lib = lib
    return _lib


class CModel:
    """Lowered model pinned in host memory in the layout vi_port.c reads."""

    def __init__(self, model):
        self.model = model
        self.keep = []
        qx, qw = quadrature(model.T)
        self.quad = np.ascontiguousarray(np.concatenate([qx, qw]))
        self.ptab = np.ascontiguousarray(model.ptab, dtype=np.float64)
        self.groups = (VoGroup * len(model.groups))()
        for d, g in zip(self.groups, model.groups):
            d.nd, d.nc, d.ng, d.ne = g.nd, g.nc, g.ng, g.ne
            for i in range(MAX_AXES):
                d.dims[i] = int(g.dims[i]) if i < len(g.dims) else 0
            d.node, d.weighted, d.pure, d.n = int(g.node), int(g.weighted), int(g.pure), int(g.n)
            d.kind = int(getattr(g, "kind", 0))
            for name, arr, dt in (("pot", g.pot, np.int32), ("poff", g.poff, np.int32),
                                  ("egval", g.egval, np.float64), ("egvar", g.egvar, np.float64),
                                  ("ecval", g.ecval, np.float64), ("wf", g.wf, np.float64),
                                  ("gam", g.gam, np.float64), ("nscale", g.nscale, np.float64)):
                a = np.ascontiguousarray(arr, dtype=dt)
                self.keep.append(a)
                setattr(d, name, a.ctypes.data if a.size else None)

    def grad_pass(self, eta, w, threads=0):
        m = self.model
        eta = np.ascontiguousarray(eta, dtype=np.float64)
        w = np.ascontiguousarray(w, dtype=np.float64)
        grad = np.zeros(m.n_param)
        g_w = np.zeros(m.K)
        energy = np.zeros(1)
        rc = load().vo_grad_pass(m.K, m.T, self.quad.ctypes.data, self.ptab.ctypes.data, eta.ctypes.data,
                                 w.ctypes.data, m.n_param, self.groups, len(m.groups), grad.ctypes.data,
                                 g_w.ctypes.data, energy.ctypes.data, int(threads))
        if rc != 0:
            raise ValueError("vi_port.c: model exceeds a compiled-in limit")
        return grad, g_w, float(energy[0])


class CRunner:
    """One VI iteration = C pass over the records (all host threads) + the Adam step."""

    def __init__(self, model, eta, tau, w_tau, threads=0):
        self.cm = CModel(model)
        self.cores = int(threads) if threads else int(load().vo_max_threads())
        self.describe = f"C/OpenMP fp64 port, {self.cores} threads"
        self.vi = NumpyVI(model)
        self.vi.eta[:], self.vi.tau[:], self.vi.w_tau = eta, tau, w_tau
        self.vi.refresh()
        vi, cm, cores = self.vi, self.cm, self.cores

        def gradients():
            g, gw, e = cm.grad_pass(vi.eta, vi.w, cores)
            return (*tau_gradients(model, g, gw, vi.eta, vi.w), e)
        vi.gradients = gradients

    def step(self, lr):
        return self.vi.adam_step(lr)
