"""ctypes binding of ``liblhvi_lift.so`` (include/lhvi_lift.h): the host-side C++ colour passing.

``load()`` returns the library or ``None`` when it has not been built (``build.build_lift()``);
``lifting.colour_passing`` then uses its numpy implementation of the same passes.  Set
``LHVI_LIFT_NATIVE=0`` to force the numpy passes (the tests compare the two)."""
from __future__ import annotations

import ctypes as C
import os
import weakref

import numpy as np

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(PKG_DIR, "liblhvi_lift.so")
ABI_VERSION = 1
SYMBOLS = ("lhvi_lift_abi_version", "lhvi_lift_colour_passing", "lhvi_lift_graph_colour_passing", "lhvi_lift_graph_create",
           "lhvi_lift_graph_destroy", "lhvi_lift_rank64", "lhvi_lift_split_evidence")
ERRORS = {-1: "null pointer, negative size or class id out of range", -2: "factor arity outside 1..16",
          -3: "variable index out of range", -4: "out of memory", -5: "class arrays too short for the new classes",
          -6: "more than 2^31 - 1 variables or factors"}


class LiftBlock(C.Structure):
    _fields_ = [("args", C.c_void_p), ("n", C.c_int64), ("arity", C.c_int32), ("symmetric", C.c_int32),
                ("colour", C.c_void_p)]


_lib = None


def load(build_if_missing=False):
    global _lib
    if _lib is not None:
        return _lib
    if os.environ.get("LHVI_LIFT_NATIVE", "1") == "0":
        return None
    if not os.path.exists(LIB_PATH):
        if not build_if_missing:
            return None
        from . import build
        build.build_lift()
    lib = C.CDLL(LIB_PATH)
    lib.lhvi_lift_abi_version.restype = C.c_int32
    if lib.lhvi_lift_abi_version() != ABI_VERSION:
        raise RuntimeError(f"liblhvi_lift.so has ABI {lib.lhvi_lift_abi_version()}, the binding expects {ABI_VERSION}; "
                           "rebuild with lifted-hybrid-variational-inference_b200/build.py")
    lib.lhvi_lift_colour_passing.restype = C.c_int64
    lib.lhvi_lift_colour_passing.argtypes = [C.c_int64, C.c_void_p, C.POINTER(LiftBlock), C.c_int32, C.c_int32,
                                             C.POINTER(C.c_int32)]
    lib.lhvi_lift_graph_create.restype = C.c_void_p
    lib.lhvi_lift_graph_create.argtypes = [C.c_int64, C.POINTER(LiftBlock), C.c_int32, C.POINTER(C.c_int32)]
    lib.lhvi_lift_graph_destroy.restype = None
    lib.lhvi_lift_graph_destroy.argtypes = [C.c_void_p]
    lib.lhvi_lift_graph_colour_passing.restype = C.c_int64
    lib.lhvi_lift_graph_colour_passing.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p), C.c_int32,
                                                   C.POINTER(C.c_int32)]
    lib.lhvi_lift_rank64.restype = C.c_int64
    lib.lhvi_lift_rank64.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
    lib.lhvi_lift_split_evidence.restype = C.c_int64
    lib.lhvi_lift_split_evidence.argtypes = [C.c_int64, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_double, C.c_int32, C.c_int32]
    _lib = lib
    return lib


def _no_graph():
    return None


class PreparedGraph:
    """``lhvi_lift_graph`` of a ground graph: incidence lists and scratch buffers built once, reused
    by every ``colour_passing`` call on the same argument arrays."""

    def __init__(self, lib, n_vars, arg_arrays, symmetric):
        self.lib, self.n_vars = lib, int(n_vars)
        self.source = list(arg_arrays)                                                # what the caller passed
        self.args = [np.ascontiguousarray(a, dtype=np.int64) for a in arg_arrays]     # what the library reads
        self.symmetric = [int(bool(x)) for x in symmetric]
        descs = (LiftBlock * max(1, len(self.args)))()
        for d, a, sym in zip(descs, self.args, self.symmetric):
            d.args, d.n, d.arity, d.symmetric, d.colour = a.ctypes.data, a.shape[0], a.shape[1], sym, None
        status = C.c_int32(0)
        self.handle = lib.lhvi_lift_graph_create(self.n_vars, descs, len(self.args), C.byref(status))
        if not self.handle:
            raise ValueError(f"lhvi_lift_graph_create: {ERRORS.get(int(status.value), status.value)}")
        self._finalizer = weakref.finalize(self, lib.lhvi_lift_graph_destroy, self.handle)

    def __reduce__(self):          # a copied / pickled holder rebuilds its own graph on first use
        return (_no_graph, ())

    def matches(self, lib, n_vars, arg_arrays, symmetric):
        return (self.lib is lib and self.n_vars == int(n_vars) and len(self.source) == len(arg_arrays)
                and self.symmetric == [int(bool(x)) for x in symmetric]
                and all(a is b for a, b in zip(self.source, arg_arrays)))


def colour_passing(lib, start, blocks, max_sweeps, holder=None):
    """``blocks``: list of ``(args int64 [n, arity], symmetric, initial colour)``.  Returns
    ``(var_colour, [factor colour per block], sweeps)``.  ``holder`` (any object with a ``__dict__``,
    e.g. the ``GroundArrays``) keeps the prepared graph between calls on the same argument arrays
    (which must not be modified in place while it does)."""
    vcol = np.ascontiguousarray(start, dtype=np.int64).copy()
    arg_arrays = [b[0] for b in blocks]
    symmetric = [b[1] for b in blocks]
    graph = getattr(holder, "_lift_graph", None) if holder is not None else None
    if graph is None or not graph.matches(lib, vcol.size, arg_arrays, symmetric):
        graph = PreparedGraph(lib, vcol.size, arg_arrays, symmetric)
        if holder is not None:
            holder._lift_graph = graph
    colours = [np.full(a.shape[0], int(b[2]), dtype=np.int64) for a, b in zip(graph.args, blocks)]
    ptrs = (C.c_void_p * max(1, len(colours)))(*[c.ctypes.data for c in colours])
    sweeps = C.c_int32(0)
    rc = lib.lhvi_lift_graph_colour_passing(graph.handle, vcol.ctypes.data, ptrs, int(max_sweeps), C.byref(sweeps))
    if rc < 0:
        raise ValueError(f"lhvi_lift_graph_colour_passing: {ERRORS.get(int(rc), rc)}")
    return vcol, colours, int(sweeps.value)


def rank64(lib, key):
    key = np.ascontiguousarray(key).view(np.uint64)
    ids = np.empty(key.size, dtype=np.int64)
    n = lib.lhvi_lift_rank64(key.ctypes.data, key.size, ids.ctypes.data)
    if n < 0:
        raise ValueError(f"lhvi_lift_rank64: {ERRORS.get(int(n), n)}")
    return ids, int(n)


def split_evidence(lib, var_colour, value, n_classes, may_split, has_centroid, centroid, epsilon, k, iterations):
    """In place on ``var_colour`` (int64) and the three class arrays (uint8, uint8, float64, all of
    the same capacity); returns the new number of classes."""
    for a, dt in ((var_colour, np.int64), (value, np.float64), (may_split, np.uint8), (has_centroid, np.uint8),
                  (centroid, np.float64)):
        if a.dtype != dt or not a.flags.c_contiguous:
            raise TypeError("split_evidence: arrays must be contiguous int64 / float64 / uint8")
    n = lib.lhvi_lift_split_evidence(var_colour.size, var_colour.ctypes.data, value.ctypes.data, int(n_classes),
                                     may_split.size, may_split.ctypes.data, has_centroid.ctypes.data,
                                     centroid.ctypes.data, float(epsilon), int(k), int(iterations))
    if n < 0:
        raise ValueError(f"lhvi_lift_split_evidence: {ERRORS.get(int(n), n)}")
    return int(n)
