"""ctypes binding of ``liblhvi.so`` (declarations mirror ``include/lhvi.h`` one to one).

There is no fallback: if the library is missing or a symbol is absent, importing the engines
fails with an explicit error instead of computing anything on the CPU.
"""
from __future__ import annotations

import ctypes as C
import os

LHVI_F32, LHVI_F64 = 0, 1
LHVI_MAX_AXES = 6
LHVI_MAX_K = 8
LHVI_MAX_T = 32
LHVI_PARTIAL_ROWS = 1184
LHVI_FOLD_TILE = 1024
LHVI_MAX_PEERS = 16
LHVI_RUN_MAX_HUBS = 16
LHVI_IPC_HANDLE_BYTES = 64
ABI_VERSION = 8

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblhvi.so")

SYMBOLS = ("lhvi_last_error", "lhvi_abi_version", "lhvi_has_specialisation",
           "lhvi_factor_expect_grad", "lhvi_elbo_reduce", "lhvi_step_tick",
           "lhvi_param_step", "lhvi_mixture_belief", "lhvi_mixture_map", "lhvi_finish", "lhvi_state_pack", "lhvi_state_unpack", "lhvi_finish_step",
           "lhvi_peer_alloc", "lhvi_peer_open", "lhvi_peer_close", "lhvi_peer_free",
           "lhvi_iterate", "lhvi_iterate_supported", "lhvi_iterate_blocks", "lhvi_category_grad_reference",
           "lhvi_gabp_sweeps", "lhvi_gabp_marginals")


class LhviGroup(C.Structure):
    _fields_ = [
        ("nd", C.c_int32), ("nc", C.c_int32), ("ng", C.c_int32), ("ne", C.c_int32),
        ("dims", C.c_int32 * LHVI_MAX_AXES),
        ("node", C.c_int32), ("weighted", C.c_int32), ("hub_mask", C.c_int32), ("pure", C.c_int32),
        ("n", C.c_int64),
        ("pot", C.c_void_p), ("poff", C.c_void_p),
        ("egval", C.c_void_p), ("egvar", C.c_void_p), ("ecval", C.c_void_p),
        ("wf", C.c_void_p), ("gam", C.c_void_p), ("nscale", C.c_void_p),
        ("fold", C.c_void_p), ("n_pad", C.c_int64),
        ("run_start", C.c_void_p), ("run_key", C.c_void_p), ("run_hid", C.c_void_p), ("hub_keys", C.c_void_p),
        ("n_runs", C.c_int64), ("n_hubs", C.c_int32), ("run_hub_arg", C.c_int32),
        ("iter_blocks", C.c_int32), ("no_category_grad", C.c_int32),
        ("pot_kind", C.c_int32), ("reserved0", C.c_int32),
        ("run_node", C.c_void_p), ("run_una_pot", C.c_void_p), ("run_una_w", C.c_void_p),
        ("cst_q", C.c_void_p), ("cst_wf", C.c_void_p), ("cst_n", C.c_int64),
    ]


class LhviGabp(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("parity", C.c_int32), ("n_vars", C.c_int64), ("n_edges", C.c_int64),
        ("src", C.c_void_p), ("dst", C.c_void_p), ("rev", C.c_void_p), ("coef", C.c_void_p),
        ("jd", C.c_void_p), ("hd", C.c_void_p),
        ("P", C.c_void_p), ("H", C.c_void_p), ("SP", C.c_void_p), ("SH", C.c_void_p),
    ]


class LhviModel(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32), ("K", C.c_int32), ("T", C.c_int32), ("rule_symmetric", C.c_int32),
        ("n_param", C.c_int64),
        ("quad", C.c_void_p), ("ptab", C.c_void_p), ("eta", C.c_void_p), ("w", C.c_void_p),
        ("grad", C.c_void_p), ("partials", C.c_void_p), ("quad_host", C.c_double * (2 * LHVI_MAX_T)),
    ]


class LhviExchange(C.Structure):
    _fields_ = [
        ("world", C.c_int32), ("rank", C.c_int32), ("blocks", C.c_int32), ("reserved", C.c_int32),
        ("n_idx", C.c_int64),
        ("idx", C.c_void_p),
        ("recv", C.c_void_p * LHVI_MAX_PEERS),
        ("flags", C.c_void_p * LHVI_MAX_PEERS),
        ("seq", C.c_void_p), ("status", C.c_void_p),
    ]


LHVI_MAX_DSTATES = 16


class LhviH2(C.Structure):
    _fields_ = [
        ("dvals", (C.c_double * LHVI_MAX_DSTATES) * LHVI_MAX_AXES),
        ("xmap", ((C.c_int32 * LHVI_MAX_DSTATES) * LHVI_MAX_AXES) * LHVI_MAX_AXES),
    ]


class LhviOptim(C.Structure):
    _fields_ = [
        ("n_vars", C.c_int64), ("n_owned", C.c_int64),
        ("var_kind", C.c_void_p), ("var_dim", C.c_void_p), ("var_off", C.c_void_p),
        ("tau", C.c_void_p), ("mom1", C.c_void_p), ("mom2", C.c_void_p), ("wstate", C.c_void_p),
        ("step", C.c_void_p), ("sm_count", C.c_void_p),
        ("lr", C.c_double), ("b1", C.c_double), ("b2", C.c_double), ("eps", C.c_double),
        ("var_threshold", C.c_double),
        ("sgd", C.c_int32), ("reserved", C.c_int32),
        ("accum", C.c_void_p), ("trace", C.c_void_p),
    ]


class LhviError(RuntimeError):
    pass


_lib = None


def load(build_if_missing: bool = False):
    """Load (once) and type the library.  Raises ``LhviError`` when it cannot be loaded."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH) and build_if_missing:
        from . import build as _build
        _build.build()
    if not os.path.exists(LIB_PATH):
        raise LhviError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(there is no CPU fallback for the variational-inference hot path)")
    lib = C.CDLL(LIB_PATH)
    missing = [s for s in SYMBOLS if not hasattr(lib, s)]
    if missing:
        raise LhviError(f"{LIB_PATH} does not export {missing}")
    lib.lhvi_last_error.restype = C.c_char_p
    lib.lhvi_last_error.argtypes = []
    lib.lhvi_abi_version.restype = C.c_int
    lib.lhvi_abi_version.argtypes = []
    if lib.lhvi_abi_version() != ABI_VERSION:
        raise LhviError(f"ABI version mismatch: library {lib.lhvi_abi_version()}, binding {ABI_VERSION}")
    lib.lhvi_has_specialisation.restype = C.c_int
    lib.lhvi_has_specialisation.argtypes = [C.POINTER(LhviModel), C.POINTER(LhviGroup)]
    lib.lhvi_factor_expect_grad.restype = C.c_int
    lib.lhvi_factor_expect_grad.argtypes = [C.POINTER(LhviModel), C.POINTER(LhviGroup), C.c_int64,
                                            C.c_int, C.c_void_p]
    lib.lhvi_elbo_reduce.restype = C.c_int
    lib.lhvi_elbo_reduce.argtypes = [C.POINTER(LhviModel), C.c_int64, C.c_void_p]
    lib.lhvi_step_tick.restype = C.c_int
    lib.lhvi_step_tick.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_void_p]
    lib.lhvi_param_step.restype = C.c_int
    lib.lhvi_param_step.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                    C.c_double, C.c_int, C.c_int, C.c_void_p]
    lib.lhvi_mixture_map.restype = C.c_int
    lib.lhvi_mixture_map.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lhvi_finish.restype = C.c_int
    lib.lhvi_finish.argtypes = [C.POINTER(LhviModel), C.c_int64, C.c_void_p, C.c_double, C.c_double,
                                C.POINTER(LhviExchange), C.c_void_p]
    lib.lhvi_peer_alloc.restype = C.c_int
    lib.lhvi_peer_alloc.argtypes = [C.c_int64, C.POINTER(C.c_void_p), C.c_char_p]
    lib.lhvi_peer_open.restype = C.c_int
    lib.lhvi_peer_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.lhvi_peer_close.restype = C.c_int
    lib.lhvi_peer_close.argtypes = [C.c_void_p]
    lib.lhvi_peer_free.restype = C.c_int
    lib.lhvi_peer_free.argtypes = [C.c_void_p]
    lib.lhvi_mixture_belief.restype = C.c_int
    lib.lhvi_mixture_belief.argtypes = [C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.lhvi_finish_step.restype = C.c_int
    lib.lhvi_finish_step.argtypes = [C.POINTER(LhviModel), C.c_int64, C.POINTER(LhviExchange), C.c_int64, C.c_int64,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double,
                                     C.c_double, C.c_int, C.c_void_p]
    lib.lhvi_iterate.restype = C.c_int
    lib.lhvi_iterate.argtypes = [C.POINTER(LhviModel), C.POINTER(LhviGroup), C.c_int32, C.POINTER(LhviExchange),
                                 C.POINTER(LhviOptim), C.c_int32, C.c_void_p]
    lib.lhvi_category_grad_reference.restype = C.c_int
    lib.lhvi_category_grad_reference.argtypes = [C.POINTER(LhviModel), C.POINTER(LhviGroup), C.POINTER(LhviH2), C.c_void_p]
    lib.lhvi_gabp_sweeps.restype = C.c_int
    lib.lhvi_gabp_sweeps.argtypes = [C.POINTER(LhviGabp), C.c_int32, C.c_void_p]
    lib.lhvi_gabp_marginals.restype = C.c_int
    lib.lhvi_gabp_marginals.argtypes = [C.POINTER(LhviGabp), C.c_void_p, C.c_void_p, C.c_void_p]
    for fn in (lib.lhvi_iterate_supported, lib.lhvi_iterate_blocks):
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(LhviModel), C.POINTER(LhviGroup), C.c_int32, C.POINTER(LhviExchange)]
    for fn in (lib.lhvi_state_pack, lib.lhvi_state_unpack):
        fn.restype = C.c_int
        fn.argtypes = [C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    _lib = lib
    return lib


def check(rc: int, lib=None):
    """Turn a negative return code into an exception (reference callers see RuntimeError /
    ValueError, SURVEY section 8 b "Errors")."""
    if rc == 0:
        return
    lib = lib or load()
    msg = lib.lhvi_last_error().decode("utf-8", "replace")
    if rc in (-1, -2):
        raise ValueError(f"lhvi: {msg}")
    raise LhviError(f"lhvi: {msg} (code {rc})")
