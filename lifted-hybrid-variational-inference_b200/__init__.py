"""B200-native variational-inference update loop (drop-in for the reference's
``VarInference`` / ``LiftedVarInference`` / ``C2FVarInference``).

Import through the repo-root alias: ``import lhvi_b200``.
"""
import importlib
import sys

__all__ = ["Graph", "Potential", "MLNPotential", "CompressedGraphWithObs", "RelationalGraph", "KalmanFilter",
           "lowering", "lifting", "install_flat_aliases"]

_FLAT = ("Graph", "Potential", "MLNPotential", "CompressedGraphWithObs", "RelationalGraph", "KalmanFilter", "utils",
         "VarInference", "LiftedVarInference", "C2FVarInference", "GaBP")


def __getattr__(name):
    # lazy submodule access: lhvi_b200.VarInference, lhvi_b200.lowering, ...
    try:
        return importlib.import_module(f"{__name__}.{name}")
    except ModuleNotFoundError as exc:
        raise AttributeError(name) from exc


def install_flat_aliases():
    """Register the reference's flat module names in ``sys.modules``."""
    for name in _FLAT:
        sys.modules[name] = importlib.import_module(f"{__name__}.{name}")
