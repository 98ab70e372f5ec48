"""Import alias for the package directory ``lifted-hybrid-variational-inference_b200/``.

The directory name is fixed by the project layout and is not a valid Python identifier, so
``import lhvi_b200`` loads it under this name.  ``install_flat_aliases()`` additionally
registers the reference's flat module names (``VarInference``, ``LiftedVarInference``,
``C2FVarInference``, ``Graph``, ``Potential``, ``MLNPotential``, ``CompressedGraphWithObs``,
``utils``) so that reference scripts (``from VarInference import VarInference``) run
unchanged on this implementation.
"""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)),
                        "lifted-hybrid-variational-inference_b200")
_NAME = "lhvi_b200"

_spec = importlib.util.spec_from_file_location(
    _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
_module = importlib.util.module_from_spec(_spec)
sys.modules[_NAME] = _module
_spec.loader.exec_module(_module)
