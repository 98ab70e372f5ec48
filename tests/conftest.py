import os
import sys
import types

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _gpu_unavailable():
    """Reason the ``gpu``-marked tests cannot run here, or None."""
    try:
        import torch
        if not torch.cuda.is_available():
            return "needs a CUDA device"
    except Exception as exc:           # pragma: no cover
        return f"torch unavailable: {exc}"
    return None      # with a GPU present a missing liblhvi.so must FAIL the tests, not skip them


def pytest_collection_modifyitems(config, items):
    """Without a CUDA device the gpu-marked tests are skipped, not failed: the product path itself
    refuses to run there (no CPU fallback)."""
    reason = _gpu_unavailable()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def repo_namespace():
    """Data-model + potential classes of this repo, in the shape ``specs`` builders expect."""
    import lhvi_b200
    ns = types.SimpleNamespace()
    for mod in (lhvi_b200.Graph, lhvi_b200.Potential, lhvi_b200.MLNPotential):
        for name in dir(mod):
            if not name.startswith("_"):
                setattr(ns, name, getattr(mod, name))
    return ns


@pytest.fixture(scope="session")
def ns():
    return repo_namespace()
