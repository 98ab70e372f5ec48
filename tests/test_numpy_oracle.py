"""``oracle/vi_numpy.py`` (fused single pass over the lowered SoA model) against the goldens
of the patched reference (intended maths).  Also covers ``lowering.py``: the potentials'
coefficient tables, argument canonicalisation, evidence folding and lifted weights all sit
between the object graph and these numbers."""
import numpy as np
import pytest

import helpers
import specs
from oracle.vi_numpy import NumpyVI


def _setup(name, engine, ns, gold):
    builder, K, T, _ = specs.CASES[name]
    g, rvs = builder(ns)
    handles, _, cg = helpers.setup_mode(g, engine)
    model = helpers.lower_for(engine, g, cg, K, T)
    cont, disc = helpers.injected_params(handles, rvs, engine, K, int(gold["seed"]))
    vi = NumpyVI(model)
    flat = helpers.flat_params(model, cont, disc)
    vi.eta[:] = flat
    vi.tau[:] = flat
    vi.w_tau = helpers.injected_w_tau(K)
    vi.refresh()
    return vi, model, rvs


@pytest.mark.parametrize("path", helpers.golden_files(), ids=helpers.golden_id)
def test_snapshot(path, ns):
    name, engine, gold = helpers.load_golden(path)
    vi, model, rvs = _setup(name, engine, ns, gold)
    g_flat, g_wtau, energy = vi.gradients()
    np.testing.assert_allclose(energy, gold["fe0_fixed"], rtol=1e-9)
    np.testing.assert_allclose(g_wtau, gold["gw0_fixed"], rtol=1e-9, atol=1e-12)
    want = gold["grad0_fixed"]
    got = helpers.rows_from_flat(model, g_flat, rvs, engine, want.shape[1])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("path", [p for p in helpers.golden_files() if not p.endswith("__c2f.npz")],
                         ids=helpers.golden_id)
def test_adam_trajectory(path, ns):
    name, engine, gold = helpers.load_golden(path)
    vi, model, rvs = _setup(name, engine, ns, gold)
    for _ in range(int(gold["steps"])):
        vi.adam_step(float(gold["lr"]))
    np.testing.assert_allclose(vi.w_tau, gold["w_tau1_fixed"], rtol=1e-8, atol=1e-10)
    want = gold["eta1_fixed"]
    got = helpers.rows_from_flat(model, vi.eta, rvs, engine, want.shape[1])
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(vi.gradients()[2], gold["fe1_fixed"], rtol=1e-8)
