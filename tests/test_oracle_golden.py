"""The oracles against golden vectors produced by the unmodified reference.

``oracle/vi_loops.py`` (object level, live potential.get calls) must reproduce the
reference's free energy, gradients and Adam trajectory; with ``h2_compat=True`` it must
reproduce the *unmodified* reference even where its category-gradient bug fires, with
``h2_compat=False`` the in-memory-patched reference (see tests/golden/make_golden.py).
"""
import numpy as np
import pytest

import helpers
import specs
from oracle.vi_loops import LoopOracle

RTOL = 1e-9   # fp64 vs fp64, different summation order only


def _oracle(name, engine, ns, h2_compat, gold):
    builder, K, T, _ = specs.CASES[name]
    g, rvs = builder(ns)
    handles, factors, _ = helpers.setup_mode(g, engine)
    orc = LoopOracle(handles, factors, K, T, mode=engine, h2_compat=h2_compat)
    cont, disc = helpers.injected_params(handles, rvs, engine, K, int(gold["seed"]))
    orc.set_params(helpers.injected_w_tau(K), cont, disc)
    return orc, rvs, handles


def _rows(orc, rvs, engine, fn, width):
    out = np.full((len(rvs), width), np.nan)
    cache = {}
    for i, rv in enumerate(rvs):
        h = helpers.handle_of(rv, engine)
        if h.value is not None:
            continue
        if h not in cache:
            cache[h] = np.asarray(fn(h)).reshape(-1)
        out[i, :cache[h].size] = cache[h]
    return out


@pytest.mark.parametrize("path", helpers.golden_files(), ids=helpers.golden_id)
@pytest.mark.parametrize("variant", ["", "_fixed"])
def test_snapshot(path, variant, ns):
    name, engine, gold = helpers.load_golden(path)
    orc, rvs, _ = _oracle(name, engine, ns, h2_compat=(variant == ""), gold=gold)
    if engine != "ground":
        assert helpers.partition_of(rvs, engine) == helpers.partition_from_ids(gold["cluster0" + variant])
    np.testing.assert_allclose(orc.free_energy(), gold["fe0" + variant], rtol=RTOL)
    np.testing.assert_allclose(orc.gradient_w_tau(), gold["gw0" + variant], rtol=RTOL, atol=1e-12)
    want = gold["grad0" + variant]

    def grad(h):
        return orc.gradient_mu_var(h) if h.domain.continuous else orc.gradient_category_tau(h)
    got = _rows(orc, rvs, engine, grad, want.shape[1])
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-11)


@pytest.mark.parametrize("path", [p for p in helpers.golden_files() if not p.endswith("__c2f.npz")],
                         ids=helpers.golden_id)
@pytest.mark.parametrize("variant", ["", "_fixed"])
def test_adam_trajectory(path, variant, ns):
    name, engine, gold = helpers.load_golden(path)
    if name in helpers.DEMO_SIZED and variant == "_fixed":
        # the pure-Python loops take ~20-40 s per trajectory at the demos' sizes: this oracle walks the
        # unmodified reference's trajectory there; the patched one is walked by the vectorised and the
        # C oracles (test_numpy_oracle.py, test_c_port.py) and by this oracle's snapshot test
        pytest.skip("demo-sized golden: patched trajectory covered by the fast oracles")
    orc, rvs, _ = _oracle(name, engine, ns, h2_compat=(variant == ""), gold=gold)
    orc.init_adam()
    for _ in range(int(gold["steps"])):
        orc.adam_step(float(gold["lr"]))
    np.testing.assert_allclose(orc.w_tau, gold["w_tau1" + variant], rtol=1e-8, atol=1e-10)
    want = gold["eta1" + variant]
    got = _rows(orc, rvs, engine, lambda h: orc.eta[h], want.shape[1])
    np.testing.assert_allclose(got, want, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(orc.free_energy(), gold["fe1" + variant], rtol=1e-8)
