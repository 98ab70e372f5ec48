"""Parity at BASELINE config 5's own size (1 M entities x 10 groups: 11.0 M factor records, K = 3, T = 3), where
the loop oracles cannot go: properties that do not depend on the size.

* the template-specialised kernels (run-major, streaming, record-major) against the generic kernel -- the
  point-by-point restatement that the small cases pin to the oracle -- on the SAME 11.0 M records;
* the oracle itself on a random sample of the records of every group, taken out of the full model with its
  parameter vector (large offsets, the full model's slots);
* additivity over the records (the free energy, G_w and every gradient are sums over records: two halves
  of the record table add up to the whole) and independence of the record order;
* linearity in the lifted weights (W_f, gamma, node scales doubled: everything doubles);
* a step of the optimiser on the persistent and on the per-group path from the same state.
"""
import dataclasses

import numpy as np
import pytest

import lhvi_b200
from oracle.vi_numpy import grad_pass

pytestmark = pytest.mark.gpu

P, G, K, T = 1_000_000, 10, 3, 3


@pytest.fixture(scope="module")
def full():
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(P, G, K, T, seed=0, order="hub", weighted=True)
    eta, tau, w_tau = syn.random_state(model, 0)
    return model, eta, tau, w_tau


def _pass(model, state, dtype="float32", **kw):
    from lhvi_b200.engine import DeviceEngine
    eta, tau, w_tau = state
    eng = DeviceEngine(model, dtype=dtype, **kw)
    eng.set_state(eta, tau, w_tau)
    grad, g_w, energy = eng.gradients()
    out = np.array(grad, dtype=np.float64), np.array(g_w, dtype=np.float64), float(energy)
    eng.close()
    return out


def _close(a, b, rel, hub_rel=None, model=None):
    """``rel`` of the largest entry for gradients, G_w and energy.  ``hub_rel``: the ten group-level variables
    collect ~10^6 records each (gradients of 5e5 .. 5e6 next to 36 for the largest other entry), so they are
    compared on their own scale: the specialised float kernels (per-thread and per-block partial sums) reach
    2e-6 of the value there, the generic float kernel -- one scalar atomicAdd per record into the same
    address -- 4e-3 (``tools/fp32_hub_probe.py``)."""
    ga, gb = np.array(a[0]), np.array(b[0])
    if hub_rel is not None:
        hubs = _hub_elements(model)
        np.testing.assert_allclose(ga[hubs], gb[hubs], rtol=hub_rel, atol=hub_rel * float(np.abs(gb[hubs]).max()))
        ga[hubs] = gb[hubs] = 0.0
    scale = max(1.0, float(np.abs(gb).max()))
    np.testing.assert_allclose(ga, gb, rtol=rel, atol=rel * scale)
    np.testing.assert_allclose(a[1], b[1], rtol=rel, atol=rel * max(1.0, float(np.abs(b[1]).max())))
    np.testing.assert_allclose(a[2], b[2], rtol=rel)


def _hub_elements(model):
    """Parameter elements of the variables that more than 10^5 records touch."""
    count = np.zeros(model.n_param, dtype=np.int64)
    for g in model.groups:
        for a in range(g.nh):
            count += np.bincount(g.poff[a], minlength=model.n_param)
    offs = np.flatnonzero(count > 100_000)
    return (offs[:, None] + np.arange(2 * model.K)[None, :]).reshape(-1)


def test_specialised_kernels_equal_the_generic_kernel_on_all_records(full):
    model, *state = full
    assert model.n_records == 11_000_090
    assert _hub_elements(model).size == G * 2 * K
    # in double precision the two paths evaluate the same expressions
    exact = _pass(model, state, "float64")
    _close(exact, _pass(model, state, "float64", force_generic=True), 1e-9)
    # in single precision both are compared with the double result
    _close(_pass(model, state), exact, 3e-5, hub_rel=2e-5, model=model)
    _close(_pass(model, state, force_generic=True), exact, 3e-5, hub_rel=1e-2, model=model)


def test_oracle_on_a_sample_of_every_group(full):
    model, *state = full
    rng = np.random.default_rng(1)
    groups = [g.take(np.sort(rng.choice(g.n, size=min(g.n, 3000), replace=False))) for g in model.groups]
    sample = dataclasses.replace(model, groups=groups)
    eta, tau, w_tau = state
    w = np.e ** w_tau / (np.e ** w_tau).sum()
    want = grad_pass(sample, eta, w)
    got = _pass(sample, state, "float64")
    _close(got, (want[0], want[1], want[2]), 1e-9)
    _close(_pass(sample, state, "float32"), (want[0], want[1], want[2]), 3e-5)


def test_the_sums_are_additive_over_the_records_and_independent_of_their_order(full):
    model, *state = full
    whole = _pass(model, state, "float64")
    halves = [_pass(model.shard(r, 2), state, "float64", shard=False) for r in (0, 1)]
    _close(tuple(a + b for a, b in zip(*halves)), whole, 1e-9)
    rng = np.random.default_rng(2)
    shuffled = dataclasses.replace(model, groups=[g.take(rng.permutation(g.n)) for g in model.groups])
    _close(_pass(shuffled, state, "float64"), whole, 1e-9)
    _close(_pass(shuffled, state, "float32"), whole, 3e-5, hub_rel=2e-5, model=model)


def test_everything_is_linear_in_the_lifted_weights(full):
    model, *state = full
    doubled = dataclasses.replace(model, groups=[dataclasses.replace(g, wf=2 * g.wf, gam=2 * g.gam, nscale=2 * g.nscale)
                                                  for g in model.groups])
    one, two = _pass(model, state, "float64"), _pass(doubled, state, "float64")
    _close(two, tuple(2 * x for x in one), 1e-12)


def test_persistent_and_per_group_paths_take_the_same_steps(full, monkeypatch):
    from lhvi_b200.engine import DeviceEngine
    model, eta, tau, w_tau = full
    out = []
    for mode in ("0", "1"):
        monkeypatch.setenv("LHVI_PERSISTENT", mode)
        eng = DeviceEngine(model, dtype="float32")
        eng.set_state(eta, tau, w_tau)
        eng.reset_moments()
        eng.iterate(3, 0.1)
        assert eng.persistent() == (mode == "1")
        out.append((eng.get_state()[0], eng.last_free_energy()))
        eng.close()
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=2e-4, atol=2e-4)
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=1e-5)
