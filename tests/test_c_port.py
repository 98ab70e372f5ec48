"""``oracle/vi_port.c`` (plain C / OpenMP restatement, the CPU baseline of bench.py) against the
goldens generated from the reference -- same checks as the numpy port -- and against the numpy
port on the benchmark generators, single- and multi-threaded."""
import subprocess

import numpy as np
import pytest

import helpers
import lhvi_b200
import specs
from oracle import c_port
from oracle.vi_numpy import grad_pass, tau_gradients
from test_numpy_oracle import _setup


@pytest.fixture(scope="module", autouse=True)
def built():
    import os
    subprocess.check_call(["make", "-s", "-C", os.path.dirname(c_port.LIB).rsplit("/_build", 1)[0]])
    assert c_port.available()


@pytest.mark.parametrize("path", helpers.golden_files(), ids=helpers.golden_id)
def test_snapshot_against_reference_goldens(path, ns):
    name, engine, gold = helpers.load_golden(path)
    vi, model, rvs = _setup(name, engine, ns, gold)
    grad, g_w, energy = c_port.CModel(model).grad_pass(vi.eta, vi.w, threads=2)
    g_flat, g_wtau = tau_gradients(model, grad, g_w, vi.eta, vi.w)
    np.testing.assert_allclose(energy, gold["fe0_fixed"], rtol=1e-9)
    np.testing.assert_allclose(g_wtau, gold["gw0_fixed"], rtol=1e-9, atol=1e-12)
    want = gold["grad0_fixed"]
    got = helpers.rows_from_flat(model, g_flat, rvs, engine, want.shape[1])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-11)


@pytest.mark.parametrize("threads", [1, 3])
def test_matches_numpy_port_on_generators(threads):
    syn = lhvi_b200.synthetic
    for model in (syn.relational_hybrid(200, 4, 3, 3, seed=2, weighted=True), syn.gaussian_grid(8, 2, 5)):
        eta, _, _ = syn.random_state(model, 1)
        w = np.linspace(1.0, 2.0, model.K)
        w /= w.sum()
        a = grad_pass(model, eta, w)
        b = c_port.CModel(model).grad_pass(eta, w, threads=threads)
        np.testing.assert_allclose(b[0], a[0], rtol=1e-10, atol=1e-11)
        np.testing.assert_allclose(b[1], a[1], rtol=1e-11)
        np.testing.assert_allclose(b[2], a[2], rtol=1e-11)


def test_runner_steps_like_numpy_runner():
    from oracle import cpu_port
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(60, 3, 2, 3, seed=5)
    eta, tau, w_tau = syn.random_state(model, 3)
    fast = cpu_port.make_runner(model, eta, tau, w_tau)
    assert isinstance(fast, c_port.CRunner)
    slow = cpu_port.NumpyRunner(model, eta, tau, w_tau)
    for _ in range(3):
        np.testing.assert_allclose(fast.step(0.1), slow.step(0.1), rtol=1e-11)
    np.testing.assert_allclose(fast.vi.eta, slow.vi.eta, rtol=1e-9, atol=1e-11)
