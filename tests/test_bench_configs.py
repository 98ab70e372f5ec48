"""The auxiliary bench workloads are the models they claim to be (CPU, numpy oracle as the engine):
config 4's grid with observations has the sparse system J mu = h as its exact posterior, and K=1 VI
converges to its solution (reference cross-check: Demo/RGM/RGMKLDivergence.py:40-46)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import bench_configs
from oracle.vi_numpy import NumpyVI

import lhvi_b200


def test_config4_grid_model_matches_its_linear_system():
    model, J, h = bench_configs.grid_with_observations(7, 1, 3, seed=3)
    exact = np.linalg.solve(J.toarray(), h)
    assert np.allclose(J.toarray(), J.toarray().T) and np.all(np.linalg.eigvalsh(J.toarray()) > 0)
    eta, tau, w_tau = lhvi_b200.synthetic.random_state(model, 0)
    vi = NumpyVI(model)
    vi.eta[:], vi.tau[:], vi.w_tau = eta, tau, w_tau
    vi.refresh()
    for _ in range(1500):
        vi.adam_step(0.05)
    np.testing.assert_allclose(vi.eta[model.var_off], exact, atol=2e-3)


def test_config_tables_name_every_baseline_config():
    assert sorted(bench_configs.ALL) == ["1", "2", "3", "4"]
