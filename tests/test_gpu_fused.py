"""Records of a run-major group's run variables evaluated inside the run-major kernel (``engine.fuse_run_extras``,
``lhvi_group::run_node / run_una_pot / run_una_w``): the node-entropy record and one pure unary factor per
variable leave their own groups; constant records (no integrated argument) ride with the streamed group
(``engine.fuse_constants``, ``lhvi_group::cst_*``).  Every sum must stay what it was."""
import numpy as np
import pytest

import lhvi_b200
from lhvi_b200 import lifting
from oracle.vi_numpy import grad_pass

pytestmark = pytest.mark.gpu


def _pass(model, state, dtype, monkeypatch, fuse, **kw):
    from lhvi_b200.engine import DeviceEngine
    monkeypatch.setenv("LHVI_FUSE_RUN_EXTRAS", "1" if fuse else "0")
    monkeypatch.setenv("LHVI_FUSE_CONSTANTS", "1" if fuse else "0")
    eng = DeviceEngine(model, dtype=dtype, **kw)
    eng.set_state(*state)
    grad, g_w, energy = eng.gradients()
    fused = {i: (k.get("fused"), k.get("fused_constants")) for i, (d, k, g) in enumerate(eng.groups)
             if k.get("fused") or k.get("fused_constants")}
    sizes = [g.n for _, _, g in eng.groups]
    out = np.array(grad, dtype=np.float64), np.array(g_w, dtype=np.float64), float(energy)
    eng.close()
    return out, fused, sizes


@pytest.mark.parametrize("weighted", [True, False])
@pytest.mark.parametrize("K", [1, 2, 3])
def test_fused_and_separate_records_give_the_same_sums(monkeypatch, weighted, K):
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(3000, 6, K, 3, seed=K, weighted=weighted)
    state = syn.random_state(model, 3)
    w = np.e ** state[2] / (np.e ** state[2]).sum()
    want = grad_pass(model, state[0], w)
    for dtype, tol in (("float64", 1e-10), ("float32", 3e-5)):
        a, fa, na = _pass(model, state, dtype, monkeypatch, True)
        b, fb, nb = _pass(model, state, dtype, monkeypatch, False)
        assert fa and not fb and sum(na) < sum(nb)             # the fusion happened: records left their groups
        assert any(c for _, c in fa.values())                  # ... the constant records among them
        scale = max(1.0, np.abs(want[0]).max())
        for got in (a, b):
            np.testing.assert_allclose(got[0], want[0], rtol=tol, atol=tol * scale)
            np.testing.assert_allclose(got[1], want[1], rtol=tol, atol=tol * np.abs(want[1]).max())
            np.testing.assert_allclose(got[2], want[2], rtol=tol)


def test_fused_records_of_a_lifted_model(monkeypatch):
    """Coarse-to-fine partition of the relational model: node scales, class sizes and neighbour counts differ
    from one record to the next (W != gradient scale on the node records)."""
    syn = lhvi_b200.synthetic
    ga = syn.relational_hybrid_arrays(4000, 6, seed=2)
    start = lifting.initial_colouring(ga, split_cont_evidence=True)
    # a partial refinement: observed entities rounded to one decimal share a class
    ga.var_value[:] = np.where(np.isnan(ga.var_value), np.nan, np.round(ga.var_value, 1))
    start = lifting.initial_colouring(ga, split_cont_evidence=True)
    vcol, fcols, _ = lifting.colour_passing(ga, start=start)
    model = lifting.lower_partition(ga, vcol, fcols, 3, 3)
    state = syn.random_state(model, 1)
    w = np.e ** state[2] / (np.e ** state[2]).sum()
    want = grad_pass(model, state[0], w)
    a, fa, _ = _pass(model, state, "float64", monkeypatch, True)
    b, fb, _ = _pass(model, state, "float64", monkeypatch, False)
    assert fa and not fb
    for got in (a, b):
        np.testing.assert_allclose(got[0], want[0], rtol=1e-10, atol=1e-10 * max(1.0, np.abs(want[0]).max()))
        np.testing.assert_allclose(got[1], want[1], rtol=1e-10)
        np.testing.assert_allclose(got[2], want[2], rtol=1e-10)


def test_fused_records_through_iterations_on_both_paths(monkeypatch):
    from lhvi_b200.engine import DeviceEngine
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(5000, 8, 3, 3, seed=4, weighted=True)
    eta, tau, w_tau = syn.random_state(model, 2)
    out = {}
    for fuse in ("1", "0"):
        for persistent in ("1", "0"):
            monkeypatch.setenv("LHVI_FUSE_RUN_EXTRAS", fuse)
            monkeypatch.setenv("LHVI_FUSE_CONSTANTS", fuse)
            monkeypatch.setenv("LHVI_PERSISTENT", persistent)
            eng = DeviceEngine(model, dtype="float64")
            eng.set_state(eta, tau, w_tau)
            eng.reset_moments()
            eng.iterate(6, 0.1)
            out[(fuse, persistent)] = (eng.get_state()[0], float(eng.free_energy()))
            eng.close()
    base = out[("0", "0")]
    for key, (e, fe) in out.items():
        np.testing.assert_allclose(e, base[0], rtol=1e-9, atol=1e-11, err_msg=str(key))
        np.testing.assert_allclose(fe, base[1], rtol=1e-11, err_msg=str(key))
