"""compat="reference": the CUDA path reproduces the UNMODIFIED reference where its
``gradient_category_tau`` differs from the intended mathematics (SURVEY hazard H2,
``VarInference.py:147-150``: the other arguments' axes come from the discrete variable's own
domain).  Goldens without the ``_fixed`` suffix were produced by the unmodified classes."""
import numpy as np
import pytest

import helpers
import lhvi_b200
import specs
from oracle.vi_numpy import NumpyVI, tau_gradients
from test_dropin_host import ENGINE_CLASS, check_trajectory, make_injector

pytestmark = pytest.mark.gpu

# goldens in which H2 fires (hidden booleans next to hidden reals) and goldens in which both variants
# agree (equal-cardinality discrete arguments, no hidden discrete argument at all)
H2_GOLDENS = ["hmln_hidden__ground", "robot_like__ground", "hmln_demo__ground", "hmln_demo__lifted", "robot_demo__ground"]
AGREEING = ["smokers__ground", "smokers__lifted", "tri_table3__ground", "chain_table__lifted", "hmln_evidence__ground"]


def _setup(gid, ns):
    path = [p for p in helpers.golden_files() if helpers.golden_id(p) == gid][0]
    name, engine, gold = helpers.load_golden(path)
    builder, K, T, _ = specs.CASES[name]
    g, rvs = builder(ns)
    handles, _, cg = helpers.setup_mode(g, engine)
    model = helpers.lower_for(engine, g, cg, K, T)
    cont, disc = helpers.injected_params(handles, rvs, engine, K, int(gold["seed"]))
    tau = helpers.flat_params(model, cont, disc)
    ref = NumpyVI(model)
    ref.eta[:] = tau
    ref.tau[:] = tau
    ref.w_tau = helpers.injected_w_tau(K)
    ref.refresh()
    return name, engine, gold, model, rvs, ref


@pytest.mark.parametrize("gid", H2_GOLDENS + AGREEING)
@pytest.mark.parametrize("force_generic", [True, False], ids=["generic", "dispatch"])
def test_snapshot_matches_the_unmodified_reference(gid, force_generic, ns):
    from lhvi_b200.engine import DeviceEngine
    name, engine, gold, model, rvs, ref = _setup(gid, ns)
    eng = DeviceEngine(model, dtype="float64", compat="reference", force_generic=force_generic)
    eng.set_state(ref.eta, ref.tau, ref.w_tau)
    grad, g_w, energy = eng.gradients()
    np.testing.assert_allclose(energy, gold["fe0"], rtol=1e-9)
    g_flat, g_wtau = tau_gradients(model, grad, g_w, ref.eta, ref.w)
    np.testing.assert_allclose(g_wtau, gold["gw0"], rtol=1e-9, atol=1e-11)
    want = gold["grad0"]
    got = helpers.rows_from_flat(model, g_flat, rvs, engine, want.shape[1])
    np.testing.assert_allclose(got, want, rtol=1e-9, atol=1e-10)
    if gid in H2_GOLDENS:        # and it is NOT what the intended mathematics gives
        assert np.nanmax(np.abs(want - gold["grad0_fixed"])) > 1e-3


@pytest.mark.parametrize("gid", ["hmln_hidden__ground", "robot_like__ground", "hmln_demo__lifted"])
def test_trajectory_matches_the_unmodified_reference(gid, ns):
    """The drop-in classes with compat="reference" walk the unmodified reference's Adam trajectory."""
    import contextlib
    import io
    path = [p for p in helpers.golden_files() if helpers.golden_id(p) == gid][0]
    name, engine, gold = helpers.load_golden(path)
    builder, K, T, _ = specs.CASES[name]
    g, rvs = builder(ns)
    vi = ENGINE_CLASS[engine]()(g, K, T, compat="reference")
    vi.init_param = make_injector(vi, rvs, engine, K, int(gold["seed"]))
    with contextlib.redirect_stdout(io.StringIO()):
        vi.run(int(gold["steps"]), lr=float(gold["lr"]), is_log=False)
    # check_trajectory reads the *_fixed keys: hand it the unmodified reference's values under those names
    view = dict(gold)
    for k in ("w_tau1", "eta1", "cluster1", "evidence1", "fe1"):
        view[k + "_fixed"] = gold[k]
    check_trajectory(vi, rvs, engine, view, rtol=1e-7, atol=1e-9)
    assert abs(float(gold["fe1"]) - float(gold["fe1_fixed"])) > 1e-6 * abs(float(gold["fe1"]))


def test_compat_refuses_what_it_cannot_reproduce(ns):
    from lhvi_b200.engine import DeviceEngine
    d3 = ns.Domain((0, 1, 2))
    dc = ns.Domain((-5, 5), continuous=True)
    a, x = ns.RV(d3), ns.RV(dc)
    pot = ns.MLNPotential(lambda v: -(v[0] - v[1]) ** 2, 1.0)
    g = ns.Graph()
    g.rvs, g.factors = {a, x}, {ns.F(pot, [a, x])}
    g.init_nb()
    model = lhvi_b200.lowering.lower_ground(g, 2, 3)
    with pytest.raises(NotImplementedError):
        DeviceEngine(model, compat="reference")
    with pytest.raises(ValueError):
        DeviceEngine(model, compat="something else")
