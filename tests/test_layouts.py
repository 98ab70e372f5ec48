"""Host-side record layouts the engine builds for the two large kernels (no GPU needed):
``engine.run_layout`` (run-major columns of lhvi_group) and ``engine.align_runs`` (hub runs of a
streamed group padded to whole tiles).  Both must leave every sum unchanged -- checked with the
numpy oracle on the re-laid-out model -- and satisfy the invariants lhvi.h documents."""
import dataclasses

import numpy as np
import pytest

import lhvi_b200
from lhvi_b200 import engine as eng_mod
from oracle.vi_numpy import grad_pass

syn = lhvi_b200.synthetic


def _full_group(model):
    return next(i for i, g in enumerate(model.groups) if not g.node and not g.pure and g.nc == 2)


def _streamed_group(model):
    return next(i for i, g in enumerate(model.groups) if g.pure and g.nc == 1 and g.ne == 1)


@pytest.mark.parametrize("P,G,K", [(300, 6, 3), (50, 3, 1), (2000, 10, 2)])
def test_run_layout_invariants_and_sums(P, G, K):
    model = syn.relational_hybrid(P, G, K, 3, seed=1, weighted=True)
    gi = _full_group(model)
    g = model.groups[gi]
    out = eng_mod.run_layout(g, K, 4)
    assert out is not None
    sg, starts, run_key, hid, hubs, hub_arg = out
    run_arg = 1 - hub_arg
    assert sg.n == g.n and starts[0] == 0 and starts[-1] == g.n and (np.diff(starts) > 0).all()
    assert hubs.size <= 16 and (np.diff(hubs) > 0).all()
    np.testing.assert_array_equal(hubs[hid], sg.poff[hub_arg])
    # every run is on one variable, and no run is longer than the split length
    for a, b, key in zip(starts[:-1], starts[1:], run_key):
        assert (sg.poff[run_arg, a:b] == key).all()
    assert np.diff(starts).max() <= eng_mod.RUN_SPLIT
    assert (np.diff(sg.poff[run_arg]) >= 0).all()
    # the re-ordered group sums to the same gradients
    eta, tau, w_tau = syn.random_state(model, 0)
    w = np.full(K, 1.0 / K)
    groups = list(model.groups)
    groups[gi] = sg
    a = grad_pass(model, eta, w)
    b = grad_pass(dataclasses.replace(model, groups=groups), eta, w)
    np.testing.assert_allclose(b[0], a[0], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(b[1], a[1], rtol=1e-12)
    np.testing.assert_allclose(b[2], a[2], rtol=1e-12)


def test_run_layout_declines_groups_it_cannot_serve():
    grid = syn.gaussian_grid(8, 2, 3)
    for g in grid.groups:
        assert eng_mod.run_layout(g, 2, 4) is None            # both arguments take many values
    model = syn.relational_hybrid(100, 4, 3, 3, seed=0)
    for g in model.groups:
        if g.node or g.pure:
            assert eng_mod.run_layout(g, 3, 4) is None
    # too many hubs for the thread-private accumulators in fp64
    big = syn.relational_hybrid(400, 12, 3, 3, seed=0)
    g = big.groups[_full_group(big)]
    assert eng_mod.run_layout(g, 3, 4) is not None
    assert eng_mod.run_layout(g, 3, 8) is None


def test_small_shards_get_shorter_runs():
    model = syn.relational_hybrid(3000, 10, 3, 3, seed=0)
    g = model.groups[_full_group(model)]
    starts = eng_mod.run_layout(g, 3, 4)[1]
    assert np.diff(starts).max() == 2          # far fewer records than resident threads: runs are cut


@pytest.mark.parametrize("P,G", [(30_000, 3), (9_000, 2)])
def test_align_runs_pads_long_runs_to_whole_tiles(P, G):
    K, tile = 3, 1024
    model = syn.relational_hybrid(P, G, K, 3, seed=0, weighted=True)
    gi = _streamed_group(model)
    g = model.groups[gi]
    null_pot = model.ptab.size
    ptab = np.concatenate([model.ptab, np.zeros(6)])
    pg = eng_mod.align_runs(g, null_pot, tile)
    assert pg.n % tile == 0 and pg.n >= g.n
    key = pg.poff[0]
    starts = np.concatenate([[0], np.flatnonzero(key[1:] != key[:-1]) + 1])
    lens = np.diff(np.append(starts, pg.n))
    long_runs = lens >= eng_mod.FOLD_ALIGN_MIN_TILES * tile
    assert (starts[long_runs] % tile == 0).all() and long_runs.any()
    assert pg.n - g.n < tile * (long_runs.sum() + 1)
    # null records: zero coefficients and zero weights
    null = pg.pot == null_pot
    assert null.sum() == pg.n - g.n and (pg.wf[null] == 0).all() and (pg.gam[:, null] == 0).all()
    eta, tau, w_tau = syn.random_state(model, 0)
    w = np.full(K, 1.0 / K)
    groups = list(model.groups)
    groups[gi] = pg
    a = grad_pass(model, eta, w)
    b = grad_pass(dataclasses.replace(model, groups=groups, ptab=ptab), eta, w)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(np.asarray(x), np.asarray(y))
    # also when the group ships no weight columns: the null potential alone must silence the records
    unweighted = syn.relational_hybrid(P, G, K, 3, seed=0, weighted=False)
    g2 = unweighted.groups[_streamed_group(unweighted)]
    pg2 = eng_mod.align_runs(g2, null_pot, tile)
    pg2 = dataclasses.replace(pg2, wf=np.ones(pg2.n), gam=np.ones((1, pg2.n)))     # what the kernels assume
    groups = list(unweighted.groups)
    groups[gi] = pg2
    a = grad_pass(unweighted, eta, w)
    b = grad_pass(dataclasses.replace(unweighted, groups=groups, ptab=ptab), eta, w)
    np.testing.assert_allclose(b[0], a[0], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(b[2], a[2], rtol=1e-12)


def test_compat_reference_tables(ns):
    """compat="reference" (SURVEY hazard H2): the lowering carries the domain values of the hidden discrete
    arguments, `engine.h2_tables` turns them into the value -> state maps of `lhvi_h2`, and refuses the
    shapes in which the unmodified reference's pairing depends on the argument order."""
    import specs
    from lhvi_b200.engine import h2_tables
    g, _ = specs.hmln_hidden(ns)
    model = lhvi_b200.lowering.lower_ground(g, 2, 3)
    full = [gr for gr in model.groups if gr.nd > 0 and not gr.node and not gr.pure]
    assert full
    for gr in full:
        assert len(gr.dvals) == gr.nd and all(tuple(v) == (0.0, 1.0) for v in gr.dvals)
        h = h2_tables(gr)
        for a in range(gr.nd):
            assert [h.dvals[a][j] for j in range(2)] == [0.0, 1.0]
            for b in range(gr.nd):
                if b != a:
                    assert [h.xmap[a][b][j] for j in range(2)] == [0, 1]
        assert gr.take(np.arange(gr.n)[::-1]).dvals == gr.dvals
    d3 = ns.Domain((0, 1, 2))
    dc = ns.Domain((-5, 5), continuous=True)
    a, x = ns.RV(d3), ns.RV(dc)
    g2 = ns.Graph()
    g2.rvs, g2.factors = {a, x}, {ns.F(ns.MLNPotential(lambda v: -(v[0] - v[1]) ** 2, 1.0), [a, x])}
    g2.init_nb()
    gr = [q for q in lhvi_b200.lowering.lower_ground(g2, 2, 3).groups if q.nd and not q.node][0]
    with pytest.raises(NotImplementedError):
        h2_tables(gr)


@pytest.mark.parametrize("weighted", [True, False])
def test_fused_records_are_the_records_that_left_their_groups(weighted):
    """``engine.fuse_run_extras`` / ``engine.fuse_constants`` (host logic; the kernels are checked under -m gpu):
    every record that leaves a group reappears exactly once as a per-run column of the run-major group (on the
    first piece of its variable's run) or as a folded constant of the streamed group, and nothing else moves."""
    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(3000, 6, 3, 3, seed=5, weighted=weighted)
    layouts = [eng_mod.run_layout(g, 3, 4) for g in model.groups]
    gi = next(i for i, l in enumerate(layouts) if l is not None)
    groups, extras = eng_mod.fuse_run_extras(model.groups, layouts)
    node, una_pot, una_w, counts = extras[gi]
    run_key = layouts[gi][2]
    assert len(groups) == len(model.groups)
    # node records: (slot offset, W, gradient scale) of what left = of what arrived
    for orig, left in zip(model.groups, groups):
        if orig.node and orig.nc == 1:
            carried = np.flatnonzero((node[0] != 0) | (node[1] != 0))
            assert counts[0] == carried.size == orig.n - left.n
            assert np.unique(run_key[carried]).size == carried.size            # once per variable
            gone = ~np.isin(orig.poff[0], left.poff[0])
            a = sorted(zip(orig.poff[0][gone], orig.wf[gone], orig.nscale[gone]))
            b = sorted(zip(run_key[carried], node[0][carried], node[1][carried]))
            assert a == b
        elif orig.pure and not orig.node and orig.nc == 1 and orig.ne == 0:
            carried = np.flatnonzero(una_pot >= 0)
            assert counts[1] == carried.size == orig.n - left.n
            gone = ~np.isin(orig.poff[0], left.poff[0])
            a = sorted(zip(orig.poff[0][gone], orig.pot[gone], orig.wf[gone], orig.gam[0][gone]))
            b = sorted(zip(run_key[carried], una_pot[carried], una_w[0][carried], una_w[1][carried]))
            assert a == b
        elif not (orig.node or orig.pure):
            assert left is orig or left.n == orig.n
    # pieces of a run that was cut carry nothing beyond the first one
    later = np.r_[False, run_key[1:] == run_key[:-1]]
    assert not np.any((node[0][later] != 0) | (node[1][later] != 0)) and np.all(una_pot[later] < 0)

    streamed = [bool(g.pure and not g.node and g.nc == 1 and g.ne == 1 and g.n > 1000) for g in groups]
    groups2, consts = eng_mod.fuse_constants(groups, streamed, np.asarray(model.ptab, dtype=float))
    (target, (q, wf)), = consts.items()
    assert streamed[target]
    orig = next(g for g in groups if g.pure and not g.node and g.nc == 0 and g.ne == 1 and g.n)
    assert q.size == orig.n and all(g.n == 0 for g in groups2 if g.pure and not g.node and g.nc == 0)
    coef = np.asarray(model.ptab)[orig.pot.astype(np.int64)[:, None] + np.arange(3)[None, :]]
    x = orig.ecval[0]
    np.testing.assert_allclose(q, coef[:, 0] + x * (coef[:, 1] + coef[:, 2] * x), rtol=1e-12, atol=1e-13)
    assert (wf is None) == (not weighted) and (wf is None or np.array_equal(wf, orig.wf))
