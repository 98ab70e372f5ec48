"""``lifting_torch``: the lifting passes of a coarse-to-fine round on resident arrays
against the host passes (``lifting.py`` + ``lhvi_lift.cpp``) -- same class ids, same statistics, same record
columns, same runs.  The torch passes run on the CPU here and on the GPU under ``-m gpu``."""
import numpy as np
import pytest
import torch

import lhvi_b200
import specs
from lhvi_b200 import lifting, lifting_torch as lt, synthetic as syn
from oracle_engine import OracleEngine


def _models(ns):
    yield "relational", syn.relational_hybrid_arrays(1500, 5, seed=1)
    yield "kalman", syn.kalman_arrays(30, 10, levels=3, seed=2)[0]
    for name in ("hmln_evidence", "rgm_split", "chain_table", "smokers", "ring_xy"):
        g, _ = specs.CASES[name][0](ns)
        yield name, lifting.arrays_from_graph(g)[0]


def test_mix_has_the_bits_of_the_host_hash():
    x = np.array([0, 1, 2, 12345678901234, 2 ** 62 + 17], dtype=np.int64)
    for seed in (1, 0xA4093822299F31D0, 0x452821E638D01377 + 259):
        want = lifting._mix(x, seed & (2 ** 64 - 1)).view(np.int64)
        assert np.array_equal(lt.mix(torch.as_tensor(x), seed).numpy(), want)


@pytest.mark.parametrize("split", [True, False])
def test_colour_passing_gives_the_host_library_s_class_ids(ns, split):
    for name, ga in _models(ns):
        start = lifting.initial_colouring(ga, split_cont_evidence=split)
        v0, f0, s0 = lifting.colour_passing(ga, start=start, use_native=True)
        v1, f1, s1 = lt.colour_passing(lt.TorchGraph(ga, "cpu"), torch.as_tensor(start))
        assert s0 == s1, name
        assert np.array_equal(v0, v1.numpy()), name
        assert all(np.array_equal(a, b.numpy()) for a, b in zip(f0, f1)), name


@pytest.mark.parametrize("split", [True, False])
def test_initial_colouring_matches_the_host_pass(ns, split):
    for name, ga in _models(ns):
        tg = lt.TorchGraph(ga, "cpu")
        cont, _ = lt.domain_tables(ga.domains, "cpu")
        want = lifting.initial_colouring(ga, split_cont_evidence=split)
        assert np.array_equal(lt.initial_colouring(tg, cont, split).numpy(), want), name


def test_colour_passing_on_random_graphs_gives_the_host_library_s_ids():
    """Random small factor graphs: several blocks of arity 1-4, symmetric and ordered potentials, repeated
    rows, blocks sharing a potential object, mixed evidence (``tests/test_lifting.py`` runs the same graphs
    through the two host implementations)."""
    class Stub:
        def __init__(self, symmetric):
            self.symmetric = symmetric
    doms = [lhvi_b200.Graph.Domain((-5, 5), continuous=True), lhvi_b200.Graph.Domain((0, 1))]
    for seed in range(60):
        rng = np.random.default_rng(seed)
        nv = int(rng.integers(3, 60))
        var_dom = rng.integers(0, 2, nv).astype(np.int32)
        val = np.full(nv, np.nan)
        ev = rng.random(nv) < 0.4
        val[ev] = np.where(var_dom[ev] == 0, rng.integers(0, 3, ev.sum()) * 0.5, rng.integers(0, 2, ev.sum()))
        blocks = []
        for _ in range(int(rng.integers(1, 5))):
            arity, n = int(rng.integers(1, 5)), int(rng.integers(1, 40))
            base = rng.integers(0, nv, (max(1, n // 3), arity))
            blocks.append(lifting.FactorBlock(Stub(bool(rng.integers(0, 2))), base[rng.integers(0, base.shape[0], n)].astype(np.int64)))
        if rng.random() < 0.3 and len(blocks) > 1:
            blocks[1] = lifting.FactorBlock(blocks[0].potential, rng.integers(0, nv, (5, blocks[0].arity)).astype(np.int64))
        ga = lifting.GroundArrays(doms, var_dom, val, blocks)
        tg = lt.TorchGraph(ga, "cpu")
        cont, _ = lt.domain_tables(doms, "cpu")
        for split in (True, False):
            v0, f0, s0 = lifting.colour_passing(ga, split_cont_evidence=split, use_native=True)
            v1, f1, s1 = lt.colour_passing(tg, lt.initial_colouring(tg, cont, split))
            assert s0 == s1 and np.array_equal(v0, v1.numpy()), (seed, split)
            assert all(np.array_equal(a, b.numpy()) for a, b in zip(f0, f1)), (seed, split)
            st0, st1 = lifting.class_stats(ga, v0), lt.class_stats(tg, v1)
            assert np.array_equal(st0["rep"], st1["rep"].numpy()) and np.array_equal(st0["degree"], st1["degree"].numpy())


def test_rank_first_survives_a_hash_collision():
    cols = [torch.tensor([5, 7, 5, 9, 7, 5]), torch.tensor([1, 1, 2, 1, 1, 1])]
    ids, n, first = lt.rank_first(torch.zeros(6, dtype=torch.int64), cols)          # every row "collides"
    assert ids.tolist() == [0, 1, 2, 3, 1, 0] and n == 4 and first.tolist() == [0, 1, 2, 3]
    # with a second hash as the verifier: a disagreement sends the rows to the exact ranking as well
    h2 = lt.mix(cols[0], 11) ^ lt.mix(cols[1], 13)
    ids, n, first = lt.rank_first(torch.zeros(6, dtype=torch.int64), lambda: cols, h2)
    assert ids.tolist() == [0, 1, 2, 3, 1, 0] and n == 4
    ids, n, _ = lt.rank_first(lt.mix(cols[0], 5) + cols[1], cols, h2)                    # no collision: fast path
    assert ids.tolist() == [0, 1, 2, 3, 1, 0] and n == 4


def test_statistics_layout_and_lowering_match_the_host_passes(ns):
    for name, ga in _models(ns):
        for split, gaussian in ((False, True), (True, False)):
            tg = lt.TorchGraph(ga, "cpu")
            start = lifting.initial_colouring(ga, split_cont_evidence=split)
            v0, f0, _ = lifting.colour_passing(ga, start=start, use_native=True)
            vt, ft = torch.as_tensor(v0), [torch.as_tensor(f) for f in f0]
            st0, st1 = lifting.class_stats(ga, v0), lt.class_stats(tg, vt)
            for k in ("size", "rep", "hidden", "degree", "dom"):
                assert np.array_equal(np.asarray(st0[k]), st1[k].numpy()), (name, k)
            for k in ("mean", "variance"):
                np.testing.assert_allclose(st0[k], st1[k].numpy(), rtol=1e-12, atol=1e-14)
            m0 = lifting.lower_partition(ga, v0, f0, 2, 3, gaussian_obs=gaussian)
            m1 = lt.lower_partition(tg, ga, vt, ft, 2, 3, gaussian_obs=gaussian)
            assert m0.n_param == m1.n_param and np.array_equal(m0.var_off, m1.var_off)
            assert np.array_equal(m0.var_kind, m1.var_kind) and np.array_equal(m0.var_dim, m1.var_dim)
            assert np.array_equal(m0.ptab, m1.ptab) and np.array_equal(m0.slot_class, m1.slot_class)
            assert [g.signature for g in m0.groups] == [g.signature for g in m1.groups], name
            for a, b in zip(m0.groups, m1.groups):
                for k in ("pot", "poff", "wf", "gam", "nscale"):
                    assert np.array_equal(getattr(a, k), getattr(b, k)), (name, a.signature, k)
                for k in ("egval", "egvar", "ecval"):
                    np.testing.assert_allclose(getattr(a, k), getattr(b, k), rtol=1e-12, atol=1e-14)


@pytest.mark.parametrize("native", [True, False])
def test_evidence_split_matches_the_host_pass(native):
    for seed in range(3):
        ga = syn.relational_hybrid_arrays(1200 + 300 * seed, 5, seed=seed)
        tg = lt.TorchGraph(ga, "cpu")
        vi = lifting.C2FArrayVI(ga, 2, 3, use_native=native)
        vi.vcol = lifting.initial_colouring(ga, split_cont_evidence=False)
        n0 = int(vi.vcol.max()) + 1
        vi.may_split = np.zeros(n0, dtype=bool)
        vi.may_split[vi.vcol[~vi.hidden & vi.cont_dom]] = True
        vi.ev_has, vi.ev_val = np.zeros(n0, dtype=bool), np.zeros(n0)
        state = [torch.as_tensor(a.copy()) for a in (vi.vcol, vi.may_split, vi.ev_has, vi.ev_val)]
        eps = float(np.sqrt(vi._evidence_stats()[2].max()))
        step = eps / 6
        for _ in range(6):
            eps = max(eps - step, 0.0)
            vi._split_evidence(eps)
            state = list(lt.split_evidence(tg, *state, eps, 2, 10))
            assert np.array_equal(vi.vcol, state[0].numpy()) and np.array_equal(vi.may_split, state[1].numpy())
            assert np.array_equal(vi.ev_has, state[2].numpy())
            np.testing.assert_allclose(vi.ev_val, state[3].numpy(), rtol=1e-12, atol=1e-13)


def _same_run(a, b):
    assert [n for n, _ in a.history] == [n for n, _ in b.history]
    np.testing.assert_allclose([f for _, f in a.history], [f for _, f in b.history], rtol=1e-10)
    assert np.array_equal(a.vcol, b.vcol) and all(np.array_equal(x, y) for x, y in zip(a.fcols, b.fcols))
    np.testing.assert_allclose(a.P, b.P, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(a.map_values(), b.map_values(), rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("its", [50, 100])
def test_coarse_to_fine_run_with_resident_passes_equals_the_host_route(its):
    ga = syn.relational_hybrid_arrays(400, 4, seed=its)
    runs = [lifting.C2FArrayVI(ga, 2, 3, engine_factory=lambda m: OracleEngine(m, var_threshold=0.1), device_passes=dp)
            .run(its, 0.1, log_fe=True) for dp in (False, "cpu")]
    _same_run(*runs)


def test_resident_passes_on_a_hybrid_object_graph(ns):
    g, _ = specs.CASES["hmln_evidence"][0](ns)
    ga = lifting.arrays_from_graph(g)[0]
    runs = [lifting.C2FArrayVI(ga, 2, 3, engine_factory=lambda m: OracleEngine(m, var_threshold=0.1), device_passes=dp)
            .run(30, 0.1, log_fe=True) for dp in (False, "cpu")]
    _same_run(*runs)


def test_lifted_engine_with_resident_passes_equals_the_host_route():
    ga = syn.kalman_arrays(30, 10, levels=3, seed=2)[0]
    a, b = [lifting.ArrayVI(ga, 2, 3, engine_factory=lambda m: OracleEngine(m, var_threshold=0.1), device_passes=dp)
            for dp in (False, "cpu")]
    assert np.array_equal(a.quotient.var_colour, b.quotient.var_colour) and np.array_equal(a.class_rep, b.class_rep)
    for ga_, gb_ in zip(a.model.groups, b.model.groups):
        assert ga_.signature == gb_.signature and np.array_equal(ga_.poff, gb_.poff) and np.array_equal(ga_.wf, gb_.wf)
    a.run(5, 0.1)
    b.run(5, 0.1)
    np.testing.assert_allclose(a.free_energy(), b.free_energy(), rtol=1e-12)


@pytest.mark.gpu
def test_coarse_to_fine_on_the_gpu_with_resident_passes_equals_the_host_route():
    ga = syn.relational_hybrid_arrays(20000, 6, seed=3)
    runs = [lifting.C2FArrayVI(ga, 3, 3, dtype="float64", device_passes=dp).run(50, 0.05, log_fe=True) for dp in (False, True)]
    a, b = runs
    assert [n for n, _ in a.history] == [n for n, _ in b.history]
    np.testing.assert_allclose([f for _, f in a.history], [f for _, f in b.history], rtol=1e-9)
    assert np.array_equal(a.vcol, b.vcol) and all(np.array_equal(x, y) for x, y in zip(a.fcols, b.fcols))
    np.testing.assert_allclose(a.P, b.P, rtol=1e-8, atol=1e-10)


@pytest.mark.gpu
def test_resident_passes_on_the_gpu_with_discrete_evidence_and_hidden_discrete_variables(ns):
    """Object graphs with observed and hidden discrete variables (table / MLN potentials) through the array
    engines: the resident passes on the GPU give the host route's partition, records and run."""
    for name, its in (("hmln_evidence", 30), ("chain_table", 20), ("rgm_split", 30)):
        g, _ = specs.CASES[name][0](ns)
        ga = lifting.arrays_from_graph(g)[0]
        a, b = [lifting.C2FArrayVI(ga, 2, 3, dtype="float64", device_passes=dp).run(its, 0.1, log_fe=True) for dp in (False, True)]
        assert [n for n, _ in a.history] == [n for n, _ in b.history], name
        np.testing.assert_allclose([f for _, f in a.history], [f for _, f in b.history], rtol=1e-9, err_msg=name)
        assert np.array_equal(a.vcol, b.vcol), name
        np.testing.assert_allclose(a.P, b.P, rtol=1e-8, atol=1e-10, err_msg=name)
        la, lb = [lifting.ArrayVI(ga, 2, 3, dtype="float64", device_passes=dp) for dp in (False, True)]
        assert np.array_equal(la.quotient.var_colour, lb.quotient.var_colour), name
        la.run(5, 0.1)
        lb.run(5, 0.1)
        np.testing.assert_allclose(la.free_energy(), lb.free_energy(), rtol=1e-10, err_msg=name)
