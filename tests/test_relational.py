"""RelationalGraph drop-in (object and array-native grounding) against what the unmodified
reference's RelationalGraph.ground_graph / add_evidence produced on the same models
(tests/golden/relational_golden.json, written by make_relational_golden.py)."""
import json
import os
import types

import numpy as np
import pytest

import lhvi_b200
import relational_specs
from oracle.vi_numpy import grad_pass

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "relational_golden.json")))
lifting = lhvi_b200.lifting


@pytest.fixture(scope="module")
def rns():
    ns = types.SimpleNamespace()
    for mod in (lhvi_b200.Graph, lhvi_b200.Potential, lhvi_b200.MLNPotential, lhvi_b200.RelationalGraph):
        for name in dir(mod):
            if not name.startswith("_"):
                setattr(ns, name, getattr(mod, name))
    return ns


@pytest.mark.parametrize("name", sorted(relational_specs.RELATIONAL))
def test_object_grounding_matches_reference(name, rns):
    rel, data = relational_specs.RELATIONAL[name](rns)
    g, rvs_dict = rel.ground_graph()
    rel.add_evidence(data)
    key_of = {id(rv): key for key, rv in rvs_dict.items()}
    pf_of = {id(pf.potential): i for i, pf in enumerate(rel.param_factors)}
    factors = sorted([pf_of[id(f.potential)], [list(key_of[id(rv)]) for rv in f.nb]] for f in g.factors)
    assert factors == GOLD[name]["factors"]
    assert sorted([list(k), rv.value] for k, rv in rvs_dict.items()) == GOLD[name]["rvs"]
    assert sorted([list(k), len(rv.nb)] for k, rv in rvs_dict.items()) == GOLD[name]["degree"]


@pytest.mark.parametrize("name", sorted(relational_specs.RELATIONAL))
def test_array_grounding_matches_reference(name, rns):
    rel, data = relational_specs.RELATIONAL[name](rns)
    ga, index = rel.ground_arrays(data)
    assert len(index) == ga.n_vars == len(GOLD[name]["rvs"])
    factors = sorted([bi, [list(index.key_of(int(v))) for v in row]]
                     for bi, b in enumerate(ga.blocks) for row in b.args)
    assert factors == GOLD[name]["factors"]
    for key, value in GOLD[name]["rvs"]:
        i = index.index_of(tuple(key))
        assert i >= 0 and index.key_of(i) == tuple(key)
        if value is None:
            assert np.isnan(ga.var_value[i])
        else:
            assert ga.var_value[i] == value
    deg = ga.degrees()
    for key, d in GOLD[name]["degree"]:
        assert deg[index.index_of(tuple(key))] == d
    assert index.index_of(("nosuchatom",) if False else (GOLD[name]["rvs"][0][0][0], "zzz")) == -1


@pytest.mark.parametrize("name", sorted(relational_specs.RELATIONAL))
def test_both_routes_lift_to_the_same_model(name, rns):
    """Array grounding -> colour passing -> lowering gives the same partition sizes and the same
    free energy as the object grounding pushed through the object-level CompressedGraph."""
    K, T = 2, 3
    rel, data = relational_specs.RELATIONAL[name](rns)
    ga, index = rel.ground_arrays(data)
    g, rvs_dict = rel.ground_graph()
    rel.add_evidence(data)
    cg = lhvi_b200.CompressedGraphWithObs.CompressedGraph(g)
    cg.run()
    vcol, fcols, _ = lifting.colour_passing(ga)
    want = {frozenset(index.index_of(k) for k, rv in rvs_dict.items() if rv.cluster is c) for c in cg.rvs}
    got = {}
    for i, c in enumerate(vcol):
        got.setdefault(int(c), set()).add(i)
    assert {frozenset(s) for s in got.values()} == want
    m_arr, q = lifting.lower_lifted(ga, K, T)
    m_obj = lhvi_b200.lowering.lower_compressed(cg, K, T)
    w = np.array([0.35, 0.65])

    def tied(model, rep):
        eta = np.zeros(model.n_param)
        for h, off, kind, dim in zip(model.handles, model.var_off, model.var_kind, model.var_dim):
            r = np.random.default_rng(rep(h))
            if kind == 0:
                eta[off:off + 2 * K:2] = r.uniform(-1, 1, K)
                eta[off + 1:off + 2 * K:2] = r.uniform(0.5, 2, K)
            else:
                p = r.uniform(0.1, 1, (K, dim))
                eta[off:off + K * dim] = (p / p.sum(axis=1, keepdims=True)).reshape(-1)
        return eta
    key_index = {id(rv): index.index_of(k) for k, rv in rvs_dict.items()}
    e_arr = grad_pass(m_arr, tied(m_arr, lambda h: h.rep), w)
    e_obj = grad_pass(m_obj, tied(m_obj, lambda h: min(key_index[id(rv)] for rv in h.rvs)), w)
    np.testing.assert_allclose(e_arr[2], e_obj[2], rtol=1e-12)
    np.testing.assert_allclose(e_arr[1], e_obj[1], rtol=1e-12)


def test_array_grounding_scales(rns):
    """100 categories x 10 banks (the reference's RGM demo size) and 300 x 40: cross products on
    index arrays, no object per ground atom."""
    for nc, nb in ((100, 10), (300, 40)):
        rel, _ = relational_specs.rgm_relational(rns, nc, nb)
        ga, index = rel.ground_arrays({("market", "c1"): 1.0})
        assert ga.n_vars == 1 + nc + nc * nb + nb
        assert ga.n_factors == nc + 2 * nc * nb
        assert index.key_of(index.index_of(("loss", f"c{nc - 1}", f"b{nb - 1}"))) == ("loss", f"c{nc - 1}", f"b{nb - 1}")
