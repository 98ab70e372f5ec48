"""Ground the models of relational_specs.py with the UNMODIFIED reference RelationalGraph and store
what it produced -- the ground atoms, every ground factor (parametric factor index + argument keys)
and the evidence assignment -- in relational_golden.json.  Build container only:

    python tests/golden/make_relational_golden.py
"""
import collections
import collections.abc
import json
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LHVI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")
collections.MutableSet = collections.abc.MutableSet      # OrderedSet.py:5 on py>=3.10
if not hasattr(np, "Inf"):
    np.Inf = np.inf

import Graph as ref_graph  # noqa: E402
import MLNPotential as ref_mln  # noqa: E402
import Potential as ref_pot  # noqa: E402
import RelationalGraph as ref_rel  # noqa: E402
import relational_specs  # noqa: E402

ns = types.SimpleNamespace()
for mod in (ref_graph, ref_pot, ref_mln, ref_rel):
    for name in dir(mod):
        if not name.startswith("_"):
            setattr(ns, name, getattr(mod, name))

out = {}
for name, builder in relational_specs.RELATIONAL.items():
    rel, data = builder(ns)
    g, rvs_dict = rel.ground_graph()
    rel.add_evidence(data)
    key_of = {id(rv): key for key, rv in rvs_dict.items()}
    pf_of = {id(pf.potential): i for i, pf in enumerate(rel.param_factors)}
    factors = sorted([pf_of[id(f.potential)], [list(key_of[id(rv)]) for rv in f.nb]] for f in g.factors)
    rvs = sorted([list(key), rv.value] for key, rv in rvs_dict.items())
    out[name] = {"factors": factors, "rvs": rvs, "degree": sorted([list(key), len(rv.nb)] for key, rv in rvs_dict.items())}
    print(name, len(rvs), "ground atoms,", len(factors), "ground factors")
with open(os.path.join(HERE, "relational_golden.json"), "w") as f:
    json.dump(out, f)
