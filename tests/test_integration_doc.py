"""The binding stub printed in INTEGRATION.md for the host-side lifting library is executed as
written (only the library path is filled in) and must give the partition of the object-level colour
passing on golden graphs."""
import os
import re

import pytest

import lhvi_b200
import specs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("case", ["rgm_split", "smokers", "hmln_evidence", "ring_xy", "edge_mix"])
def test_lift_stub_from_the_integration_notes(case, ns):
    from lhvi_b200 import _lift_native, build
    build.build_lift()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"# lift_stub.py.*?```", text, flags=re.S).group(0)[:-3]
    scope = {}
    exec(code.replace("/path/to/liblhvi_lift.so", _lift_native.LIB_PATH), scope)
    g, _ = specs.CASES[case][0](ns)
    order = list(g.rvs)
    vcol, fcols = scope["colour_passing"](g)
    got = {}
    for rv, c in zip(order, vcol):
        got.setdefault(int(c), set()).add(id(rv))
    cg = lhvi_b200.CompressedGraphWithObs.CompressedGraph(g)
    cg.run()
    assert {frozenset(s) for s in got.values()} == {frozenset(id(rv) for rv in c.rvs) for c in cg.rvs}
    assert sum(len(c) for c in fcols) == len(g.factors)
