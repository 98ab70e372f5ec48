"""Golden vectors for Gaussian belief propagation from the UNMODIFIED reference ``GaBP`` (GaBP.py).

Run in the build container only (needs /root/reference):

    python tests/golden/make_gabp_golden.py

The model (``specs.gabp_grid``) is a loopy 4 x 4 grid: ``GaussianPotential`` and ``XYPotential`` edges
between hidden variables, one ``LinearGaussianPotential`` observation and one ``X2Potential`` prior per
variable.  ``tests/golden/gabp_grid.json`` stores the belief parameters ``get_belief_params`` returns
after ``run(n)`` for several n (the flooding schedule makes them depend on n until convergence) and
the exact posterior means / variances of the same Gaussian for reference.  Only numbers are stored.
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LHVI_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

import collections  # noqa: E402
import collections.abc  # noqa: E402

collections.MutableSet = collections.abc.MutableSet
if not hasattr(np, "Inf"):
    np.Inf = np.inf                                      # GaBP.py:2 on numpy 2

import GaBP as ref_gabp  # noqa: E402
import make_golden  # noqa: E402  (reference_namespace)
import specs  # noqa: E402

ITERATIONS = (1, 2, 3, 6, 12, 40)


def main():
    out = {"iterations": list(ITERATIONS), "mean": {}, "var": {}}
    for n in ITERATIONS:
        g, rvs = specs.gabp_grid(make_golden.reference_namespace())
        bp = ref_gabp.GaBP(g)
        with contextlib.redirect_stdout(io.StringIO()):
            bp.run(n)
        params = [bp.get_belief_params(rv) for rv in rvs if rv.value is None]
        out["mean"][str(n)] = [float(np.asarray(m).reshape(-1)[0]) for m, _ in params]
        out["var"][str(n)] = [float(np.asarray(v).reshape(-1)[0]) for _, v in params]
    with open(os.path.join(HERE, "gabp_grid.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("gabp_grid: means after 40 iterations", np.round(out["mean"]["40"][:4], 6), "...")


if __name__ == "__main__":
    main()
