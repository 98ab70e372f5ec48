"""Fixture for the reference's relational Kalman filter demo (Demo/RKF/LRKFDemoCycle.py: two groups of
three wells, 6 state variables x 20 steps, dense transition matrix, five parameter settings).  Build
container only: imports the UNMODIFIED reference from /root/reference and reads its data files
(Demo/Data/RKF/well_t.mat, cluster_NcutDiscrete.mat, LRKF_cycle.mat).

Written to rkf_cycle.json:
  data         the 6 x 20 observation slice the demo feeds to KalmanFilter.grounded_graph (after its
               own preprocessing of well_t.mat)
  param        3 x 5: transition variance, observation variance, transition coefficient per setting
  lrkf_res     6 x 5: the means at the last step that the demo compares against (LRKF_cycle.mat 'res',
               computed outside the repository; the demo prints the average error, it does not assert)
  exact        per setting: exact posterior means of the last step's variables on the reference's own
               ground graph (solve of J mu = h assembled from its potentials' quadratic parameters),
               and the graph's size
  tree         the same four entries (data 76 x 20, param, lrkf_res 76 x 5, exact) for
               Demo/RKF/LRKFDemoTree.py: the 76 wells of cluster 1 with a diagonal transition matrix
  lvi          for the settings in RUN: free energy and last-step means of the reference's
               LiftedVarInference(g, 1, 3) after 400 Adam iterations at lr 0.1, numpy seed 0 (K=1 over
               Gaussian factors is convex: the end point does not depend on the draw)
"""
import collections
import collections.abc
import contextlib
import io
import json
import os
import sys
from concurrent.futures import ProcessPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("LHVI_REFERENCE", "/root/reference")
RUN = (0, 3)
T = 20


def setup():
    import numpy as np
    collections.MutableSet = collections.abc.MutableSet
    np.Inf = np.inf
    sys.path.insert(0, REF)
    import scipy.io
    d = os.path.join(REF, "Demo", "Data", "RKF")
    cluster_mat = scipy.io.loadmat(os.path.join(d, "cluster_NcutDiscrete.mat"))["NcutDiscrete"]
    well_t = scipy.io.loadmat(os.path.join(d, "well_t.mat"))["well_t"]
    mat = scipy.io.loadmat(os.path.join(d, "LRKF_cycle.mat"))
    # the demo's own preprocessing (LRKFDemoCycle.py:17-41)
    idx = np.where(cluster_mat[:, 1] == 1)[0]
    cluster_mat[idx[3:], 1] = 0
    idx = np.where(cluster_mat[:, 2] == 1)[0]
    cluster_mat[idx[:49], 2] = 0
    cluster_mat[idx[52:], 2] = 0
    well_t = well_t[:, 199:]
    well_t[well_t[:, 0] == 5000, 0] = 0
    well_t[well_t == 5000] = 1
    rvs_id = np.concatenate([np.where(cluster_mat[:, i] == 1)[0] for i in (1, 2)], axis=None)
    return well_t[rvs_id, :T].astype(float), mat["param"].astype(float), mat["res"].astype(float)


def graph_for(i, data, param):
    import numpy as np
    from Graph import Domain
    from KalmanFilter import KalmanFilter
    from numpy import linspace
    n = data.shape[0]
    domain = Domain((-4, 4), continuous=True, integral_points=linspace(-4, 4, 30))
    kmf = KalmanFilter(domain, np.eye(n) * param[2, i] + 0.01, param[0, i], np.eye(n), param[1, i])
    return kmf.grounded_graph(T, data)


def exact_means(g, table):
    import numpy as np
    hidden = [rv for rv in g.rvs if rv.value is None]
    pos = {rv: j for j, rv in enumerate(hidden)}
    J = np.zeros((len(hidden), len(hidden)))
    h = np.zeros(len(hidden))
    for f in g.factors:
        A, b, _ = f.potential.get_quadratic_params()
        S = np.asarray(A, float) + np.asarray(A, float).T
        b = np.asarray(b, float).reshape(-1)
        for a, rv in enumerate(f.nb):
            if rv not in pos:
                continue
            h[pos[rv]] += b[a]
            for c, other in enumerate(f.nb):
                if other in pos:
                    J[pos[rv], pos[other]] -= S[a, c]
                else:
                    h[pos[rv]] += S[a, c] * other.value
    mu = np.linalg.solve(J, h)
    return [float(mu[pos[rv]]) for rv in table[T - 1]]


def lifted_run(i):
    import numpy as np
    data, param, _ = setup()
    g, table = graph_for(i, data, param)
    from LiftedVarInference import VarInference
    np.random.seed(0)
    vi = VarInference(g, 1, 3)
    with contextlib.redirect_stdout(io.StringIO()):
        vi.run(400, lr=0.1)
    return i, {"free_energy": float(vi.free_energy()),
               "means": [float(vi.eta[rv.cluster][0, 0]) for rv in table[T - 1]],
               "classes": len(vi.g.rvs)}


def setup_tree():
    """Demo/RKF/LRKFDemoTree.py: the 76 wells of cluster 1, diagonal transition matrix."""
    import numpy as np
    import scipy.io
    d = os.path.join(REF, "Demo", "Data", "RKF")
    cluster_mat = scipy.io.loadmat(os.path.join(d, "cluster_NcutDiscrete.mat"))["NcutDiscrete"]
    well_t = scipy.io.loadmat(os.path.join(d, "well_t.mat"))["well_t"]
    mat = scipy.io.loadmat(os.path.join(d, "LRKF_tree.mat"))
    well_t = well_t[:, 199:]
    well_t[well_t[:, 0] == 5000, 0] = 0
    well_t[well_t == 5000] = 1
    rvs_id = np.where(cluster_mat[:, 1] == 1)[0]
    return well_t[rvs_id, :T].astype(float), mat["param"].astype(float), mat["res"].astype(float)


def tree_part():
    import numpy as np
    from Graph import Domain
    from KalmanFilter import KalmanFilter
    from numpy import linspace
    data, param, res = setup_tree()
    n = data.shape[0]
    out = {"data": data.tolist(), "param": param.tolist(), "lrkf_res": res.tolist(), "exact": []}
    for i in range(param.shape[1]):
        domain = Domain((-4, 4), continuous=True, integral_points=linspace(-4, 4, 30))
        kmf = KalmanFilter(domain, np.eye(n) * param[2, i], param[0, i], np.eye(n), param[1, i])
        g, table = kmf.grounded_graph(T, data)
        out["exact"].append({"means": exact_means(g, table), "rvs": len(g.rvs), "factors": len(g.factors)})
    return out


def main():
    if "--tree-only" in sys.argv:          # add the tree demo to an existing fixture without the long runs
        setup()
        out = json.load(open(os.path.join(HERE, "rkf_cycle.json")))
        out["tree"] = tree_part()
        json.dump(out, open(os.path.join(HERE, "rkf_cycle.json"), "w"))
        return
    data, param, res = setup()
    out = {"data": data.tolist(), "param": param.tolist(), "lrkf_res": res.tolist(), "exact": [], "lvi": {}}
    for i in range(param.shape[1]):
        g, table = graph_for(i, data, param)
        out["exact"].append({"means": exact_means(g, table), "rvs": len(g.rvs), "factors": len(g.factors)})
    with ProcessPoolExecutor(len(RUN)) as pool:
        for i, r in pool.map(lifted_run, RUN):
            out["lvi"][str(i)] = r
            print(i, r["free_energy"], r["classes"], flush=True)
    out["tree"] = tree_part()
    json.dump(out, open(os.path.join(HERE, "rkf_cycle.json"), "w"))


if __name__ == "__main__":
    main()
