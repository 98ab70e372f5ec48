"""Relational Kalman filter builder: same constructor and ``grounded_graph`` as the reference's
``KalmanFilter.py`` (``:7-104``), plus ``grounded_arrays`` for the array-native engines.

Model: states ``x_t`` (n per step), ``x_{t+1} = A^T x_t + noise(transition_variance)``, observation
``y_t = C x_t + noise(observation_variance)`` where a value is recorded.  The reference does not use
one Gaussian factor per transition; it expands ``sum_y (x_{t+1,y} - sum_x A[x,y] x_{t,x})^2`` into
unary ``X2`` and pairwise ``XY`` potentials so that equal coefficients share a potential (and a
colour).  The pieces, all with variance ``transition_variance``:

* ``X2(1)`` on every state of step t >= 1,
* ``XY(-2 A[x,y])`` on ``(x_{t,x}, x_{t+1,y})`` for every non-zero ``A[x,y]``,
* for 0 < t < T-1 (step 0 is observed, the last step has no successor):
  ``X2(S[x,x])`` on ``x_{t,x}`` and ``XY(2 S[x,x'])`` on ``(x_{t,x}, x_{t,x'})``, x < x', with
  ``S = A A^T``,
* ``LinearGaussian(C[x,x], observation_variance)`` on ``(x_{t,x}, y_{t,x})`` for t >= 1 wherever
  ``data[x,t] != 5000`` (the reference's missing-value mark); step 0 carries ``data[:,0]`` as evidence.
"""
from __future__ import annotations

import numpy as np

try:
    from .Graph import F, RV, Graph
    from .Potential import LinearGaussianPotential, X2Potential, XYPotential
except ImportError:                       # flat import (install_flat_aliases)
    from Graph import F, RV, Graph
    from Potential import LinearGaussianPotential, X2Potential, XYPotential

MISSING = 5000


class KalmanFilter:
    def __init__(self, domain, transition_coeff, transition_variance, observation_coeff, observation_variance):
        self.domain = domain
        self.transition_coeff = np.asarray(transition_coeff, dtype=float)
        self.transition_variance = transition_variance
        self.observation_coeff = np.asarray(observation_coeff, dtype=float)
        self.observation_variance = observation_variance

    # ---- the factor list, independent of how variables are represented ----------------------
    def _pieces(self, num_t_steps, data):
        """Yields ``(kind, coefficient, variance, [(t, x) or ('obs', t, x), ...])`` for every ground
        factor, in a fixed order."""
        A = self.transition_coeff
        n = A.shape[0]
        # S = A A^T accumulated column by column like the reference (KalmanFilter.py:52-55): equal
        # entries must round equally, the potentials -- and the colours -- are keyed by these floats
        S = np.zeros((n, n))
        for y in range(n):
            S += np.outer(A[:, y], A[:, y])
        T = int(num_t_steps)
        tv, ov = self.transition_variance, self.observation_variance
        for t in range(1, T):
            for x in range(n):
                if data[x, t] != MISSING:
                    yield "lg", float(self.observation_coeff[x, x]), ov, [(t, x), ("obs", t, x)]
        for t in range(T - 1):
            for x in range(n):
                if t > 0 and S[x, x] != 0:
                    yield "x2", float(S[x, x]), tv, [(t, x)]
                for y in range(n):
                    if A[x, y] != 0:
                        yield "xy", float(-2 * A[x, y]), tv, [(t, x), (t + 1, y)]
                    if t > 0 and x < y and S[x, y] != 0:
                        yield "xy", float(2 * S[x, y]), tv, [(t, x), (t, y)]
        for t in range(1, T):
            for x in range(n):
                yield "x2", 1.0, tv, [(t, x)]

    @staticmethod
    def _potential(cache, kind, coeff, var):
        key = (kind, coeff, var)
        if key not in cache:
            cls = {"lg": LinearGaussianPotential, "x2": X2Potential, "xy": XYPotential}[kind]
            cache[key] = cls(coeff, var)
        return cache[key]

    # ---- objects ---------------------------------------------------------------------------
    def grounded_graph(self, num_t_steps, data):
        """``(Graph, table)`` with ``table[t][x]`` the state variable of step t (``g.rvs`` is a
        list, as in the reference)."""
        data = np.asarray(data)
        n, T = self.transition_coeff.shape[0], int(num_t_steps)
        table = [[RV(self.domain, data[x, 0] if t == 0 else None) for x in range(n)] for t in range(T)]
        node = {(t, x): table[t][x] for t in range(T) for x in range(n)}
        rvs = [rv for row in table for rv in row]
        factors, pots = [], {}
        for kind, coeff, var, where in self._pieces(T, data):
            nb = []
            for w in where:
                if w[0] == "obs":
                    leaf = RV(self.domain, data[w[2], w[1]])
                    rvs.append(leaf)
                    nb.append(leaf)
                else:
                    nb.append(node[w])
            factors.append(F(self._potential(pots, kind, coeff, var), nb))
        g = Graph()
        g.rvs, g.factors = rvs, factors
        g.init_nb()
        return g, table

    # ---- arrays ----------------------------------------------------------------------------
    def grounded_arrays(self, num_t_steps, data):
        """``(lifting.GroundArrays, state_index [T, n])``: the same ground model as index arrays, one
        factor block per distinct potential."""
        from .lifting import FactorBlock, GroundArrays
        data = np.asarray(data)
        n, T = self.transition_coeff.shape[0], int(num_t_steps)
        state = np.arange(T * n, dtype=np.int64).reshape(T, n)
        values = [np.nan] * (T * n)
        for x in range(n):
            values[x] = float(data[x, 0])
        pots, rows = {}, {}
        for kind, coeff, var, where in self._pieces(T, data):
            idx = []
            for w in where:
                if w[0] == "obs":
                    idx.append(len(values))
                    values.append(float(data[w[2], w[1]]))
                else:
                    idx.append(int(state[w]))
            rows.setdefault(id(self._potential(pots, kind, coeff, var)), (self._potential(pots, kind, coeff, var), []))[1].append(idx)
        blocks = [FactorBlock(p, np.asarray(r, dtype=np.int64)) for p, r in rows.values()]
        ga = GroundArrays([self.domain], np.zeros(len(values), dtype=np.int32), np.asarray(values, dtype=float), blocks)
        return ga, state
