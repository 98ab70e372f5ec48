"""Host-side logic shared by the three drop-in engines.

The public surface follows the reference classes (SURVEY section 8 b): constructor
``(g, num_mixtures=5, num_quadrature_points=3)``, ``run(iteration, lr, is_log, log_fe)``,
``ADAM_update``, ``GD_update``, ``free_energy``, ``gradient_w_tau``, ``gradient_mu_var``,
``gradient_category_tau``, ``belief``, ``rvs_belief``, ``map``, and the state attributes
``K, T, quad_x, quad_w, w, w_tau, eta, eta_tau, w_tau_g, eta_g, eta_tau_g, t, time_log,
total_time``.  ``eta[rv]`` / ``eta_tau[rv]`` are host ``numpy`` views of the device state,
refreshed after every update; assigning to them before a call is honoured (they are pushed
to the device first), which is how callers inject parameters.

All numerical work (expected log-potentials, gradients, free energy, optimiser step, belief
queries) runs in the CUDA kernels behind ``engine.DeviceEngine``.  Extra keyword-only
arguments (``dtype``, ``device``) do not exist in the reference.
"""
from __future__ import annotations

import time

import numpy as np

from . import lowering
from .utils import log_likelihood

SQRT_2PI = 2.506628274631   # the reference's literal (VarInference.py:29)


def norm_pdf(x, mu, var):
    u = x - mu
    return np.exp(-u * u * 0.5 / var) / (SQRT_2PI * var)


def softmax(x, axis=0):
    """Un-shifted softmax with the reference's axis convention (VarInference.py:32-38)."""
    r = np.e ** np.asarray(x, dtype=float)
    return r / np.sum(r, axis=axis, keepdims=True)


def mixture_mode(w, mu, var, x0, iters=60):
    """Arg-max of sum_k w_k q_k(x) near ``x0`` by safeguarded Newton ascent (replaces the
    per-variable scipy BFGS of VarInference.py:363-367).  Vectorised: ``mu, var`` are
    [n, K], ``x0`` is [n]."""
    x = np.array(x0, dtype=float, copy=True)
    w = np.asarray(w, dtype=float)[None, :]

    def f(z):
        return (w * norm_pdf(z[:, None], mu, var)).sum(axis=1)

    fx = f(x)
    for _ in range(iters):
        p = w * norm_pdf(x[:, None], mu, var)
        d = x[:, None] - mu
        g1 = (-p * d / var).sum(axis=1)
        g2 = (p * (d * d / (var * var) - 1.0 / var)).sum(axis=1)
        step = np.where(g2 < 0, -g1 / np.where(g2 < 0, g2, -1.0), np.sign(g1) * 0.1)
        # backtrack so that the density never decreases
        for _ in range(30):
            cand = x + step
            fc = f(cand)
            worse = fc < fx
            if not worse.any():
                break
            step = np.where(worse, step * 0.5, step)
        x, fx = x + step, f(x + step)
        if np.max(np.abs(step)) < 1e-13:
            break
    return x


class VIBase:
    var_threshold = 0.1

    # ---- construction ---------------------------------------------------------------------
    def _init_common(self, num_mixtures, num_quadrature_points, dtype, device, compat=None):
        self.K = int(num_mixtures)
        self.T = int(num_quadrature_points)
        self.quad_x, self.quad_w = np.polynomial.hermite.hermgauss(self.T)
        self.quad_w = self.quad_w / np.sqrt(np.pi)
        self.w_tau = np.zeros(self.K)
        self.w = np.zeros(self.K)
        self.eta_tau = dict()
        self.eta = dict()
        self.dtype = dtype
        self.device = device
        # compat="reference": gradient_category_tau as the unmodified reference computes it (its axes of
        # the other arguments come from the wrong domain, VarInference.py:147-150; SURVEY hazard H2);
        # None: the intended mathematics
        if compat not in (None, "reference"):
            raise ValueError(f"compat must be None or 'reference', got {compat!r}")
        self.compat = compat
        self.t = 0
        self.alpha = 0.1
        self.b1, self.b2, self.eps = 0.9, 0.999, 1e-8
        self.w_tau_g = [np.zeros(self.K), np.zeros(self.K)]
        self.eta_g = [dict(), dict()]
        self.eta_tau_g = [dict(), dict()]
        self.time_log = list()
        self.total_time = 0
        self.is_log, self.log_fe = True, True
        self._engine = None
        self._model = None
        self._pushed = None
        self._grad_cache = None
        self._map_cache = None

    # hooks the three engines specialise
    def _handles(self):
        raise NotImplementedError

    def _lower(self):
        raise NotImplementedError

    def _ground_graph(self):
        raise NotImplementedError

    def _handle_of(self, rv):
        return rv

    def _make_engine(self, model):
        from .engine import DeviceEngine
        return DeviceEngine(model, dtype=self.dtype, device=self.device,
                            var_threshold=self.var_threshold, compat=self.compat)

    # ---- parameters -----------------------------------------------------------------------
    def init_param(self):
        """Same distributions as the reference (VarInference.py:197-213) but drawn in a
        deterministic order (sorted ids) instead of set-iteration order."""
        self.w_tau = np.zeros(self.K)
        self.eta, self.eta_tau = dict(), dict()
        for rv in self._handles():
            if rv.value is not None:
                continue
            if rv.domain.continuous:
                table = np.ones((self.K, 2))
                table[:, 0] = np.random.rand(self.K) * 3 - 1.5
                self.eta[rv] = table
            else:
                self.eta_tau[rv] = np.random.rand(self.K, len(rv.domain.values)) * 10
        self.w = softmax(self.w_tau)
        for rv, table in self.eta_tau.items():
            self.eta[rv] = softmax(table, 1)

    def _zero_moments(self):
        self.w_tau_g = [np.zeros(self.K), np.zeros(self.K)]
        self.eta_g = [dict(), dict()]
        self.eta_tau_g = [dict(), dict()]
        for rv in self._handles():
            if rv.value is not None:
                continue
            if rv.domain.continuous:
                self.eta_g[0][rv] = np.zeros((self.K, 2))
                self.eta_g[1][rv] = np.zeros((self.K, 2))
            else:
                shape = (self.K, len(rv.domain.values))
                self.eta_tau_g[0][rv] = np.zeros(shape)
                self.eta_tau_g[1][rv] = np.zeros(shape)
        self.t = 0

    # ---- host <-> device ------------------------------------------------------------------
    def _rebuild(self):
        """(Re)lower the current graph and create the device engine."""
        self._model = self._lower()
        self._drop_engine()
        self._engine = self._make_engine(self._model)
        self._pushed = None
        self._grad_cache = None

    def _drop_engine(self):
        """Release the current engine's graphs / peer buffers before it is replaced."""
        eng, self._engine = self._engine, None
        if eng is not None and hasattr(eng, "close"):
            eng.close()

    def _flat(self, cont_dict, disc_dict, fill=0.0):
        m, K = self._model, self.K
        flat = np.full(m.n_param, fill, dtype=float)
        for h, i in m.index.items():
            off, d = int(m.var_off[i]), int(m.var_dim[i])
            src = cont_dict.get(h) if m.var_kind[i] == 0 else disc_dict.get(h)
            if src is not None:
                flat[off:off + K * d] = np.asarray(src, dtype=float).reshape(-1)
        return flat

    def _unflat(self, flat, cont_dict, disc_dict):
        m, K = self._model, self.K
        for h, i in m.index.items():
            off, d = int(m.var_off[i]), int(m.var_dim[i])
            block = flat[off:off + K * d].reshape(K, d)        # a view of the pulled vector: no copy per variable
            if m.var_kind[i] == 0:
                cont_dict[h] = block
            elif disc_dict is not None:
                disc_dict[h] = block

    def _push(self, moments=False):
        if self._engine is None:
            self._rebuild()
        eta = self._flat(self.eta, self.eta)
        tau = self._flat({}, self.eta_tau)
        key = (eta.tobytes(), tau.tobytes(), np.asarray(self.w_tau, dtype=float).tobytes())
        if key != self._pushed:
            self._engine.set_state(eta, tau, self.w_tau)
            self._pushed = key
            self._grad_cache = None
        if moments:
            self._engine.set_moments(self._flat(self.eta_g[0], self.eta_tau_g[0]),
                                     self._flat(self.eta_g[1], self.eta_tau_g[1]),
                                     self.w_tau_g[0], self.w_tau_g[1], self.t)

    def _pull(self, moments=False):
        eta, tau, w_tau, w = self._engine.get_state()
        self.w_tau, self.w = w_tau, w
        self._unflat(eta, self.eta, self.eta)
        self._unflat(tau, {}, self.eta_tau)
        self._pushed = (self._flat(self.eta, self.eta).tobytes(), self._flat({}, self.eta_tau).tobytes(),
                        np.asarray(self.w_tau, dtype=float).tobytes())
        self._grad_cache = None
        if moments:
            m1, m2, mw, uw, t = self._engine.get_moments()
            self._unflat(m1, self.eta_g[0], self.eta_tau_g[0])
            self._unflat(m2, self.eta_g[1], self.eta_tau_g[1])
            self.w_tau_g = [mw, uw]
            self.t = int(round(t))

    # ---- objective and gradients (one fused device pass serves all of them) -----------------
    def _grads(self):
        self._push()
        if self._grad_cache is None:
            self._grad_cache = self._engine.gradients()
        return self._grad_cache

    def free_energy(self):
        return self._grads()[2]

    def gradient_w_tau(self):
        g_w = self._grads()[1]
        w = softmax(self.w_tau)
        return w * (g_w - np.sum(g_w * w))

    def _slot(self, rv):
        m = self._model
        i = m.index[rv]
        return int(m.var_off[i]), int(m.var_dim[i])

    def gradient_mu_var(self, rv):
        grad = self._grads()[0]
        off, _ = self._slot(rv)
        return grad[off:off + 2 * self.K].reshape(self.K, 2).copy()

    def gradient_category_tau(self, rv):
        grad = self._grads()[0]
        off, D = self._slot(rv)
        g_c = grad[off:off + self.K * D].reshape(self.K, D)
        eta = self.eta[rv]
        return eta * (g_c - np.sum(g_c * eta, 1)[:, np.newaxis])

    # ---- optimisation ---------------------------------------------------------------------
    def _start_run(self, lr, is_log, log_fe):
        self.is_log, self.log_fe = is_log, log_fe
        self.alpha = lr
        self.b1, self.b2, self.eps = 0.9, 0.999, 1e-8

    def run(self, iteration=100, lr=0.1, is_log=True, log_fe=True):
        self._start_run(lr, is_log, log_fe)
        self._rebuild()
        self.init_param()
        self._zero_moments()
        if self.is_log:
            self.time_log = list()
            self.total_time = 0
        self.ADAM_update(iteration)

    def _objective_for_log(self):
        if self.log_fe:
            return self.free_energy()
        g = self._ground_graph()
        assignment = {rv: self.map(rv) for rv in g.rvs}
        return log_likelihood(g, assignment)

    def _sync(self):
        import torch
        torch.cuda.synchronize()

    def _update_loop(self, iteration, lr, sgd):
        if self._engine is None:
            self._rebuild()
        eng = self._engine
        # before the moments go up: set_moments seeds the bias corrections 1 - b^t from these
        eng.b1, eng.b2, eng.eps = self.b1, self.b2, self.eps
        eng.var_threshold = float(self.var_threshold)
        self._push(moments=not sgd)
        if not (self.is_log or sgd):
            eng.iterate(iteration, lr, sgd=False)
            self._pull(moments=True)
            return
        # Logging costs one scalar read per iteration.  The reference prints the objective AFTER each
        # update (VarInference.py:290-300); the fused pass of an iteration evaluates the free energy
        # at the parameters BEFORE its step, i.e. after the previous update -- so the entry of
        # iteration i is completed by the pass of iteration i + 1 (printed one iteration late, same
        # values and order), and only the last one needs a pass of its own.  The host copies of the
        # parameters are refreshed once, at the end.
        def emit(t, fe):
            if sgd:
                print(fe)
            else:
                print(fe, t)
                self.time_log.append([t, fe])
        by_pass = sgd or self.log_fe                     # the logged objective is the free energy
        stamp = None
        for i in range(int(iteration)):
            self._sync()
            start = time.perf_counter()
            eng.iterate(1, lr, sgd=sgd)
            self._sync()
            self.total_time += time.perf_counter() - start
            self._grad_cache = None
            self._map_cache = None
            if by_pass:
                if i > 0:
                    emit(stamp, eng.last_free_energy())
                stamp = self.total_time
            else:
                emit(self.total_time, self._objective_for_log())      # MAP assignment from the device state
        if by_pass and int(iteration) > 0:
            emit(stamp, eng.free_energy())
        self._pull(moments=not sgd)

    def ADAM_update(self, iteration):
        """``iteration`` Jacobi Adam steps (VarInference.py:249-300); ``self.t`` persists."""
        self._update_loop(iteration, self.alpha, sgd=False)

    def GD_update(self, iteration, lr):
        """Plain gradient descent (VarInference.py:302-331); prints the free energy each step."""
        self._update_loop(iteration, lr, sgd=True)

    # ---- checkpoint / resume (not in the reference, whose run() re-randomises every call) ----
    def state_dict(self):
        """Everything an interrupted optimisation needs to go on: mixture logits, per-variable
        parameters (continuous: (mu, var); discrete: logits) and the Adam moments and step counter,
        keyed by the position of the variable in this engine's sorted handle list (ground: ``rv.id``
        order; lifted: class order), as plain numpy arrays."""
        order = {h: i for i, h in enumerate(self._handles())}

        def by_index(d):
            return {order[h]: np.array(v, dtype=float, copy=True) for h, v in d.items() if h in order}
        cont = {h: v for h, v in self.eta.items() if h.domain.continuous}
        return {"K": self.K, "T": self.T, "t": int(self.t), "w_tau": np.array(self.w_tau, dtype=float),
                "eta": by_index(cont), "eta_tau": by_index(self.eta_tau),
                "w_tau_g": [np.array(v, dtype=float) for v in self.w_tau_g],
                "eta_g": [by_index(d) for d in self.eta_g], "eta_tau_g": [by_index(d) for d in self.eta_tau_g]}

    def load_state_dict(self, state):
        """Inverse of ``state_dict`` on an engine over the same graph; continue with ``ADAM_update``."""
        if int(state["K"]) != self.K or int(state["T"]) != self.T:
            raise ValueError("state_dict belongs to an engine with other K / T")
        handles = self._handles()

        def by_handle(d):
            return {handles[int(i)]: np.array(v, dtype=float, copy=True) for i, v in d.items()}
        self.w_tau = np.array(state["w_tau"], dtype=float)
        self.w = softmax(self.w_tau)
        self.eta = by_handle(state["eta"])
        self.eta_tau = by_handle(state["eta_tau"])
        for h, table in self.eta_tau.items():
            self.eta[h] = softmax(table, 1)
        self.w_tau_g = [np.array(v, dtype=float) for v in state["w_tau_g"]]
        self.eta_g = [by_handle(d) for d in state["eta_g"]]
        self.eta_tau_g = [by_handle(d) for d in state["eta_tau_g"]]
        self.t = int(state["t"])
        self._pushed = None
        self._grad_cache = None
        self._map_cache = None

    # ---- queries --------------------------------------------------------------------------
    def _gauss_evidence(self, h):
        return None

    def rvs_belief(self, x, rvs):
        """b(x_S) = sum_k w_k prod_i q_ik(x_i) for handles of *this* engine's graph
        (VarInference.py:336-353)."""
        b = np.array(self.w, dtype=float, copy=True)
        for xi, rv in zip(x, rvs):
            if rv.value is not None:
                ge = self._gauss_evidence(rv)
                if ge is not None:
                    b = b * norm_pdf(xi, ge[0], ge[1])
                elif xi != rv.value:
                    return 0
            elif rv.domain.continuous:
                eta = self.eta[rv]
                b = b * norm_pdf(xi, eta[:, 0], eta[:, 1])
            else:
                b = b * self.eta[rv][:, rv.domain.values.index(xi)]
        return np.sum(b)

    def beliefs(self, xs, rvs):
        """Batched ``belief``: one device launch for many (x, rv) queries on hidden variables."""
        self._push()
        m = self._model
        off, dim, kind, xv = [], [], [], []
        for x, rv in zip(xs, rvs):
            h = self._handle_of(rv)
            i = m.index[h]
            off.append(int(m.var_off[i]))
            dim.append(int(m.var_dim[i]))
            kind.append(int(m.var_kind[i]))
            xv.append(float(x) if m.var_kind[i] == 0 else float(h.domain.values.index(x)))
        return self._engine.mixture_belief(off, dim, kind, xv).double().cpu().numpy()

    def belief(self, x, rv):
        h = self._handle_of(rv)
        if h.value is not None:
            return self.rvs_belief((x,), (h,))
        if not h.domain.continuous and x not in h.domain.values:
            raise ValueError(f"{x!r} is not in the variable's domain")
        return float(self.beliefs([x], [rv])[0])

    def map_all(self):
        """MAP of every hidden variable in one device launch: ``{handle: value}`` (discrete
        values are domain values).  ``map(rv)`` serves itself from this table, so the usual
        ``for rv in g.rvs: infer.map(rv)`` loop of the reference's demos costs one launch."""
        self._push()
        if self._map_cache is not None and self._map_cache[0] == self._pushed:
            return self._map_cache[1]
        m = self._model
        res = self._engine.mixture_map().double().cpu().numpy()
        table = {}
        for h, i in m.index.items():
            table[h] = float(res[i]) if m.var_kind[i] == 0 else h.domain.values[int(res[i])]
        self._map_cache = (self._pushed, table)
        return table

    def map(self, rv):
        """MAP value of one variable's marginal belief (VarInference.py:355-376)."""
        h = self._handle_of(rv)
        if h.value is not None:
            return rv.value
        return self.map_all()[h]
