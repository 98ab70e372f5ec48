"""Test double with ``engine.DeviceEngine``'s interface, backed by the numpy oracle.

Lets the CPU suite exercise the host logic of the drop-in classes (parameter push/pull,
C2F refinement with inheritance, logging, queries) without a GPU.  Product code never
imports this; ``use_oracle_engine(vi)`` patches one instance inside a test.
"""
from __future__ import annotations

import numpy as np
import torch

from oracle.vi_numpy import NumpyVI, norm_pdf_ref


class OracleEngine:
    def __init__(self, model, var_threshold=0.1):
        self.model = model
        self.K = model.K
        self.n_param = model.n_param
        self.vi = NumpyVI(model, var_threshold=var_threshold)
        self.b1, self.b2, self.eps = 0.9, 0.999, 1e-8
        self.var_threshold = var_threshold

    def set_state(self, eta, tau, w_tau):
        self.vi.eta[:] = eta
        self.vi.tau[:] = tau
        self.vi.w_tau = np.array(w_tau, dtype=float)
        e = np.e ** self.vi.w_tau
        self.vi.w = e / e.sum()

    def get_state(self):
        return self.vi.eta.copy(), self.vi.tau.copy(), self.vi.w_tau.copy(), self.vi.w.copy()

    def get_moments(self):
        return self.vi.m.copy(), self.vi.u.copy(), self.vi.m_w.copy(), self.vi.u_w.copy(), float(self.vi.t)

    def set_moments(self, m1, m2, m_w, u_w, t):
        self.vi.m[:] = m1
        self.vi.u[:] = m2
        self.vi.m_w = np.array(m_w, dtype=float)
        self.vi.u_w = np.array(u_w, dtype=float)
        self.vi.t = int(round(t))

    def reset_moments(self):
        self.vi.m[:] = 0
        self.vi.u[:] = 0
        self.vi.m_w = np.zeros(self.K)
        self.vi.u_w = np.zeros(self.K)
        self.vi.t = 0

    def iterate(self, n, lr, sgd=False):
        self.vi.var_threshold = self.var_threshold
        for _ in range(int(n)):
            if sgd:
                self._last_fe = self.vi.sgd_step(lr)
            else:
                self._last_fe = self.vi.adam_step(lr, self.b1, self.b2, self.eps)

    def last_free_energy(self):
        """Free energy evaluated by the most recent pass (at the parameters before its step)."""
        return float(self._last_fe)

    def gradients(self):
        from oracle.vi_numpy import grad_pass
        return grad_pass(self.model, self.vi.eta, self.vi.w)

    def free_energy(self):
        return self.gradients()[2]

    def mixture_belief(self, q_off, q_dim, q_kind, x):
        K = self.K
        out = np.zeros(len(q_off))
        for i, (o, d, kd, xi) in enumerate(zip(q_off, q_dim, q_kind, x)):
            if kd == 0:
                p = self.vi.eta[o:o + 2 * K].reshape(K, 2)
                out[i] = (self.vi.w * norm_pdf_ref(xi, p[:, 0], p[:, 1])).sum()
            else:
                p = self.vi.eta[o:o + K * d].reshape(K, d)
                out[i] = (self.vi.w * p[:, int(xi)]).sum()
        return torch.as_tensor(out)


    def mixture_map(self, q_off=None, q_dim=None, q_kind=None):
        from lhvi_b200._vi_base import mixture_mode
        m, K = self.model, self.K
        if q_off is None:
            q_off, q_dim, q_kind = m.var_off, m.var_dim, m.var_kind
        out = np.zeros(len(q_off))
        for i, (o, d, kd) in enumerate(zip(q_off, q_dim, q_kind)):
            if kd == 0:
                p = self.vi.eta[o:o + 2 * K].reshape(K, 2)
                dens = [(self.vi.w * norm_pdf_ref(x, p[:, 0], p[:, 1])).sum() for x in p[:, 0]]
                x0 = p[int(np.argmax(dens)), 0]
                out[i] = mixture_mode(self.vi.w, p[None, :, 0], p[None, :, 1], np.array([x0]))[0]
            else:
                p = self.vi.eta[o:o + K * d].reshape(K, d)
                out[i] = int(np.argmax((self.vi.w[:, None] * p).sum(axis=0)))
        return torch.as_tensor(out)


def use_oracle_engine(vi):
    vi._make_engine = lambda model: OracleEngine(model, var_threshold=vi.var_threshold)
    vi._sync = lambda: None
    return vi
