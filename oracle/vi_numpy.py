"""ORACLE (test infrastructure, not product code) -- vectorised numpy fp64 restatement.

Computes, on the lowered structure-of-arrays model (``lowering.LoweredModel``), exactly what
the CUDA kernels must produce: the fused single-pass form of the reference's
``gradient_w_tau`` / ``gradient_mu_var`` / ``gradient_category_tau`` / ``free_energy``
(``VarInference.py:57-195``; lifted weights ``LiftedVarInference.py:74,90,131-132,162``;
Gaussian evidence ``C2FVarInference.py:110-113,266-267``) followed by the Adam / SGD step
(``VarInference.py:249-329``).  SURVEY section 9 is the numerical spec.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline leg may import
this file.  It is pinned to the reference through ``oracle/vi_loops.py`` and the golden
vectors under ``tests/golden/`` (``tests/test_oracle_golden.py``).
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial.hermite import hermgauss

EPS = 1e-100
SQRT_2PI = 2.506628274631      # the reference's literal (VarInference.py:29)


def quadrature(T):
    x, w = hermgauss(T)
    return x, w / np.sqrt(np.pi)


def _pdf(x, mu, var):
    u = x - mu
    return np.exp(-u * u * 0.5 / var) / (SQRT_2PI * var)


norm_pdf_ref = _pdf


def _on_axis(arr, lead, axis, n_axes):
    """View ``arr`` (shape lead-dims + (n,)) with its last dim placed on grid axis ``axis``."""
    shape = list(arr.shape[:lead]) + [1] * n_axes
    shape[lead + axis] = arr.shape[-1]
    return arr.reshape(shape)


def grad_pass(model, eta, w, chunk_elems=1 << 24):
    """One pass over every record group.

    ``eta``: flat parameter vector (continuous slots hold (mu, var), discrete slots hold
    probabilities); ``w``: mixture weights.  Returns ``(grad, g_w, energy)`` where ``grad``
    has eta's layout and holds g_mu / g_var / raw G_c, ``g_w`` is the raw G_w (before the
    softmax Jacobian) and ``energy`` the Bethe free energy."""
    K, T = model.K, model.T
    qx, qw = quadrature(T)
    grad = np.zeros_like(eta, dtype=float)
    g_w = np.zeros(K)
    energy = 0.0
    karange = np.arange(K)

    for g in model.groups:
        n_ax = g.nd + g.nc + g.ng
        sizes = list(g.dims) + [T] * (g.nc + g.ng)
        G = int(np.prod(sizes)) if sizes else 1
        step = max(1, chunk_elems // max(1, K * K * G))
        ncoef = 1 if g.nct == 0 else (g.nct + 1) * (g.nct + 2) // 2
        for lo in range(0, g.n, step):
            sl = slice(lo, min(g.n, lo + step))
            m = sl.stop - sl.start
            X, Wt, Q = [], [], []          # per axis
            mus, vars_ = [], []
            for a in range(g.nd):
                D = g.dims[a]
                idx = g.poff[a, sl][:, None, None] + karange[None, :, None] * D + np.arange(D)[None, None, :]
                p = eta[idx]                                   # [m, K, D]
                X.append(None)
                Wt.append(p)
                Q.append(np.broadcast_to(p[:, None, :, :], (m, K, K, D)))
            for c in range(g.nc):
                base = g.poff[g.nd + c, sl][:, None] + 2 * karange[None, :]
                mu, var = eta[base], eta[base + 1]             # [m, K]
                x = np.sqrt(2 * var)[:, :, None] * qx[None, None, :] + mu[:, :, None]   # [m,K,T]
                X.append(x)
                Wt.append(np.broadcast_to(qw[None, None, :], (m, K, T)))
                Q.append(_pdf(x[:, :, None, :], mu[:, None, :, None], var[:, None, :, None]))
                mus.append(mu)
                vars_.append(var)
            for j in range(g.ng):
                val, var = g.egval[j, sl], g.egvar[j, sl]
                x = np.sqrt(2 * var)[:, None] * qx[None, :] + val[:, None]              # [m,T]
                x = np.broadcast_to(x[:, None, :], (m, K, T))
                X.append(x)
                Wt.append(np.broadcast_to(qw[None, None, :], (m, K, T)))
                q = _pdf(x, val[:, None, None], var[:, None, None])
                Q.append(np.broadcast_to(q[:, :, None, :], (m, K, K, T)))

            # belief on the grid: [m, K, grid...]
            if g.pure:
                lb = 0.0         # the -log b part of unary factors lives in the node records
            else:
                prod = np.broadcast_to(w.reshape([1, 1, K] + [1] * n_ax), [m, K, K] + sizes).copy()
                for a in range(n_ax):
                    prod *= _on_axis(Q[a], 3, a, n_ax)
                lb = np.log(prod.sum(axis=2) + EPS)

            gscale = 1.0
            if g.node:
                F = lb            # energy / G_w weighted by wf, gradients by nscale
                gscale = g.nscale[sl]
            else:
                cfg = np.zeros([m] + [1] * n_ax, dtype=np.int64)
                stride = 1
                for a in reversed(range(g.nd)):
                    cfg = cfg + _on_axis(np.arange(g.dims[a])[None, :] * stride, 1, a, n_ax)
                    stride *= g.dims[a]
                base = g.pot[sl].astype(np.int64).reshape([m] + [1] * n_ax) + cfg * ncoef
                base = base[:, None]                           # add the k axis
                if g.nct == 0:
                    lpsi = model.ptab[base]
                else:
                    xs = [_on_axis(X[g.nd + c], 2, g.nd + c, n_ax) for c in range(g.nc + g.ng)]
                    xs += [g.ecval[j, sl].reshape([m, 1] + [1] * n_ax) for j in range(g.ne)]
                    kind = getattr(g, "kind", 0)
                    if kind == 2:
                        # ImageEdgePotential (Potential.py:419-424): block = [distant, scaling, threshold, v]
                        d = np.abs(xs[0] - xs[1])
                        tail = np.where(d > model.ptab[base + 2], model.ptab[base + 3], np.exp(-d / model.ptab[base + 1]))
                        lpsi = np.log(d * model.ptab[base] + tail + EPS)
                    else:
                        q = model.ptab[base]
                        for i in range(g.nct):
                            q = q + model.ptab[base + 1 + i] * xs[i]
                        p = 1 + g.nct
                        for i in range(g.nct):
                            for j in range(i, g.nct):
                                q = q + model.ptab[base + p] * xs[i] * xs[j]
                                p += 1
                        if kind == 1:
                            # MLNHardPotential (MLNPotential.py:48-49): psi = 1 where the formula is > 0, else 0
                            lpsi = np.log(np.where(q > 0, 1.0, 0.0) + EPS)
                        else:
                            with np.errstate(over="ignore"):
                                lpsi = np.log(np.exp(q) + EPS)
                F = lpsi - lb

            Wgrid = np.ones([m, K] + [1] * n_ax)
            for a in range(n_ax):
                Wgrid = Wgrid * _on_axis(Wt[a], 2, a, n_ax)
            S = Wgrid * F
            grid_axes = tuple(range(2, 2 + n_ax))
            Ek = S.sum(axis=grid_axes) if n_ax else S.reshape(m, K)
            wf = g.wf[sl]
            g_w -= (wf[:, None] * Ek).sum(axis=0)
            energy -= float((wf[:, None] * Ek * w[None, :]).sum())

            for c in range(g.nc):
                a = g.nd + c
                dx = _on_axis(X[a] - mus[c][:, :, None], 2, a, n_ax)
                var = vars_[c]
                gam = (g.gam[a, sl] * gscale)[:, None]
                gmu = -gam * (S * dx).sum(axis=grid_axes) / var
                gvar = -gam * (S * (dx * dx - var.reshape([m, K] + [1] * n_ax))).sum(axis=grid_axes) \
                    / (2 * var * var)
                base = g.poff[a, sl][:, None] + 2 * karange[None, :]
                np.add.at(grad, base, gmu)
                np.add.at(grad, base + 1, gvar)
            for a in range(g.nd):
                D = g.dims[a]
                Wo = np.ones([m, K] + [1] * n_ax)
                for a2 in range(n_ax):
                    if a2 != a:
                        Wo = Wo * _on_axis(Wt[a2], 2, a2, n_ax)
                other = tuple(ax for ax in grid_axes if ax != 2 + a)
                gc = -((g.gam[a, sl] * gscale)[:, None, None]) * (Wo * F).sum(axis=other)     # [m, K, D]
                idx = g.poff[a, sl][:, None, None] + karange[None, :, None] * D + np.arange(D)[None, None, :]
                np.add.at(grad, idx, gc)
    return grad, g_w, energy


def free_energy(model, eta, w):
    return grad_pass(model, eta, w)[2]


# ---- parameter-space helpers (flat layout) -------------------------------------------------

def slot_masks(model):
    """Index arrays over the flat vector: continuous mu / var elements and, per discrete
    variable, its [K, D] block."""
    K = model.K
    cont = model.var_off[model.var_kind == 0]
    mu_idx = (cont[:, None] + 2 * np.arange(K)[None, :]).reshape(-1)
    disc = [(int(o), int(d)) for o, d, k in zip(model.var_off, model.var_dim, model.var_kind) if k == 1]
    return mu_idx, mu_idx + 1, disc


def softmax_unshifted(x, axis=-1):
    r = np.e ** x
    return r / r.sum(axis=axis, keepdims=True)


def tau_gradients(model, grad, g_w, eta, w):
    """Chain rule through the two softmaxes (``VarInference.py:90,160``): returns the
    gradient w.r.t. the optimised quantities (mu, var, eta_tau) in flat layout and w_tau."""
    out = grad.copy()
    K = model.K
    _, _, disc = slot_masks(model)
    for off, D in disc:
        blk = slice(off, off + K * D)
        p = eta[blk].reshape(K, D)
        gc = grad[blk].reshape(K, D)
        out[blk] = (p * (gc - (gc * p).sum(axis=1, keepdims=True))).reshape(-1)
    return out, w * (g_w - np.sum(g_w * w))


class NumpyVI:
    """Flat-vector optimiser loop mirroring ``ADAM_update`` / ``GD_update``."""

    def __init__(self, model, var_threshold=0.1):
        self.model = model
        self.var_threshold = var_threshold
        self.K = model.K
        n = model.n_param
        self.eta = np.zeros(n)
        self.tau = np.zeros(n)          # logits for discrete slots
        self.w_tau = np.zeros(self.K)
        self.w = softmax_unshifted(self.w_tau)
        self.mu_idx, self.var_idx, self.disc = slot_masks(model)
        self.eta[self.var_idx] = 1.0
        self.reset_moments()

    def reset_moments(self):
        n = self.model.n_param
        self.t = 0
        self.m = np.zeros(n)
        self.u = np.zeros(n)
        self.m_w = np.zeros(self.K)
        self.u_w = np.zeros(self.K)

    def refresh(self):
        """w = softmax(w_tau); eta = softmax(tau) on discrete slots."""
        self.w = softmax_unshifted(self.w_tau)
        K = self.K
        for off, D in self.disc:
            blk = slice(off, off + K * D)
            self.eta[blk] = softmax_unshifted(self.tau[blk].reshape(K, D)).reshape(-1)

    def gradients(self):
        grad, g_w, energy = grad_pass(self.model, self.eta, self.w)
        g_flat, g_wtau = tau_gradients(self.model, grad, g_w, self.eta, self.w)
        return g_flat, g_wtau, energy

    def _apply(self, step_flat, step_w):
        self.w_tau = self.w_tau - step_w
        cont = np.concatenate([self.mu_idx, self.var_idx])
        self.eta[cont] -= step_flat[cont]
        self.eta[self.var_idx] = np.clip(self.eta[self.var_idx], self.var_threshold, np.inf)
        K = self.K
        for off, D in self.disc:
            blk = slice(off, off + K * D)
            self.tau[blk] -= step_flat[blk]
        self.refresh()

    def adam_step(self, lr, b1=0.9, b2=0.999, eps=1e-8):
        g, gw, energy = self.gradients()
        self.t += 1
        self.m = b1 * self.m + (1 - b1) * g
        self.u = b2 * self.u + (1 - b2) * g * g
        self.m_w = b1 * self.m_w + (1 - b1) * gw
        self.u_w = b2 * self.u_w + (1 - b2) * gw * gw
        c1, c2 = 1 - b1 ** self.t, 1 - b2 ** self.t
        self._apply((lr * (self.m / c1)) / (np.sqrt(self.u / c2) + eps),
                    (lr * (self.m_w / c1)) / (np.sqrt(self.u_w / c2) + eps))
        return energy          # free energy at the parameters *before* the step

    def sgd_step(self, lr):
        g, gw, energy = self.gradients()
        self._apply(lr * g, lr * gw)
        return energy
