"""Potential plugin types understood by the B200 VI engines.

Same class names, constructor arguments and ``get(x)`` values as the reference's
``Potential.py`` (Table ``:15-24``, Gaussian ``:35-60``, Quadratic ``:69-110``,
HybridQuadratic ``:273-308``, LinearGaussian ``:311-338``, X2 ``:341-368``,
XY ``:371-397``, ImageNode/ImageEdge ``:400-424``).  Every exp-quadratic class derives
from one base that stores ``log psi(x) = x'Ax + b'x + c``; that triple is exactly what
the lowering layer ships to the GPU (``lowering.py``), so ``get`` here and the kernels
evaluate the same polynomial.
"""
from __future__ import annotations

import math

import numpy as np

try:  # package import
    from .Graph import Potential
except ImportError:  # flat import, reference style (`from Potential import ...`)
    from Graph import Potential


def mu_prec_to_quad_params(mu, prec):
    """(A, b, c) with x'Ax + b'x + c == -0.5 (x-mu)' prec (x-mu)."""
    mu = np.asarray(mu, dtype=float)
    prec = np.asarray(prec, dtype=float)
    lin = prec @ mu
    return -0.5 * prec, lin, -0.5 * float(mu @ lin)


class TablePotential(Potential):
    """psi(x) = table[x]; ``table`` is a dict keyed by value tuples or an ndarray."""

    def __init__(self, table, symmetric=False):
        super().__init__(symmetric=symmetric)
        self.table = table

    def get(self, parameters):
        return self.table[tuple(parameters)]


class _ExpQuadratic(Potential):
    """psi(x) = exp(x'Ax + b'x + c) over continuous arguments only."""

    def __init__(self, A, b, c, symmetric=False):
        super().__init__(symmetric=symmetric)
        self._A = np.atleast_2d(np.asarray(A, dtype=float))
        self._b = np.atleast_1d(np.asarray(b, dtype=float))
        self._c = float(c)

    def log_value(self, x):
        x = np.asarray(x, dtype=float).reshape(-1)
        return float(x @ self._A @ x + self._b @ x + self._c)

    def get(self, parameters):
        return math.exp(self.log_value(parameters))

    def get_quadratic_params(self):
        return self._A, self._b, self._c

    def dim(self):
        return self._b.size


class GaussianPotential(_ExpQuadratic):
    """exp{-0.5 (x-mu)' sig^{-1} (x-mu)} (the normaliser is *not* part of ``get``)."""

    def __init__(self, mu, sig, w=1):
        self.mu = np.asarray(mu, dtype=float)
        self.sig = np.asarray(sig, dtype=float)
        det = np.linalg.det(self.sig)
        if det == 0:
            raise ValueError("covariance matrix is singular")
        self.prec = np.linalg.inv(self.sig)
        self.coefficient = w / ((2 * math.pi) ** (0.5 * len(self.mu)) * math.sqrt(det))
        super().__init__(*mu_prec_to_quad_params(self.mu, self.prec))

    def get(self, parameters, use_coef=False):
        val = math.exp(self.log_value(parameters))
        return self.coefficient * val if use_coef else val


class QuadraticPotential(_ExpQuadratic):
    """exp(x'Ax + b'x + c) with user-supplied coefficients."""

    def __init__(self, A, b, c):
        super().__init__(A, b, c)
        self.A, self.b, self.c = self._A, self._b, self._c

    def get(self, args, ignore_const=False):
        lv = self.log_value(args)
        return math.exp(lv - self._c if ignore_const else lv)

    __call__ = get


class _CoeffSigPotential(_ExpQuadratic):
    """Potentials parameterised by (coeff, sig) that hash/compare by value, so that
    equal-parameter instances share a colour in colour passing
    (reference ``Potential.py:320-328,350-358,380-388``)."""

    def __init__(self, coeff, sig, prec, symmetric=False):
        self.coeff = coeff
        self.sig = sig
        super().__init__(*mu_prec_to_quad_params(np.zeros(len(prec)), prec), symmetric=symmetric)

    def __hash__(self):
        return hash((self.coeff, self.sig))

    def __eq__(self, other):
        return (self.__class__ == other.__class__
                and self.coeff == other.coeff and self.sig == other.sig)


class LinearGaussianPotential(_CoeffSigPotential):
    """exp(-(x1 - coeff*x0)^2 / (2 sig))."""

    def __init__(self, coeff, sig):
        a = coeff
        super().__init__(coeff, sig, np.array([[a * a, -a], [-a, 1.0]]) / sig)


class X2Potential(_CoeffSigPotential):
    """exp(-coeff * x0^2 / (2 sig))."""

    def __init__(self, coeff, sig):
        super().__init__(coeff, sig, np.array([[coeff / sig]]))


class XYPotential(_CoeffSigPotential):
    """exp(-coeff * x0 * x1 / (2 sig)); symmetric in its arguments."""

    def __init__(self, coeff, sig):
        super().__init__(coeff, sig, np.array([[0.0, 0.5], [0.5, 0.0]]) * coeff / sig,
                         symmetric=True)


class HybridQuadraticPotential(Potential):
    """exp(x_c' A[x_d] x_c + b[x_d]' x_c + c[x_d]); arguments ordered ``[x_d..., x_c...]``
    and discrete states are the integers 0..v-1 (reference ``Potential.py:273-305``)."""

    def __init__(self, A, b, c):
        super().__init__(symmetric=False)
        self.A = np.asarray(A, dtype=float)
        self.b = np.asarray(b, dtype=float)
        self.c = np.asarray(c, dtype=float)
        self.Nd = self.c.ndim
        self.Nc = int(self.b.shape[-1])

    def get_quadratic_params_given_x_d(self, x_d):
        key = tuple(int(v) for v in x_d)
        return self.A[key], self.b[key], self.c[key]

    def get(self, args, ignore_const=False):
        A, b, c = self.get_quadratic_params_given_x_d(args[:self.Nd])
        x = np.asarray(args[self.Nd:], dtype=float)
        lv = float(x @ A @ x + b @ x)
        return math.exp(lv if ignore_const else lv + float(c))


class ImageNodePotential(Potential):
    """Gaussian pdf of (x0 - x1 - mu) (reference ``Potential.py:400-408``): exp-quadratic, lowered like
    any other potential."""

    def __init__(self, mu, sig):
        super().__init__(symmetric=True)
        self.mu = mu
        self.sig = sig

    def get(self, parameters):
        u = (parameters[0] - parameters[1] - self.mu) / self.sig
        return math.exp(-0.5 * u * u) / (2.506628274631 * self.sig)


class ImageEdgePotential(Potential):
    """Truncated-Laplacian smoothness prior of the denoising demo (reference ``Potential.py:411-424``).
    Not exp-quadratic: its three coefficients travel in the coefficient block and the device evaluates
    ``log(psi + 1e-100)`` at every grid point (``LHVI_POT_IMAGE_EDGE``, generic kernel)."""

    def __init__(self, distant_cof, scaling_cof, max_threshold):
        super().__init__(symmetric=True)
        self.distant_cof = distant_cof
        self.scaling_cof = scaling_cof
        self.max_threshold = max_threshold
        self.v = math.exp(-max_threshold / scaling_cof)

    def get(self, parameters):
        d = abs(parameters[0] - parameters[1])
        tail = self.v if d > self.max_threshold else math.exp(-d / self.scaling_cof)
        return d * self.distant_cof + tail
