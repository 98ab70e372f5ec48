"""Markov-logic potentials: psi(x) = exp(w * formula(x)).

Same surface as the reference's ``MLNPotential.py`` (soft-logic operators ``:6-27``,
``MLNPotential`` ``:30-40``, ``MLNHardPotential`` ``:43-49``).  ``formula`` is an arbitrary
Python callable over the argument tuple; the lowering layer (``lowering.py``) turns it into
a per-discrete-configuration quadratic in the continuous arguments by exact finite
differences and refuses formulas that are not of that shape.
"""
from __future__ import annotations

import math

try:
    from .Graph import Potential
except ImportError:
    from Graph import Potential


# ---- soft-logic connectives ----------------------------------------------------------------------
# Truth values are 0 / 1 (or soft values in [0, 1]); every connective is a polynomial in its
# arguments, which is what lets ``lowering.py`` tabulate a formula per discrete configuration and fit
# the remaining continuous part exactly.

def and_op(x, y):
    """Conjunction as a product."""
    return x * y


def neg_op(x):
    """Negation: the complement to one."""
    return 1 - x


def or_op(x, y):
    """Disjunction by inclusion-exclusion."""
    both = x * y
    return x + y - both


def imp_op(x, y):
    """Implication ``x -> y`` = ``not x or y``, expanded: false only for x = 1, y = 0."""
    return 1 - x + x * y


def bic_op(x, y):
    """Biconditional: both implications hold."""
    forward, backward = imp_op(x, y), imp_op(y, x)
    return forward * backward


def eq_op(x, y):
    """Soft equality of two reals: minus the squared difference (0 at equality, the log of a
    Gaussian bump once scaled by the formula weight)."""
    d = x - y
    return -(d * d)


# ---- potentials ------------------------------------------------------------------------------------

class _FormulaPotential(Potential):
    """A potential defined by a Python callable over the argument tuple; never symmetric, compared
    by identity (two factors share a colour only if they share the object)."""

    def __init__(self, formula):
        super().__init__(symmetric=False)
        self.formula = formula


class MLNPotential(_FormulaPotential):
    """``psi(x) = exp(w * formula(x))`` (reference ``MLNPotential.py:30-40``)."""

    def __init__(self, formula, w=1):
        super().__init__(formula)
        self.w = w

    def get(self, parameters):
        return math.e ** (self.formula(parameters) * self.w)


class MLNHardPotential(_FormulaPotential):
    """Indicator of ``formula(x) > 0`` (reference ``MLNPotential.py:43-49``).  Over discrete arguments
    it is tabulated; with continuous arguments the formula (at most quadratic in them, like every
    formula of the reference) is lowered to its coefficient block and the device tests its sign at
    every grid point (``LHVI_POT_HARD``, generic kernel)."""

    def get(self, parameters):
        return 1 if self.formula(parameters) > 0 else 0
