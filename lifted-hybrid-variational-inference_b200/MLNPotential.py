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


def and_op(x, y):
    return x * y


def or_op(x, y):
    return x + y - x * y


def neg_op(x):
    return 1 - x


def imp_op(x, y):
    return 1 - x + x * y


def bic_op(x, y):
    return imp_op(x, y) * imp_op(y, x)


def eq_op(x, y):
    d = x - y
    return -(d * d)


class MLNPotential(Potential):
    def __init__(self, formula, w=1):
        super().__init__(symmetric=False)
        self.formula = formula
        self.w = w

    def get(self, parameters):
        return math.e ** (self.formula(parameters) * self.w)


class MLNHardPotential(Potential):
    """Indicator of ``formula(x) > 0``; only lowerable when every argument is discrete."""

    def __init__(self, formula):
        super().__init__(symmetric=False)
        self.formula = formula

    def get(self, parameters):
        return 1 if self.formula(parameters) > 0 else 0
