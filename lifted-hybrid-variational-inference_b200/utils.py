"""Evaluation helper used by the engines' ``log_fe=False`` logging mode."""
from __future__ import annotations

import math

import numpy as np


def log_likelihood(g, assignment):
    """Negative log of the unnormalised joint at ``assignment`` (reference ``utils.py:6-15``);
    ``-inf`` as soon as one potential vanishes."""
    total = 0.0
    for f in g.factors:
        value = float(np.asarray(f.potential.get([assignment[rv] for rv in f.nb])).reshape(-1)[0])
        if value == 0:
            return -np.inf
        total += math.log(value)
    return -total
