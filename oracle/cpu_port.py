"""ORACLE (test infrastructure) -- CPU runner used as bench.py's ``cpu_baseline`` / reference arm.

Wraps the fastest available CPU restatement of one VI iteration over a lowered model:
the C/OpenMP port (``oracle/_build/libvi_oracle.so``, all host threads) when it has been built
by ``__graft_entry__.build()``, else the single-threaded numpy port.  Never imported by the
product path.
"""
from __future__ import annotations

import os

from oracle.vi_numpy import NumpyVI


class NumpyRunner:
    cores = 1
    describe = "numpy fp64 port, 1 thread"

    def __init__(self, model, eta, tau, w_tau):
        self.vi = NumpyVI(model)
        self.vi.eta[:], self.vi.tau[:], self.vi.w_tau = eta, tau, w_tau
        self.vi.refresh()

    def step(self, lr):
        return self.vi.adam_step(lr)


def make_runner(model, eta, tau, w_tau):
    try:
        from oracle import c_port
        if c_port.available():
            return c_port.CRunner(model, eta, tau, w_tau)
    except ImportError:
        pass
    return NumpyRunner(model, eta, tau, w_tau)
