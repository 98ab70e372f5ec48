"""Gaussian belief propagation on the config-4 grid (n x n), a few sweeps: the launch ncu profiles."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from lhvi_b200 import GaBP as gabp
from tools.bench_configs import grid_with_observations

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
model, J, h = grid_with_observations(n, 1, 3)
bp = gabp.DeviceGaBP(gabp.gabp_from_model(model), "float32")
bp.sweeps(10)
torch.cuda.synchronize()
print("directed edges", int(bp.a.src.size), "sweeps", bp.sweeps_done)
