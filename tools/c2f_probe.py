"""Coarse-to-fine run of the config-5 model family on arrays, with the time of every phase:
evidence split, colour passing, lowering (host), engine construction + upload, the iterations on
the device, and the read-back.

    python tools/c2f_probe.py [entities=100000] [iterations=50] [groups=10] [dtype=float32]

Prints one JSON line.  Needs a CUDA device (the engine has no CPU fallback)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np

import lhvi_b200


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    P = int(args[0]) if len(args) > 0 else 100_000
    its = int(args[1]) if len(args) > 1 else 50
    G = int(args[2]) if len(args) > 2 else 10
    dtype = args[3] if len(args) > 3 else "float32"
    lifting, syn = lhvi_b200.lifting, lhvi_b200.synthetic
    t0 = time.perf_counter()
    ga = syn.relational_hybrid_arrays(P, G, observed_frac=0.7, seed=0)
    t_gen = time.perf_counter() - t0
    # CUDA context, library load and first launches on a toy model, outside the timed run
    t0 = time.perf_counter()
    toy = lifting.C2FArrayVI(syn.relational_hybrid_arrays(50, 3, seed=1), 3, 3, dtype=dtype, device_passes=False)
    toy.run(10, 0.05)
    t_warm = time.perf_counter() - t0
    passes = False if "--host-passes" in sys.argv else None            # None: resident (lifting_torch) on the GPU
    if passes is None:                                                   # (graph upload + first torch launches, untimed)
        lifting.C2FArrayVI(syn.relational_hybrid_arrays(50, 3, seed=1), 3, 3, dtype=dtype).run(10, 0.05)
    vi = lifting.C2FArrayVI(ga, 3, 3, dtype=dtype, device_passes=passes)
    t0 = time.perf_counter()
    vi.run(its, 0.05)
    total = time.perf_counter() - t0
    fe = float(vi.free_energy())
    out = {"probe": "c2f_arrays", "entities": P, "groups": G, "ground_variables": ga.n_vars,
           "ground_factors": ga.n_factors, "K": 3, "T": 3, "dtype": dtype, "iterations": its,
           "rounds": len(vi.history), "classes_per_round": [n for n, _ in vi.history],
           "records_last_round": vi.model.n_records, "generate_s": round(t_gen, 3), "run_s": round(total, 3),
           "phases_s": {k: round(v, 4) for k, v in vi.timing.items()}, "device_warmup_s": round(t_warm, 3),
           "phases_per_round_s": [{k: round(v, 4) for k, v in r.items()} for r in vi.timing_rounds],
           "native_lifting": lhvi_b200._lift_native.load() is not None,
           "lifting_passes": "host (lhvi_lift.cpp / numpy)" if passes is False else "resident on the GPU (lifting_torch)",
           "free_energy": fe, "finite": bool(np.isfinite(fe))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
