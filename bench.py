#!/usr/bin/env python
"""Benchmark of the VI update loop at 10 M ground factors (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    torchrun --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A *step* is one Jacobi VI iteration over the whole record table: node-entropy + factor
expectation/gradient kernels for every record group, lhvi_finish (ELBO reduction, step counter
and -- N>1 -- the cross-GPU sum of G_w, the free energy and the shared variables' gradients over
NVLink peer memory) and the Adam parameter step.

Workload (``config.workload``): the relational hybrid model of SURVEY section 8 d config 5 in its
fully refined C2F state -- 1 M entities x 10 groups = 10 M hybrid link factors + 1 M priors +
90 session factors, lifted record format (W_f, gamma columns shipped), K=3 mixtures, Gauss-
Hermite degree 3, fp32 arithmetic.  With N GPUs the records are partitioned owner-computes
(dist.py; strong scaling: the total stays 10 M): entity variables live on one rank together with
their records, only the group variables are shared.

Printed JSON (one line, rank 0): see the task contract; ``roofline`` is for the dominant
kernel (the largest record group's launch), ``cpu_baseline`` times the CPU port of the same
pass (oracle/) on a bounded sample on this box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "vi_iterations_per_sec_at_10M_ground_factors"
UNIT = "it/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--entities", type=int, default=1_000_000)
    ap.add_argument("--groups", type=int, default=10)
    ap.add_argument("--K", type=int, default=3)
    ap.add_argument("--T", type=int, default=3)
    ap.add_argument("--dtype", default="float32", choices=["float32", "float64"])
    ap.add_argument("--order", default="hub", choices=["hub", "entity"])
    ap.add_argument("--generic", action="store_true", help="force the generic kernel")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c2f", action="store_true", help="skip the auxiliary coarse-to-fine figure")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the auxiliary figures of BASELINE configs 1-4, the fp64 line and the drop-in classes")
    ap.add_argument("--config", default="all",
                    help="which auxiliary configs to run beside the headline (config 5): all, or a list like 1,3")
    ap.add_argument("--cpu-sample-entities", type=int, default=200_000)
    ap.add_argument("--write-reference", action="store_true",
                    help="N=1 only: store the free energy of the initial state for the checks of the N>1 runs")
    return ap.parse_args()


def workload_config(a, n_records, parallelism=None):
    return {
        "workload": f"config5 relational hybrid MLN, C2F fully-refined state: {a.entities} entities x "
                    f"{a.groups} groups = {a.entities * a.groups} link factors (+priors, sessions), "
                    f"lifted record format, K={a.K}, T={a.T}",
        "factor_records": int(n_records),
        "K": a.K, "T": a.T,
        "record_order": f"generator: {a.order}-major; the engine re-sorts the two-hidden-argument group run-major "
                        "and pads hub runs of the streamed group to whole tiles (results are order-independent)",
        "l2_policy": "inputs larger than L2 (record table ~1 GB >> 126 MB), no explicit flush",
        # (the same text in both arms; what the partition came out as is reported beside `config`)
        "parallelism": f"{a.gpus} rank(s), owner-computes record partition",
    }


# ---- algorithmic bytes (DESIGN.md "Roofline accounting") -------------------------------------

def is_streamed(g):
    """Pure group with one hidden continuous argument in a hub run: served by the streaming
    unary kernel from the folded columns (c0, l0, a0)."""
    return (not g.node) and g.pure and g.nd == 0 and g.nc == 1 and g.ng == 0


def group_bytes(g, K, s, hub=False):
    """Algorithmic bytes one launch over record group ``g`` must move.

    General groups (SURVEY section 8 d): the record columns once, plus per hidden argument one
    parameter gather and one gradient write of K*P elements.  Streamed hub groups: the folded
    record columns only -- (c0, l0, a0), the slot offset and the two lifted weights; the
    variable's parameters and gradient are touched once per run, not per record."""
    if hub and is_streamed(g):
        return (3 * s + 4 + (2 * s if g.weighted else 0)) * g.n
    per = 0
    if not g.node:
        per += 4                                    # pot
    per += 4 * g.nh                                 # parameter offsets
    per += s * (2 * g.ng + g.ne)                    # evidence columns
    if g.node:
        per += 2 * s                                # energy scale, gradient scale
    elif g.weighted:
        per += s * (1 + g.nh)                       # W_f, gamma
    elems = sum(K * d for d in g.dims) + 2 * K * g.nc
    per += 2 * s * elems                            # gather + gradient write
    return per * g.n


def variable_bytes(model, s):
    """Optimiser step: read grad, theta, m, v; write theta, m, v (7 * s per element)."""
    return 7 * s * model.n_param


def traffic_from_profile(persistent, group_name, a, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel, per launch, read from the
    ncu --set full summary committed under profiles/ (profiles/traffic.json: written by
    tools/ncu_traffic.py from the capture of this same command); None when there is no capture of
    this workload (other sizes, several GPUs)."""
    if world != 1 or a.entities != 1_000_000 or a.groups != 10:
        return None
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(path):
        return None
    table = json.load(open(path))
    key = "iterate_kernel" if persistent else group_name.split(" n=")[0].strip()
    entry = table.get(f"{key} {a.dtype} K={a.K}")
    if entry is None:
        return None
    if persistent:          # captured per launch of `steps_per_launch` iterations
        return entry["dram_bytes"] / entry["steps_per_launch"] * a.steps
    return entry["dram_bytes"]
NOTE = ("the full (two hidden arguments) group runs in the run-major kernel, which also evaluates the node-entropy and "
        "unary records of its run variables (one launch instead of three): 571 instructions per link record, issue "
        "slots 66% busy, FMA pipe 43%, XU 29%, 16 resident warps per SM (128 registers), DRAM at 15% -- its "
        "limiter is instruction issue / latency at that occupancy, not HBM: the entity slots are hit ten times each "
        "and stay in L1/L2, so physical traffic (92 MB) is a quarter of the algorithmic bytes SURVEY 8d counts "
        "(398 MB with the fused records) and the fraction of the HBM roofline is reported as asked "
        "(profiles/r2_ncu_summary.md)")


# ---- clocks ------------------------------------------------------------------------------------

class ClockSampler:
    QUERY = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._nvml = None

    def _init_nvml(self):
        """NVML polling (about a microsecond per query, so that even a timed region of a few
        milliseconds is sampled many times).  Set up before the thread starts: loading the
        library takes longer than the timed region.  False if NVML is not usable here."""
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
                pynvml.nvmlDeviceGetCurrentClocksThrottleReasons
            float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            int(reasons_fn(h))
        except Exception:
            return False
        self._nvml = (pynvml, h, reasons_fn)
        return True

    def _run_nvml(self):
        if self._nvml is None:
            return False
        pynvml, h, reasons_fn = self._nvml
        bits = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
        while not self._stop.is_set():
            try:
                self.samples.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                mask = int(reasons_fn(h))
                for bit, name in bits.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.0005)
        return True

    def _run(self):
        if self._run_nvml():
            return
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                self.samples.append(float(parts[0]))
                self.max_mhz = float(parts[1])
                for n, v in zip(names, parts[2:]):
                    if v.lower().startswith("active"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def __enter__(self):
        self._init_nvml()
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---- CPU baseline (the oracle port, bounded sample) ---------------------------------------------

def cpu_baseline(a, n_records_full, seconds=12.0):
    """Time the CPU restatement of one iteration (oracle/) on a sub-sampled model of the same
    generator and scale linearly in the number of factor records (SURVEY section 8 d)."""
    import lhvi_b200
    from oracle import cpu_port
    syn = lhvi_b200.synthetic
    sample_P = min(a.entities, a.cpu_sample_entities)
    model = syn.relational_hybrid(sample_P, a.groups, a.K, a.T, seed=0, order=a.order, weighted=True)
    eta, tau, w_tau = syn.random_state(model, 0)
    runner = cpu_port.make_runner(model, eta, tau, w_tau)
    runner.step(0.1)                                    # warm-up
    t0 = time.perf_counter()
    its = 0
    while its < 3 or (time.perf_counter() - t0 < seconds and its < 200):
        runner.step(0.1)
        its += 1
    dt = (time.perf_counter() - t0) / its
    rec_per_s = model.n_records / dt
    sample = (f"{its} iterations of the same generator at {sample_P} entities ({model.n_records} factor "
              f"records, {dt * 1e3:.1f} ms/iteration, {runner.describe}); scaled linearly to "
              f"{n_records_full} records")
    # the unmodified Python reference cannot travel to the GPU box: its figure for this generator was
    # measured once in the build container (tools/ref_time.py) and is quoted here for scale
    ref_path = os.path.join(ROOT, "profiles", "r2_reference_python_timing.json")
    if os.path.exists(ref_path):
        ref = json.load(open(ref_path))
        parts = [f"{k} {v['ground_factors_per_s']:.0f} ground factors/s ({v['s_per_iter']:.2f} s/iteration at "
                 f"{v['ground_factors']} factors)" for k, v in ref.items() if isinstance(v, dict)]
        hours = n_records_full / min(v["ground_factors_per_s"] for v in ref.values() if isinstance(v, dict)) / 3600
        sample += ("; the unmodified Python reference on the object-graph twin of this generator, one core of the "
                   "build container: " + ", ".join(parts) + f" -- about {hours:.1f} hours per iteration at this size")
    return {
        "value": rec_per_s / n_records_full, "unit": UNIT, "cores": runner.cores, "kind": "port",
        "sample": sample,
        "records_per_s": rec_per_s,
    }


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # all host threads, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1 for N > 1, which
    # made the reference arm 16x slower at N >= 2 than at N = 1 in round 1): must be set before the
    # OpenMP runtime of the C port starts
    threads = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(threads)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    n_full = a.entities * a.groups + a.entities + a.groups * (a.groups - 1)
    t0 = time.perf_counter()
    base = cpu_baseline(a, n_full, seconds=max(5.0, 2.0 * (a.steps + a.warmup)))
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": a.gpus,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 / base["value"], "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(a, n_full),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": time.perf_counter() - t0,
    }
    print(json.dumps(line))


# ---- our arm ------------------------------------------------------------------------------------

def c2f_probe(a, iterations=100):
    """Auxiliary figure beside the headline (not part of `value`): the engine BASELINE config 5 names, the
    coarse-to-fine loop (C2FVarInference.py:301-352: evidence split, colour passing, re-lowering and upload
    between blocks of ten iterations) over the WHOLE workload -- every entity, 100 iterations = 10 refinement
    rounds -- through `lifting.C2FArrayVI`, whose lifting passes run on the GPU next to the kernels
    (`lifting_torch`: sort / unique / prefix-sum passes on the resident ground graph).  The same run with the
    passes in the host library (lhvi_lift.cpp, all host threads) is in profiles/r2_c2f_host_passes.jsonl."""
    import lhvi_b200
    lifting, syn = lhvi_b200.lifting, lhvi_b200.synthetic
    entities = max(1000, a.entities)
    # warm-up on a small model: the first launch of a kernel pays its module load (the persistent
    # iteration kernel is tens of megabytes of SASS), which is not part of a refinement round
    # (large enough that the sort / unique / scan kernels the lifting passes use at full size are loaded too)
    warm_entities = 50_000 if entities >= 500_000 else 2000
    warm = lifting.C2FArrayVI(syn.relational_hybrid_arrays(warm_entities, a.groups, observed_frac=0.7, seed=1), a.K, a.T, dtype=a.dtype)
    warm.run(20, 0.05)
    ga = syn.relational_hybrid_arrays(entities, a.groups, observed_frac=0.7, seed=0)
    vi = lifting.C2FArrayVI(ga, a.K, a.T, dtype=a.dtype)
    t0 = time.perf_counter()
    vi.run(iterations, 0.05)
    total = time.perf_counter() - t0
    passes = sum(vi.timing[k] for k in ("split", "refine", "lower"))
    return {"engine": "C2FArrayVI", "ground_factors": ga.n_factors, "iterations": iterations, "rounds": len(vi.history),
            "lifting_passes": "GPU-resident (lifting_torch)" if vi._passes_device() is not None else "host library (lhvi_lift.cpp)",
            "classes_per_round": [n for n, _ in vi.history], "records_last_round": vi.model.n_records,
            "run_s": round(total, 4), "setup_s": round(vi.timing.get("setup", 0.0), 4),
            "lifting_passes_s": round(passes, 4), "upload_s": round(vi.timing["upload"], 4),
            "device_iterations_s": round(vi.timing["iterate"], 4), "readback_s": round(vi.timing["pull"], 4),
            "per_round_s": [{k: round(v, 4) for k, v in r.items()} for r in vi.timing_rounds],
            "host_route_run_s": {"value": 17.301, "source": "profiles/r2_c2f_host_passes.jsonl (same model and iterations, B200 box host)"}
            if (entities == 1_000_000 and iterations == 100 and a.groups == 10) else None,
            "free_energy_finite": bool(np.isfinite(vi.free_energy()))}


def run_ours(a):
    import torch
    import torch.distributed as dist

    import lhvi_b200
    from lhvi_b200.engine import DeviceEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    # rank 0 prints ONE JSON line on stdout: anything a library writes to file descriptor 1 meanwhile
    # (NCCL's version banner is a C printf) goes to stderr instead
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    syn = lhvi_b200.synthetic
    model = syn.relational_hybrid(a.entities, a.groups, a.K, a.T, seed=0, order=a.order, weighted=True)
    eta, tau, w_tau = syn.random_state(model, 0)
    eng = DeviceEngine(model, dtype=a.dtype, device=f"cuda:{local}", force_generic=a.generic)
    eng.set_state(eta, tau, w_tau)
    eng.reset_moments()
    s = 4 if a.dtype == "float32" else 8
    lr = 0.1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def per_record(pred):
        """Algorithmic bytes per record of the (unfused) groups of the model that match ``pred``."""
        gs = [g for g in eng.full_model.groups if g.n > 0 and pred(g)]
        return group_bytes(gs[0], a.K, s) / gs[0].n if gs else 0.0

    def gbytes(i):
        # records that the engine moved into another group's kernel (fuse_run_extras / fuse_constants) are
        # counted where they are evaluated, with the bytes SURVEY 8d gives them (the accounting does not
        # depend on the layout)
        d, keep, g = eng.groups[i]
        total = group_bytes(g, a.K, s, hub=bool(d.fold) and bool((d.hub_mask >> g.nd) & 1))
        if keep.get("fused"):
            total += keep["fused"][0] * per_record(lambda h: h.node and h.nc == 1 and h.nd == 0)
            total += keep["fused"][1] * per_record(lambda h: h.pure and not h.node and h.nc == 1 and h.ne == 0 and h.nd == 0)
        if keep.get("fused_constants"):
            total += keep["fused_constants"] * per_record(lambda h: h.pure and not h.node and h.nc == 0 and h.nd == 0 and h.ng == 0)
        return int(total)

    def gname(i):
        d, _, g = eng.groups[i]
        kind = "node" if g.node else ("pure" if g.pure else "full")
        stream = " streamed" if (d.fold and (d.hub_mask >> g.nd) & 1) else ""
        fused = eng.groups[i][1].get("fused")
        tail = " [run-major]" if d.run_start else ""
        if fused:
            tail = f" [run-major, with {fused[0]} node-entropy and {fused[1]} unary records of its run variables]"
        if eng.groups[i][1].get("fused_constants"):
            tail += f" [with {eng.groups[i][1]['fused_constants']} constant records]"
        return f"{kind}{stream} nd={g.nd} nc={g.nc} ng={g.ng} ne={g.ne} n={g.n}" + tail

    # the free energy of the (deterministic) initial state must not depend on how the records are
    # sharded: every run computes it once (dense sum over the ranks) and compares it with the value an
    # N=1 run stored under profiles/ (fp32 sums of 11 M terms, double partial sums: 1e-5 relative)
    fe0 = float(eng.free_energy())
    ref_key = f"{a.entities}x{a.groups} K={a.K} T={a.T} {a.dtype} {a.order}"
    ref_path = os.path.join(ROOT, "profiles", "r2_free_energy_initial.json")
    ref_table = json.load(open(ref_path)) if os.path.exists(ref_path) else {}
    if a.write_reference and world == 1:
        ref_table[ref_key] = fe0
        for path in (ref_path, os.path.join(ROOT, "gpurun_out", "r2_free_energy_initial.json")):
            if os.path.isdir(os.path.dirname(path)):      # (gpurun_out/ is what travels back from a GPU box)
                json.dump(ref_table, open(path, "w"), indent=1, sort_keys=True)
    fe0_ref = ref_table.get(ref_key)
    if fe0_ref is not None and not abs(fe0 - fe0_ref) <= 1e-5 * abs(fe0_ref):
        raise SystemExit(f"bench.py: free energy of the initial state {fe0!r} differs from the single-GPU value "
                         f"{fe0_ref!r} ({world} ranks): the sharded pass does not compute the same sums")

    # a step = one Jacobi iteration; `iterate(n)` is the reference's `for itr in range(iteration)` of
    # ADAM_update (VarInference.py:249-300): on the persistent path the K timed steps are ONE
    # cooperative launch that loops K times (grid barriers instead of launch boundaries), else K
    # replays of the captured per-group launches
    eng.iterate(a.warmup, lr)
    barrier()
    persistent = eng.persistent()

    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    repeats = []
    with ClockSampler(local) as clocks:
        for rep in range(5):                   # the timed region, five times: min / median reported beside it
            barrier()
            l0 = eng.launch_count
            ev0.record()
            eng.iterate(a.steps, lr)
            ev1.record()
            barrier()
            repeats.append(ev0.elapsed_time(ev1))
            launches = eng.launch_count - l0
        # one call per step (what a caller who reads the state after every ADAM_update(1) gets: calls of a
        # single iteration replay the captured per-group launches)
        eng.iterate(1, lr)                      # (the graph is captured at its first use)
        barrier()
        ev0.record()
        for _ in range(a.steps):
            eng.iterate(1, lr)
        ev1.record()
        barrier()
        ms_single = ev0.elapsed_time(ev1) / a.steps
        # the same steps again, launched one kernel after the other with CUDA events around every
        # group's launch (each kernel timed alone -> burst peak applies)
        eng.profile_group = "all"
        eng.dom_events = []
        for _ in range(a.steps):
            eng.iterate(1, lr)
        barrier()
    ms = repeats[0]                            # the headline is the FIRST timed region of exactly K steps
    per_group = {}
    for i, b, e in eng.dom_events:
        per_group.setdefault(i, []).append(b.elapsed_time(e))
    kernels = [{"group": gname(i), "ms": float(np.mean(v)), "bytes": gbytes(i),
                "gbs": gbytes(i) / (float(np.mean(v)) * 1e-3) / 1e9} for i, v in sorted(per_group.items())]
    dom = max(per_group, key=lambda i: np.mean(per_group[i])) if per_group else 0
    dom_ms = float(np.mean(per_group[dom])) if per_group else None
    eng.profile_group = None

    # ---- end to end through the public engine API with host-resident parameters: the state a
    # caller of ADAM_update holds (eta, and the logits when the model has discrete variables)
    # goes up from pinned memory, one iteration runs, the new state and [G_w | free energy] come back
    # The host holds the reference's compact arrays (eta[rv]: K x 2 / K x D per variable, no slot
    # padding); lhvi_state_unpack / lhvi_state_pack move between them and the device slots.
    n = model.n_param
    np_t = np.float32 if s == 4 else np.float64
    has_disc = bool((model.var_kind == 1).any())
    eng.packed_map(local=world > 1)        # N ranks: each holds the variables it steps (owned + shared)
    pidx = eng.packed_index
    n_packed = int(pidx.size)
    host_eta = torch.from_numpy(eta[pidx].astype(np_t)).pin_memory()
    host_tau = torch.from_numpy(tau[pidx].astype(np_t)).pin_memory() if has_disc else None
    dev_eta = torch.empty(n_packed, dtype=host_eta.dtype, device=eng.device)
    dev_tau = torch.empty(n_packed, dtype=host_eta.dtype, device=eng.device) if has_disc else None
    host_tail = torch.empty(a.K + 1, dtype=host_eta.dtype).pin_memory()
    e2e_steps = max(3, a.steps // 2)

    def e2e_step():
        dev_eta.copy_(host_eta, non_blocking=True)
        eng.unpack_state(dev_eta, "eta")
        if has_disc:
            dev_tau.copy_(host_tau, non_blocking=True)
            eng.unpack_state(dev_tau, "tau")
        eng.iterate(1, lr)
        eng.pack_state(dev_eta, "eta")
        host_eta.copy_(dev_eta, non_blocking=True)
        if has_disc:
            eng.pack_state(dev_tau, "tau")
            host_tau.copy_(dev_tau, non_blocking=True)
        host_tail.copy_(eng.grad[n:], non_blocking=True)          # G_w and the free energy
        torch.cuda.synchronize()
        return float(host_tail[-1])

    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fe_last = e2e_step()
    barrier()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    h2d = (2 if has_disc else 1) * n_packed * s
    d2h = h2d + (a.K + 1) * s
    eng.check_exchange()

    t = torch.tensor([ms, e2e_ms, ms_single] + repeats, dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, ms_single = float(t[0]), float(t[1]), float(t[2])
    repeats = [float(v) / a.steps for v in t[3:]]
    if world > 1:                               # bytes of the whole job: sum over the ranks
        b = torch.tensor([h2d, d2h], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        h2d, d2h = int(b[0]), int(b[1])

    if rank == 0:
        ms_per_step = ms / a.steps
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = json.load(open(peaks_path))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        step_bytes = sum(gbytes(i) for i in range(len(eng.groups))) + variable_bytes(eng.model, s)
        if persistent:
            # the dominant (only) kernel of the timed region is the persistent iteration kernel: one launch
            # covers K whole steps, so bytes per launch = K x the step's algorithmic bytes and the launch
            # duration is the timed region itself
            dom_bytes, dom_kernel_ms = step_bytes * a.steps, ms
            dom_name = ("iterate_kernel (persistent: every record group + ELBO reduction + optimiser step, "
                        f"{a.steps} iterations per launch)")
        else:
            dom_bytes, dom_kernel_ms = gbytes(dom), dom_ms
            dom_name = f"longest launch of the step: {gname(dom)} (rank 0)"
        achieved = dom_bytes / (dom_kernel_ms * 1e-3) / 1e9 if dom_kernel_ms else None
        roofline = {
            "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak if achieved else None,
            "traffic": traffic_from_profile(persistent, gname(dom), a, world),
            "kernel": dom_name,
            "kernel_ms": dom_kernel_ms, "kernel_bytes": dom_bytes, "peak_source": peak_src,
            "note": NOTE,
            "kernels_launched_alone": kernels,
            "step_bytes": step_bytes, "step_achieved": step_bytes / (ms_per_step * 1e-3) / 1e9,
            "step_frac": step_bytes / (ms_per_step * 1e-3) / 1e9 / peak,
        }
        line = {
            "metric": METRIC, "value": 1e3 / ms_per_step, "unit": UNIT, "n_gpus": world, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32" if s == 4 else "f64", "data": "synthetic",
            "config": workload_config(a, model.n_records),
            "partition": eng.plan.describe() + (f"; exchange={eng.exchange}" if eng.exchange else ""),
            "clocks": clocks.summary(),
            "e2e": {"value": 1e3 / e2e_ms, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "note": "per step: variational parameters (compact eta[rv] arrays) host->device from pinned "
                            "memory, lhvi_state_unpack, one iteration, lhvi_state_pack, new parameters + G_w + "
                            "free energy device->host; record table resident; with N ranks every rank "
                            "moves the variables it owns plus the shared ones (bytes are summed over the ranks)",
                    "free_energy_last": fe_last},
            "free_energy_initial": {"value": fe0, "single_gpu_reference": fe0_ref,
                                    "checked": fe0_ref is not None, "rtol": 1e-5},
            "gpu_launches": launches,
            "launch_mode": ("persistent: the K timed steps are one cooperative launch of lhvi_iterate" if persistent
                            else "CUDA graph of per-group launches, replayed K times"),
            "ms_per_step_repeats": {"all": repeats, "min": min(repeats), "median": float(np.median(repeats))},
            "ms_per_step_one_launch_per_step": ms_single,
            "roofline": roofline,
        }
        if world == 1 and not a.no_cpu_baseline:
            line["cpu_baseline"] = {k: v for k, v in cpu_baseline(a, model.n_records).items()
                                    if k in ("value", "unit", "cores", "kind", "sample")}
        if world == 1 and not a.no_c2f:
            try:
                line["c2f"] = c2f_probe(a)
            except Exception as exc:            # an auxiliary figure must not cost the bench line
                line["c2f"] = {"error": repr(exc)[:300]}
        if world == 1 and not a.no_configs:
            # free the headline engine's record table first (config 4 and the fp64 line need the room)
            eng.close()
            del eng
            torch.cuda.empty_cache()
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_configs
            wanted = list(bench_configs.ALL) if a.config == "all" else [c.strip() for c in a.config.split(",")]
            line["configs"] = {}
            for c in wanted:
                if c not in bench_configs.ALL:
                    continue
                t0 = time.perf_counter()
                try:
                    line["configs"][f"config{c}"] = bench_configs.ALL[c]()
                except Exception as exc:
                    line["configs"][f"config{c}"] = {"error": repr(exc)[:300]}
                line["configs"][f"config{c}"]["wall_s"] = round(time.perf_counter() - t0, 2)
                torch.cuda.empty_cache()
            for key, fn in (("fp64", lambda: bench_configs.fp64_headline(a)), ("dropin", bench_configs.dropin_rgm)):
                try:
                    line[key] = fn()
                except Exception as exc:
                    line[key] = {"error": repr(exc)[:300]}
                torch.cuda.empty_cache()
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        # leave without tearing NCCL down: destroy_process_group() after CUDA-graph capture of a
        # collective was seen to hang at exit; every rank has passed its last collective here
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
