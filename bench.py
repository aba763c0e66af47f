#!/usr/bin/env python
"""bench.py -- throughput of the FieldCalculations hot path on B200 (contract: task brief, section 4).

Workload (BASELINE.json configs[1], "meps65_alevel_chain"): the MEPS 65-level atmospheric-level
temperature/humidity conversion chain.  One STEP = one lead time = 65 levels of 949x1069 points
through the reference call sequence
    aleveltemp(c=3: T -> theta), alevelhum(c=1: T,q -> RH), alevelhum(c=5: T,q -> Td), alevelthe(c=1: T,q -> theta_e)
i.e. three input variables (t, q, p) and four outputs per grid point.  `value` = grid points pushed
through the whole chain per second (whole job, all ranks), inputs resident in HBM.  `e2e` = the same
chain called through the C-ABI with HOST (pinned) buffers, host<->device copies inside the timed
region.  Every step streams > 700 MB of inputs, far more than the 126 MB L2, and consecutive steps
alternate between two input sets, so nothing is served from cache ("inputs larger than L2").

    python bench.py [--gpus N] [--steps K] [--warmup W]            # product arm
    python bench.py --impl reference [--gpus N] [--steps K] ...    # the reference's CPU code (oracle/_ref)
    torchrun --nproc-per-node N bench.py --gpus N ...              # one rank per GPU, weak scaling by batch
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

NX, NY, NLEV, NTIMES = 949, 1069, 65, 67
N = NX * NY
UNDEF = 1.0e35
METRIC = "grid points/s (MEPS 65-level alevel T/q conversion chain: theta, RH, Td, theta_e)"
WORKLOAD = "meps65_alevel_chain"
# algorithmic HBM bytes per grid point: unfused = the four reference calls
# (12 + 16 + 16 + 16), fused = t, q, p read once + four outputs
BYTES_UNFUSED = {"aleveltemp": 12, "alevelhum_rh": 16, "alevelhum_td": 16, "alevelthe": 16}
BYTES_FUSED = 28
# dram__bytes_read.sum + dram__bytes_write.sum of one fused-chain launch (65 levels), from the ncu --set full
# capture committed under profiles/ (None until captured)
TRAFFIC_NCU = 1800.5e6  # profiles/r01_ncu_full_chain_final_summary.csv: 792.6 MB read + 1007.9 MB written (algorithmic: 1846 MB)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------- synthetic input
def synth_level_set(rng, nlev, dtype=np.float32):
    """t, q, p for `nlev` MEPS-sized levels: hybrid-level pressure a_k + b_k*ps,
    lapse-rate temperature + noise in [200, 310] K, q in [1e-6, 2e-2] kg/kg.  All defined."""
    y, x = np.mgrid[0:NY, 0:NX].astype(np.float32)
    ps = (1000.0 + 30.0 * np.sin(x / 97.0) * np.cos(y / 131.0)).astype(np.float32)
    t = np.empty((nlev, NY, NX), dtype)
    q = np.empty((nlev, NY, NX), dtype)
    p = np.empty((nlev, NY, NX), dtype)
    for k in range(nlev):
        eta = (k + 0.5) / nlev
        a_k, b_k = 200.0 * (1 - eta) * eta * 2.0 + 10.0 * (1 - eta), eta ** 1.5
        p[k] = (a_k + b_k * ps).astype(np.float32)
        tk = 288.0 * (p[k] / 1000.0) ** 0.19 + 4.0 * np.sin(x / 53.0 + k) + rng.standard_normal((NY, NX)).astype(np.float32)
        t[k] = np.clip(tk, 200.0, 310.0)
        q[k] = np.clip(2e-2 * eta ** 3 * (1.0 + 0.3 * np.cos(y / 71.0)) + 1e-6, 1e-6, 2e-2)
    return t, q, p


# ------------------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def samples_since(self, t0):
        return sum(1 for ts, _ in self.rows if ts >= t0)

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 6:
                continue
            try:
                smax = float(parts[1])
                if t0 - 0.05 <= ts <= t1 + 0.15:
                    sm.append(float(parts[0]))
                    for nme, val in zip(names, parts[2:6]):
                        if val.lower().startswith("active"):
                            reasons.add(nme)
            except ValueError:
                continue
        if not sm:  # region shorter than the sampling period: take what we have
            sm = [float(l.split(",")[0]) for _, l in self.rows if l and l.split(",")[0].strip().replace(".", "").isdigit()]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------- reference arm
def pick_reference_lib():
    """Serial (the reference's default build) or OpenMP (its optional build, <= 8 threads): whichever is
    faster on this host for one aleveltemp call -- some containers expose cores that do not run in parallel."""
    import fclibs
    threads = min(8, os.cpu_count() or 1)
    os.environ.setdefault("OMP_NUM_THREADS", str(threads))
    cands = []
    for omp in (True, False):
        lib = fclibs.reference(openmp=omp)
        if lib is not None:
            cands.append((lib, "reference", threads if omp else 1))
    if not cands:
        return fclibs.oracle(), "port", 1
    if len(cands) == 1:
        return cands[0]
    rng = np.random.default_rng(1)
    t = rng.uniform(220, 300, (NY, NX)).astype(np.float32)
    p = rng.uniform(300, 1000, (NY, NX)).astype(np.float32)
    o = np.empty_like(t)
    best = None
    for lib, kind, thr in cands:
        f = np.array([0], np.int32)
        lib.call("aleveltemp", NX, NY, t, p, "kelvin", 3, o, f, UNDEF)
        t0 = time.perf_counter()
        lib.call("aleveltemp", NX, NY, t, p, "kelvin", 3, o, f, UNDEF)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, lib, kind, thr)
    return best[1], best[2], best[3]


def run_reference(args):
    """The reference's own CPU implementation of the chain (oracle/_ref, OpenMP build: the reference
    never uses more than 8 threads, openmp_tools.cc:58-65), on a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    lib, kind, threads = pick_reference_lib()
    nlev = args.ref_levels
    rng = np.random.default_rng(2000)
    t, q, p = synth_level_set(rng, nlev)
    outs = [np.empty((NY, NX), np.float32) for _ in range(4)]

    def step():
        for k in range(nlev):
            f = np.array([0], np.int32)
            lib.call("aleveltemp", NX, NY, t[k], p[k], "kelvin", 3, outs[0], f, UNDEF)
            f[0] = 0
            lib.call("alevelhum", NX, NY, t[k], q[k], p[k], "celsius", 1, outs[1], f, UNDEF)
            f[0] = 0
            lib.call("alevelhum", NX, NY, t[k], q[k], p[k], "celsius", 5, outs[2], f, UNDEF)
            f[0] = 0
            lib.call("alevelthe", NX, NY, t[k], q[k], p[k], 1, outs[3], f, UNDEF)

    for _ in range(max(1, args.warmup)):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = nlev * N * args.steps / dt
    sample = "%d of %d levels per step (%d grid points), chain of 4 reference calls per level" % (nlev, NLEV, nlev * N)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "grid points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "grid": [NX, NY], "levels_per_step": nlev, "host": "cpu"},
        "cpu_baseline": {"value": value, "unit": "grid points/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": "grid points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def cpu_baseline_sample(nlev=6, reps=2):
    """Bounded CPU sample timed next to the GPU number (rank 0, N=1 only)."""
    lib, kind, threads = pick_reference_lib()
    rng = np.random.default_rng(2000)
    t, q, p = synth_level_set(rng, nlev)
    outs = [np.empty((NY, NX), np.float32) for _ in range(4)]
    best = None
    for _ in range(reps + 1):
        t0 = time.perf_counter()
        for k in range(nlev):
            f = np.array([0], np.int32)
            lib.call("aleveltemp", NX, NY, t[k], p[k], "kelvin", 3, outs[0], f, UNDEF)
            f[0] = 0
            lib.call("alevelhum", NX, NY, t[k], q[k], p[k], "celsius", 1, outs[1], f, UNDEF)
            f[0] = 0
            lib.call("alevelhum", NX, NY, t[k], q[k], p[k], "celsius", 5, outs[2], f, UNDEF)
            f[0] = 0
            lib.call("alevelthe", NX, NY, t[k], q[k], p[k], 1, outs[3], f, UNDEF)
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": nlev * N / best, "unit": "grid points/s", "cores": threads, "kind": kind,
            "sample": "%d of %d levels, best of %d after one warm-up pass, chain of 4 reference calls per level" % (nlev, NLEV, reps)}


class stdout_to_stderr:
    """NCCL prints its version banner on stdout when a communicator is created; the contract is ONE JSON line
    on stdout, so file descriptor 1 points at stderr while the process group comes up."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


# ------------------------------------------------------------------------------------- product arm
def chain_calls(gpu, t, q, p, outs, flags):
    """the reference call sequence, one batched launch per operator; returns the 4 kernel names"""
    nlev = t.shape[0]
    flags[:] = 0
    gpu.call("aleveltemp_batched", NX, NY, nlev, t, p, "kelvin", 3, outs[0], flags[0], UNDEF)
    gpu.call("alevelhum_batched", NX, NY, nlev, t, q, p, "celsius", 1, outs[1], flags[1], UNDEF)
    gpu.call("alevelhum_batched", NX, NY, nlev, t, q, p, "celsius", 5, outs[2], flags[2], UNDEF)
    gpu.call("alevelthe_batched", NX, NY, nlev, t, q, p, 1, outs[3], flags[3], UNDEF)


def run_product(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        with stdout_to_stderr():
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()  # creates the communicator (and prints NCCL's banner) now
            torch.cuda.synchronize()
    pkg = importlib.import_module("mi-fieldcalc_b200")
    gpu = pkg.load()
    gpu.set_device(local)
    dev = torch.device("cuda", local)
    peak, peak_src = peaks()

    # two input sets (alternated between steps), one output set; every rank owns its own lead times
    rng = np.random.default_rng(2000 + rank)
    nlev = args.levels
    t_h, q_h, p_h = synth_level_set(rng, nlev)
    sets = []
    for s in range(2):
        sets.append([torch.from_numpy(a).to(dev) for a in (t_h, q_h, p_h)])
        if s == 0:
            t_h = t_h[::-1].copy()  # second set: levels in reverse order
            q_h = q_h[::-1].copy()
            p_h = p_h[::-1].copy()
    outs = [torch.empty((nlev, NY, NX), dtype=torch.float32, device=dev) for _ in range(4)]
    flags = np.zeros((4, nlev), np.int32)
    stream = torch.cuda.current_stream()
    gpu.set_stream(stream.cuda_stream, True)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: K steps, CUDA events on the launching stream.
    # One step = ONE launch of the fused chain kernel (fcb200_alevel_chain_batched): t, q, p read once,
    # theta / RH / Td / theta_e written -- the same values, masks and flags as the four reference calls.
    fin = np.zeros(nlev, np.int32)            # ALL_DEFINED: what the reference call sequence would be given
    fout = np.full((4, nlev), -1, np.int32)

    def fused_step(t, q, p):
        gpu.call("alevel_chain_batched", NX, NY, nlev, t, q, p, "celsius", outs[0], outs[1], outs[2], outs[3], fin, fout, UNDEF)

    gpu.begin_deferred()
    for w in range(args.warmup):
        fused_step(*sets[w % 2])
    gpu.end_deferred()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    launches0 = gpu.launch_count()
    wall0 = time.time()
    gpu.begin_deferred()
    for k in range(args.steps):
        evs[k][0].record(stream)
        fused_step(*sets[k % 2])
        evs[k][1].record(stream)
    gpu.end_deferred()
    barrier()
    wall1 = time.time()
    launches = gpu.launch_count() - launches0
    total_ms = evs[0][0].elapsed_time(evs[-1][1])
    if world > 1:
        tt = torch.tensor([total_ms], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    assert (fout == 0).all(), "synthetic input is fully defined: every output flag must be ALL_DEFINED"
    # The timed region (K steps of ~0.4 ms) is shorter than nvidia-smi's 100 ms sampling period: keep the SAME load
    # running (identical untimed steps) until at least three samples have been taken under it.
    clocks = None
    if sampler:
        tail_steps = 0
        while sampler.samples_since(wall0) < 3 and time.time() - wall1 < 2.0:
            gpu.begin_deferred()
            for k in range(50):
                fused_step(*sets[k % 2])
            gpu.end_deferred()
            torch.cuda.synchronize()
            tail_steps += 50
        clocks = sampler.stop(wall0, time.time() if tail_steps else wall1)
        clocks["window"] = "timed region" if not tail_steps else "timed region + %d identical untimed steps right after it (the timed region is shorter than the 100 ms sampling period)" % tail_steps
    if world > 1:
        dist.barrier()
    kern_ms = float(np.mean([evs[k][0].elapsed_time(evs[k][1]) for k in range(args.steps)]))
    points_per_step = nlev * N
    value = world * points_per_step * args.steps / (total_ms * 1e-3)
    achieved = BYTES_FUSED * points_per_step / (kern_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": "ew_kernel<AlevelChainOpT<2, 2, 4>, 4> (fused chain)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": TRAFFIC_NCU, "peak_source": peak_src, "algorithmic_bytes_per_point": BYTES_FUSED,
                "kernel_ms": kern_ms, "note": "issue-bound, not HBM-bound: ~157 instructions per point (246 at the start of the round), 67-72 % issue-slot utilisation, DRAM 50 % busy (ncu, profiles/r01_ncu_full_chain_final_summary.csv; steps in profiles/r01_chain_tuning.txt)"}

    # for the record: the same step as the UNFUSED reference call sequence (four batched launches, 60 B/point)
    gpu.begin_deferred()
    chain_calls(gpu, *sets[0], outs, flags)
    gpu.end_deferred()
    barrier()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gpu.begin_deferred()
    u0.record(stream)
    for k in range(3):
        chain_calls(gpu, *sets[k % 2], outs, flags)
    u1.record(stream)
    gpu.end_deferred()
    barrier()
    unfused_ms = u0.elapsed_time(u1) / 3

    # ---- end to end: host (pinned) buffers through the C-ABI, copies inside the timed region
    e2e_lev = min(nlev, args.e2e_levels)
    hin = [torch.from_numpy(a[:e2e_lev].copy()).pin_memory() for a in synth_level_set(np.random.default_rng(3000 + rank), e2e_lev)]
    hout = [torch.empty((e2e_lev, NY, NX), dtype=torch.float32).pin_memory() for _ in range(4)]
    hfin = np.zeros(e2e_lev, np.int32)
    hfout = np.full((4, e2e_lev), -1, np.int32)

    def e2e_step():
        gpu.call("alevel_chain_batched", NX, NY, e2e_lev, hin[0], hin[1], hin[2], "celsius", hout[0], hout[1], hout[2], hout[3], hfin, hfout, UNDEF)

    for _ in range(min(3, max(1, args.warmup))):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        e2e_step()
    barrier()
    e2e_dt = time.perf_counter() - t0
    if world > 1:
        tt = torch.tensor([e2e_dt], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_dt = float(tt.item())
    assert (hfout == 0).all()
    e2e_value = world * e2e_lev * N * args.e2e_steps / e2e_dt
    e2e = {"value": e2e_value, "unit": "grid points/s", "h2d_bytes_per_step": 3 * 4 * e2e_lev * N, "d2h_bytes_per_step": 4 * 4 * e2e_lev * N,
           "levels_per_step": e2e_lev, "ms_per_step": 1e3 * e2e_dt / args.e2e_steps,
           "api": "fcb200_alevel_chain_batched with pinned host buffers (chunked, copy-in / kernel / copy-out pipelined over 3 streams)"}

    cpu = cpu_baseline_sample() if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "grid points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "grid": [NX, NY], "levels_per_step": nlev, "points_per_step_per_gpu": points_per_step,
                       "chain": "aleveltemp c3 + alevelhum c1 + alevelhum c5 + alevelthe c1, fused into one launch per step (t, q, p read once)",
                       "unfused_ms_per_step": unfused_ms,
                       "cache": "inputs larger than L2 (>= 790 MB streamed per step, two alternating input sets)", "sharding": "by field batch, no collective"},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="product", choices=["product", "reference"])
    ap.add_argument("--levels", type=int, default=NLEV, help="levels per step (default: the full 65)")
    ap.add_argument("--e2e-levels", type=int, default=NLEV)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-levels", type=int, default=4, help="levels per step of the bounded CPU sample (--impl reference)")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
