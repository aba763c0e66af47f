#!/usr/bin/env python
"""Per-operator device-resident throughput against the HBM roofline (run on a B200 via gpurun) -- the development
companion of bench.py's per_operator record: the same rows (tools/oplib.py), plus the rest of the operator surface.

    python tools/opbench.py [--ops relvort,advection,...] [--all] [--mask 0.3] [--seconds 0.3] [--json out.json]

Every operator runs on a batch far larger than the 126 MB L2; time = CUDA events on the launching stream around deferred
launches; a clock sample (nvidia-smi, 100 ms) is taken under each operator's own load.
"""
import argparse
import importlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import torch
    import oplib
    from bench import ClockSampler, peaks
    ap = argparse.ArgumentParser()
    ap.add_argument("--ops", default="")
    ap.add_argument("--all", action="store_true", help="also the operators outside BASELINE.json's configs (SURVEY.md 8a rest, 8f)")
    ap.add_argument("--seconds", type=float, default=0.3)
    ap.add_argument("--cfg3-levels", type=int, default=32)
    ap.add_argument("--json", default="")
    ap.add_argument("--mask", type=float, default=0.0, help="fraction of undefined points in every maskable input (flag SOME_DEFINED)")
    args = ap.parse_args()
    gpu = importlib.import_module("mi-fieldcalc_b200").load()
    dev = torch.device("cuda", 0)
    peak, _ = peaks()
    stream = torch.cuda.current_stream()
    gpu.set_stream(stream.cuda_stream, True)
    inputs = oplib.Inputs(torch, dev)
    rows = oplib.rows(levels_cfg3=args.cfg3_levels) + (oplib.extra_rows() if args.all or args.ops else [])
    wanted = [o for o in args.ops.split(",") if o]
    sampler = ClockSampler(0)
    results = []
    print("%-32s %-5s %-10s %6s %5s %9s %9s %8s %6s %6s" % ("operator", "cfg", "grid", "fields", "mask", "ms", "Gpt/s", "GB/s", "frac", "MHz"))
    for row in rows:
        if wanted and row.name not in wanted:
            continue
        b = oplib.Built(row, inputs, args.mask)
        ms, launches, w0, w1 = oplib.time_row(gpu, torch, stream, b, args.seconds, sampler)
        gbs = row.bpp * b.points / (ms * 1e-3) / 1e9
        clk = sampler.window(w0, w1)
        res = {"operator": row.name, "config": row.cfg, "grid": list(row.grid), "fields": row.nf, "ms": ms, "gpts": b.points / (ms * 1e-3) / 1e9, "gbs": gbs,
               "frac": gbs / peak, "bytes_per_point": row.bpp, "mask": b.mask, "clocks": clk, "launches_timed": launches}
        results.append(res)
        print("%-32s %-5s %-10s %6d %5.2f %9.4f %9.2f %8.1f %6.3f %6s" % (row.name, row.cfg, "%dx%d" % row.grid, row.nf, b.mask, ms, res["gpts"], gbs, res["frac"],
                                                                         "%.0f" % clk["sm_mhz"] if clk["sm_mhz"] else "-"))
        sys.stdout.flush()
        del b
        torch.cuda.empty_cache()
    sampler.stop()
    if args.json:
        json.dump({"peak_gbs": peak, "results": results}, open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
