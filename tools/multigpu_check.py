#!/usr/bin/env python
"""Row-slab decomposition of one large grid over N GPUs through the library's C++ NCCL path.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_check.py [--nx 3600 --ny 1801 --levels 8] [--json out.json]

Runs tools/slab_run.py twice -- all defined, and 30 % of the field undefined with SOME_DEFINED flags -- and fails unless every
rank's owned rows and the combined flags equal the single-GPU result bit for bit.
"""
import argparse
import importlib
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def main():
    import slab_run
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=3600)
    ap.add_argument("--ny", type=int, default=1801)
    ap.add_argument("--levels", type=int, default=8)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--json", default="")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved = os.dup(1)  # NCCL prints its banner on stdout when the communicator comes up: keep stdout for the JSON
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    torch.cuda.synchronize()
    gpu = importlib.import_module("mi-fieldcalc_b200").load()
    gpu.set_device(local)
    stream = torch.cuda.current_stream()
    gpu.set_stream(stream.cuda_stream, True)
    out = [slab_run.slab_record(gpu, torch, dist, dev, stream, rank, world, nx=args.nx, ny=args.ny, levels=args.levels, steps=args.steps, mask=m) for m in (0.0, 0.3, -1.0)]
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)
    if rank == 0:
        text = json.dumps(out, indent=1)
        print(text)
        if args.json:
            open(args.json, "w").write(text)
    dist.destroy_process_group()
    assert all(s["bit_identical_to_single_gpu"] for rec in out for s in rec["steps"]), "slab result differs from the single-GPU result"


if __name__ == "__main__":
    main()
