#!/usr/bin/env python
"""Row-slab decomposition of one large grid over N GPUs with an NCCL halo exchange.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multigpu_check.py [--nx 3600 --ny 1801 --levels 8]

Every rank owns a band of rows of an ECMWF-sized batch, receives `halo` rows from rank-1 / rank+1 over
NCCL (mi-fieldcalc_b200/distributed.py: grouped isend/irecv), runs the ordinary single-GPU operator on
its extended slab and keeps the rows it owns.  Rank 0 also computes the whole grid alone; the script
checks that the assembled slabs equal it bit for bit (values, undefined mask) and that the combined
flag equals the single-GPU flag, then prints timings (max over ranks, CUDA events).
"""
import argparse
import importlib
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
UNDEF = 1.0e35


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nx", type=int, default=3600)
    ap.add_argument("--ny", type=int, default=1801)
    ap.add_argument("--levels", type=int, default=8)
    ap.add_argument("--reps", type=int, default=10)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    saved = os.dup(1)  # NCCL prints its banner on stdout when the communicator comes up: keep stdout for the JSON
    os.dup2(2, 1)
    dist.init_process_group("nccl", device_id=dev)
    dist.barrier()
    torch.cuda.synchronize()
    sys.stdout.flush()
    os.dup2(saved, 1)
    os.close(saved)
    pkg = importlib.import_module("mi-fieldcalc_b200")
    D = importlib.import_module("mi-fieldcalc_b200.distributed")
    gpu = pkg.load()
    gpu.set_device(local)
    stream = torch.cuda.current_stream()
    gpu.set_stream(stream.cuda_stream, True)
    nx, ny, nl = args.nx, args.ny, args.levels

    # the same full-grid inputs on every rank (seeded), 30 % of one variant masked
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    y, x = torch.meshgrid(torch.arange(ny, device=dev, dtype=torch.float32), torch.arange(nx, device=dev, dtype=torch.float32), indexing="ij")
    base = 280.0 + 15.0 * torch.sin(x / 97.0) * torch.cos(y / 61.0)
    f = (base[None] + torch.randn((nl, ny, nx), device=dev, generator=g)).contiguous()
    u = (20.0 * torch.cos(x / 131.0)[None] + torch.randn((nl, ny, nx), device=dev, generator=g)).contiguous()
    v = (15.0 * torch.sin(y / 89.0)[None] + torch.randn((nl, ny, nx), device=dev, generator=g)).contiguous()
    ym = torch.full((ny, nx), 4.497e-5, device=dev)
    xm = (ym / torch.clamp(torch.cos((y / (ny - 1) - 0.5) * 3.14159), min=0.01)).contiguous()
    f_masked = f.clone()
    f_masked[torch.rand((nl, ny, nx), device=dev, generator=g) < 0.3] = UNDEF

    def call(name, nrows, fields, flag_in, extra=()):
        """one batched single-GPU call on `nrows` rows; returns (out, flags)"""
        out = torch.empty((nl, nrows, nx), device=dev)
        flags = np.full(nl, flag_in, np.int32)
        if name == "advection":
            r = gpu.call("advection_batched", nx, nrows, nl, fields[0], fields[1], fields[2], fields[3], fields[4], 1.0, out, flags, UNDEF)
        elif name == "thermalFrontParameter":
            r = gpu.call("thermalFrontParameter_batched", nx, nrows, nl, fields[0], fields[1], fields[2], out, flags, UNDEF)
        else:
            r = gpu.call("shapiro2_filter_batched", nx, nrows, nl, fields[0], out, flags, UNDEF)
        assert r == 1, gpu.last_error()
        return out, flags

    results = []
    r0, r1 = D.partition_rows(ny, world)[rank]
    for name, per_field, shared, flag_in in (("advection", [f, u, v], [xm, ym], 0), ("advection", [f_masked, u, v], [xm, ym], 2),
                                             ("thermalFrontParameter", [f], [xm, ym], 0), ("thermalFrontParameter", [f_masked], [xm, ym], 2),
                                             ("shapiro2_filter", [f], [], 0), ("shapiro2_filter", [f_masked], [], 2)):
        halo = D.HALO[name]
        lo, hi = D.slab_bounds(r0, r1, ny, halo)

        # persistent extended slabs: the owned rows are in place (a previous sharded operator would have written
        # them there), the halo rows are refreshed from the neighbours before every call
        ext_fields = []
        for a in per_field:
            e = torch.zeros((nl, hi - lo, nx), device=dev)
            e[:, r0 - lo:r1 - lo, :] = a[:, r0:r1, :]
            ext_fields.append(e)
        ext_shared = [a[lo:hi].contiguous() for a in shared]  # grid-constant arrays: every rank holds its rows

        def slab_step():
            for e in ext_fields:
                D.exchange_halo_inplace(e, halo, rank, world)
            out, flags = call(name, hi - lo, ext_fields + ext_shared, flag_in)
            return out[:, r0 - lo:r1 - lo, :], flags

        owned, flags = slab_step()
        torch.cuda.synchronize()
        gflags = [D.combine_flags(int(fl), device=dev) for fl in flags]
        # reference: the whole grid on one GPU (every rank computes it; rank 0 compares)
        want, wflags = call(name, ny, per_field + shared, flag_in)
        torch.cuda.synchronize()
        gathered = [torch.empty((nl, b - a, nx), device=dev) for a, b in D.partition_rows(ny, world)]
        if world > 1:
            dist.all_gather(gathered, owned.contiguous()) if len({t.shape for t in gathered}) == 1 else None
        ok_values = bool(torch.equal(owned, want[:, r0:r1, :]) or torch.equal(torch.nan_to_num(owned), torch.nan_to_num(want[:, r0:r1, :])))
        ok_flags = list(map(int, wflags)) == gflags
        okt = torch.tensor([1 if (ok_values and ok_flags) else 0], device=dev)
        dist.all_reduce(okt, op=dist.ReduceOp.MIN)

        # timing: slab step (halo exchange + operator) vs the whole grid on one GPU
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(args.reps):
            slab_step()
        e1.record(stream)
        torch.cuda.synchronize()
        t_slab = torch.tensor([e0.elapsed_time(e1) / args.reps], device=dev)
        dist.all_reduce(t_slab, op=dist.ReduceOp.MAX)
        e0.record(stream)
        for _ in range(args.reps):
            call(name, ny, per_field + shared, flag_in)
        e1.record(stream)
        torch.cuda.synchronize()
        t_one = e0.elapsed_time(e1) / args.reps
        results.append({"operator": name, "flag_in": flag_in, "bit_identical_to_single_gpu": bool(okt.item()), "ms_slab_step_max_over_ranks": float(t_slab.item()),
                        "ms_whole_grid_one_gpu": t_one, "halo_rows": halo})
    if rank == 0:
        print(json.dumps({"world": world, "grid": [nx, ny], "levels": nl, "results": results}, indent=1))
    dist.destroy_process_group()
    assert all(r["bit_identical_to_single_gpu"] for r in results)


if __name__ == "__main__":
    main()
