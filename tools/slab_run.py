"""Row-slab run of BASELINE.json configs[2] over the ranks of a torch.distributed job: ONE ECMWF-sized grid (all levels in a
batch) cut into row slabs, halo rows moved by the library's C++ NCCL path (fcb200_slab_exchange, csrc/slab.cu), the ordinary
batched operators on the extended slabs, global flags by fcb200_slab_combine_flags.  Used by bench.py (the `slab` record)
and by tools/multigpu_check.py (the bit-identity check on its own).

Every rank holds the SAME full-grid inputs (same seed) so that it can (a) time the whole grid on one GPU -- the strong-scaling
denominator -- and (b) compare the rows it owns with the single-GPU result bit for bit, without gathering anything.

Steps (one "step" = what a caller does per time step, all enqueued on one stream in deferred mode, CUDA events around it):
  advection               halo 1: refresh f's halo rows from the neighbours (exchange), advection on the extended slab
  thermalFrontParameter   halo 2: exchange T, TFP on the extended slab
  shapiro2 -> TFP         halo 2: shapiro2_filter on the extended slab (its halo rows come out wrong by construction),
                          exchange of the SMOOTHED field's halo rows -- the exchange a chained stencil genuinely needs --, TFP
"""
from __future__ import annotations

import numpy as np

UNDEF = 1.0e35


def _mismatch(torch, a, b):
    """number of elements that differ bit for bit (NaN == NaN)"""
    return int(((a != b) & ~(torch.isnan(a) & torch.isnan(b))).sum().item())


def slab_record(gpu, torch, dist, dev, stream, rank, world, nx=3600, ny=1801, levels=137, steps=5, mask=0.0, seed=777):
    uid = [gpu.slab_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(uid, src=0)
    gpu.slab_init(rank, world, uid[0])
    try:
        return _run(gpu, torch, dist, dev, stream, rank, world, nx, ny, levels, steps, mask, seed)
    finally:
        gpu.slab_finalize()


def _run(gpu, torch, dist, dev, stream, rank, world, nx, ny, nf, steps, mask, seed):
    g = torch.Generator(device=dev)
    g.manual_seed(seed)  # the same on every rank
    y = torch.arange(ny, device=dev, dtype=torch.float32)[:, None]
    x = torch.arange(nx, device=dev, dtype=torch.float32)[None, :]
    base = 280.0 + 15.0 * torch.sin(x / 97.0) * torch.cos(y / 61.0)
    f = torch.randn((nf, ny, nx), device=dev, generator=g).add_(base)
    u = torch.randn((nf, ny, nx), device=dev, generator=g).add_(20.0 * torch.cos(x / 131.0))
    v = torch.randn((nf, ny, nx), device=dev, generator=g).add_(15.0 * torch.sin(y / 89.0))
    del base
    if mask > 0:
        f[torch.rand((nf, ny, nx), device=dev, generator=g) < mask] = UNDEF
    elif mask < 0:  # flags that differ between the ranks: a few undefined points in the LAST rank's rows only, one field undefined everywhere
        f[:, ny - 4:ny - 2, 5:25] = UNDEF
        f[nf - 1] = UNDEF
    ym = torch.full((ny, nx), 4.497e-5, device=dev)
    xm = (ym / torch.clamp(torch.cos((y / (ny - 1) - 0.5) * 3.14159), min=0.01)).expand(ny, nx).contiguous()
    flag_in = 2 if mask != 0 else 0

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, reps, reduce_flags=False):
        """ms per call of fn (deferred mode, events on the launching stream), max over ranks; reduce_flags: the step ends with the
        on-stream combination of its flags over the ranks (fcb200_slab_reduce_flags), inside the timed region"""
        gpu.begin_deferred()  # warm-up: arena growth, NCCL connection set-up (of the flag all-reduce too)
        fn()
        if reduce_flags:
            gpu.slab_reduce_flags()
        gpu.end_deferred()
        gpu.synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        for _ in range(reps):  # one drain per step: the unfused TFP's scratch field is recycled
            gpu.begin_deferred()
            e0.record(stream)
            fn()
            if reduce_flags:
                gpu.slab_reduce_flags()
            e1.record(stream)
            gpu.end_deferred()
            torch.cuda.synchronize()
            total += e0.elapsed_time(e1)
        t = torch.tensor([total / reps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def ext_of(a, lo, hi, r0, r1, zero_halo):
        e = a[:, lo:hi, :].contiguous()
        if zero_halo:  # only the owned rows are in place: the exchange has to deliver the rest
            e[:, : r0 - lo, :] = 0
            e[:, r1 - lo:, :] = 0
        return e

    out_full = torch.empty((nf, ny, nx), device=dev)
    tmp_full = torch.empty((nf, ny, nx), device=dev)
    results = []

    def record(name, halo, full_fn, slab_fn, out_ext, part, exchanges_per_step, exchange_fn=None, compute_fn=None):
        r0, r1, lo, hi = part
        flags_full = np.full(nf, flag_in, np.int32)
        flags_slab = np.full(nf, flag_in, np.int32)
        t1 = timed(lambda: full_fn(flags_full), max(2, steps // 2))
        flags_full[:] = flag_in
        full_fn(flags_full)
        gpu.synchronize()
        sent0 = gpu.slab_bytes_sent()
        tn = timed(lambda: slab_fn(flags_slab), steps, reduce_flags=True)
        sent = (gpu.slab_bytes_sent() - sent0) / (steps + 1)
        flags_global = flags_slab.copy()  # what the last timed step left: the GLOBAL flags (combined on the stream)
        # for comparison: the host-side combination (fcb200_slab_combine_flags) of the local flags, as wall time
        flags_slab[:] = flag_in
        slab_fn(flags_slab)
        gpu.synchronize()
        import time
        warm = flags_slab.copy()
        gpu.slab_combine_flags(warm)  # the first all-reduce of a communicator sets its connections up: not part of a step
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        gpu.slab_combine_flags(flags_slab)
        combine_us = (time.perf_counter() - t0) * 1e6
        bad = (_mismatch(torch, out_ext[:, r0 - lo:r1 - lo, :], out_full[:, r0:r1, :]) + int((flags_slab != flags_full).sum())
               + int((flags_global != flags_full).sum()))
        stats = torch.tensor([bad, sent, combine_us], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(stats, op=dist.ReduceOp.SUM)
        bad_all, sent_all, combine_sum = stats.tolist()
        step_ms = tn  # (flags included: combined on the stream inside the timed region)
        # where the step goes: the halo exchange alone and the operators on the extended slab alone (each max over ranks)
        parts = {}
        if exchange_fn is not None:
            parts["ms_exchange_only"] = timed(exchange_fn, steps)
        if compute_fn is not None:
            parts["ms_operators_only"] = timed(lambda: compute_fn(flags_slab), steps)
        parts["ms_ideal"] = t1 / world
        results.append({"breakdown": parts,"step": name, "halo_rows": halo, "ms_whole_grid_one_gpu": t1, "ms_slab_step": tn, "host_flag_combine_us_not_in_step": combine_sum / world,
                        "ms_slab_step_with_flags": step_ms, "speedup": t1 / step_ms, "efficiency_vs_one_gpu": t1 / (world * step_ms),
                        "nvlink_payload_bytes_per_step_all_ranks": sent_all, "exchanges_per_step": exchanges_per_step,
                        "bit_identical_to_single_gpu": bad_all == 0, "mismatches": int(bad_all),
                        "gpts": nf * nx * ny / (step_ms * 1e-3) / 1e9})

    # ---- advection, halo 1
    p1 = gpu.slab_partition(ny, 1, rank, world)
    assert p1 is not None, "grid too small for %d slabs" % world
    r0, r1, lo, hi = p1
    ef, eu, ev = ext_of(f, lo, hi, r0, r1, True), ext_of(u, lo, hi, r0, r1, False), ext_of(v, lo, hi, r0, r1, False)
    exm, eym = xm[lo:hi].contiguous(), ym[lo:hi].contiguous()
    eo = torch.empty_like(ef)

    def adv_full(flags):
        flags[:] = flag_in
        gpu.call("advection_batched", nx, ny, nf, f, u, v, xm, ym, 1.0, out_full, flags, UNDEF)

    def adv_slab(flags):
        flags[:] = flag_in
        gpu.slab_exchange(ef, nx, hi - lo, nf, 1)
        gpu.call("advection_batched", nx, hi - lo, nf, ef, eu, ev, exm, eym, 1.0, eo, flags, UNDEF)

    def adv_compute(flags):
        flags[:] = flag_in
        gpu.call("advection_batched", nx, hi - lo, nf, ef, eu, ev, exm, eym, 1.0, eo, flags, UNDEF)

    record("advection", 1, adv_full, adv_slab, eo, p1, 1, lambda: gpu.slab_exchange(ef, nx, hi - lo, nf, 1), adv_compute)
    del ef, eu, ev, eo, u, v
    torch.cuda.empty_cache()

    # ---- thermalFrontParameter, halo 2
    p2 = gpu.slab_partition(ny, 2, rank, world)
    assert p2 is not None
    r0, r1, lo, hi = p2
    et = ext_of(f, lo, hi, r0, r1, True)
    exm, eym = xm[lo:hi].contiguous(), ym[lo:hi].contiguous()
    eo = torch.empty_like(et)

    def tfp_full(flags):
        flags[:] = flag_in
        gpu.call("thermalFrontParameter_batched", nx, ny, nf, f, xm, ym, out_full, flags, UNDEF)

    def tfp_slab(flags):
        flags[:] = flag_in
        gpu.slab_exchange(et, nx, hi - lo, nf, 2)
        gpu.call("thermalFrontParameter_batched", nx, hi - lo, nf, et, exm, eym, eo, flags, UNDEF)

    def tfp_compute(flags):
        flags[:] = flag_in
        gpu.call("thermalFrontParameter_batched", nx, hi - lo, nf, et, exm, eym, eo, flags, UNDEF)

    record("thermalFrontParameter", 2, tfp_full, tfp_slab, eo, p2, 1, lambda: gpu.slab_exchange(et, nx, hi - lo, nf, 2), tfp_compute)

    # ---- shapiro2_filter -> thermalFrontParameter: the smoothed field's halo rows have to be exchanged
    et = ext_of(f, lo, hi, r0, r1, False)  # static input scattered with its halo: no exchange needed for it
    es = torch.empty_like(et)
    sflags = np.zeros(nf, np.int32)

    def chain_full(flags):
        sflags[:] = flag_in
        gpu.call("shapiro2_filter_batched", nx, ny, nf, f, tmp_full, sflags, UNDEF)
        flags[:] = 0  # shapiro2_filter always returns ALL_DEFINED (FC.cc:2176): what the caller passes on
        gpu.call("thermalFrontParameter_batched", nx, ny, nf, tmp_full, xm, ym, out_full, flags, UNDEF)

    def chain_slab(flags):
        sflags[:] = flag_in
        gpu.call("shapiro2_filter_batched", nx, hi - lo, nf, et, es, sflags, UNDEF)
        gpu.slab_exchange(es, nx, hi - lo, nf, 2)
        flags[:] = 0
        gpu.call("thermalFrontParameter_batched", nx, hi - lo, nf, es, exm, eym, eo, flags, UNDEF)

    def chain_compute(flags):
        sflags[:] = flag_in
        gpu.call("shapiro2_filter_batched", nx, hi - lo, nf, et, es, sflags, UNDEF)
        flags[:] = 0
        gpu.call("thermalFrontParameter_batched", nx, hi - lo, nf, es, exm, eym, eo, flags, UNDEF)

    record("shapiro2_filter->thermalFrontParameter", 2, chain_full, chain_slab, eo, p2, 1, lambda: gpu.slab_exchange(es, nx, hi - lo, nf, 2), chain_compute)
    rows = [gpu.slab_partition(ny, 2, r, world) for r in range(world)]
    return {"grid": [nx, ny], "levels": nf, "mask": mask, "ranks": world, "rows_per_rank": [p[1] - p[0] for p in rows],
            "transport": "ncclSend/ncclRecv from C++ (fcb200_slab_exchange): pack kernel -> one grouped send/recv pair per neighbour -> unpack kernel, "
                         "on the operators' stream; flags by ncclAllReduce on the same stream (fcb200_slab_reduce_flags)",
            "timing": "CUDA events on the launching stream around one deferred step INCLUDING the on-stream flag combination (fcb200_slab_reduce_flags: one kernel + one ncclAllReduce), max over ranks; the host-side combination is timed beside it for comparison",
            "steps": results}
