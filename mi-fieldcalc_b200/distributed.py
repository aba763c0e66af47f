"""Multi-GPU partitioning of the FieldCalculations hot path: one process per GPU,
torch.distributed for the plumbing (NCCL over NVLink on the GPUs, gloo in the CPU tests).

1. Batched fields (levels x times x members) shard naturally: every field is independent, so each
   rank takes a contiguous range of the field index -- `shard_range` -- and no collective touches
   the data path.  The flags are per field.

2. One very large grid is cut into ROW SLABS.  A five-point stencil needs `halo` rows of every
   neighbour-read input above and below the rows a rank owns (1 row for relvort / absvort /
   divergence / advection / gradient c2-c4 / jacobian / ilevelgwind, 2 rows for
   thermalFrontParameter and shapiro2_filter, which look two rows away).  The rank then calls the
   ordinary single-GPU operator on its extended slab and keeps the rows it owns:
     * the flat-index wrap of the reference (i-1 of column 0 is the last element of the previous
       row) stays inside slab + halo, so values, masks and counted points are identical;
     * global rows 0 and ny-1 belong to the first / last rank, whose slab has no halo on that side,
       so the operator's own fillEdges clamp produces exactly the global border;
     * the operator's flag on a slab is ALL / NONE / SOME of the rows it counted; the global flag is
       ALL iff every rank says ALL, NONE iff every rank says NONE, else SOME -- one tiny all-reduce.
   `exchange_halo` moves the halo rows with grouped isend/irecv to rank-1 / rank+1 (h*nx floats per
   field, latency bound: 14.4 KB per row at nx = 3600).  Static inputs can skip the exchange by
   scattering overlapped slabs (`slab_bounds`); the exchange is needed when a stencil consumes the
   OUTPUT of a previous sharded operator (shapiro2_filter -> thermalFrontParameter ...).

Not exact in slab mode, by construction of the reference: gradient compute=1 (its flag counts the
flat range [1, N-1), including rows 0 and ny-1, against N-2nx -- use the single-GPU call), and the
pathological thermalFrontParameter case where the first pass produces NaN/inf from defined inputs.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple

ALL_DEFINED, NONE_DEFINED, SOME_DEFINED = 0, 1, 2

# halo rows needed by each slab-capable operator
HALO = {"relvort": 1, "absvort": 1, "divergence": 1, "advection": 1, "gradient": 1, "jacobian": 1, "ilevelgwind": 1,
        "thermalFrontParameter": 2, "shapiro2_filter": 2}


def shard_range(nitems: int, rank: int, world: int) -> Tuple[int, int]:
    """contiguous [begin, end) of `nitems` independent fields for `rank`; sizes differ by at most one"""
    base, extra = divmod(nitems, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def partition_rows(ny: int, world: int) -> List[Tuple[int, int]]:
    """owned row ranges [r0, r1) per rank"""
    return [shard_range(ny, r, world) for r in range(world)]


def slab_bounds(r0: int, r1: int, ny: int, halo: int) -> Tuple[int, int]:
    """rows [lo, hi) a rank must hold to compute its owned rows [r0, r1)"""
    return max(0, r0 - halo), min(ny, r1 + halo)


def min_rows_ok(ny: int, world: int, halo: int) -> bool:
    """every extended slab must be a valid grid for the reference (ny >= 3) and every halo must come from
    the direct neighbour"""
    return all((r1 - r0) >= max(halo, 1) and (slab_bounds(r0, r1, ny, halo)[1] - slab_bounds(r0, r1, ny, halo)[0]) >= 3
               for r0, r1 in partition_rows(ny, world))


def exchange_halo(owned, halo: int, rank: int, world: int, group=None):
    """owned: tensor [..., rows, nx] of the rows this rank owns (any leading batch dims).
    Returns the extended slab [..., halo_up + rows + halo_down, nx] with the neighbours' boundary rows
    (no halo on the outer side of the first / last rank).  One grouped isend/irecv per direction."""
    import torch
    import torch.distributed as dist

    rows = owned.shape[-2]
    assert rows >= halo, "a slab must own at least `halo` rows"
    up = torch.empty(owned.shape[:-2] + (halo, owned.shape[-1]), dtype=owned.dtype, device=owned.device) if rank > 0 else None
    down = torch.empty(owned.shape[:-2] + (halo, owned.shape[-1]), dtype=owned.dtype, device=owned.device) if rank < world - 1 else None
    ops = []
    if rank > 0:
        ops.append(dist.P2POp(dist.isend, owned[..., :halo, :].contiguous(), rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, up, rank - 1, group))
    if rank < world - 1:
        ops.append(dist.P2POp(dist.isend, owned[..., rows - halo:, :].contiguous(), rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, down, rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    parts = [p for p in (up, owned, down) if p is not None]
    return torch.cat(parts, dim=-2).contiguous()


def exchange_halo_inplace(ext, halo: int, rank: int, world: int, group=None):
    """ext: the EXTENDED slab [..., halo_up + rows + halo_down, nx] (halo_up = 0 on rank 0, halo_down = 0 on the
    last rank) whose owned rows are already in place -- typically written there by the previous sharded
    operator.  Only the halo rows move: each rank sends its first / last `halo` owned rows to rank-1 / rank+1
    and receives the neighbours' straight into its halo rows.  No concatenation, no extra pass over the slab."""
    import torch.distributed as dist

    up = halo if rank > 0 else 0
    down = halo if rank < world - 1 else 0
    total = ext.shape[-2]
    rows = total - up - down
    assert rows >= halo, "a slab must own at least `halo` rows"
    ops, stage = [], []
    if rank > 0:
        send = ext[..., up:up + halo, :].contiguous()
        recv = ext[..., :up, :] if ext[..., :up, :].is_contiguous() else None
        buf = recv if recv is not None else ext.new_empty(ext.shape[:-2] + (halo, ext.shape[-1]))
        ops += [dist.P2POp(dist.isend, send, rank - 1, group), dist.P2POp(dist.irecv, buf, rank - 1, group)]
        stage.append((buf, recv is None, slice(0, up)))
    if rank < world - 1:
        send = ext[..., up + rows - halo:up + rows, :].contiguous()
        recv = ext[..., up + rows:, :] if ext[..., up + rows:, :].is_contiguous() else None
        buf = recv if recv is not None else ext.new_empty(ext.shape[:-2] + (halo, ext.shape[-1]))
        ops += [dist.P2POp(dist.isend, send, rank + 1, group), dist.P2POp(dist.irecv, buf, rank + 1, group)]
        stage.append((buf, recv is None, slice(up + rows, total)))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    for buf, copy_needed, where in stage:
        if copy_needed:  # batched slabs: the halo rows of every field are strided, received via one small buffer
            ext[..., where, :].copy_(buf)
    return ext


def combine_flags(local_flag: int, group=None, device=None) -> int:
    """global ValuesDefined of a slab-sharded operator from the per-rank flags (one 8-byte all-reduce)"""
    import torch
    import torch.distributed as dist

    t = torch.tensor([1 if local_flag == ALL_DEFINED else 0, 1 if local_flag == NONE_DEFINED else 0], dtype=torch.int32, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    if int(t[0]) == 1:
        return ALL_DEFINED
    if int(t[1]) == 1:
        return NONE_DEFINED
    return SOME_DEFINED


def run_slab(op: Callable, name: str, ny: int, rank: int, world: int, slab_inputs: Sequence, flag_in: int, group=None, device=None):
    """Run one slab-sharded stencil.

    op(nx_rows, *slab_inputs, flag_in) -> (ret, [outputs over the SAME rows as the slab], flag) is the
    single-device operator: the CUDA library on the GPUs, the oracle in the CPU tests.  `slab_inputs`
    already hold the rows slab_bounds(...) (scattered with overlap, or assembled by exchange_halo).
    Returns (ret, [outputs restricted to the owned rows], global flag)."""
    halo = HALO[name]
    r0, r1 = partition_rows(ny, world)[rank]
    lo, hi = slab_bounds(r0, r1, ny, halo)
    ret, outs, flag = op(hi - lo, *slab_inputs, flag_in)
    owned = [o[..., r0 - lo:r1 - lo, :] for o in outs]
    return ret, owned, combine_flags(flag, group, device)
