"""CPU tests of the N > 1 paths (world_size 2 and 3, gloo): batch sharding and the row-slab
decomposition with halo exchange.  The per-slab compute is injected; here it is the oracle (the
product has no CPU path), so the test proves that slab partitioning + halo exchange + owner-rank
border handling + flag combination reproduce the single-grid result exactly."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import matrix  # noqa: E402

D = importlib.import_module("mi-fieldcalc_b200.distributed")


def test_shard_range_covers_everything():
    for n in (1, 7, 65, 4355):
        for world in (1, 2, 3, 8):
            spans = [D.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


SLAB_OPS = [("relvort", {}), ("absvort", {}), ("divergence", {}), ("advection", {}), ("gradient", dict(compute=2)), ("gradient", dict(compute=3)),
            ("gradient", dict(compute=4)), ("jacobian", {}), ("ilevelgwind", {}), ("thermalFrontParameter", {}), ("shapiro2_filter", {})]


def _worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fclibs
    oracle = fclibs.oracle()
    nx, ny = 37, 41
    failures = []
    for name, params in SLAB_OPS:
        halo = D.HALO[name]
        for mask, flag_in in (("none", cases.ALL), ("bernoulli", cases.SOME), ("edge", cases.SOME), ("all", cases.SOME)):
            case = cases.build(name, nx, ny, seed=99, flag_in=flag_in, mask=mask, **params)
            want = cases.run(oracle, case)  # the single-grid answer (every rank computes it: tiny)
            r0, r1 = D.partition_rows(ny, world)[rank]
            lo, hi = D.slab_bounds(r0, r1, ny, halo)
            field_idx = [k for k, a in enumerate(case.args) if isinstance(a, np.ndarray) and a.ndim == 2 and k not in case.out_idx]

            # every rank starts from the rows it OWNS and gets its halo from the neighbours
            slabs = {}
            for k in field_idx:
                owned = torch.from_numpy(np.ascontiguousarray(case.args[k][r0:r1]))
                ext = D.exchange_halo(owned, halo, rank, world)
                assert ext.shape[0] == hi - lo
                assert np.array_equal(ext.numpy(), case.args[k][lo:hi]), "halo exchange delivered wrong rows"
                slabs[k] = ext.numpy()

            def op(rows, *inputs_and_flag):
                args = list(case.args)
                for k, a in zip(field_idx, inputs_and_flag[:-1]):
                    args[k] = np.ascontiguousarray(a)
                args[1] = rows  # ny of the slab
                for k in case.out_idx:
                    args[k] = np.full((rows, nx), cases.SENTINEL, np.float32)
                args[case.flag_idx] = np.array([inputs_and_flag[-1]], np.int32)
                ret = oracle.call(name, *args)
                return ret, [args[k] for k in case.out_idx], int(args[case.flag_idx][0])

            ret, owned_outs, flag = D.run_slab(op, name, ny, rank, world, [slabs[k] for k in field_idx], flag_in)
            if ret != want[0] or flag != want[2]:
                failures.append("%s %s %s: ret/flag %r/%r != %r/%r" % (name, params, mask, ret, flag, want[0], want[2]))
            for o, w in zip(owned_outs, want[1]):
                w = w[r0:r1]
                same = (o.view(np.uint32) == w.view(np.uint32)) | (np.isnan(o) & np.isnan(w))
                if not same.all():
                    failures.append("%s %s %s: %d values differ on rank %d" % (name, params, mask, int((~same).sum()), rank))
    results[rank] = failures
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_row_slab_decomposition_matches_single_grid(world):
    port = 29500 + world + (os.getpid() % 200)
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_worker, args=(world, port, results), nprocs=world, join=True)
        for r in range(world):
            assert results[r] == [], "\n".join(results[r])


def _batch_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fclibs
    oracle = fclibs.oracle()
    nf, nx, ny = 7, 19, 11
    rng = np.random.default_rng(5)
    t = np.stack([cases.field(rng, "tk", nx, ny) for _ in range(nf)])
    p = np.stack([cases.field(rng, "p", nx, ny) for _ in range(nf)])
    cases.apply_mask(rng, t[3], "all", cases.UNDEF)
    b, e = D.shard_range(nf, rank, world)
    flags = torch.full((nf,), -1, dtype=torch.int32)
    outs = torch.zeros((nf, ny, nx))
    for k in range(b, e):
        o = np.empty((ny, nx), np.float32)
        f = np.array([cases.SOME], np.int32)
        oracle.call("aleveltemp", nx, ny, t[k], p[k], "kelvin", 3, o, f, float(cases.UNDEF))
        flags[k] = int(f[0])
        outs[k] = torch.from_numpy(o)
    # gather for the check only: the data path itself needs no collective
    dist.all_reduce(flags, op=dist.ReduceOp.MAX)
    results[rank] = flags.tolist()
    dist.destroy_process_group()


def test_batch_sharding_needs_no_collective():
    world = 2
    port = 29800 + (os.getpid() % 200)
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_batch_worker, args=(world, port, results), nprocs=world, join=True)
        assert results[0] == results[1] == [0, 0, 0, 1, 0, 0, 0]


def _inplace_worker(rank, world, port, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    for (nl, ny, nx, h) in ((1, 11, 5, 1), (3, 11, 5, 2), (4, 23, 7, 2)):
        full = torch.arange(nl * ny * nx, dtype=torch.float32).reshape(nl, ny, nx)
        r0, r1 = D.partition_rows(ny, world)[rank]
        lo, hi = D.slab_bounds(r0, r1, ny, h)
        ext = torch.full((nl, hi - lo, nx), -1.0)
        ext[:, r0 - lo:r1 - lo] = full[:, r0:r1]
        D.exchange_halo_inplace(ext, h, rank, world)
        ok = ok and bool(torch.equal(ext, full[:, lo:hi]))
        if nl == 1:  # a single field: the halo rows are contiguous and received in place
            e2 = torch.full((hi - lo, nx), -1.0)
            e2[r0 - lo:r1 - lo] = full[0, r0:r1]
            D.exchange_halo_inplace(e2, h, rank, world)
            ok = ok and bool(torch.equal(e2, full[0, lo:hi]))
    results[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_in_place(world):
    """only the halo rows move; the owned rows stay where the previous operator wrote them"""
    port = 29300 + world + (os.getpid() % 200)
    with mp.Manager() as mgr:
        results = mgr.dict()
        mp.spawn(_inplace_worker, args=(world, port, results), nprocs=world, join=True)
        assert all(results[r] for r in range(world))
