"""GPU suite, multi-GPU part (SURVEY.md 8e case 2): the C++ row-slab path -- fcb200_slab_exchange over NCCL, the batched
operators on extended slabs, fcb200_slab_combine_flags -- against the single-GPU result, bit for bit, on two GPUs
(one process per GPU under torch.distributed.run).  Skipped on a one-GPU box; the partition logic and the exactness of
the slab decomposition itself are covered on the CPU (tests/test_abi.py, tests/test_distributed_cpu.py)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_slab_path_is_bit_identical_on_two_gpus(gpu, tmp_path):
    n = gpu.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    out = tmp_path / "slab.json"
    for nx, ny, levels in ((3600, 257, 5), (949, 131, 3)):  # 16-byte aligned rows (float4 halo kernels) and odd rows (scalar)
        r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29631",
                            os.path.join(ROOT, "tools", "multigpu_check.py"), "--nx", str(nx), "--ny", str(ny), "--levels", str(levels), "--steps", "2", "--json", str(out)],
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        recs = json.load(open(out))
        assert all(s["bit_identical_to_single_gpu"] for rec in recs for s in rec["steps"])
        assert all(s["nvlink_payload_bytes_per_step_all_ranks"] > 0 for rec in recs for s in rec["steps"])


def test_slab_calls_are_no_ops_on_one_rank(gpu):
    """nranks == 1: init needs no NCCL, exchange and combine leave everything untouched"""
    import numpy as np
    import torch
    gpu.slab_init(0, 1, b"\0" * 128)
    try:
        a = torch.arange(3 * 7 * 8, dtype=torch.float32, device="cuda").reshape(3, 7, 8).contiguous()
        b = a.clone()
        gpu.slab_exchange(a, 8, 7, 3, 2)
        flags = np.array([0, 1, 2], np.int32)
        gpu.slab_combine_flags(flags)
        gpu.synchronize()
        assert torch.equal(a, b) and flags.tolist() == [0, 1, 2]
        assert gpu.slab_partition(1801, 2, 0, 1) == (0, 1801, 0, 1801)
    finally:
        gpu.slab_finalize()
