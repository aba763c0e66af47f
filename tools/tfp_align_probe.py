"""thermalFrontParameter on grids that differ only in the parity of the row length: python tools/tfp_align_probe.py [nx ...]
(an odd nx -- MEPS, 949 -- costs the tile kernel ~20 %: 238 against 291-296 Gpt/s for 948 / 950 / 952)"""
import importlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gpu = importlib.import_module("mi-fieldcalc_b200").load()
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream()
gpu.set_stream(stream.cuda_stream, True)


def run(nx, ny, nf, reps=10):
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    f = torch.randn((nf, ny, nx), device=dev, generator=g).add_(280.0)
    xm = torch.full((ny, nx), 6e-5, device=dev)
    ym = torch.full((ny, nx), 4.497e-5, device=dev)
    out = torch.empty_like(f)
    flags = np.zeros(nf, np.int32)

    def call():
        flags[:] = 0
        gpu.call("thermalFrontParameter_batched", nx, ny, nf, f, xm, ym, out, flags, 1e35)

    call()
    gpu.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gpu.begin_deferred()
    e0.record(stream)
    for _ in range(reps):
        call()
    e1.record(stream)
    gpu.end_deferred()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("tfp %5d x %5d x %4d  %8.4f ms  %7.1f Gpt/s" % (nx, ny, nf, ms, nf * nx * ny / ms / 1e6), flush=True)


sizes = [int(a) for a in sys.argv[1:]] or [949, 948, 952, 950, 960, 3600]
for nx in sizes:
    run(nx, 1801 if nx > 2000 else 1069, 16 if nx > 2000 else 96, reps=2 if len(sys.argv) > 1 else 10)
