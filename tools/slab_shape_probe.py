"""Per-point throughput of the slab operators on ONE GPU for the shapes a rank of an 8-way row-slab run sees (3600 x ~229 rows x 137 levels)
against the whole-grid shape: isolates the shape penalty from the exchange and the flag combine.  python tools/slab_shape_probe.py"""
import importlib
import sys
import os

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
gpu = importlib.import_module("mi-fieldcalc_b200").load()
dev = torch.device("cuda:0")
stream = torch.cuda.current_stream()
gpu.set_stream(stream.cuda_stream, True)
UNDEF = 1.0e35


def run(name, nx, ny, nf, reps=10):
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    f = torch.randn((nf, ny, nx), device=dev, generator=g).add_(280.0)
    u = torch.randn((nf, ny, nx), device=dev, generator=g)
    v = torch.randn((nf, ny, nx), device=dev, generator=g)
    xm = torch.full((ny, nx), 6e-5, device=dev)
    ym = torch.full((ny, nx), 4.497e-5, device=dev)
    out = torch.empty_like(f)
    flags = np.zeros(nf, np.int32)

    def call():
        flags[:] = 0
        if name == "advection":
            gpu.call("advection_batched", nx, ny, nf, f, u, v, xm, ym, 1.0, out, flags, UNDEF)
        elif name == "tfp":
            gpu.call("thermalFrontParameter_batched", nx, ny, nf, f, xm, ym, out, flags, UNDEF)
        else:
            gpu.call("shapiro2_filter_batched", nx, ny, nf, f, out, flags, UNDEF)

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
    print("%-10s %5d x %5d x %4d  %8.4f ms  %7.1f Gpt/s" % (name, nx, ny, nf, ms, nf * nx * ny / ms / 1e6), flush=True)
    del f, u, v, out
    torch.cuda.empty_cache()


for name in ("advection", "tfp", "shapiro2"):
    run(name, 3600, 1801, 17)
    run(name, 3600, 229, 137)
    run(name, 3600, 232, 137)
    run(name, 3600, 227, 137)
