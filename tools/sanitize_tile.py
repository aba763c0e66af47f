#!/usr/bin/env python
"""A small pass over the kernels with hand-rolled synchronisation, meant to run under compute-sanitizer
(memcheck / racecheck / synccheck): the TMA-staged tile engine's producer / consumer mbarrier pipeline (relvort, advection,
jacobian, thermalFrontParameter, ilevelgwind: 1-3 staged arrays, 1-4 stages), the register-marching shapiro2 kernel
(warp shuffles) and the elementwise engine's shared-memory tables.  Small grids: the sanitizer slows kernels ~100x.

    compute-sanitizer --tool racecheck python tools/sanitize_tile.py
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    gpu = importlib.import_module("mi-fieldcalc_b200").load()
    dev = torch.device("cuda", 0)
    nx, ny, nf = 301, 41, 6
    g = torch.Generator(device=dev)
    g.manual_seed(1)

    def f(lo, hi, *shape, mask=0.0):
        a = torch.rand(shape, device=dev, generator=g) * (hi - lo) + lo
        if mask:
            a[torch.rand(shape, device=dev, generator=g) < mask] = 1e35
        return a

    xm, ym, fc = f(1.9e-4, 2.1e-4, ny, nx), f(1.9e-4, 2.1e-4, ny, nx), f(1.1e-4, 1.4e-4, ny, nx)
    n = 0
    for mask, flag in ((0.0, 0), (0.3, 2)):
        t, u, v = f(250, 300, nf, ny, nx, mask=mask), f(-30, 30, nf, ny, nx, mask=mask), f(-30, 30, nf, ny, nx, mask=mask)
        o, o2 = torch.empty_like(t), torch.empty_like(t)
        fl = lambda: np.full(nf, flag, np.int32)  # noqa: E731
        calls = [("relvort_batched", (nx, ny, nf, u, v, xm, ym, o, fl(), 1e35)), ("advection_batched", (nx, ny, nf, t, u, v, xm, ym, 1.0, o, fl(), 1e35)),
                 ("jacobian_batched", (nx, ny, nf, t, u, xm, ym, o, fl(), 1e35)), ("thermalFrontParameter_batched", (nx, ny, nf, t, xm, ym, o, fl(), 1e35)),
                 ("ilevelgwind_batched", (nx, ny, nf, t, xm, ym, fc, o, o2, fl(), 1e35)), ("gradient_batched", (nx, ny, nf, t, xm, ym, 3, o, fl(), 1e35)),
                 ("shapiro2_filter_batched", (nx, ny, nf, t, o, fl(), 1e35)), ("shapiro2_filter_batched", (300, ny, nf, t[:, :, :300].contiguous(), o[:, :, :300].contiguous(), fl(), 1e35)),
                 ("alevelhum_batched", (nx, ny, nf, t, f(1e-6, 2e-2, nf, ny, nx, mask=mask), f(300, 1040, nf, ny, nx, mask=mask), "celsius", 5, o, fl(), 1e35)),
                 ("meanValue_batched", (nx, ny, 2, [t[:2].contiguous(), u[:2].contiguous(), v[:2].contiguous()], 3, np.full(6, flag, np.int32), o[:2].contiguous(), np.zeros(2, np.int32), 1e35))]
        for name, args in calls:
            rc = gpu.call(name, *args)
            assert rc == 1, (name, rc, gpu.last_error())
            n += 1
    torch.cuda.synchronize()
    print("sanitize pass: %d calls, %d kernels launched" % (n, gpu.launch_count()))


if __name__ == "__main__":
    main()
