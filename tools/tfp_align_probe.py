import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
gpu = importlib.import_module("mi-fieldcalc_b200").load()
dev = torch.device("cuda:0"); stream = torch.cuda.current_stream(); gpu.set_stream(stream.cuda_stream, True)
def run(nx, ny, nf, reps=10):
    g = torch.Generator(device=dev); g.manual_seed(1)
    f = torch.randn((nf, ny, nx), device=dev, generator=g).add_(280.0)
    xm = torch.full((ny, nx), 6e-5, device=dev); ym = torch.full((ny, nx), 4.497e-5, device=dev)
    out = torch.empty_like(f); flags = np.zeros(nf, np.int32)
    def call():
        flags[:] = 0
        gpu.call("thermalFrontParameter_batched", nx, ny, nf, f, xm, ym, out, flags, 1e35)
    call(); gpu.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gpu.begin_deferred(); e0.record(stream)
    for _ in range(reps): call()
    e1.record(stream); gpu.end_deferred(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("tfp %5d x %5d x %4d  %8.4f ms  %7.1f Gpt/s" % (nx, ny, nf, ms, nf * nx * ny / ms / 1e6), flush=True)
for nx, ny in ((949, 1069), (948, 1069), (952, 1069), (950, 1069), (960, 1069), (962, 1068), (3600, 1801)):
    run(nx, ny, 96 if nx < 2000 else 16)
