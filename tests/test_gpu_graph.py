"""GPU suite, graph API (include/fcb200.h fcb200_graph_*; SURVEY.md 8f "fused chains / graph API"): chains of calls on
device-resident fields captured once into a CUDA graph and replayed with one launch.  The chains are the ones the
reference's callers run call by call -- cfg1: pleveltemp -> relvort -> divergence (FC.cc:400, 1875, 1942); the stencil chain
theta -> shapiro2_filter -> thermalFrontParameter (FC.cc:1355, 2181, 2351).  Every replay is checked bit for bit (flags
included) against the same calls made one by one on the oracle with the data the buffers hold at that launch."""
import numpy as np
import pytest

import cases

pytestmark = pytest.mark.gpu

UNDEF = float(cases.UNDEF)


def _arbiter():
    import fclibs
    return fclibs.reference() or fclibs.oracle()


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _inputs(seed, nx, ny, masked):
    rng = np.random.default_rng(seed)
    d = {k: cases.field(rng, "wind" if k in ("u", "v") else k, nx, ny) for k in ("tk", "u", "v", "xm", "ym", "p")}
    if masked:
        for k in ("tk", "u", "v"):
            cases.apply_mask(rng, d[k], "bernoulli", cases.UNDEF)
    return d


def _cfg1_reference(arb, nx, ny, d, flag_in):
    out = {}
    for name, args in (("pleveltemp", (d["tk"], 850.0, "kelvin", 3)), ("relvort", (d["u"], d["v"], d["xm"], d["ym"])),
                       ("divergence", (d["u"], d["v"], d["xm"], d["ym"]))):
        o, f = np.empty((ny, nx), np.float32), np.array([flag_in], np.int32)
        assert arb.call(name, nx, ny, *args, o, f, UNDEF) == 1
        out[name] = (o, int(f[0]))
    return out


@pytest.mark.parametrize("masked", [False, True])
def test_cfg1_chain_replayed(gpu, masked):
    """capture pleveltemp -> relvort -> divergence on one MEPS field, then launch it three times with NEW data in the same buffers"""
    import torch
    arb = _arbiter()
    nx, ny = 949, 1069 if not masked else 211
    flag_in = cases.SOME if masked else cases.ALL
    d0 = _inputs(11, nx, ny, masked)
    bufs = {k: _dev(d0[k]) for k in ("tk", "u", "v", "xm", "ym")}
    outs = {k: torch.empty((ny, nx), dtype=torch.float32, device="cuda") for k in ("pleveltemp", "relvort", "divergence")}
    flags = {k: np.array([flag_in], np.int32) for k in outs}
    gpu.graph_begin()
    assert gpu.call("pleveltemp", nx, ny, bufs["tk"], 850.0, "kelvin", 3, outs["pleveltemp"], flags["pleveltemp"], UNDEF) == 1
    assert gpu.call("relvort", nx, ny, bufs["u"], bufs["v"], bufs["xm"], bufs["ym"], outs["relvort"], flags["relvort"], UNDEF) == 1
    assert gpu.call("divergence", nx, ny, bufs["u"], bufs["v"], bufs["xm"], bufs["ym"], outs["divergence"], flags["divergence"], UNDEF) == 1
    g = gpu.graph_end()
    try:
        assert gpu.graph_kernels(g) >= 3
        for rep in range(3):
            d = _inputs(11 + rep, nx, ny, masked)
            for k in bufs:
                bufs[k].copy_(torch.from_numpy(d[k]))
            torch.cuda.synchronize()
            for k in outs:
                outs[k].fill_(-7.0)
                flags[k][0] = flag_in  # (an operator that has nothing to test leaves the flag as it is, like the reference)
            torch.cuda.synchronize()
            before = gpu.launch_count()
            gpu.graph_launch(g)
            assert gpu.launch_count() - before == gpu.graph_kernels(g)
            want = _cfg1_reference(arb, nx, ny, d, flag_in)
            for k in outs:
                o, f = want[k]
                got = outs[k].cpu().numpy()
                if k == "pleveltemp":  # powf on the host differs in the last place between libms: north_star's 1e-5
                    assert np.array_equal(got == cases.UNDEF, o == cases.UNDEF)
                    assert np.allclose(got, o, rtol=1e-5, atol=0.0), (k, rep)
                else:
                    assert np.array_equal(got, o), (k, rep)
                assert int(flags[k][0]) == f, (k, rep, int(flags[k][0]), f)
    finally:
        gpu.graph_destroy(g)


@pytest.mark.parametrize("masked", [False, True])
def test_stencil_chain_batched_replayed(gpu, masked):
    """aleveltemp (theta) -> shapiro2_filter -> thermalFrontParameter on a batch of levels, the intermediate fields staying on the device"""
    import torch
    arb = _arbiter()
    nx, ny, nf = 300, 181, 5
    flag_in = cases.SOME if masked else cases.ALL

    def make(seed):
        rng = np.random.default_rng(seed)
        t = np.stack([cases.field(rng, "tk", nx, ny) for _ in range(nf)])
        p = np.stack([cases.field(rng, "p", nx, ny) for _ in range(nf)])
        if masked:
            for k in range(nf):
                cases.apply_mask(rng, t[k], "sparse", cases.UNDEF)
        return t, p

    rng = np.random.default_rng(3)
    xm, ym = cases.field(rng, "xm", nx, ny), cases.field(rng, "ym", nx, ny)
    t0, p0 = make(100)
    dt, dp, dxm, dym = _dev(t0), _dev(p0), _dev(xm), _dev(ym)
    theta, smooth, tfp = (torch.empty((nf, ny, nx), dtype=torch.float32, device="cuda") for _ in range(3))
    f_theta, f_smooth, f_tfp = (np.full(nf, flag_in, np.int32) for _ in range(3))
    gpu.graph_begin()
    assert gpu.call("aleveltemp_batched", nx, ny, nf, dt, dp, "kelvin", 3, theta, f_theta, UNDEF) == 1
    assert gpu.call("shapiro2_filter_batched", nx, ny, nf, theta, smooth, f_smooth, UNDEF) == 1
    assert gpu.call("thermalFrontParameter_batched", nx, ny, nf, smooth, dxm, dym, tfp, f_tfp, UNDEF) == 1
    g = gpu.graph_end()
    try:
        for rep in range(2):
            t, p = make(100 + rep)
            dt.copy_(torch.from_numpy(t))
            dp.copy_(torch.from_numpy(p))
            for f in (f_theta, f_smooth, f_tfp):
                f[:] = flag_in
            torch.cuda.synchronize()
            gpu.graph_launch(g)
            got_theta, got_tfp = theta.cpu().numpy(), tfp.cpu().numpy()
            for k in range(nf):
                # the oracle's chain starts from the product's own theta (theta itself is judged to 1e-5 like everywhere else): the
                # stencil part of the chain is then bit for bit
                th, f1 = np.empty((ny, nx), np.float32), np.array([flag_in], np.int32)
                assert arb.call("aleveltemp", nx, ny, t[k], p[k], "kelvin", 3, th, f1, UNDEF) == 1
                assert np.array_equal(got_theta[k] == cases.UNDEF, th == cases.UNDEF)
                assert np.allclose(got_theta[k], th, rtol=1e-5, atol=0.0)
                assert int(f_theta[k]) == int(f1[0])
                src = got_theta[k].copy()
                sm, f2 = np.empty((ny, nx), np.float32), np.array([flag_in], np.int32)
                assert arb.call("shapiro2_filter", nx, ny, src, sm, f2, UNDEF) == 1
                out, f3 = np.empty((ny, nx), np.float32), np.array([flag_in], np.int32)
                assert arb.call("thermalFrontParameter", nx, ny, sm, xm, ym, out, f3, UNDEF) == 1
                assert np.array_equal(smooth[k].cpu().numpy(), sm), (rep, k)
                assert np.array_equal(got_tfp[k], out), (rep, k)
                assert (int(f_smooth[k]), int(f_tfp[k])) == (int(f2[0]), int(f3[0])), (rep, k)
    finally:
        gpu.graph_destroy(g)


def test_graph_in_deferred_mode_and_beside_ordinary_calls(gpu):
    """a graph launched between begin_deferred / end_deferred finalises with the other deferred calls; ordinary calls still work
    before, between and after launches"""
    import torch
    arb = _arbiter()
    nx, ny = 257, 67
    d = _inputs(5, nx, ny, True)
    bufs = {k: _dev(d[k]) for k in d}
    out_g, out_d = (torch.empty((ny, nx), dtype=torch.float32, device="cuda") for _ in range(2))
    f_g, f_d = np.array([cases.SOME], np.int32), np.array([cases.SOME], np.int32)
    gpu.graph_begin()
    gpu.call("relvort", nx, ny, bufs["u"], bufs["v"], bufs["xm"], bufs["ym"], out_g, f_g, UNDEF)
    g = gpu.graph_end()
    try:
        want = _cfg1_reference(arb, nx, ny, d, cases.SOME)
        gpu.begin_deferred()
        gpu.graph_launch(g)
        gpu.call("divergence", nx, ny, bufs["u"], bufs["v"], bufs["xm"], bufs["ym"], out_d, f_d, UNDEF)
        gpu.end_deferred()
        assert np.array_equal(out_g.cpu().numpy(), want["relvort"][0]) and int(f_g[0]) == want["relvort"][1]
        assert np.array_equal(out_d.cpu().numpy(), want["divergence"][0]) and int(f_d[0]) == want["divergence"][1]
        # an ordinary call with HOST pointers, then the graph again
        o, f = np.empty((ny, nx), np.float32), np.array([cases.SOME], np.int32)
        assert gpu.call("relvort", nx, ny, d["u"], d["v"], d["xm"], d["ym"], o, f, UNDEF) == 1
        assert np.array_equal(o, want["relvort"][0])
        out_g.fill_(0)
        torch.cuda.synchronize()
        gpu.graph_launch(g)
        assert np.array_equal(out_g.cpu().numpy(), want["relvort"][0])
    finally:
        gpu.graph_destroy(g)


def test_what_cannot_be_captured_fails_loudly(gpu):
    """host-memory fields and operators with a host-side decision make the capture fail: graph_end raises with the reason, and the
    library works normally afterwards"""
    import torch
    nx, ny = 64, 33
    d = _inputs(9, nx, ny, False)
    o, f = np.empty((ny, nx), np.float32), np.array([cases.ALL], np.int32)
    gpu.graph_begin()
    with pytest.raises(RuntimeError, match="host-memory"):
        gpu.call("relvort", nx, ny, d["u"], d["v"], d["xm"], d["ym"], o, f, UNDEF)
    with pytest.raises(RuntimeError):
        gpu.graph_end()
    with pytest.raises(RuntimeError, match="without"):
        gpu.graph_end()
    # nested / misplaced calls
    gpu.graph_begin()
    with pytest.raises(RuntimeError):
        gpu.graph_begin()
    with pytest.raises(RuntimeError):
        gpu.synchronize()
    with pytest.raises(RuntimeError):
        gpu.graph_end()
    # the library is intact
    assert gpu.call("relvort", nx, ny, d["u"], d["v"], d["xm"], d["ym"], o, f, UNDEF) == 1
    arb = _arbiter()
    o2, f2 = np.empty((ny, nx), np.float32), np.array([cases.ALL], np.int32)
    arb.call("relvort", nx, ny, d["u"], d["v"], d["xm"], d["ym"], o2, f2, UNDEF)
    assert np.array_equal(o, o2)
    # an empty graph is valid and launches nothing
    gpu.graph_begin()
    g = gpu.graph_end()
    assert gpu.graph_kernels(g) == 0
    gpu.graph_launch(g)
    gpu.graph_destroy(g)
    torch.cuda.synchronize()


@pytest.mark.parametrize("which", ["legacy_default", "torch_side_stream"])
def test_capture_with_a_caller_stream_installed(gpu, which):
    """fcb200_set_stream(torch's current stream) -- the legacy default stream cannot be captured: the capture runs on the library's
    own stream, the launch on the caller's (bench.py's set-up)"""
    import torch
    arb = _arbiter()
    nx, ny = 300, 67
    d = _inputs(21, nx, ny, False)
    side = torch.cuda.Stream() if which == "torch_side_stream" else None
    stream = side if side is not None else torch.cuda.current_stream()
    bufs = {k: _dev(d[k]) for k in d}
    out = torch.empty((ny, nx), dtype=torch.float32, device="cuda")
    f = np.array([cases.ALL], np.int32)
    torch.cuda.synchronize()
    gpu.set_stream(stream.cuda_stream, True)
    try:
        gpu.graph_begin()
        gpu.call("relvort", nx, ny, bufs["u"], bufs["v"], bufs["xm"], bufs["ym"], out, f, UNDEF)
        g = gpu.graph_end()
        for _ in range(2):
            out.fill_(0)
            torch.cuda.synchronize()
            gpu.graph_launch(g)
            want = _cfg1_reference(arb, nx, ny, d, cases.ALL)["relvort"]
            assert np.array_equal(out.cpu().numpy(), want[0]) and int(f[0]) == want[1]
        gpu.graph_destroy(g)
    finally:
        gpu.set_stream(None, False)
