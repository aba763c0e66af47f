"""GPU suite, runtime side of the boundary: re-entrancy from several host threads, one host
thread per device, counts and grids beyond the limits a single launch used to impose, deferred mode under
the C++ shim.  Every result is checked against the oracle like the parity tests proper."""
import os
import subprocess
import threading
import zlib

import numpy as np
import pytest

import cases
import matrix

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "mi-fieldcalc_b200", "lib")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _arbiter():
    import fclibs
    return fclibs.reference() or fclibs.oracle()


def _to_device(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


MIXED = [("relvort", {}, 949, 41), ("thermalFrontParameter", {}, 257, 67), ("alevelhum", dict(compute=5), 949, 23), ("meanValue", {}, 131, 37),
         ("stddevValue", {}, 131, 37), ("shapiro2_filter", {}, 128, 45), ("fieldOPERfield", dict(compute=4), 949, 17), ("advection", {}, 300, 33),
         ("vesselIcingOverland", {}, 211, 19), ("pleveltemp", dict(compute=4, unit=""), 949, 29), ("probability", {}, 67, 21), ("ilevelgwind", {}, 300, 40)]


def test_concurrent_host_threads(gpu):
    """the reference is re-entrant (its only global state is OpenMP's, openmp_tools.h:43-44): four host threads call a mix of
    operators at once -- host and device pointers, masked and unmasked -- each thread on its own stream and arena"""
    arb = _arbiter()
    work = []
    for k, (name, params, nx, ny) in enumerate(MIXED):
        for mask, flag in (("none", cases.ALL), ("bernoulli", cases.SOME)):
            case = cases.build(name, nx, ny, seed=900 + k, flag_in=flag, mask=mask, **params)
            work.append((case, cases.run(arb, case)))
    errors = []

    def worker(tid):
        try:
            import torch
            torch.cuda.set_device(0)
            for rep in range(3):
                for j in range(len(work)):
                    case, want = work[(j * 5 + tid * 7 + rep) % len(work)]
                    dev = (j + tid + rep) % 2 == 1
                    got = cases.run(gpu, case, to_device=_to_device if dev else None)
                    problems = cases.compare(case, got, want, rtol=cases.TRANSCENDENTAL.get(case.name, 0.0))
                    if problems:
                        errors.append("thread %d %s device=%s: %s" % (tid, case, dev, problems))
        except Exception as e:  # noqa: BLE001
            errors.append("thread %d: %r" % (tid, e))

    threads = [threading.Thread(target=worker, args=(t,)) for t in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, "\n".join(errors[:10])


def test_one_host_thread_per_device(gpu):
    """INTEGRATION.md's multi-GPU set-up inside one process: a host thread per device (fcb200_set_device).  Batched stencils
    need the > 48 KB shared-memory opt-in, which is a per-DEVICE kernel attribute: both devices must get it."""
    import torch
    if gpu.device_count() < 2:
        pytest.skip("needs two devices")
    arb = _arbiter()
    errors = []

    def worker(device):
        try:
            gpu.set_device(device)
            torch.cuda.set_device(device)
            for name in ("relvort", "advection", "thermalFrontParameter", "jacobian", "ilevelgwind"):
                nx, ny, nf = 949, 67, 8
                singles = [cases.build(name, nx, ny, seed=50 + k, flag_in=cases.SOME if k % 2 else cases.ALL, mask="bernoulli" if k % 2 else "none",
                                       **matrix.VARIANTS[name][0]) for k in range(nf)]
                spec = cases.SPECS[name]
                args, out_pos, flags = [], [], np.array([c.args[c.flag_idx][0] for c in singles], np.int32)
                for pos, d in enumerate(spec):
                    a0 = singles[0].args[pos]
                    if d == "ny":
                        args += [a0, nf]
                    elif d == "flag":
                        args.append(flags)
                    elif d == "out" or (isinstance(d, tuple) and d[0] == "in"):
                        st = torch.from_numpy(np.stack([c.args[pos] for c in singles])).cuda(device)
                        if d == "out":
                            out_pos.append(len(args))
                        args.append(st)
                    elif isinstance(d, tuple) and d[0] == "in!":
                        for c in singles[1:]:
                            c.args[pos] = a0
                        args.append(torch.from_numpy(a0).cuda(device))
                    else:
                        args.append(a0)
                r = gpu.call(name + "_batched", *args)
                if r != 1:
                    errors.append("device %d %s: rc %d %s" % (device, name, r, gpu.last_error()))
                    continue
                for k, c in enumerate(singles):
                    outs = [args[p][k].cpu().numpy() for p in out_pos]
                    problems = cases.compare(c, (1, outs, int(flags[k])), cases.run(arb, c))
                    if problems:
                        errors.append("device %d %s field %d: %s" % (device, name, k, problems))
        except Exception as e:  # noqa: BLE001
            errors.append("device %d: %r" % (device, e))

    threads = [threading.Thread(target=worker, args=(d,)) for d in (1, 0)]  # device 1 first: it must not inherit device 0's opt-in
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    gpu.set_device(0)
    torch.cuda.set_device(0)
    assert not errors, "\n".join(errors[:10])


def test_counts_beyond_the_former_caps(gpu):
    """sumFields with more than 64 fields, values2classes with more than 64 limits, ensembles of more than 2048 (tables in
    device memory) and more than 4096 members (no reciprocal Welford): the reference takes any count"""
    arb = _arbiter()
    todo = [("sumFields", dict(), dict(nmembers=70), 61, 9), ("sumFields", dict(), dict(nmembers=200), 19, 7),
            ("values2classes", dict(limits=tuple(np.linspace(200.0, 320.0, 90))), {}, 131, 23),
            ("meanValue", {}, dict(nmembers=2100), 19, 5), ("stddevValue", {}, dict(nmembers=2100), 19, 5), ("stddevValue", {}, dict(nmembers=4200), 9, 5),
            ("extremeValue", dict(compute=3), dict(nmembers=2100), 19, 5), ("probability", dict(compute=3), dict(nmembers=2100), 19, 5)]
    for name, params, kw, nx, ny in todo:
        for mask, flag in (("none", cases.ALL), ("bernoulli", cases.SOME)):
            for device in (False, True):
                case = cases.build(name, nx, ny, seed=zlib.crc32(name.encode()), flag_in=flag, mask=mask, **kw, **params)
                got = cases.run(gpu, case, to_device=_to_device if device else None)
                problems = cases.compare(case, got, cases.run(arb, case))
                assert not problems, "%s %s %s device=%s: %s (%s)" % (name, kw, mask, device, problems, gpu.last_error())


def test_tile_engine_beyond_65535_tiles(gpu):
    """a tall grid with more than 65 535 tiles (256 x 8 points each): the tile engine's grid is one-dimensional, so any
    nx*ny < 2^31 fits one launch (it used to fail with 'grid too large')"""
    import torch
    arb = _arbiter()
    nx, ny = 300, 8 * 33000 + 2
    rng = np.random.default_rng(3)
    u = rng.uniform(-30, 30, (ny, nx)).astype(np.float32)
    v = rng.uniform(-30, 30, (ny, nx)).astype(np.float32)
    xm = rng.uniform(1.9e-4, 2.1e-4, (ny, nx)).astype(np.float32)
    ym = rng.uniform(1.9e-4, 2.1e-4, (ny, nx)).astype(np.float32)
    u[rng.random(u.shape) < 0.01] = cases.UNDEF
    want = np.empty_like(u)
    fw = np.array([cases.SOME], np.int32)
    assert arb.call("relvort", nx, ny, u, v, xm, ym, want, fw, float(cases.UNDEF)) == 1
    out = torch.empty((ny, nx), dtype=torch.float32, device="cuda")
    fg = np.array([cases.SOME], np.int32)
    r = gpu.call("relvort", nx, ny, _to_device(u), _to_device(v), _to_device(xm), _to_device(ym), out, fg, float(cases.UNDEF))
    assert r == 1, gpu.last_error()
    assert int(fg[0]) == int(fw[0])
    assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))


DEFERRED_SRC = r"""
// the C++ drop-in called while the thread is in the C-ABI's deferred mode: outputs and flags must still be final on return
#include "mi_fieldcalc/FieldCalculations.h"
#include "fcb200.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>
namespace fc = miutil::fieldcalc;
static int run(bool deferred, std::vector<float>& rv, std::vector<float>& th, std::vector<float>& mean, int* flags)
{
  const int nx = 300, ny = 41, n = nx * ny;
  const float undef = 1e35f;
  std::vector<float> u(n), v(n), xm(n, 2e-4f), ym(n, 2e-4f), t(n);
  for (int i = 0; i < n; ++i) {
    u[i] = 10.f * std::sin(0.01f * i);
    v[i] = 7.f * std::cos(0.013f * i);
    t[i] = 250.f + 30.f * std::sin(0.002f * i);
  }
  u[5 * nx + 17] = undef;
  t[3] = undef;
  rv.assign(n, -1.f), th.assign(n, -1.f), mean.assign(n, -1.f);
  if (deferred)
    fcb200_begin_deferred();
  miutil::ValuesDefined f1 = miutil::SOME_DEFINED, f2 = miutil::SOME_DEFINED, f3 = miutil::SOME_DEFINED;
  bool ok = fc::relvort(nx, ny, u.data(), v.data(), xm.data(), ym.data(), rv.data(), f1, undef);
  flags[0] = f1; // read right after the call, as a reference caller would
  ok = fc::pleveltemp(nx, ny, t.data(), 500.f, "kelvin", 4, th.data(), f2, undef) && ok;
  flags[1] = f2;
  std::vector<float*> members = {u.data(), v.data(), t.data()};
  std::vector<miutil::ValuesDefined> fin = {miutil::SOME_DEFINED, miutil::ALL_DEFINED, miutil::SOME_DEFINED};
  ok = fc::meanValue(nx, ny, members, fin, mean.data(), f3, undef) && ok;
  flags[2] = f3;
  const float probe = rv[7 * nx + 9] + th[11] + mean[12]; // outputs are final on return
  if (deferred)
    fcb200_end_deferred();
  return ok && probe == probe ? 0 : 1;
}
int main()
{
  std::vector<float> a1, a2, a3, b1, b2, b3;
  int fa[3], fb[3];
  if (run(false, a1, a2, a3, fa) || run(true, b1, b2, b3, fb))
    return std::printf("call failed\n"), 1;
  int bad = std::memcmp(fa, fb, sizeof fa) != 0;
  bad += std::memcmp(a1.data(), b1.data(), a1.size() * 4) != 0;
  bad += std::memcmp(a2.data(), b2.data(), a2.size() * 4) != 0;
  bad += std::memcmp(a3.data(), b3.data(), a3.size() * 4) != 0;
  // and from four threads at once, two of them deferred
  std::vector<std::thread> th;
  int tbad[4] = {0, 0, 0, 0};
  for (int k = 0; k < 4; ++k)
    th.emplace_back([&, k] {
      for (int rep = 0; rep < 5; ++rep) {
        std::vector<float> c1, c2, c3;
        int fc_[3];
        if (run(k % 2 == 1, c1, c2, c3, fc_) || std::memcmp(fa, fc_, sizeof fa) || std::memcmp(a1.data(), c1.data(), a1.size() * 4) ||
            std::memcmp(a2.data(), c2.data(), a2.size() * 4) || std::memcmp(a3.data(), c3.data(), a3.size() * 4))
          tbad[k] += 1;
      }
    });
  for (auto& t : th)
    t.join();
  bad += tbad[0] + tbad[1] + tbad[2] + tbad[3];
  std::printf("flags %d %d %d, %d mismatches\n", fa[0], fa[1], fa[2], bad);
  return bad ? 1 : 0;
}
"""


def test_shim_is_final_on_return_in_deferred_mode(tmp_path):
    """the shim's flag lives in the wrapper's stack frame, so a shim call made while the thread is in deferred mode
    must drain before it returns (it used to write the flag through a dangling pointer at fcb200_end_deferred())"""
    src = tmp_path / "deferred.cc"
    src.write_text(DEFERRED_SRC)
    exe = str(tmp_path / "deferred")
    subprocess.run([CXX, "-std=c++11", "-O2", "-pthread", "-I", os.path.join(ROOT, "include"), str(src), "-o", exe, "-L", LIB, "-l:libmi-fieldcalc.so.0",
                    "-lfcb200", "-Wl,-rpath," + LIB], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
