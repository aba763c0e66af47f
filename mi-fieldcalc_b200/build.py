"""Build the in-tree native libraries of the B200 FieldCalculations hot path.

    python mi-fieldcalc_b200/build.py [--force] [--no-shim] [--no-pybind]

Outputs (git-ignored, shipped to the GPU box by gpurun):
    mi-fieldcalc_b200/lib/libfcb200.so            C-ABI + sm_100a kernels (include/fcb200.h)
    mi-fieldcalc_b200/lib/libmi-fieldcalc.so.0    drop-in C++ shim, the reference's mangled API
    mi-fieldcalc_b200/lib/mi_fieldcalc*.so        pybind11 module mirroring python/py_mi_fieldcalc.cc

Everything is compiled for sm_100a only (-gencode arch=compute_100a,code=sm_100a), with
-fmad=false: the reference build has no FMA contraction and several operators differ by 1e-3
relative if the compiler contracts a*b - c*d (hard part 2).  nvcc
cross-compiles without a GPU.
"""
from __future__ import annotations

import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "lib")

CUDA_SOURCES = ["runtime.cu", "slab.cu", "ops_elementwise.cu", "ops_diagnostics.cu", "ops_arith.cu", "ops_neighbour.cu", "ops_stencil.cu", "ops_ensemble.cu", "ops_icing.cu"]

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=default",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 FieldCalculations library cannot be built")


def _host_cxx() -> str:
    for cand in ("/usr/bin/g++", shutil.which("g++")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("g++ not found")


def _newer(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s\n%s" % (" ".join(cmd), r.stdout, r.stderr))


def _headers() -> list[str]:
    hs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    inc = os.path.join(ROOT, "include")
    hs += [os.path.join(inc, f) for f in os.listdir(inc)]
    return hs


def build_cuda(force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIB, exist_ok=True)
    nvcc = _nvcc()
    headers = _headers()
    jobs = []
    objs = []
    for src in CUDA_SOURCES:
        s = os.path.join(CSRC, src)
        o = os.path.join(OBJ, src.replace(".cu", ".o"))
        objs.append(o)
        if force or _newer(o, [s] + headers):
            jobs.append([nvcc] + NVCC_FLAGS + ["-ccbin", _host_cxx(), "-I", os.path.join(ROOT, "include"), "-c", s, "-o", o])
    with concurrent.futures.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        list(ex.map(_run, jobs))
    out = os.path.join(LIB, "libfcb200.so")
    if force or jobs or _newer(out, objs):
        _run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", _host_cxx(), "-o", out] + objs + ["-lcudart_static", "-ldl", "-lrt", "-lpthread"])
    return out


def build_shim(force: bool = False) -> str:
    """The drop-in C++ library: same mangled symbols as the reference's libmi-fieldcalc.so.0."""
    src = os.path.join(CSRC, "shim.cc")
    out = os.path.join(LIB, "libmi-fieldcalc.so.0")
    if not os.path.exists(src):
        return ""
    if force or _newer(out, [src, os.path.join(LIB, "libfcb200.so")] + _headers()):
        _run([_host_cxx(), "-std=c++11", "-O2", "-fPIC", "-shared", "-Wall", "-I", os.path.join(ROOT, "include"), "-o", out, src,
              "-Wl,-soname,libmi-fieldcalc.so.0", "-L", LIB, "-lfcb200", "-Wl,-rpath,$ORIGIN"])
        link = os.path.join(LIB, "libmi-fieldcalc.so")
        if os.path.lexists(link):
            os.remove(link)
        os.symlink("libmi-fieldcalc.so.0", link)
    return out


def build_pybind(force: bool = False) -> str:
    src = os.path.join(CSRC, "py_module.cc")
    if not os.path.exists(src):
        return ""
    import pybind11

    ext = sysconfig.get_config_var("EXT_SUFFIX") or ".so"
    out = os.path.join(LIB, "mi_fieldcalc" + ext)
    if force or _newer(out, [src, os.path.join(LIB, "libmi-fieldcalc.so.0")] + _headers()):
        _run([_host_cxx(), "-std=c++17", "-O2", "-fPIC", "-shared", "-fvisibility=hidden", "-ftemplate-depth=2048",
              "-I", pybind11.get_include(), "-I", sysconfig.get_paths()["include"], "-I", os.path.join(ROOT, "include"),
              "-o", out, src, "-L", LIB, "-l:libmi-fieldcalc.so.0", "-Wl,-rpath,$ORIGIN"])
    return out


def build_all(force: bool = False, shim: bool = True, pybind: bool = True) -> dict:
    res = {"cuda": build_cuda(force)}
    if shim:
        res["shim"] = build_shim(force)
    if shim and pybind:
        res["pybind"] = build_pybind(force)
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--no-shim", action="store_true")
    ap.add_argument("--no-pybind", action="store_true")
    a = ap.parse_args()
    for k, v in build_all(a.force, not a.no_shim, not a.no_pybind).items():
        print("%-7s %s" % (k, v))
