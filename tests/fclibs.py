"""Loaders for the three libraries that share the signatures of include/fcb200_api.inc.

  product()    mi-fieldcalc_b200/lib/libfcb200.so   fcb200_*  the thing under test (CUDA, no fallback)
  oracle()     oracle/libfcoracle.so                fco_*     plain-C restatement (test infrastructure)
  reference()  oracle/_ref/libfcref.so              fcref_*   the unmodified reference, when built

Nothing here reads /root/reference at run time: the reference library is prebuilt by
`make -C oracle ref` (done by __graft_entry__.build() in the container that has /root/reference)
and travels to the GPU box as a git-ignored binary.
"""
import importlib
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

pkg = importlib.import_module("mi-fieldcalc_b200")
capi = pkg.capi

_cache = {}


def product():
    if "product" not in _cache:
        _cache["product"] = capi.load()
    return _cache["product"]


def oracle():
    if "oracle" not in _cache:
        lib = os.path.join(ORACLE_DIR, "libfcoracle.so")
        src = os.path.join(ORACLE_DIR, "fc_oracle.c")
        if not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(src):
            subprocess.run(["make", "-C", ORACLE_DIR, "oracle"], check=True, capture_output=True)
        _cache["oracle"] = capi.Api(lib, "fco_", batched=False)
    return _cache["oracle"]


def reference(openmp=False):
    key = "ref_omp" if openmp else "ref"
    if key not in _cache:
        lib = os.path.join(ORACLE_DIR, "_ref", "libfcref_omp.so" if openmp else "libfcref.so")
        _cache[key] = capi.Api(lib, "fcref_", batched=False) if os.path.exists(lib) else None
    return _cache[key]
