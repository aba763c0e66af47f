"""The reference's own unit tests for the hot path (restated in tests/cpp/test_fieldcalc_api.cc)
compiled against the drop-in header and linked against
  - the unmodified reference (CPU, validates the test source itself), and
  - libmi-fieldcalc.so.0 from this repository (the CUDA product, -m gpu)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_fieldcalc_api.cc")
LIB = os.path.join(ROOT, "mi-fieldcalc_b200", "lib")
REF = os.path.join(ROOT, "oracle", "_ref")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"


def _build(tmp_path, libdir, libname):
    exe = str(tmp_path / "test_api")
    subprocess.run([CXX, "-std=c++11", "-O2", "-I", os.path.join(ROOT, "include"), SRC, "-o", exe, "-L", libdir, "-l:" + libname,
                    "-Wl,-rpath," + libdir], check=True)
    return exe


def test_reference_tests_pass_on_reference(tmp_path):
    if not os.path.exists(os.path.join(REF, "libfcref.so")):
        pytest.skip("oracle/_ref not built")
    r = subprocess.run([_build(tmp_path, REF, "libfcref.so")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr


def test_shim_exports_the_reference_symbols():
    """every miutil::fieldcalc function of the shim has a same-named mangled symbol in the reference build"""
    shim = os.path.join(LIB, "libmi-fieldcalc.so.0")
    ref = os.path.join(REF, "libfcref.so")
    if not (os.path.exists(shim) and os.path.exists(ref)):
        pytest.skip("shim or oracle/_ref not built")

    def syms(path):
        out = subprocess.run(["nm", "-D", "--defined-only", path], capture_output=True, text=True, check=True).stdout
        return {l.split()[-1] for l in out.splitlines() if " T " in l or " B " in l or " R " in l or " D " in l}

    s, r = syms(shim), syms(ref)
    ours = {x for x in s if "6miutil" in x or x == "fieldUndef"}
    assert len(ours) >= 35
    missing = sorted(ours - r)
    assert not missing, "symbols that the reference does not export under the same mangled name: %s" % missing
    # and the other way round: the whole exported ABI of the reference (72 miutil::fieldcalc functions,
    # fieldUndef / UNDEF, checkDefined x2, combineDefined, compute_num_threads, four ICAO conversions) except the
    # saturation-table helper class, which is an implementation detail of the field functions
    theirs = {x for x in r if ("6miutil" in x or x == "fieldUndef") and "ewt_calculator" not in x}
    assert len(theirs) >= 80
    absent = sorted(theirs - s)
    assert not absent, "exported by the reference but not by the drop-in: %s" % absent


def test_icao_host_utilities_match_the_reference():
    """MetConstants.cc:81-131 restated in the shim (host scalars): bit-identical on a sweep of pressures / altitudes"""
    import ctypes

    import numpy as np
    shim = os.path.join(LIB, "libmi-fieldcalc.so.0")
    ref = os.path.join(REF, "libfcref.so")
    if not (os.path.exists(shim) and os.path.exists(ref)):
        pytest.skip("shim or oracle/_ref not built")
    a, b = ctypes.CDLL(shim), ctypes.CDLL(ref)
    xs = np.concatenate([np.linspace(-500, 90000, 2001), np.geomspace(1e-4, 1100, 2001)])
    for name, restype in (("_ZN6miutil9constants31ICAO_geo_altitude_from_pressureEd", ctypes.c_double),
                          ("_ZN6miutil9constants31ICAO_pressure_from_geo_altitudeEd", ctypes.c_double),
                          ("_ZN6miutil9constants20FL_from_geo_altitudeEd", ctypes.c_int),
                          ("_ZN6miutil9constants20geo_altitude_from_FLEd", ctypes.c_double)):
        fa, fb = getattr(a, name), getattr(b, name)
        fa.restype = fb.restype = restype
        fa.argtypes = fb.argtypes = [ctypes.c_double]
        assert all(fa(float(x)) == fb(float(x)) for x in xs), name


@pytest.mark.gpu
def test_reference_tests_pass_on_cuda_shim(tmp_path):
    assert os.path.exists(os.path.join(LIB, "libmi-fieldcalc.so.0")), "shim not built: run __graft_entry__.build()"
    r = subprocess.run([_build(tmp_path, LIB, "libmi-fieldcalc.so.0")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "0 failed" in r.stdout
