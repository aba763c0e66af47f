"""The Python module `mi_fieldcalc` (pybind11 subset of the reference, python/py_mi_fieldcalc.cc:189-207):
import and argument conventions on the CPU, values against the oracle on the GPU."""
import os
import sys

import numpy as np
import pytest

import cases

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mi-fieldcalc_b200", "lib")

# all 15 functions of the reference module (py_mi_fieldcalc.cc:189-207)
ON_PATH = ["kIndex", "ductingIndex", "showalterIndex", "boydenIndex", "sweatIndex", "seaSoundSpeed", "cvtemp", "cvhum", "abshum", "windCooling",
           "underCooledRain", "vesselIcingOverland", "vesselIcingMertins", "vesselIcingModStall", "vesselIcingMincog"]


def _module():
    if LIB not in sys.path:
        sys.path.insert(0, LIB)
    import mi_fieldcalc
    return mi_fieldcalc


def test_module_exports_the_reference_names():
    m = _module()
    for name in ON_PATH:
        assert callable(getattr(m, name)), name
    assert int(m.ValuesDefined.ALL_DEFINED) == 0 and int(m.ValuesDefined.NONE_DEFINED) == 1 and int(m.ValuesDefined.SOME_DEFINED) == 2


def test_shape_mismatch_and_wrong_rank_return_none():
    """reference py_mi_fieldcalc.cc:82-84 -- checked before any device work, so this runs without a GPU"""
    m = _module()
    a, b = np.ones((3, 4)), np.ones((3, 5))
    assert m.windCooling(a, b, a, 1, 1e35) is None
    assert m.windCooling(np.ones(12), np.ones(12), np.ones(12), 1, 1e35) is None
    assert m.vesselIcingOverland(a, a, a, a, a, b, 1e35) is None


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["windCooling", "vesselIcingOverland", "vesselIcingMertins", "kIndex", "ductingIndex", "showalterIndex", "boydenIndex",
                                  "sweatIndex", "seaSoundSpeed", "cvtemp", "cvhum", "abshum", "underCooledRain"])
def test_values_match_the_oracle(gpu, name):
    import fclibs
    import matrix
    arb = fclibs.reference() or fclibs.oracle()
    m = _module()
    # showalterIndex leaves points with an undefined input unwritten (fresh allocation in the module): no mask there
    case = cases.build(name, 61, 37, seed=9, flag_in=cases.SOME, mask="none" if name == "showalterIndex" else "bernoulli", **matrix.VARIANTS[name][0])
    want = cases.run(arb, case)
    spec = cases.SPECS[name]
    args = [a for d, a in zip(spec, case.args) if isinstance(d, tuple)]  # fields and scalars, in API order
    # the module takes nx = shape[0], ny = shape[1]; the operators are point-wise, any 2-D shape with the same data works
    got = getattr(m, name)(*args, float(case.undef))
    assert got is not None and got.dtype == np.float32 and got.shape == case.args[case.out_idx[0]].shape
    problems = cases.compare(case, (1, [got], want[2]), want, rtol=cases.TRANSCENDENTAL.get(name, 0.0))
    assert not problems, problems


@pytest.mark.gpu
def test_invalid_compute_returns_none(gpu):
    m = _module()
    a = np.full((5, 7), 280.0, np.float32)
    assert m.windCooling(a, a, a, 7, 1e35) is None  # reference: compute outside 1..3 -> false -> None
