"""GPU suite: the CUDA product (through the C-ABI of include/fcb200.h) against the oracle on the
same seeded inputs.  When the compiled reference (oracle/_ref) travelled to the box it is the
arbiter, otherwise the plain-C restatement is.

Bar (BASELINE.json north_star): return value, ValuesDefined flag and undefined mask bit-exact;
values bit-exact for every operator without a device transcendental, and within the relative
tolerance of cases.TRANSCENDENTAL (<= 1e-5; MINCOG 1e-4, its bisection amplifies expf ulps)
for the others.
"""
import zlib

import numpy as np
import pytest

import cases
import matrix

pytestmark = pytest.mark.gpu

MATRIX = matrix.small_matrix()


def _arbiter():
    import fclibs
    return fclibs.reference() or fclibs.oracle()


def _to_device(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name", sorted(matrix.VARIANTS))
def test_small_matrix_host_pointers(gpu, name):
    """every variant x mask x flag of one operator, host pointers (the drop-in path)"""
    arb = _arbiter()
    failures = []
    for c in MATRIX:
        if c[0] != name:
            continue
        _, params, nx, ny, mask, flag = c
        case = cases.build(name, nx, ny, seed=zlib.crc32(matrix.case_id(c).encode()), flag_in=flag, mask=mask, **params)
        problems = cases.compare(case, cases.run(gpu, case), cases.run(arb, case), rtol=cases.TRANSCENDENTAL.get(name, 0.0))
        if problems:
            failures.append("%s: %s" % (matrix.case_id(c), "; ".join(problems)))
    assert not failures, "%d failing cases:\n%s" % (len(failures), "\n".join(failures[:20]))


@pytest.mark.parametrize("name", sorted(matrix.VARIANTS))
def test_device_pointers_meps_rows(gpu, name):
    """device-resident fields (the roofline path) on an odd-sized grid: nx = 949 like MEPS"""
    arb = _arbiter()
    nx, ny = (949, 37) if name not in matrix.SLOW else (61, 9)
    for params in matrix.VARIANTS[name][:3]:
        for mask, flag in [("none", cases.ALL), ("bernoulli", cases.SOME), ("edge", cases.SOME)]:
            case = cases.build(name, nx, ny, seed=17, flag_in=flag, mask=mask, **params)
            problems = cases.compare(case, cases.run(gpu, case, to_device=_to_device), cases.run(arb, case), rtol=cases.TRANSCENDENTAL.get(name, 0.0))
            assert not problems, "%s %s %s: %s" % (name, params, mask, "\n".join(problems))


@pytest.mark.parametrize("name", ["pleveltemp", "plevelhum", "aleveltemp", "alevelhum", "fieldOPERfield", "windCooling", "shapiro2_filter"])
@pytest.mark.parametrize("device", [False, True])
def test_output_may_alias_input(gpu, name, device):
    arb = _arbiter()
    case = cases.build(name, 131, 67, seed=3, flag_in=cases.SOME, mask="bernoulli", alias=True, **matrix.VARIANTS[name][0])
    got = cases.run(gpu, case, to_device=_to_device if device else None)
    problems = cases.compare(case, got, cases.run(arb, case), rtol=cases.TRANSCENDENTAL.get(name, 0.0))
    assert not problems, "\n".join(problems)


def test_other_undef_values(gpu):
    arb = _arbiter()
    for undef in (12356789.0, 1e30, 123456.0, -999.0):
        for name in ("relvort", "alevelhum", "meanValue", "extremeValue", "probability", "fieldOPERfield"):
            case = cases.build(name, 19, 11, seed=11, undef=undef, flag_in=cases.SOME, mask="bernoulli", **matrix.VARIANTS[name][0])
            problems = cases.compare(case, cases.run(gpu, case), cases.run(arb, case), rtol=cases.TRANSCENDENTAL.get(name, 0.0))
            assert not problems, "%s undef=%g: %s" % (name, undef, "\n".join(problems))


def test_ensemble_member_flags(gpu):
    arb = _arbiter()
    flags = [cases.ALL, cases.NONE, cases.SOME, cases.SOME, cases.NONE, cases.ALL, cases.SOME]
    for name in ("meanValue", "stddevValue", "probability"):
        for params in matrix.VARIANTS[name]:
            case = cases.build(name, 23, 9, seed=5, mask="bernoulli", nmembers=len(flags), member_flags=flags, **params)
            problems = cases.compare(case, cases.run(gpu, case), cases.run(arb, case))
            assert not problems, "%s %s: %s" % (name, params, "\n".join(problems))


@pytest.mark.parametrize("name,grid", [("pleveltemp", (949, 1069)), ("relvort", (949, 1069)), ("divergence", (949, 1069)), ("aleveltemp", (949, 1069)),
                                       ("alevelhum", (949, 1069)), ("advection", (3600, 1801)), ("thermalFrontParameter", (3600, 1801)),
                                       ("shapiro2_filter", (3600, 1801)), ("meanValue", (949, 1069)), ("stddevValue", (949, 1069)),
                                       ("extremeValue", (949, 1069)), ("probability", (949, 1069)), ("vesselIcingOverland", (949, 1069)),
                                       ("fieldOPERfield", (949, 1069))])
def test_full_size_fields(gpu, name, grid):
    """BASELINE.json grids (MEPS 949x1069, ECMWF 3600x1801), one field, masked and unmasked"""
    arb = _arbiter()
    nx, ny = grid
    params = matrix.VARIANTS[name][0]
    if name == "fieldOPERfield":
        params = dict(compute=4)
    for mask, flag in [("none", cases.ALL), ("bernoulli", cases.SOME)]:
        case = cases.build(name, nx, ny, seed=1000, flag_in=flag, mask=mask, nmembers=30 if name in matrix.ENSEMBLE else 5, **params)
        problems = cases.compare(case, cases.run(gpu, case, to_device=_to_device), cases.run(arb, case), rtol=cases.TRANSCENDENTAL.get(name, 0.0))
        assert not problems, "%s %s: %s" % (name, mask, "\n".join(problems))
