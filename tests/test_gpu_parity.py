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


@pytest.mark.parametrize("mask,flag", [("none", cases.ALL), ("none", cases.SOME), ("bernoulli", cases.SOME), ("nan", cases.SOME), ("nan", cases.ALL), ("all", cases.SOME)])
@pytest.mark.parametrize("unit", ["celsius", "kelvin"])
def test_fused_alevel_chain_equals_four_reference_calls(gpu, mask, flag, unit):
    """fcb200_alevel_chain_batched == aleveltemp(c3) + alevelhum(c1) + alevelhum(c5/9) + alevelthe(c1), field by field"""
    arb = _arbiter()
    nx, ny, nf = 949, 23, 5
    rng = np.random.default_rng(42)
    t = np.stack([cases.field(rng, "tk", nx, ny) for _ in range(nf)])
    q = np.stack([cases.field(rng, "q", nx, ny) for _ in range(nf)])
    p = np.stack([cases.field(rng, "p", nx, ny) for _ in range(nf)])
    for a in (t, q, p):
        for k in range(nf):
            cases.apply_mask(rng, a[k], mask, cases.UNDEF)
    outs = [np.full((nf, ny, nx), cases.SENTINEL, np.float32) for _ in range(4)]
    fin = np.full(nf, flag, np.int32)
    fout = np.full((4, nf), -1, np.int32)
    for device in (False, True):
        args = [t, q, p] + outs
        if device:
            args = [_to_device(a) for a in args]
        r = gpu.call("alevel_chain_batched", nx, ny, nf, args[0], args[1], args[2], unit, args[3], args[4], args[5], args[6], fin, fout, float(cases.UNDEF))
        assert r == 1
        got = [a.cpu().numpy() if device else a for a in args[3:]]
        calls = [("aleveltemp", lambda k, o, f: (nx, ny, t[k], p[k], "kelvin", 3, o, f, float(cases.UNDEF)), 1e-5, 0.0),
                 ("alevelhum", lambda k, o, f: (nx, ny, t[k], q[k], p[k], unit, 1, o, f, float(cases.UNDEF)), 0.0, 0.0),
                 ("alevelhum", lambda k, o, f: (nx, ny, t[k], q[k], p[k], unit, 5, o, f, float(cases.UNDEF)), 0.0, 273.15),
                 ("alevelthe", lambda k, o, f: (nx, ny, t[k], q[k], p[k], 1, o, f, float(cases.UNDEF)), 1e-5, 0.0)]
        for oi, (name, mk, rtol, floor) in enumerate(calls):
            for k in range(nf):
                o = np.full((ny, nx), cases.SENTINEL, np.float32)
                f = np.array([flag], np.int32)
                assert arb.call(name, *mk(k, o, f)) == 1
                case = cases.Case(name, [], [], None, cases.UNDEF, {})
                case.floor = floor
                problems = cases.compare(case, (1, [got[oi][k]], int(fout[oi, k])), (1, [o], int(f[0])), rtol=rtol)
                assert not problems, "output %d (%s) field %d device=%s: %s" % (oi, name, k, device, problems)
