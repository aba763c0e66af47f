"""CPU suite, part 1: the plain-C restatement (oracle/fc_oracle.c) against the UNMODIFIED reference
compiled into oracle/_ref (when present).  Everything must be bit-identical -- values, undefined
mask, ValuesDefined flag and return value -- because both sides use the same libm and the same
expression types.  This is what pins the oracle."""
import zlib

import numpy as np
import pytest

import cases
import matrix

MATRIX = matrix.small_matrix()


@pytest.mark.parametrize("c", MATRIX, ids=[matrix.case_id(c) for c in MATRIX])
def test_oracle_matches_reference(oracle, ref, c):
    name, params, nx, ny, mask, flag = c
    seed = zlib.crc32(matrix.case_id(c).encode())
    case = cases.build(name, nx, ny, seed=seed, flag_in=flag, mask=mask, **params)
    want = cases.run(ref, case)
    got = cases.run(oracle, case)
    problems = cases.compare(case, got, want, rtol=0.0)
    assert not problems, "\n".join(problems)


@pytest.mark.parametrize("name", sorted(matrix.VARIANTS))
def test_oracle_matches_reference_meps_row(oracle, ref, name):
    """one larger, odd-sized grid per operator (nx = 949 like MEPS, a few rows)"""
    nx, ny = (949, 12) if name not in matrix.SLOW else (61, 9)
    for mask, flag in [("none", cases.ALL), ("bernoulli", cases.SOME)]:
        case = cases.build(name, nx, ny, seed=7, flag_in=flag, mask=mask, **matrix.VARIANTS[name][0])
        problems = cases.compare(case, cases.run(oracle, case), cases.run(ref, case), rtol=0.0)
        assert not problems, "%s %s: %s" % (name, mask, "\n".join(problems))


@pytest.mark.parametrize("name", ["pleveltemp", "plevelhum", "aleveltemp", "alevelhum", "fieldOPERfield", "windCooling", "shapiro2_filter"])
def test_output_may_alias_input(oracle, ref, name):
    case = cases.build(name, 31, 17, seed=3, flag_in=cases.SOME, mask="bernoulli", alias=True, **matrix.VARIANTS[name][0])
    problems = cases.compare(case, cases.run(oracle, case), cases.run(ref, case), rtol=0.0)
    assert not problems, "\n".join(problems)


def test_other_undef_values(oracle, ref):
    for undef in (12356789.0, 1e30, 123456.0, -999.0):
        for name in ("relvort", "alevelhum", "meanValue", "extremeValue", "probability", "fieldOPERfield"):
            case = cases.build(name, 19, 11, seed=11, undef=undef, flag_in=cases.SOME, mask="bernoulli", **matrix.VARIANTS[name][0])
            problems = cases.compare(case, cases.run(oracle, case), cases.run(ref, case), rtol=0.0)
            assert not problems, "%s undef=%g: %s" % (name, undef, "\n".join(problems))


def test_ensemble_member_flags(oracle, ref):
    """mixed per-member flags: ALL (skip tests), NONE (probability skips the member), SOME"""
    flags = [cases.ALL, cases.NONE, cases.SOME, cases.SOME, cases.NONE, cases.ALL, cases.SOME]
    for name in ("meanValue", "stddevValue", "probability"):
        for params in matrix.VARIANTS[name]:
            case = cases.build(name, 23, 9, seed=5, mask="bernoulli", nmembers=len(flags), member_flags=flags, **params)
            problems = cases.compare(case, cases.run(oracle, case), cases.run(ref, case), rtol=0.0)
            assert not problems, "%s %s: %s" % (name, params, "\n".join(problems))
    case = cases.build("probability", 5, 4, seed=5, nmembers=3, member_flags=[cases.NONE] * 3)
    assert not cases.compare(case, cases.run(oracle, case), cases.run(ref, case))
