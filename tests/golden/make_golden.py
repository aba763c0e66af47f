#!/usr/bin/env python
"""Generate tests/golden/reference_outputs.npz from the UNMODIFIED reference (oracle/_ref).

Run in the container that has /root/reference (after `make -C oracle ref`):
    python tests/golden/make_golden.py
For every operator of the hot path it stores the reference's return value, output flag and output
field(s) for a few seeded cases (tests/cases.py builds the inputs deterministically from the seed,
so only the outputs are stored).  tests/test_golden.py then pins the oracle (CPU) and the CUDA
product (GPU) against these vectors on machines where the reference itself is not available.
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import cases  # noqa: E402
import fclibs  # noqa: E402
import matrix  # noqa: E402


def golden_cases():
    out = []
    for name, variants in matrix.VARIANTS.items():
        nx, ny = (37, 23) if name not in matrix.SLOW else (9, 7)
        for v in variants[:4]:
            for mask, flag in (("none", cases.ALL), ("bernoulli", cases.SOME), ("edge", cases.SOME)):
                out.append((name, v, nx, ny, mask, flag))
    return out


def key(c):
    return matrix.case_id(c)


def main():
    ref = fclibs.reference()
    assert ref is not None, "build oracle/_ref first"
    store = {}
    for c in golden_cases():
        name, params, nx, ny, mask, flag = c
        case = cases.build(name, nx, ny, seed=zlib.crc32(key(c).encode()), flag_in=flag, mask=mask, **params)
        ret, outs, f = cases.run(ref, case)
        store[key(c) + "|meta"] = np.array([ret, f], np.int32)
        for k, o in enumerate(outs):
            store[key(c) + "|out%d" % k] = o
    path = os.path.join(HERE, "reference_outputs.npz")
    np.savez_compressed(path, **store)
    print("wrote %s: %d cases, %.1f KB" % (path, len(golden_cases()), os.path.getsize(path) / 1024))


if __name__ == "__main__":
    main()
