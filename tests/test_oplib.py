"""bench.py's operator work-lists (tools/oplib.py) on the CPU: every row's batched argument list matches the C-ABI signature
it will be called with, and the single-field twin is accepted by the oracle (so the CPU timing beside each GPU number times
the right call on the right arguments).  No GPU, no product compute."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

import fclibs  # noqa: E402
import oplib  # noqa: E402

ROWS = oplib.rows(levels_cfg3=3) + oplib.extra_rows()


@pytest.mark.parametrize("row", ROWS, ids=[r.name + "-" + r.cfg for r in ROWS])
def test_row_matches_the_abi_and_runs_on_the_oracle(row):
    row.grid, row.nf = (37, 23), min(row.nf, 3)
    if row.cpu_rows:
        row.cpu_rows = 9
    row.items = [(it[0], it[1][: row.nf]) if isinstance(it, tuple) and it[0] == "SV" else it for it in row.items]
    b = oplib.Built(row, oplib.Inputs(torch, torch.device("cpu"), seed=1), 0.0)
    sigs = fclibs.capi.parse_inc(os.path.join(ROOT, "include", "fcb200_batched.inc"), "FCB_FN")
    assert row.call in sigs, row.call
    assert len(b.sets[0]) == len(sigs[row.call]), (row.call, len(b.sets[0]), [n for _, n in sigs[row.call]])
    single, pts = b.single_host_args()
    arb = fclibs.reference() or fclibs.oracle()
    name = row.call.replace("_batched", "")
    assert len(single) == len(arb.sigs[name])
    assert arb.call(name, *single) == 1
    out = [a for a in single if isinstance(a, np.ndarray) and a.dtype == np.float32 and a.ndim == 2][-1]
    assert pts == out.size
    if b.mask == 0.0 and row.cfg != "cfg5":
        flags = [a for a in single if isinstance(a, np.ndarray) and a.dtype == np.int32]
        assert flags, "no flag argument"
