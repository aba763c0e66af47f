"""CPU suite: the float shortcuts of the CUDA kernels against the reference's expressions, on random and adversarial operands.

Where a kernel replaces an expression of the reference (a double-precision evaluation, an IEEE division) by cheaper float
arithmetic, it does so only on operands for which the result is provably the same float, and falls back to the reference's own
expression elsewhere.  tests/proofs/shortcuts.c restates those acceptance tests and shortcuts in plain C
(gcc -ffp-contract=off, glibc's correctly rounded fmaf) next to the reference's expressions (FC.cc:2024-2057, 2113-2121, 2147-2155, 2295-2296, 2741-2744, MetConstants.cc:44)
and counts accepted operands whose results differ: there must be none.  The GPU parity tests check the same thing end to end on
fields; this checks the arithmetic itself on 10^8 operand tuples, zeros, subnormals, infinities and NaNs included."""
import os
import re
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def shortcuts_binary(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("proofs") / "shortcuts")
    src = os.path.join(HERE, "proofs", "shortcuts.c")
    r = subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-o", exe, src, "-lm"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_float_shortcuts_reproduce_the_reference_expressions(shortcuts_binary):
    r = subprocess.run([shortcuts_binary, "3"], capture_output=True, text=True, timeout=600)
    lines = dict((m.group(1), (int(m.group(2)), int(m.group(3)), int(m.group(4))))
                 for m in re.finditer(r"^(\w+) cases=(\d+) accepted=(\d+) mismatches=(\d+)", r.stdout, re.M))
    assert set(lines) == {"shapiro_all", "shapiro_masked", "welford_quotient", "welford_replacement", "tfp_quotient", "half_map_diff", "table_inverse_tail",
                          "midrange_div"}, r.stdout + r.stderr
    for name, (cases, accepted, mismatches) in lines.items():
        assert mismatches == 0, "%s: %d of %d accepted operands differ from the reference's expression" % (name, mismatches, accepted)
        assert accepted > cases // 4, "%s: the shortcut is hardly ever taken (%d of %d)" % (name, accepted, cases)
    assert r.returncode == 0
