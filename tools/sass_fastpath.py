#!/usr/bin/env python
"""Static instruction count of a kernel's hot path from its SASS (no GPU needed).

    python tools/sass_fastpath.py mi-fieldcalc_b200/build/ops_stencil.o shapiro2_kernel_vec_all            # loops (backward branches)
    python tools/sass_fastpath.py mi-fieldcalc_b200/build/ops_stencil.o shapiro2_kernel_vec_all 0860 86a0  # count one loop body

The count walks the instructions from <start> to <end> (addresses as cuobjdump prints them), treating the branch that follows a
`VOTE.ANY` within three instructions as TAKEN -- the kernels of this library skip their rare slow paths with
`if (__any_sync(mask, doubt)) { slow path }`, which compiles to VOTE.ANY + `@!P BRA past_the_slow_path` -- and every other
forward branch as not taken.  Divide by the points a thread produces per trip for instructions per point (shapiro2: 24 points per
trip of the six-row body).  This is how the 44 -> 26 instructions per point of DESIGN.md 9.4 were found and checked before any
GPU time was spent; `ncu`'s `smsp__inst_executed.sum` on the device agrees (31 per stored point with the halo lanes and rows).
"""
import collections
import re
import subprocess
import sys


def kernel_sass(obj, name):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout.splitlines()
    lines, take = [], False
    for l in out:
        if "Function :" in l:
            take = name in l
            continue
        if take:
            m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?)\s*;", l)
            if m:
                lines.append((int(m.group(1), 16), m.group(2)))
    if not lines:
        raise SystemExit("no kernel matching %r in %s" % (name, obj))
    return lines


def main():
    obj, name = sys.argv[1], sys.argv[2]
    lines = kernel_sass(obj, name)
    if len(sys.argv) < 5:
        print("%d instructions; backward branches (loops):" % len(lines))
        for a, ins in lines:
            m = re.search(r"BRA(?:\.U)? (?:!?U?P\d, )?0x([0-9a-f]+)", ins)
            if m and int(m.group(1), 16) < a:
                print("  %05x  %s" % (a, ins))
        return
    start, end = int(sys.argv[3], 16), int(sys.argv[4], 16)
    index = {a: i for i, (a, _) in enumerate(lines)}
    i, n, mix, hist = index[start], 0, collections.Counter(), []
    while True:
        a, ins = lines[i]
        n += 1
        op = ins.split()[1] if ins.startswith("@") else ins.split()[0]
        mix[op.split(".")[0]] += 1
        if a == end:
            break
        m = re.match(r"@!?U?P\d BRA(?:\.U)? 0x([0-9a-f]+)", ins)
        if m and int(m.group(1), 16) > a and any("VOTE" in h for h in hist[-3:]):
            i = index[int(m.group(1), 16)]
        else:
            i += 1
        hist.append(ins)
    print("%d instructions on the fast path %05x..%05x" % (n, start, end))
    print("  " + "  ".join("%s %d" % kv for kv in mix.most_common()))


if __name__ == "__main__":
    main()
