#!/usr/bin/env python
"""Condense an .ncu-rep (read with `ncu -i ... --page raw --csv`) into one line per profiled launch.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--csv profiles/name.csv]
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "us"),
    ("dram__bytes_read.sum", "rdMB"),
    ("dram__bytes_write.sum", "wrMB"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("smsp__inst_executed.sum", "Minst"),
    ("launch__grid_size", "grid"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64c%"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    cols = [(hdr.index(m), m, short) for m, short in WANT if m in hdr]
    name_i = hdr.index("Kernel Name")
    table = [["kernel"] + [s for _, _, s in cols]]
    for r in rows[2:]:
        line = [r[name_i].replace("fcb200::", "").replace("<unnamed>::", "")[:60]]
        for i, m, short in cols:
            v = r[i].replace(",", "")
            try:
                f = float(v)
                u = units[i]
                if short in ("rdMB", "wrMB"):
                    f = f * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                if short == "us":
                    f = f * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                if short == "Minst":
                    f = f * 1e-6
                v = "%.1f" % f if short not in ("regs", "grid") else "%d" % f
            except ValueError:
                pass
            line.append(v)
        table.append(line)
    w = [max(len(r[c]) for r in table) for c in range(len(table[0]))]
    for r in table:
        print("  ".join(x.ljust(w[c]) if c == 0 else x.rjust(w[c]) for c, x in enumerate(r)))
    if "--traffic" in sys.argv:
        # dram__bytes_read.sum + dram__bytes_write.sum of the LAST profiled launch, for bench.py's roofline.traffic
        import json
        import os
        out_path = sys.argv[sys.argv.index("--traffic") + 1]
        rd_i, wr_i = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        r = rows[-1]
        total = float(r[rd_i].replace(",", "")) * scale.get(units[rd_i], 1.0) + float(r[wr_i].replace(",", "")) * scale.get(units[wr_i], 1.0)
        json.dump({"bytes_per_launch": total, "kernel": r[name_i][:120],
                   "source": "%s (ncu --set full --clock-control none, dram__bytes_read.sum + dram__bytes_write.sum of one launch)" % os.path.basename(rep)},
                  open(out_path, "w"), indent=1)
        print("traffic: %.1f MB per launch -> %s" % (total / 1e6, out_path))
    if "--csv" in sys.argv:
        with open(sys.argv[sys.argv.index("--csv") + 1], "w", newline="") as f:
            csv.writer(f).writerows(table)


if __name__ == "__main__":
    main()
