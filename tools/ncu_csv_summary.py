#!/usr/bin/env python
"""One line per profiled launch from `ncu --page raw --csv` files (the compact form tools/ncu_masked.sh brings back).

    python tools/ncu_csv_summary.py gpurun_out/prof_*_raw.csv [--csv profiles/name.csv]
"""
import csv
import sys

WANT = [("gpu__time_duration.sum", "us"), ("dram__bytes_read.sum", "rdMB"), ("dram__bytes_write.sum", "wrMB"),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"), ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%"),
        ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lsuwf%"), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"), ("launch__registers_per_thread", "regs"), ("smsp__inst_executed.sum", "Minst"),
        ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"), ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%"),
        ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu%"), ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%"),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%"), ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%"),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"), ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"), ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
        ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "st_br"), ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "st_noinst"),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"), ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
        ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio")]


def main():
    files = [a for a in sys.argv[1:] if not a.startswith("--") and a.endswith(".csv") and (sys.argv.index(a) == 0 or sys.argv[sys.argv.index(a) - 1] != "--csv")]
    table = [["file", "kernel"] + [s for _, s in WANT]]
    for f in files:
        rows = list(csv.reader(open(f)))
        if len(rows) < 3:
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            line = [f.split("/")[-1].replace("_raw.csv", "").replace("prof_", ""), r[hdr.index("Kernel Name")].replace("fcb200::", "").replace("<unnamed>::", "")[:44]]
            for m, short in WANT:
                if m not in hdr:
                    line.append("-")
                    continue
                i = hdr.index(m)
                v = r[i].replace(",", "")
                try:
                    x = float(v)
                    u = units[i]
                    if short in ("rdMB", "wrMB"):
                        x *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1.0)
                    if short == "us":
                        x *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
                    if short == "Minst":
                        x *= 1e-6
                    v = "%d" % x if short == "regs" else "%.2f" % x if short.startswith("st_") else "%.1f" % x
                except ValueError:
                    pass
                line.append(v)
            table.append(line)
    w = [max(len(r[c]) for r in table) for c in range(len(table[0]))]
    for r in table:
        print("  ".join(x.ljust(w[c]) if c < 2 else x.rjust(w[c]) for c, x in enumerate(r)))
    if "--csv" in sys.argv:
        with open(sys.argv[sys.argv.index("--csv") + 1], "w", newline="") as fh:
            csv.writer(fh).writerows(table)


if __name__ == "__main__":
    main()
