#!/usr/bin/env python
"""Markdown table of tools/opbench.py results for DESIGN.md section 6.

    python tools/opbench_table.py profiles/r01_opbench_f.json profiles/r01_opbench_f_masked.json [profiles/r01_opbench_c_all_operators.json]
"""
import json
import sys


def main():
    plain = json.load(open(sys.argv[1]))["results"]
    masked = {x["operator"]: x for x in json.load(open(sys.argv[2]))["results"]} if len(sys.argv) > 2 else {}
    start = {x["operator"]: x for x in json.load(open(sys.argv[3]))["results"]} if len(sys.argv) > 3 else {}
    print("| operator | grid × fields | B/pt | Gpt/s | frac | masked | r1 start |")
    print("|---|---|---|---|---|---|---|")
    for x in plain:
        g = "MEPS" if x["grid"][0] == 949 else "ECMWF"
        m, c = masked.get(x["operator"]), start.get(x["operator"])
        print("| %s | %s × %d | %d | %.1f | **%.2f** | %s | %s |" % (x["operator"], g, x["fields"], x["bytes_per_point"], x["gpts"], x["frac"],
                                                                    "%.2f" % m["frac"] if m else "—", "%.2f" % c["frac"] if c else "new"))


if __name__ == "__main__":
    main()
