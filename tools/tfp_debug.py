#!/usr/bin/env python
"""Development aid: where does thermalFrontParameter differ from the arbiter?  (run on a B200 via gpurun)"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import cases
    import fclibs
    gpu = importlib.import_module("mi-fieldcalc_b200").load()
    arb = fclibs.reference() or fclibs.oracle()
    name = sys.argv[1] if len(sys.argv) > 1 else "thermalFrontParameter"
    for nx, ny in ((949, 37), (300, 40), (3600, 64)):
        for mask, flag in (("none", cases.ALL), ("bernoulli", cases.SOME)):
            case = cases.build(name, nx, ny, seed=17, flag_in=flag, mask=mask)
            got = cases.run(gpu, case, to_device=lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda())
            want = cases.run(arb, case)
            g, w = got[1][0], want[1][0]
            same = (g.view(np.uint32) == w.view(np.uint32)) | (np.isnan(g) & np.isnan(w)) | (g == w)
            bad = np.argwhere(~same)
            print("%s %dx%d mask=%s: %d differ; flags %r %r" % (name, nx, ny, mask, len(bad), got[2], want[2]))
            if len(bad):
                ys, cy = np.unique(bad[:, 0], return_counts=True)
                print("  rows:", dict(zip(ys.tolist(), cy.tolist())))
                xs, cx = np.unique((bad[:, 1] - 1) % 60, return_counts=True)
                print("  (x-1)%60:", dict(zip(xs.tolist(), cx.tolist())))
                with np.errstate(all="ignore"):
                    rel = np.abs(g[~same].astype(np.float64) - w[~same]) / np.abs(w[~same])
                print("  rel err: min %.3g median %.3g max %.3g" % (rel.min(), np.median(rel), rel.max()))
                for y, x in bad[:8]:
                    print("   (y=%d,x=%d) got %r want %r" % (y, x, g[y, x], w[y, x]))


if __name__ == "__main__":
    main()
