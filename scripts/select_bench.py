"""Bandwidth of the region-selection kernel (csrc/select.cu) on device-resident positions.

Algorithmic bytes per particle: 3*sizeof(T) read + 1 written.  Kernel time from CUDA events on
the library's stream (pnb_last_kernel_ms(3)); best of `reps` launches after warm-up.  Usage:
    python scripts/select_bench.py [--reps 8]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import torch

from pynbody_b200.filt import geometry_selection as g


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=8)
    a = ap.parse_args()
    for dt, n in ((torch.float32, 1 << 26), (torch.float64, 1 << 25)):
        pos = torch.rand((n, 3), device="cuda", dtype=dt) - 0.5
        for region, prm, wrap in (("sphere", (0.45, -0.4, 0.3, 0.25), 1.0), ("cube", (0.2, 0.3, -0.1, 0.8, 0.7, 0.4), 1.0),
                                  ("sphere", (0.0, 0.0, 0.0, 0.3), -1.0)):
            ms = []
            for _ in range(a.reps + 2):
                out = g.selection(pos, region, prm, wrap)
                ms.append(g.last_selection_ms()[1])
            best = min(ms[2:])
            nbytes = n * (3 * pos.element_size() + 1)
            print(json.dumps({"kernel": "k_select", "dtype": str(dt).split(".")[1], "region": region, "wrap": wrap, "n": n,
                              "selected": int(out.sum()), "ms": round(best, 4), "GB_per_s": round(nbytes / best / 1e6, 1)}))


if __name__ == "__main__":
    main()
