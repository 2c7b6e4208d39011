#!/usr/bin/env python
"""The stable stream-compaction primitive (k_compact_u32: warp ballot -> block scan -> decoupled look-back) on its
own: device-timed throughput against the HBM roofline.  Algorithmic bytes per element: 4 (value) + 1 (flag) read,
4 written per kept element."""
import importlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pt = importlib.import_module("project3-pathtracer_b200")


def main():
    peak = 6552.3
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    rng = np.random.default_rng(1)
    for n in (1 << 24, 1 << 27, 1 << 28):
        v = rng.integers(0, 2**32, n, dtype=np.uint32)
        for keep in (0.1, 0.5, 0.7, 1.0):
            f = (rng.random(n) < keep).astype(np.uint8)
            for mode, name in ((0, "count/scan/scatter"), (1, "single pass")):
                pt.set_compact_mode(mode)
                out, ms = pt.compact_u32_timed(v, f, iters=10)
                assert (out == v[f != 0]).all()
                bytes_alg = 5.0 * n + 4.0 * len(out)
                print(json.dumps({"mode": name, "n": n, "keep": keep, "kept": int(len(out)), "kernel_ms": ms,
                                  "Gelem_per_s": n / ms / 1e6, "GB_per_s": bytes_alg / ms / 1e6,
                                  "frac_of_hbm_peak": bytes_alg / ms / 1e6 / peak}), flush=True)
    pt.set_compact_mode(0)


if __name__ == "__main__":
    main()
