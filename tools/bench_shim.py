#!/usr/bin/env python
"""The reference's own calling pattern through the drop-in symbol: one cudaRaytraceCore() call per sample
(src/main.cpp:88-116), host `renderCam->image` updated with the running mean on every call.  Prints the call rate and the
segments/s it amounts to, next to the batched pt_render path of the same frame."""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pt = importlib.import_module("project3-pathtracer_b200")
compat = importlib.import_module("project3-pathtracer_b200.compat")
from scenes_for_tests import _sample  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    g, m, cam = _sample(pt)
    rs = compat.RefScene([(g, cam)], m, iterations=iters + 8)  # camera.iterations as a scene file would set it (ITERATIONS)
    compat.reset(); compat.set_trace_depth(8); compat.set_seed(565); compat.set_exit_on_error(False)
    for k in range(1, 4):  # warm-up
        compat.cudaRaytraceCore(None, rs.camera, 0, k, rs.materials, len(rs.materials), rs.geoms, len(rs.geoms))
    compat.reset()
    t0 = time.perf_counter()
    for k in range(1, iters + 1):
        compat.cudaRaytraceCore(None, rs.camera, 0, k, rs.materials, len(rs.materials), rs.geoms, len(rs.geoms))
    dt = time.perf_counter() - t0
    with pt.Context(g, m, cam) as ctx:
        ctx.render(0, iters, 8, 565)
        ms = ctx.last_render_ms()
        _, segs, _ = ctx.counters()
        mean = ctx.download_mean(iters)
    err = float(np.abs(mean - rs.image).max())
    print(json.dumps({"calls": iters, "seconds": dt, "calls_per_s": iters / dt, "ms_per_call": 1e3 * dt / iters,
                      "Mseg_per_s_through_the_shim": segs / dt / 1e6, "Mseg_per_s_batched": segs / ms / 1e3,
                      "max_abs_diff_running_mean_vs_batched_mean": err,
                      "note": "per call: scene flatten + upload check, 1 spp render (640 k paths), D2H of 7.7 MB, host running mean"}))
    compat.reset()


if __name__ == "__main__":
    main()
