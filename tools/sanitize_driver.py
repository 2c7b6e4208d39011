#!/usr/bin/env python
"""Workload for tools/sanitize.sh: every kernel family of libpt_b200.so at sizes compute-sanitizer finishes in minutes.
  renders   64x64 x 2 spp, 8 bounces: few geoms (k_bounce first depth + k_bounce_q) with and without direct light
            sampling; the fused kernel at every depth (PT_B200_FUSED=1 is set by the script for a second pass);
            200 random geoms (k_bounce_bvh) with and without direct light sampling; mirror + glass + thin lens
  lists     pt_raygen, pt_intersect (filtered / exact scan, both scene sizes), sampling entry points
  compact   pt_compact_u32 in both modes at n = 1, 4097, 10^6
  shim      cudaRaytraceCore (reference signature), three iterations, with a device PBO
Prints one line per item; any CUDA error raises."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pt = importlib.import_module("project3-pathtracer_b200")
compat = importlib.import_module("project3-pathtracer_b200.compat")
import scenes_for_tests as S  # noqa: E402


def small_cam(cam, w, h):
    c = cam.copy()
    c["resolution"][0] = [w, h]
    return c


def main():
    g, m, cam = S._sample(pt)
    cam = small_cam(cam, 64, 64)
    # few geoms
    for nee in (False, True):
        with pt.Context(g, m, cam) as ctx:
            ctx.set_direct_lighting(nee)
            ctx.render(0, 2, 8, 7)
            img = ctx.download_mean(2)
            ctx.resolve_rgba8(2)
            print("render sample scene nee=%d: mean %.4f, counters %s" % (nee, float(img.mean()), ctx.counters()[:2]))
            o, d = ctx.raygen(7, np.arange(64, dtype=np.uint32), np.zeros(64, np.uint32))
            for mode in (pt.HIT_FILTERED, pt.HIT_EXACT_SCAN):
                gid = ctx.intersect(o, d, mode=mode)[0]
            print("  raygen + intersect: %d hits of 64" % int((gid >= 0).sum()))
    # mirror + glass + thin lens, one and many depths (first-and-last kernel variant included)
    m2 = m.copy()
    m2[3]["hasReflective"] = 1.0
    m2[4]["hasRefractive"] = 1.0
    m2[4]["indexOfRefraction"] = 1.5
    m2[4]["absorptionCoefficient"] = [0.1, 0.2, 0.3]
    for depth in (1, 2, 12):
        with pt.Context(g, m2, cam, lens=(0.2, 11.0)) as ctx:
            ctx.render(0, 2, depth, 3)
            print("render optics depth %d: mean %.4f" % (depth, float(ctx.download_mean(2).mean())))
    # many geoms: the hierarchy
    gb = S.random_scene(pt, 200, 5)
    gb["materialid"] = np.arange(200) % 9
    for nee in (False, True):
        with pt.Context(gb, m, cam) as ctx:
            ctx.set_direct_lighting(nee)
            ctx.render(0, 2, 8, 11)
            print("render 200 geoms nee=%d: mean %.4f, fallbacks %d" % (nee, float(ctx.download_mean(2).mean()), ctx.filter_stats()))
            o, d = ctx.raygen(11, np.arange(64, dtype=np.uint32), np.zeros(64, np.uint32))
            ctx.intersect(o, d)
    # sampling entry points
    pt.random_points_on_geom(g[8], np.arange(64, dtype=np.float32))
    pt.points_on_geom_u(g[5], np.random.default_rng(1).random((64, 3), dtype=np.float32))
    pt.random_directions_in_sphere(np.linspace(0, 0.99, 64, dtype=np.float32), np.linspace(0.99, 0, 64, dtype=np.float32))
    pt.calculate_transmission(np.ones((64, 3), np.float32), np.linspace(0, 5, 64, dtype=np.float32))
    pt.reference_stub_image(32, 32, 3)
    print("sampling entry points ok")
    # stream compaction
    rng = np.random.default_rng(2)
    for mode in (0, 1):
        pt.set_compact_mode(mode)
        for n in (1, 4097, 1000000):
            v = rng.integers(0, 2 ** 32, n, dtype=np.uint32)
            f = (rng.random(n) < 0.6).astype(np.uint8)
            out = pt.compact_u32(v, f)
            assert (out == v[f != 0]).all()
            print("compact mode %d n %d: kept %d" % (mode, n, len(out)))
    pt.set_compact_mode(0)
    # the shim
    rs = compat.RefScene([(g, cam)], m, iterations=3)
    compat.reset(); compat.set_trace_depth(8); compat.set_seed(3); compat.set_exit_on_error(False)
    for k in (1, 2, 3):
        compat.cudaRaytraceCore(None, rs.camera, 0, k, rs.materials, len(rs.materials), rs.geoms, len(rs.geoms))
        assert compat.last_status() == 0
    print("shim: 3 iterations, mean %.4f" % float(rs.image.mean()))
    compat.reset()


if __name__ == "__main__":
    main()
