#!/usr/bin/env python
"""Randomised parity campaign: for many random scenes (1..300 geoms, random sizes, anisotropy, rotations, offsets) and
four ray populations each, the filtered closest hit (pair scan or hierarchy) must equal the exact scan bit for bit --
geom id, distance, point, normal.  Prints one line per scene and a summary; exits non-zero on the first mismatch.

usage: python tools/stress_parity.py [n_scenes] [rays_per_population] [first_seed]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
pt = importlib.import_module("project3-pathtracer_b200")
from scenes_for_tests import _sample, random_scene, ray_sets  # noqa: E402


def main():
    n_scenes = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    n_rays = int(sys.argv[2]) if len(sys.argv) > 2 else 200_000
    seed0 = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    _, m, cam = _sample(pt)
    total = fallbacks = 0
    for seed in range(seed0, seed0 + n_scenes):
        rng = np.random.default_rng(1000 + seed)
        n = int(rng.choice([1, 2, 3, 5, 9, 17, 32, 33, 64, 150, 300]))
        extent = float(np.exp(rng.uniform(np.log(2.0), np.log(300.0))))
        smin = float(np.exp(rng.uniform(np.log(0.01), np.log(1.0))))
        smax = smin * float(np.exp(rng.uniform(0.0, np.log(50.0))))
        aniso = float(rng.choice([1.0, 1.0, 3.0, 30.0]))
        offset = rng.choice([0.0, 0.0, 1e3, 1e4]) * rng.normal(size=3)
        g = random_scene(pt, n, seed, extent=extent, smin=smin, smax=smax, aniso=aniso, offset=offset)
        with pt.Context(g, m, cam) as ctx:
            fb_scene = 0
            for rname, (o, d) in ray_sets(pt, ctx, g, n_rays).items():
                want = ctx.intersect(o, d, mode=pt.HIT_EXACT_SCAN)
                gid, t, p, nr, fb = ctx.intersect(o, d, with_stats=True)
                hit = want[0] >= 0
                ok = ((gid == want[0]).all() and (t.view(np.uint32) == want[1].view(np.uint32)).all()
                      and (p[hit].view(np.uint32) == want[2][hit].view(np.uint32)).all()
                      and (nr[hit].view(np.uint32) == want[3][hit].view(np.uint32)).all())
                if not ok:
                    print("MISMATCH seed %d population %s (n=%d extent=%g smin=%g smax=%g aniso=%g)" % (seed, rname, n, extent, smin, smax, aniso))
                    sys.exit(1)
                total += len(gid)
                fb_scene += fb
            fallbacks += fb_scene
        print("seed %3d  geoms %3d  extent %7.1f  scale %.3g..%.3g  aniso %4.0f  offset %8.0f  fallback %.3f%%  ok"
              % (seed, n, extent, smin, smax, aniso, float(np.abs(offset).max()), 100.0 * fb_scene / (4 * n_rays)), flush=True)
    print("all %d rays over %d scenes identical to the exact scan; fallback rate %.3f%%" % (total, n_scenes, 100.0 * fallbacks / max(total, 1)))


if __name__ == "__main__":
    main()
