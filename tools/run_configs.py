#!/usr/bin/env python
"""Measure the BASELINE.json configs other than the headline on one GPU (device-timed, CUDA events on the render
stream) and write one JSON line per config.  Scene files come from scenes/gen_scenes.py (reference text format) and
go through the product's own loader (pt_scene_load).

usage: python tools/run_configs.py [--frac F] [--out profiles/r01_configs.jsonl] [--png-dir gpurun_out] [--direct]
  --frac   fraction of each config's sample count to render (1 = the config as stated)"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pt = importlib.import_module("project3-pathtracer_b200")

CONFIGS = [
    # name, scene file (generated if missing), spp, depth, wavefront spp
    ("config1_sample_800", "scenes/sample.txt", 5000, 8, 50),
    ("config2_cornell_glass_dof_1080p", "scenes/cornell_glass_dof.txt", 4096, 12, 8),
    ("config3_procedural_10k_1080p", "procedural:10000", 1024, 8, 32),
    ("config4_sample_4k", "scenes/sample_4k.txt", 16384, 8, 4),
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frac", type=float, default=1.0)
    ap.add_argument("--out", default=os.path.join(ROOT, "profiles", "r01_configs.jsonl"))
    ap.add_argument("--png-dir", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--only", default="")
    ap.add_argument("--wf-spp", type=int, default=0, help="samples per wavefront (0: the config's own)")
    ap.add_argument("--band", type=int, default=0, help="pixels per wavefront band (pt_set_band_pixels; 0 = automatic)")
    ap.add_argument("--direct", action="store_true", help="direct light sampling on (pt_set_direct_lighting)")
    ap.add_argument("--filter-scale", type=float, default=1.0, help="experiment: scale of the filter's rounding-error bounds")
    args = ap.parse_args()
    lines = []
    for name, path, spp_full, depth, wf_spp in CONFIGS:
        wf_spp = args.wf_spp or wf_spp
        if args.only and args.only not in name:
            continue
        if path.startswith("procedural:"):
            n = int(path.split(":")[1])
            tmp = os.path.join(tempfile.gettempdir(), "procedural_%d.txt" % n)
            subprocess.check_call([sys.executable, os.path.join(ROOT, "scenes", "gen_scenes.py"), "--procedural", str(n), tmp])
            path = tmp
        else:
            path = os.path.join(ROOT, path)
        t0 = time.perf_counter()
        sc = pt.Scene(path)
        g, m, cam, lens = sc.frame(0)
        t_load = time.perf_counter() - t0
        spp = max(1, int(round(spp_full * args.frac)))
        t0 = time.perf_counter()
        with pt.Context(g, m, cam, lens=lens if lens[0] > 0 else None) as ctx:
            t_ctx = time.perf_counter() - t0
            ctx.set_wavefront_paths(sc.width * sc.height * wf_spp)
            if args.filter_scale != 1.0:
                ctx.set_filter_scale(args.filter_scale)
            ctx.set_direct_lighting(args.direct)
            ctx.set_band_pixels(args.band)
            ctx.render(0, min(spp, wf_spp), depth, 565)  # warm-up
            ctx.sync()
            ctx.clear()
            ctx.render(0, spp, depth, 565)
            ms = ctx.last_render_ms()
            paths, segs, live = ctx.counters()
            fb = ctx.filter_stats()
            retries = ctx.filter_retries()
            shadow, n_lights = ctx.shadow_rays()
            img = ctx.download_mean(spp)
        out_png = pt.save_image(img, sc.width, sc.height, os.path.join(args.png_dir, name + ".png"), 0, True)
        line = {"config": name, "scene": os.path.basename(path), "geoms": int(sc.n_geoms), "resolution": [sc.width, sc.height],
                "spp": spp, "spp_of_config": spp_full, "depth": depth, "paths": int(paths), "segments": int(segs),
                "render_ms": ms, "Mseg_per_s": segs / ms / 1e3, "spp_per_s": spp / (ms * 1e-3),
                "live_per_depth": [int(x) for x in live[:depth]], "exact_scan_fallbacks": int(fb),
                "fallback_fraction": fb / max(1, segs), "retry_fraction": retries / max(1, segs), "scene_load_s": t_load, "context_create_s": t_ctx,
                "mean_luminance": float(img.mean()), "image": os.path.basename(out_png),
                "direct_lighting": bool(args.direct), "lights": int(n_lights), "shadow_rays": int(shadow),
                "Mrays_per_s": (segs + shadow) / ms / 1e3}
        print(json.dumps(line), flush=True)
        lines.append(line)
    with open(args.out, "w") as f:
        for ln in lines:
            f.write(json.dumps(ln) + "\n")


if __name__ == "__main__":
    main()
