"""Multi-GPU (needs >= 2 devices; skipped otherwise): sharding by sample index + one NCCL reduce gives the same image
as one GPU rendering all samples, up to float summation order; segment counts add up exactly."""
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import with_resolution

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ndev(pt):
    try:
        return pt.device_count()
    except Exception:
        return 0


def test_in_process_nccl_reduce_matches_single_gpu(pt, sample_scene):
    if _ndev(pt) < 2:
        pytest.skip("needs 2 GPUs")
    sh = __import__("importlib").import_module("project3-pathtracer_b200.sharding")
    cam = with_resolution(sample_scene["camera"], 200, 200)
    spp, depth, seed = 10, 8, 31
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam, device=0) as one:
        one.render(0, spp, depth, seed)
        want = one.download_sum()
        _, want_segs, want_live = one.counters()
    ctxs = [pt.Context(sample_scene["geoms"], sample_scene["materials"], cam, device=d) for d in range(2)]
    try:
        segs, live = 0, np.zeros(64, np.uint64)
        for r, c in enumerate(ctxs):
            b, n = sh.sample_range(r, 2, spp)
            c.render(b, n, depth, seed)
        for c in ctxs:
            c.sync()
            _, s, l = c.counters()
            segs += s
            live += l
        pt.reduce_to_first(ctxs)
        got = ctxs[0].download_sum()
    finally:
        for c in ctxs:
            c.close()
    assert segs == want_segs and (live == want_live).all()
    rmse = float(np.sqrt(np.mean((got - want) ** 2)))
    assert rmse <= 1e-6 * max(1.0, float(np.abs(want).max())) and np.allclose(got, want, rtol=1e-5, atol=1e-5)


def test_driver_gpus2_matches_gpus1(pt, tmp_path):
    if _ndev(pt) < 2:
        pytest.skip("needs 2 GPUs")
    exe = os.path.join(ROOT, "project3-pathtracer_b200", "pt_render")
    scene = os.path.join(ROOT, "scenes", "sample.txt")
    infos = []
    for g in (1, 2):
        out = subprocess.check_output([exe, "scene=" + scene, "spp=12", "depth=8", "seed=2", "gpus=%d" % g,
                                       "out=%s" % (tmp_path / ("g%d.png" % g)), "json=1"], text=True)
        infos.append(json.loads(out.strip().splitlines()[-1]))
    assert infos[0]["segments"] == infos[1]["segments"] and infos[0]["paths"] == infos[1]["paths"] == 800 * 800 * 12
    from PIL import Image
    a = np.asarray(Image.open(infos[0]["file"]).convert("RGB")).astype(int)
    b = np.asarray(Image.open(infos[1]["file"]).convert("RGB")).astype(int)
    assert np.abs(a - b).max() <= 1 and (a != b).mean() < 1e-3
