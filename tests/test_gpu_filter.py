"""The filtered closest hit (csrc/pt_filter.cuh) must return the exact scan's answer bit for bit -- at the shipped
error bounds and, as a margin check, with every rounding-error term of the bounds cut to a quarter."""
import os

import numpy as np
import pytest

from scenes_for_tests import all_scenes, ray_sets

pytestmark = pytest.mark.gpu

N = 300_000


def _same(got, want):
    gid, t, p, nr = got[:4]
    hit = want[0] >= 0
    return ((gid == want[0]).all() and (t.view(np.uint32) == want[1].view(np.uint32)).all()
            and (p[hit].view(np.uint32) == want[2][hit].view(np.uint32)).all()
            and (nr[hit].view(np.uint32) == want[3][hit].view(np.uint32)).all())


@pytest.mark.parametrize("scene", ["sample", "random64", "tiny_far", "aniso100", "offset1e4", "touching"])
def test_filtered_equals_exact_scan(pt, scene):
    g, m, cam = all_scenes(pt)[scene]
    with pt.Context(g, m, cam) as ctx:
        for rname, (o, d) in ray_sets(pt, ctx, g, N).items():
            want = ctx.intersect(o, d, mode=pt.HIT_EXACT_SCAN)
            for scale in (1.0, 0.25):
                ctx.set_filter_scale(scale)
                got = ctx.intersect(o, d, with_stats=True)
                assert _same(got, want), (scene, rname, scale)
            ctx.set_filter_scale(1.0)


def test_exact_scan_equals_oracle(pt, oracle, sample_scene):
    """the exact scan itself is the oracle's closest hit (so filtered == exact scan == oracle)"""
    g = sample_scene["geoms"]
    with pt.Context(g, sample_scene["materials"], sample_scene["camera"]) as ctx:
        o, d = ray_sets(pt, ctx, g, 100_000)["aimed"]
        gid, t, p, nr = ctx.intersect(o, d, mode=pt.HIT_EXACT_SCAN)
    wid, wt, wp, wn = oracle.intersect_rays(g, o, d)
    hit = wid >= 0
    assert (gid == wid).all() and (t.view(np.uint32) == wt.view(np.uint32)).all()
    assert (p[hit].view(np.uint32) == wp[hit].view(np.uint32)).all()
    assert (nr[hit].view(np.uint32) == wn[hit].view(np.uint32)).all()


def test_fallback_is_rare_on_the_sample_scene(pt, sample_scene):
    """the filter resolves almost every segment of a real render without the exact scan"""
    from conftest import with_resolution
    cam = with_resolution(sample_scene["camera"], 200, 200)
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as ctx:
        ctx.render(0, 8, 8, 7)
        _, segs, _ = ctx.counters()
        fb = ctx.filter_stats()
    assert segs > 0 and fb / segs < 0.01, (fb, segs)


@pytest.mark.parametrize("n_geoms,n_rays", [(200, 200_000), (10_000, 60_000)])
def test_hierarchy_equals_exact_scan(pt, sample_scene, n_geoms, n_rays):
    """scenes with many geoms go through the hierarchy (csrc/pt_bvh.cuh): same answers as the exact scan, bit for bit"""
    from scenes_for_tests import random_scene
    g = random_scene(pt, n_geoms, 21, extent=12.0, smin=0.05, smax=0.6)
    with pt.Context(g, sample_scene["materials"], sample_scene["camera"]) as ctx:
        for rname, (o, d) in ray_sets(pt, ctx, g, n_rays).items():
            want = ctx.intersect(o, d, mode=pt.HIT_EXACT_SCAN)
            for scale in (1.0, 0.25):
                ctx.set_filter_scale(scale)
                got = ctx.intersect(o, d, with_stats=True)
                assert _same(got, want), (n_geoms, rname, scale)
            ctx.set_filter_scale(1.0)


def test_hierarchy_render_matches_oracle(pt, oracle, sample_scene):
    """a whole render of a 600-geom scene (BASELINE config 3 in small): image and live counts identical to the oracle"""
    from conftest import with_resolution
    from scenes_for_tests import random_scene
    g = random_scene(pt, 600, 33, extent=6.0, smin=0.2, smax=0.9)
    rng = np.random.default_rng(5)
    g["materialid"] = rng.integers(0, len(sample_scene["materials"]), len(g))
    g["translation"] += np.float32(0)  # keep transforms as built
    cam = with_resolution(sample_scene["camera"], 64, 64)
    g = np.concatenate([g, sample_scene["geoms"]])  # walls and the light around the cloud
    scn = oracle.make_scene(g, sample_scene["materials"], cam)
    want_sum, want_live, _ = oracle.render(scn, 0, 2, 6, 9)
    with pt.Context(g, sample_scene["materials"], cam) as c:
        c.render(0, 2, 6, 9)
        got = c.download_sum()
        _, _, live = c.counters()
    assert live[:6].tolist() == want_live.tolist()
    assert (got.view(np.uint32) == want_sum.view(np.uint32)).all()


def test_procedural_10k_render_matches_oracle(pt, oracle, tmp_path):
    """BASELINE configs[3] itself -- the 10 000-object procedural scene file, through the product's loader -- on a
    96x54 crop of the frame (same view, 2 spp, 8 bounces): image and per-depth live counts identical to the oracle,
    which tests all 10 000 geoms exactly for every segment."""
    import subprocess
    import sys
    from conftest import ROOT, with_resolution
    path = tmp_path / "procedural_10000.txt"
    subprocess.check_call([sys.executable, os.path.join(ROOT, "scenes", "gen_scenes.py"), "--procedural", "10000", str(path)])
    sc = pt.Scene(str(path))
    g, m, cam, lens = sc.frame(0)
    assert len(g) == 10000
    cam = with_resolution(cam, 96, 54)
    scn = oracle.make_scene(g, m, cam)
    want_sum, want_live, _ = oracle.render(scn, 0, 2, 8, 565)
    with pt.Context(g, m, cam) as c:
        c.render(0, 2, 8, 565)
        got = c.download_sum()
        _, segs, live = c.counters()
    assert live[:8].tolist() == want_live.tolist()
    assert (got.view(np.uint32) == want_sum.view(np.uint32)).all()
    assert segs > 96 * 54 * 2 * 1.5  # the frame looks at the cloud: most paths bounce
