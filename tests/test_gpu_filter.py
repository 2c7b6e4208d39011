"""The filtered closest hit (csrc/pt_filter.cuh) must return the exact scan's answer bit for bit -- at the shipped
error bounds and, as a margin check, with every rounding-error term of the bounds cut to a quarter."""
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
