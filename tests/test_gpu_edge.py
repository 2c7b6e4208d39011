"""Edge cases of the hot path through the C ABI: empty and degenerate inputs, class / pairing corner cases of the
closest-hit filter, the hierarchy threshold, axis-parallel and non-finite rays.  The yardstick is always the exact
scan (PT_HIT_EXACT_SCAN), which tests/test_gpu_filter.py ties to the CPU oracle."""
import numpy as np
import pytest

from conftest import with_resolution
from scenes_for_tests import build_geom, random_scene

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _same_as_exact(ctx, pt, o, d):
    want = ctx.intersect(o, d, mode=pt.HIT_EXACT_SCAN)
    got = ctx.intersect(o, d)
    hit = want[0] >= 0
    assert (got[0] == want[0]).all()
    assert (_bits(got[1]) == _bits(want[1])).all()
    assert (_bits(got[2][hit]) == _bits(want[2][hit])).all() and (_bits(got[3][hit]) == _bits(want[3][hit])).all()
    return want


def test_zero_rays_and_zero_samples(pt, sample_scene):
    cam = with_resolution(sample_scene["camera"], 32, 32)
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as ctx:
        gid, t, p, n = ctx.intersect(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
        assert gid.shape == (0,) and t.shape == (0,)
        ctx.render(0, 0, 8, 1)  # no samples: nothing happens
        paths, segs, live = ctx.counters()
        assert paths == 0 and segs == 0 and not ctx.download_sum().any()


def test_scene_without_geometry_renders_black(pt, sample_scene):
    """MESH objects have no geometry (src/scene.cpp:57-66): every path leaves at depth 0"""
    g = sample_scene["geoms"].copy()
    g["type"] = pt.MESH
    cam = with_resolution(sample_scene["camera"], 40, 24)
    with pt.Context(g, sample_scene["materials"], cam) as ctx:
        ctx.render(0, 3, 8, 5)
        paths, segs, live = ctx.counters()
        assert paths == 3 * 40 * 24 and segs == paths and live[1] == 0
        assert not ctx.download_sum().any()
        gid, t, _, _ = ctx.intersect(np.zeros((5, 3), np.float32), np.ones((5, 3), np.float32))
        assert (gid == -1).all() and (t == -1).all()


@pytest.mark.parametrize("n_geoms", [1, 2, 3, 31, 32, 33, 34])
def test_pairing_and_hierarchy_threshold(pt, sample_scene, n_geoms):
    """odd counts get a never-hit partner; 33 geoms is where the hierarchy takes over from the pair scan"""
    rng = np.random.default_rng(n_geoms)
    g = random_scene(pt, n_geoms, 100 + n_geoms, extent=4.0, smin=0.5, smax=2.0)
    o = rng.uniform(-8, 8, (50_000, 3)).astype(np.float32)
    d = rng.normal(size=(50_000, 3)).astype(np.float32)
    with pt.Context(g, sample_scene["materials"], sample_scene["camera"]) as ctx:
        want = _same_as_exact(ctx, pt, o, d)
    assert (want[0] >= 0).any()


def test_all_four_filter_classes_in_one_scene(pt, sample_scene):
    gs = [build_geom(pt, 0, 0, (0, 0, 0), (10, 20, 30), (2, 2, 2)),        # class 0: uniformly scaled sphere
          build_geom(pt, 0, 0, (3, 0, 0), (10, 20, 30), (2, 1, 0.5)),      # class 1: ellipsoid
          build_geom(pt, 1, 0, (0, 3, 0), (0, 90, 180), (2, 1, 0.5)),      # class 2: world-axis-aligned cube
          build_geom(pt, 1, 0, (0, 0, 3), (10, 20, 30), (2, 1, 0.5)),      # class 3: rotated cube
          build_geom(pt, 1, 0, (-3, 0, 0), (0, 0, 0), (1, 1, 1))]          # class 2 again (pairs with the other)
    g = np.zeros(len(gs), pt.GEOM_DTYPE)
    for i, x in enumerate(gs):
        g[i] = x
    rng = np.random.default_rng(2)
    o = rng.uniform(-7, 7, (200_000, 3)).astype(np.float32)
    d = (rng.uniform(-3, 3, (200_000, 3)) - o).astype(np.float32)
    with pt.Context(g, sample_scene["materials"], sample_scene["camera"]) as ctx:
        want = _same_as_exact(ctx, pt, o, d)
    assert set(np.unique(want[0])) == {-1, 0, 1, 2, 3, 4}


def test_axis_parallel_degenerate_and_nonfinite_rays(pt, sample_scene):
    """direction components that are exactly 0 (1/0 = inf, 0*inf = NaN in the slab tests), the zero direction,
    NaN / inf origins: the filter must stay conservative, i.e. still agree with the exact scan"""
    rng = np.random.default_rng(3)
    n = 60_000
    o = rng.uniform([-6, -1, -6], [6, 11, 13], (n, 3)).astype(np.float32)
    d = np.zeros((n, 3), np.float32)
    axis = rng.integers(0, 3, n)
    d[np.arange(n), axis] = rng.choice(np.float32([-1, 1]), n)          # exactly axis-parallel
    two = rng.random(n) < 0.3
    d[two, (axis[two] + 1) % 3] = rng.normal(size=two.sum()).astype(np.float32)  # one zero component
    o[::7] = np.round(o[::7])                                            # origins on integer planes (walls at 0, +-5, 10)
    d[::101] = 0                                                         # the zero direction
    o[5::997, 0] = np.nan
    o[6::997, 1] = np.inf
    d[7::997, 2] = -np.inf
    for g in (sample_scene["geoms"], random_scene(pt, 40, 8, extent=5.0, smin=0.5, smax=2.0)):
        with pt.Context(g, sample_scene["materials"], sample_scene["camera"]) as ctx:
            _same_as_exact(ctx, pt, o, d)


def test_tiny_wavefront_and_deep_paths(pt, oracle, sample_scene):
    """wavefront of one sample of the frame, 40 bounces: same image as the oracle"""
    cam = with_resolution(sample_scene["camera"], 48, 48)
    scn = oracle.make_scene(sample_scene["geoms"], sample_scene["materials"], cam)
    want, want_live, _ = oracle.render(scn, 0, 2, 40, 11)
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as ctx:
        ctx.set_wavefront_paths(1)
        ctx.render(0, 2, 40, 11)
        got = ctx.download_sum()
        _, _, live = ctx.counters()
    assert live[:40].tolist() == want_live.tolist()
    assert (_bits(got) == _bits(want)).all()


def test_update_scene_switches_between_pair_scan_and_hierarchy(pt, sample_scene):
    """pt_update_scene (next animation frame) may change the number of geoms across the hierarchy threshold"""
    cam = with_resolution(sample_scene["camera"], 64, 64)
    small = sample_scene["geoms"]
    big = np.concatenate([random_scene(pt, 80, 77, extent=4.0, smin=0.3, smax=1.0), small])
    big["materialid"][:80] = 1

    def fresh(g):
        with pt.Context(g, sample_scene["materials"], cam) as c:
            c.render(0, 2, 6, 13)
            return c.download_sum(), c.counters()[2][:6].tolist()

    want_small, want_big = fresh(small), fresh(big)
    with pt.Context(small, sample_scene["materials"], cam) as ctx:
        for g, want in ((big, want_big), (small, want_small), (big, want_big)):
            ctx.update_scene(g, sample_scene["materials"], cam)
            ctx.clear()
            ctx.render(0, 2, 6, 13)
            assert ctx.counters()[2][:6].tolist() == want[1]
            assert (_bits(ctx.download_sum()) == _bits(want[0])).all()


def test_banded_wavefronts_change_nothing(pt, oracle, sample_scene):
    """large frames are rendered band by band (pt_set_band_pixels; automatic above 48 MB of accumulation image): the image
    and the live counts are those of the unbanded render and of the oracle -- ragged last band, bands smaller than a
    wavefront, bands larger than a wavefront, two wavefronts in flight"""
    from conftest import same_bits, with_resolution
    cam = with_resolution(sample_scene["camera"], 128, 96)
    g, m = sample_scene["geoms"], sample_scene["materials"]
    want, want_live, _ = oracle.render(oracle.make_scene(g, m, cam), 3, 2, 8, 9)
    with pt.Context(g, m, cam) as c:
        for band, wf in ((0, 0), (5000, 0), (5000, 3000), (1, 4096), (12287, 12288 * 2), (128 * 96, 128 * 96)):
            c.set_band_pixels(band)
            if wf:
                c.set_wavefront_paths(max(wf, 128 * 96))  # (capacity is rounded to whole frames)
            c.clear()
            c.render(3, 2, 8, 9)
            _, _, live = c.counters()
            assert live[:8].tolist() == want_live.tolist(), (band, wf)
            assert same_bits(c.download_sum(), want), (band, wf)

