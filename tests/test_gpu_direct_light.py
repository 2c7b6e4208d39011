"""GPU parity: direct light sampling through the C ABI against the oracle.
Bars: per-depth live counts and the number of shadow rays identical; the image bit-identical where a pixel receives
at most two contributions (float addition is commutative), otherwise within 1e-5 relative of the pixel's sum (the
GPU adds contributions with atomics in arbitrary order)."""
import numpy as np
import pytest

from conftest import same_bits, with_resolution
from scenes_for_tests import build_geom, random_scene

pytestmark = pytest.mark.gpu

SUM_TOL = 1e-5


def lit_scene(pt, sample_scene):
    """the sample scene with its ceiling light replaced by a floating, rotated cube light and a sphere light"""
    g, m = sample_scene["geoms"].copy(), sample_scene["materials"].copy()
    light = int(g[8]["materialid"])
    g[8] = build_geom(pt, 1, light, (0, 8, 0), (20, 30, 40), (3, 0.5, 2))
    g[5] = build_geom(pt, 0, light, (-2, 5, 2), (0, 0, 0), (1.5, 1.5, 1.5))
    m[3]["hasReflective"] = 1.0  # sphere 6 becomes a mirror: specular events clear the no-emission flag
    m[3]["specularColor"] = [0.9, 0.9, 0.9]
    return g, m


def _render(pt, oracle, g, m, cam, spp, depth, seed, nee=True, first=0):
    want, want_live, _ = oracle.render(oracle.make_scene(g, m, cam, direct_lighting=nee), first, spp, depth, seed)
    want_shadow = oracle.last_shadow_rays
    with pt.Context(g, m, cam) as c:
        c.set_direct_lighting(nee)
        c.render(first, spp, depth, seed)
        got = c.download_sum()
        _, _, live = c.counters()
        shadow, n_lights = c.shadow_rays()
    assert live[:depth].tolist() == want_live.tolist()
    assert shadow == want_shadow
    return got, want, shadow, n_lights


def test_two_segment_paths_bit_identical(pt, oracle, sample_scene):
    g, m = lit_scene(pt, sample_scene)
    cam = with_resolution(sample_scene["camera"], 160, 160)
    got, want, shadow, n_lights = _render(pt, oracle, g, m, cam, 1, 2, 19)
    assert n_lights == 2 and shadow > 10000
    assert same_bits(got, want)


def test_deep_paths_match_oracle(pt, oracle, sample_scene):
    g, m = lit_scene(pt, sample_scene)
    cam = with_resolution(sample_scene["camera"], 128, 128)
    got, want, shadow, _ = _render(pt, oracle, g, m, cam, 3, 8, 23, first=4)
    assert shadow > 50000
    assert np.abs(got - want).max() <= SUM_TOL * max(1.0, float(np.abs(want).max()))
    assert np.allclose(got, want, rtol=SUM_TOL, atol=SUM_TOL)


def test_hierarchy_mode_matches_oracle(pt, oracle, sample_scene):
    """64 random geoms (k_bounce_bvh) around two lights"""
    m = sample_scene["materials"].copy()
    light = int(sample_scene["geoms"][8]["materialid"])
    g = random_scene(pt, 66, 7, extent=6.0, smin=0.2, smax=1.5)
    g["materialid"] = np.arange(66) % 5
    g[64] = build_geom(pt, 1, light, (0, 9, 0), (0, 0, 0), (4, 0.3, 4))
    g[65] = build_geom(pt, 0, light, (5, 4, 5), (0, 0, 0), (2, 2, 2))
    cam = with_resolution(sample_scene["camera"], 96, 96)
    got, want, shadow, n_lights = _render(pt, oracle, g, m, cam, 1, 2, 5)
    assert n_lights == 2 and shadow > 1000 and same_bits(got, want)
    got, want, _, _ = _render(pt, oracle, g, m, cam, 2, 6, 5)
    assert np.allclose(got, want, rtol=SUM_TOL, atol=SUM_TOL)


def test_off_by_default_and_no_lights(pt, oracle, sample_scene):
    g, m = lit_scene(pt, sample_scene)
    cam = with_resolution(sample_scene["camera"], 64, 64)
    with pt.Context(g, m, cam) as c:
        c.render(0, 2, 8, 3)
        base = c.download_sum()
        assert c.shadow_rays() == (0, 2)
        c.set_direct_lighting(True)
        c.clear()
        c.render(0, 2, 8, 3)
        assert c.shadow_rays()[0] > 0
        c.set_direct_lighting(False)
        c.clear()
        c.render(0, 2, 8, 3)
        assert same_bits(c.download_sum(), base) and c.shadow_rays()[0] == 0
    m2 = m.copy()
    m2["emittance"] = 0  # nothing emits: the switch must change nothing
    got, want, shadow, n_lights = _render(pt, oracle, g, m2, cam, 1, 4, 3)
    assert (shadow, n_lights) == (0, 0) and same_bits(got, want) and not got.any()


def test_same_expected_image_less_noise(pt, sample_scene):
    """the estimator is unbiased (same mean image as plain path tracing) and, with small lights that few sampled
    directions find by chance, converges much faster.

    Noise = difference of two independent estimates of the frame.  Its RMS is dominated by a few fireflies and swings
    between 0.4 and 0.7 of plain path tracing's from one sample count to the next (round 1 asserted < 0.6, measured 0.69
    and loosened the bound to fit); the robust statistic is the MEAN ABSOLUTE difference: the CPU oracle gives
    0.19 / 0.18 of plain path tracing's at 128 / 512 spp (same estimator, same RNG streams), the bar is 0.35."""
    g, m = lit_scene(pt, sample_scene)
    light = int(g[8]["materialid"])
    g[8] = build_geom(pt, 1, light, (0, 8, 0), (20, 30, 40), (0.6, 0.1, 0.4))
    g[5] = build_geom(pt, 0, light, (-2, 5, 2), (0, 0, 0), (0.3, 0.3, 0.3))
    m[light]["emittance"] = 400.0
    m[3]["hasReflective"] = 0.0  # all diffuse here (the oracle figures above are for this scene)
    cam = with_resolution(sample_scene["camera"], 64, 64)
    spp = 4096
    out = {}
    for nee in (False, True):
        with pt.Context(g, m, cam) as c:
            c.set_direct_lighting(nee)
            c.render(0, spp, 8, 1)
            a = c.download_mean(spp)
            c.clear()
            c.render(spp, spp, 8, 1)  # an independent second estimate: their difference measures the noise
            b = c.download_mean(spp)
        d = np.abs(a - b).sum(axis=1)
        out[nee] = (a, float(d.mean()), float(np.sqrt(np.mean((a - b) ** 2))))
    lum = float(out[False][0].mean())
    assert abs(float(out[True][0].mean()) - lum) < 0.02 * lum
    assert out[True][1] < 0.35 * out[False][1], (out[True][1:], out[False][1:])
    assert out[True][2] < out[False][2]  # and the firefly-dominated RMS is at least not worse


def test_shadow_queue_follows_scene_and_capacity_changes(pt, oracle, sample_scene):
    """The shadow-ray queues are sized like the wavefront and served by a different launch in few-geom and hierarchy
    scenes: one context that switches between the two kinds of scene and changes its wavefront capacity in between must
    keep producing what fresh contexts produce (two-segment paths: bit-identical to the oracle)."""
    cam = with_resolution(sample_scene["camera"], 96, 96)
    g_few, m = lit_scene(pt, sample_scene)
    light = int(sample_scene["geoms"][8]["materialid"])
    g_many = random_scene(pt, 66, 7, extent=6.0, smin=0.2, smax=1.5)
    g_many["materialid"] = np.arange(66) % 5
    g_many[64] = build_geom(pt, 1, light, (0, 9, 0), (0, 0, 0), (4, 0.3, 4))
    g_many[65] = build_geom(pt, 0, light, (5, 4, 5), (0, 0, 0), (2, 2, 2))
    want = {}
    for name, g in (("few", g_few), ("many", g_many)):
        want[name], _, _ = oracle.render(oracle.make_scene(g, m, cam, direct_lighting=True), 0, 1, 2, 11)
    with pt.Context(g_few, m, cam) as c:
        c.set_direct_lighting(True)
        for name, g, paths in (("few", g_few, 96 * 96), ("many", g_many, 96 * 96), ("many", g_many, 4 * 96 * 96),
                               ("few", g_few, 4 * 96 * 96), ("many", g_many, 96 * 96)):
            c.update_scene(g, m, cam)
            c.set_wavefront_paths(paths)
            c.clear()
            c.render(0, 1, 2, 11)
            assert same_bits(c.download_sum(), want[name]), (name, paths)
