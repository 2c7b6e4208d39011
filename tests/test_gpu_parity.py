"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars (BASELINE.json north_star): hit flag and geom id bit-exact; t and normal within 1e-5 relative; images within
a stated RMSE / mean-luminance bound.  Because kernels and oracle share one arithmetic contract (binary32, unfused,
same order) these tests assert the stronger property -- identical bits -- and state the tolerance they would
otherwise fall back to."""
import numpy as np
import pytest

from conftest import f32, optics_scene, same_bits, with_resolution

pytestmark = pytest.mark.gpu

REL_TOL = 1e-5  # north-star tolerance for t / normal (the tests below achieve 0)


@pytest.fixture(scope="module")
def ctx(pt, sample_scene):
    c = pt.Context(sample_scene["geoms"], sample_scene["materials"], sample_scene["camera"])
    yield c
    c.close()


def test_raygen_matches_oracle_bitwise(pt, oracle, sample_scene, ctx):
    rng = np.random.default_rng(3)
    pix = rng.integers(0, 800 * 800, 50000).astype(np.uint32)
    smp = rng.integers(0, 5000, 50000).astype(np.uint32)
    o, d = ctx.raygen(7, pix, smp)
    oo, od = oracle.raygen(sample_scene["camera"], (0.0, 0.0), 7, pix, smp)
    assert same_bits(o, oo) and same_bits(d, od)
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], sample_scene["camera"], lens=(0.25, 11.5)) as c2:
        o, d = c2.raygen(7, pix, smp)
    oo, od = oracle.raygen(sample_scene["camera"], (0.25, 11.5), 7, pix, smp)
    assert same_bits(o, oo) and same_bits(d, od)


def test_raygen_golden(spec_gold, ctx):
    g = spec_gold["raygen"]
    o, d = ctx.raygen(g["seed"], g["pixel"], g["sample"])
    assert same_bits(o.ravel(), f32(g["o"])) and same_bits(d.ravel(), f32(g["d"]))


def _check_hits(got, want):
    gid, t, p, n = got
    wid, wt, wp, wn = want
    assert (gid == wid).all(), "geom id / hit flag must be bit-exact"
    hit = wid >= 0
    assert np.allclose(t[hit], wt[hit], rtol=REL_TOL, atol=0)
    assert np.allclose(n[hit], wn[hit], rtol=REL_TOL, atol=REL_TOL)
    # stronger: identical bits
    assert same_bits(t, wt) and same_bits(p[hit], wp[hit]) and same_bits(n[hit], wn[hit])


def test_closest_hit_primary_rays(oracle, sample_scene, ctx):
    """every 3rd primary ray of the 800x800 frame"""
    pix = np.arange(0, 800 * 800, 3, dtype=np.uint32)
    o, d = oracle.raygen(sample_scene["camera"], (0.0, 0.0), 1, pix, np.zeros_like(pix))
    _check_hits(ctx.intersect(o, d), oracle.intersect_rays(sample_scene["geoms"], o, d))


def test_closest_hit_random_and_surface_rays(oracle, sample_scene, ctx):
    rng = np.random.default_rng(11)
    n = 200000
    o = rng.uniform([-6, -1, -6], [6, 11, 13], (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    want = oracle.intersect_rays(sample_scene["geoms"], o, d)
    _check_hits(ctx.intersect(o, d), want)
    # rays that start on surfaces (second-bounce like), including unnormalised directions
    hit = want[0] >= 0
    o2 = (want[2][hit] + want[3][hit] * np.float32(2e-4)).astype(np.float32)
    d2 = rng.normal(size=o2.shape).astype(np.float32)
    _check_hits(ctx.intersect(o2, d2), oracle.intersect_rays(sample_scene["geoms"], o2, d2))


def test_closest_hit_golden(spec_gold, ctx):
    g, r = spec_gold["closest_hit"], spec_gold["raygen"]
    gid, t, p, n = ctx.intersect(f32(r["o"]).reshape(-1, 3), f32(r["d"]).reshape(-1, 3))
    assert gid.tolist() == g["id"]
    assert same_bits(t, f32(g["t"]))
    hit = gid >= 0
    assert same_bits(p[hit].ravel(), f32(g["p"]).reshape(-1, 3)[hit].ravel())
    assert same_bits(n[hit].ravel(), f32(g["n"]).reshape(-1, 3)[hit].ravel())


def test_many_geoms_chunked(pt, oracle, sample_scene):
    """more geoms than one shared-memory chunk holds (1024): exercises the chunk loop, ids still exact"""
    rng = np.random.default_rng(5)
    n = 1500
    g = np.zeros(n, pt.GEOM_DTYPE)
    base = sample_scene["geoms"]
    for i in range(n):
        src = base[5 + (i % 4)] if i % 4 != 3 else base[8]
        g[i] = src
        tr = rng.uniform([-5, 0, -5], [5, 10, 5]).astype(np.float32)
        s = np.float32(rng.uniform(0.1, 0.5))
        fwd = np.eye(4, dtype=np.float32) * s
        fwd[3, 3] = 1
        fwd[:3, 3] = tr
        inv = np.eye(4, dtype=np.float32) / s
        inv[3, 3] = 1
        inv[:3, 3] = -tr / s
        g[i]["transform"] = fwd.ravel()
        g[i]["inverseTransform"] = inv.ravel()
        g[i]["type"] = i % 2
    o = rng.uniform([-6, -1, 6], [6, 11, 13], (20000, 3)).astype(np.float32)
    d = (rng.uniform([-5, 0, -5], [5, 10, 5], (20000, 3)) - o).astype(np.float32)
    with pt.Context(g, sample_scene["materials"], sample_scene["camera"]) as c:
        _check_hits(c.intersect(o, d), oracle.intersect_rays(g, o, d))


@pytest.mark.parametrize("spp,depth", [(1, 8), (1, 1), (3, 5)])
def test_image_and_counters_match_oracle(pt, oracle, sample_scene, spp, depth):
    cam = with_resolution(sample_scene["camera"], 160, 160)
    scn = oracle.make_scene(sample_scene["geoms"], sample_scene["materials"], cam)
    want_sum, want_live, _ = oracle.render(scn, 0, spp, depth, 42)
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as c:
        c.render(0, spp, depth, 42)
        got_sum = c.download_sum()
        got_mean = c.download_mean(spp)
        paths, segs, live = c.counters()
    assert paths == 160 * 160 * spp
    assert live[:depth].tolist() == want_live.tolist(), "per-depth live counts must be identical"
    assert segs == int(want_live.sum())
    want_mean = want_sum / np.float32(spp)
    rmse = float(np.sqrt(np.mean((got_mean - want_mean) ** 2)))
    lum = lambda im: float((im @ np.array([0.2126, 0.7152, 0.0722], np.float32)).mean())
    assert rmse <= 1e-6 and abs(lum(got_mean) - lum(want_mean)) <= 1e-6 * max(1.0, lum(want_mean))
    if spp <= 2:  # one or two float contributions per pixel: order cannot matter
        assert same_bits(got_sum, want_sum) and same_bits(got_mean, want_mean)
    else:
        assert np.allclose(got_sum, want_sum, rtol=1e-6, atol=1e-6)


# the kernels that can trace depths >= 1 of a few-geom wavefront: chosen per depth (default), always re-batched, always fused
KERNEL_PATHS = {"auto": (0,), "rebatched": (1,), "fused": (2,)}


@pytest.mark.parametrize("path", sorted(KERNEL_PATHS))
@pytest.mark.parametrize("scene", ["sample", "optics", "direct"])
def test_every_kernel_path_matches_oracle(pt, oracle, sample_scene, path, scene):
    """image bits and per-depth live counts of the oracle, whichever kernels trace the wavefront
    (pt_set_kernel_policy): diffuse sample scene, mirror + glass + thin lens, direct light sampling"""
    cam = with_resolution(sample_scene["camera"], 128, 96)
    g, m, lens, nee = sample_scene["geoms"], sample_scene["materials"], None, False
    if scene == "optics":
        g, m = optics_scene(pt, sample_scene)
        lens = (0.15, 11.0)
    nee = scene == "direct"
    spp, depth, seed = 2, 7, 77
    scn = oracle.make_scene(g, m, cam, lens=lens if lens else (0.0, 0.0), direct_lighting=nee)
    want_sum, want_live, _ = oracle.render(scn, 0, spp, depth, seed)
    with pt.Context(g, m, cam, lens=lens) as c:
        c.set_kernel_policy(*KERNEL_PATHS[path])
        c.set_direct_lighting(nee)
        l0 = c.launch_count()
        c.render(0, spp, depth, seed)
        got = c.download_sum()
        _, segs, live = c.counters()
        launches = c.launch_count() - l0
    assert live[:depth].tolist() == want_live.tolist()
    if nee:  # several contributions per pixel and sample: float summation order
        assert np.allclose(got, want_sum, rtol=2e-6, atol=2e-6)
    else:
        assert same_bits(got, want_sum)
    # one launch per depth, the count fold, the resolve of download_sum; with direct lighting a shadow launch behind every
    # depth but the last
    assert launches == depth + 2 + (depth - 1 if nee else 0), launches


def test_small_wavefronts_give_same_image(pt, sample_scene):
    """splitting the samples over many wavefronts / calls changes nothing but float summation order"""
    cam = with_resolution(sample_scene["camera"], 96, 96)
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as c:
        c.render(0, 6, 8, 9)
        a = c.download_sum()
        ca = c.counters()
        c.clear()
        c.set_wavefront_paths(96 * 96 * 2)
        c.render(0, 4, 8, 9)
        c.render(4, 2, 8, 9)
        b = c.download_sum()
        cb = c.counters()
    assert ca[0] == cb[0] and ca[1] == cb[1] and (ca[2] == cb[2]).all()
    assert np.allclose(a, b, rtol=1e-6, atol=1e-6)


def test_golden_render_fixture(pt, spec_gold, sample_scene):
    g = spec_gold["render_40x40_2spp_d8_seed11"]
    cam = sample_scene["camera"].copy()
    cam["resolution"][0] = [40, 40]
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as c:
        c.render(0, 2, 8, 11)
        s = c.download_sum()
        _, _, live = c.counters()
    assert live[:8].tolist() == g["live"]
    assert same_bits(s.ravel(), f32(g["sum_rgb"]))


def test_mirror_glass_and_dof_paths_match_oracle(pt, oracle, sample_scene):
    g, m = optics_scene(pt, sample_scene)
    cam = with_resolution(sample_scene["camera"], 128, 128)
    lens = (0.3, 9.0)
    scn = oracle.make_scene(g, m, cam, lens)
    want_sum, want_live, _ = oracle.render(scn, 5, 1, 12, 77)
    with pt.Context(g, m, cam, lens=lens) as c:
        c.render(5, 1, 12, 77)
        got = c.download_sum()
        _, _, live = c.counters()
    assert live[:12].tolist() == want_live.tolist()
    assert same_bits(got, want_sum)


def test_resolve_rgba8_semantics(pt, sample_scene):
    """sendImageToPBO (reference src/raytraceKernel.cu:58-89): min(mean*255, 255) truncated, alpha 0"""
    cam = with_resolution(sample_scene["camera"], 64, 64)
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as c:
        c.render(0, 4, 8, 1)
        mean = c.download_mean(4)
        px = c.resolve_rgba8(4)
    want = np.minimum(mean * np.float32(255.0), np.float32(255.0)).astype(np.uint8)
    assert (px[:, :3] == want).all() and (px[:, 3] == 0).all()


def test_upload_sum_roundtrip(pt, sample_scene):
    cam = with_resolution(sample_scene["camera"], 32, 32)
    rng = np.random.default_rng(0)
    img = rng.random((32 * 32, 3), dtype=np.float32)
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as c:
        c.upload_sum(img)
        assert same_bits(c.download_sum(), img)
        c.clear()
        assert not c.download_sum().any()


@pytest.mark.parametrize("n", [0, 1, 3, 4, 5, 31, 32, 33, 255, 256, 257, 511, 512, 513, 1023, 1024, 1025, 4095, 4096, 4097,
                               8191, 8193, 100003, 10_000_000])
@pytest.mark.parametrize("mode", [0, 1])  # 0: count / scan / scatter, 1: single pass with the look-back inside
def test_compaction_is_a_stable_partition(pt, n, mode):
    pt.set_compact_mode(mode)
    rng = np.random.default_rng(n)
    v = rng.integers(0, 2**32, n, dtype=np.uint64).astype(np.uint32)
    for density in (0.0, 0.5, 1.0, 0.03):
        f = (rng.random(n) < density).astype(np.uint8)
        f = f * rng.integers(1, 256, n).astype(np.uint8)  # any non-zero byte means "keep"
        out = pt.compact_u32(v, f)
        assert out.shape[0] == int((f != 0).sum())
        assert (out == v[f != 0]).all()
    if n > 8:  # unaligned views of the host arrays end up aligned on the device; ragged tails are covered by the sizes
        out = pt.compact_u32(v[1:], (v[1:] & 1).astype(np.uint8))
        assert (out == v[1:][(v[1:] & 1) != 0]).all()
    pt.set_compact_mode(0)


def test_single_guard_ieee_math_exhaustive(pt):
    """sqrt_ieee / rcp_ieee / inv_sqrt_ieee (csrc/pt_device.cuh) equal sqrtf, 1/x, 1/sqrtf(x) on all 2^32 inputs"""
    assert pt.selftest_math() == (0, 0, 0)


def test_full_size_config_properties(pt, oracle, sample_scene):
    """BASELINE configs[1] at FULL size (800x800, 5000 spp, 8 bounces) through size-independent properties:
    splitting the samples in two renders adds up (live counts exactly, image up to float summation order), the
    wavefront size does not matter, paths only ever die, and no pixel exceeds the light's radiance"""
    g, m, cam = sample_scene["geoms"], sample_scene["materials"], sample_scene["camera"]
    spp, depth, seed = 5000, 8, 565
    with pt.Context(g, m, cam) as ctx:
        ctx.set_wavefront_paths(800 * 800 * 50)
        ctx.render(0, spp, depth, seed)
        whole = ctx.download_sum()
        paths, segs, live = ctx.counters()
        assert paths == 800 * 800 * spp and segs == int(live[:depth].sum())
        assert all(live[i + 1] <= live[i] for i in range(depth - 1)) and live[depth] == 0
        ctx.clear()
        ctx.set_wavefront_paths(800 * 800 * 16)  # another wavefront size, and the samples in two calls
        ctx.render(0, 2000, depth, seed)
        _, _, live_a = ctx.counters()
        ctx.render(2000, 3000, depth, seed)
        parts = ctx.download_sum()
        _, _, live_ab = ctx.counters()
    assert live_ab[:depth].tolist() == live[:depth].tolist() and (live_a[:depth] < live[:depth]).all()
    assert np.allclose(parts, whole, rtol=2e-5, atol=1e-3)  # same samples, different atomic summation order
    mean = whole / np.float32(spp)
    assert mean.max() <= 15.0 * 1.0001 and mean.min() >= 0.0  # the only emitter has emittance 15, albedos <= 1
    # mean luminance against the CPU oracle on the same frame (8 of the 5000 samples per pixel: 5.1 M paths, whose
    # mean is within a few 1e-3 of the converged one)
    want8, _, _ = oracle.render(oracle.make_scene(g, m, cam), 0, 8, depth, seed)
    assert abs(float(mean.mean()) - float(want8.mean()) / 8.0) < 0.01 * float(mean.mean())
