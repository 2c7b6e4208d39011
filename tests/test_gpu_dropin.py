"""GPU tests of the drop-in layers around the kernels: the cudaRaytraceCore shim (reference
src/raytraceKernel.h:17), the headless driver (src/main.cpp runCuda loop) and scene-file -> image end to end."""
import importlib
import json
import os
import subprocess

import numpy as np
import pytest
from PIL import Image

from conftest import same_bits, with_resolution

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def compat():
    return importlib.import_module("project3-pathtracer_b200.compat")


def test_cudaRaytraceCore_running_mean_matches_oracle(pt, compat, oracle, sample_scene):
    """k calls with iterations = 1..k leave the running mean of samples 0..k-1 in renderCam->image"""
    cam = with_resolution(sample_scene["camera"], 96, 96)
    rs = compat.RefScene([(sample_scene["geoms"], cam)], sample_scene["materials"], iterations=5)
    compat.reset(); compat.set_trace_depth(8); compat.set_seed(3); compat.set_exit_on_error(False)
    scn = oracle.make_scene(sample_scene["geoms"], sample_scene["materials"], cam)
    want_sum = np.zeros((96 * 96, 3), np.float32)
    for k in range(1, 6):
        compat.cudaRaytraceCore(None, rs.camera, 0, k, rs.materials, len(rs.materials), rs.geoms, len(rs.geoms))
        assert compat.last_status() == 0
        oracle.render(scn, k - 1, 1, 8, 3, sum_rgb=want_sum)
        want = want_sum / np.float32(k)
        if k <= 2:
            assert same_bits(rs.image, want)
        assert np.allclose(rs.image, want, rtol=1e-6, atol=1e-6)
    # out-of-sequence call: resumes from the caller's running mean (image*(k-1) + L_k)/k
    img4 = rs.image.copy()
    compat.reset()
    rs.image[:] = img4
    compat.cudaRaytraceCore(None, rs.camera, 0, 6, rs.materials, len(rs.materials), rs.geoms, len(rs.geoms))
    oracle.render(scn, 5, 1, 8, 3, sum_rgb=want_sum)
    assert np.allclose(rs.image, want_sum / np.float32(6), rtol=1e-5, atol=1e-5)
    # iterations == 1 restarts the image (src/main.cpp:147-157 zeroes it between frames)
    compat.cudaRaytraceCore(None, rs.camera, 0, 1, rs.materials, len(rs.materials), rs.geoms, len(rs.geoms))
    first, _, _ = oracle.render(scn, 0, 1, 8, 3)
    assert same_bits(rs.image, first)
    compat.reset()


def test_cudaRaytraceCore_sample_traced_ahead_is_dropped_when_the_sequence_changes(pt, compat, oracle, sample_scene):
    """While a call's image travels to the host the shim already traces the next iteration's sample.  If the next call is
    not that iteration with those settings, the sample traced ahead must not count."""
    cam = with_resolution(sample_scene["camera"], 64, 64)
    rs = compat.RefScene([(sample_scene["geoms"], cam)], sample_scene["materials"], iterations=4)
    compat.reset(); compat.set_trace_depth(8); compat.set_seed(3); compat.set_exit_on_error(False)
    scn = oracle.make_scene(sample_scene["geoms"], sample_scene["materials"], cam)
    call = lambda k: compat.cudaRaytraceCore(None, rs.camera, 0, k, rs.materials, len(rs.materials), rs.geoms, len(rs.geoms))
    want_sum = np.zeros((64 * 64, 3), np.float32)
    for k in (1, 2):
        call(k)
        oracle.render(scn, k - 1, 1, 8, 3, sum_rgb=want_sum)
    assert same_bits(rs.image, want_sum / np.float32(2))
    # the seed changes: sample 2 was traced ahead with seed 3 and is dropped; the sum restarts from the caller's mean
    compat.set_seed(4)
    call(3)
    oracle.render(scn, 2, 1, 8, 4, sum_rgb=want_sum)
    assert np.allclose(rs.image, want_sum / np.float32(3), rtol=1e-6, atol=1e-6)
    # in sequence again (sample 3 traced ahead with seed 4 and used); iteration 4 = camera.iterations: nothing is traced ahead
    call(4)
    oracle.render(scn, 3, 1, 8, 4, sum_rgb=want_sum)
    assert np.allclose(rs.image, want_sum / np.float32(4), rtol=1e-6, atol=1e-6)
    # past the scene's iteration count the calls still work (no sample traced ahead)
    call(5)
    oracle.render(scn, 4, 1, 8, 4, sum_rgb=want_sum)
    assert np.allclose(rs.image, want_sum / np.float32(5), rtol=1e-6, atol=1e-6)
    # a jump back: iteration 2 again restarts from the caller's image, which the test resets to the 1-sample mean
    first, _, _ = oracle.render(scn, 0, 1, 8, 4)
    rs.image[:] = first
    call(2)
    second = first.copy()
    oracle.render(scn, 1, 1, 8, 4, sum_rgb=second)
    assert np.allclose(rs.image, second / np.float32(2), rtol=1e-6, atol=1e-6)
    assert compat.last_status() == 0
    compat.reset()


@pytest.mark.parametrize("group", [1, 3, 8])
def test_sample_stream_equals_one_render_per_sample(pt, sample_scene, group):
    """pt_stream_*: samples traced ahead in groups and folded in one per call leave exactly the sums that one pt_render per
    sample leaves (one radiance contribution per pixel and sample: the same float adds in the same order)"""
    cam = with_resolution(sample_scene["camera"], 96, 64)
    depth, seed, n = 8, 12, 19  # 19 samples: the groups wrap around several times, the last one stays partly unused
    with pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as a, \
         pt.Context(sample_scene["geoms"], sample_scene["materials"], cam) as b:
        a.render(0, 2, depth, seed)  # the stream starts on top of an existing sum
        b.render(0, 2, depth, seed)
        b.stream_begin(2, 2, depth, seed, group)
        for k in range(2, n):
            a.render(k, 1, depth, seed)
            mean, spp = b.stream_next()
            assert spp == k + 1 and same_bits(mean, a.download_mean(k + 1)), k
        # anything else that touches the sum ends the stream: the samples traced ahead are dropped, the sum is untouched
        assert same_bits(b.download_sum(), a.download_sum())
        b.render(n, 1, depth, seed)
        a.render(n, 1, depth, seed)
        assert same_bits(b.download_sum(), a.download_sum())
        with pytest.raises(pt.PtError):
            b.stream_next()  # no stream open any more


def test_sample_stream_with_direct_lighting_and_scene_change(pt, oracle, sample_scene):
    """a stream under direct light sampling (several contributions per pixel and sample: float summation order differs,
    the sums agree to 1e-5), then a scene update in the middle of a group: the samples traced ahead with the old scene are
    dropped, the samples handed out so far stay in the sum, and a new stream continues with the new scene"""
    cam = with_resolution(sample_scene["camera"], 64, 48)
    g, m = sample_scene["geoms"], sample_scene["materials"]
    depth, seed = 6, 21
    with pt.Context(g, m, cam) as a, pt.Context(g, m, cam) as b:
        for c in (a, b):
            c.set_direct_lighting(True)
        b.stream_begin(0, 0, depth, seed, 4)
        for k in range(6):  # one and a half groups
            a.render(k, 1, depth, seed)
            mean, spp = b.stream_next()
            want = a.download_mean(k + 1)
            assert spp == k + 1 and np.allclose(mean, want, rtol=1e-5, atol=1e-5), k
        g2 = g.copy()
        g2[5]["materialid"] = 1  # the big sphere turns red
        for c in (a, b):
            c.update_scene(g2, m, cam)
        assert np.allclose(b.download_sum(), a.download_sum(), rtol=1e-5, atol=1e-5)  # six samples of the old scene, no more
        b.stream_begin(6, 6, depth, seed, 4)
        for k in range(6, 9):
            a.render(k, 1, depth, seed)
            mean, spp = b.stream_next()
            assert spp == k + 1 and np.allclose(mean, a.download_mean(k + 1), rtol=1e-5, atol=1e-5), k
        b.stream_end()
        assert np.allclose(b.download_sum(), a.download_sum(), rtol=1e-5, atol=1e-5)
        # the oracle agrees with the mixed sequence: 6 samples of the first scene + 3 of the second
        want = np.zeros((64 * 48, 3), np.float32)
        oracle.render(oracle.make_scene(g, m, cam, direct_lighting=True), 0, 6, depth, seed, sum_rgb=want)
        oracle.render(oracle.make_scene(g2, m, cam, direct_lighting=True), 6, 3, depth, seed, sum_rgb=want)
        assert np.allclose(a.download_sum(), want, rtol=1e-5, atol=1e-5)


def test_cudaRaytraceCore_frame_selects_per_frame_arrays_and_writes_pbo(pt, compat, oracle, sample_scene):
    import torch
    cam = with_resolution(sample_scene["camera"], 64, 64)
    g1 = sample_scene["geoms"].copy()
    # frame 1: sphere 5 moved up by 2 (forward/inverse translation columns)
    g1[5]["translation"][1] += 2
    g1[5]["transform"][7] += 2
    inv = np.linalg.inv(g1[5]["transform"].reshape(4, 4).astype(np.float64)).astype(np.float32)
    g1[5]["inverseTransform"] = inv.ravel()
    rs = compat.RefScene([(sample_scene["geoms"], cam), (g1, cam)], sample_scene["materials"])
    compat.reset(); compat.set_trace_depth(4); compat.set_seed(9); compat.set_exit_on_error(False)
    pbo = torch.zeros(64 * 64 * 4, dtype=torch.uint8, device="cuda")
    compat.cudaRaytraceCore(pbo.data_ptr(), rs.camera, 1, 1, rs.materials, len(rs.materials), rs.geoms, len(rs.geoms))
    want, _, _ = oracle.render(oracle.make_scene(g1, sample_scene["materials"], cam), 0, 1, 4, 9)
    assert same_bits(rs.image, want)
    px = pbo.cpu().numpy().reshape(-1, 4)
    assert (px[:, :3] == np.minimum(want * np.float32(255), np.float32(255)).astype(np.uint8)).all() and (px[:, 3] == 0).all()
    compat.reset()


def test_headless_driver_scene_file_to_png(pt, oracle, tmp_path):
    """pt_render scene=... -> PNG; pixels equal the oracle's image pushed through the reference's 8-bit rules"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scenes"))
    import gen_scenes
    text = gen_scenes.sample_scene((120, 80), 3)
    p = tmp_path / "small.txt"
    p.write_text(text)
    out = subprocess.check_output([os.path.join(ROOT, "project3-pathtracer_b200", "pt_render"), "scene=%s" % p,
                                   "depth=6", "seed=21", "out=%s" % (tmp_path / "o.bmp"), "json=1"], text=True)
    info = json.loads(out.strip().splitlines()[-1])
    assert info["file"].endswith("o.0.png") and info["spp"] == 3 and info["paths"] == 120 * 80 * 3
    s = pt.Scene(p)
    g, m, cam, lens = s.frame(0)
    want_sum, live, _ = oracle.render(oracle.make_scene(g, m, cam, lens), 0, 3, 6, 21)
    assert info["segments"] == int(live.sum())
    want8 = pt.image_to_rgb8(want_sum / np.float32(3), 120, 80)
    got8 = np.asarray(Image.open(info["file"]).convert("RGB"))
    assert np.abs(got8.astype(int) - want8.astype(int)).max() <= 1  # 3 samples: summation order may move one LSB
    assert (got8 != want8).mean() < 1e-3


def test_cornell_scene_file_end_to_end(pt, oracle):
    """config 2 (glass + mirror + depth of field) from its scene file, reduced frame, bit-exact at 1 spp"""
    s = pt.Scene(os.path.join(ROOT, "scenes", "cornell_glass_dof.txt"))
    g, m, cam, lens = s.frame(0)
    assert lens[0] > 0
    cam = with_resolution(cam, 192, 108)
    want, live, _ = oracle.render(oracle.make_scene(g, m, cam, lens), 0, 1, 12, 4)
    with pt.Context(g, m, cam, lens=lens) as c:
        c.render(0, 1, 12, 4)
        got = c.download_sum()
        _, segs, glive = c.counters()
    assert glive[:12].tolist() == live.tolist() and same_bits(got, want)
    assert live[11] > 0.05 * live[0], "closed box: paths stay alive"


def test_headless_driver_renders_every_frame_of_an_animation(pt, oracle, tmp_path):
    """no frame= token: the driver walks all frames (src/main.cpp:147-157), re-uploading the per-frame transforms and
    camera and clearing the image in between; every saved frame equals the oracle's render of that frame"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scenes"))
    import gen_scenes
    p = tmp_path / "anim.txt"
    p.write_text(gen_scenes.sample_animated(3, (96, 64), 2))
    out = subprocess.check_output([os.path.join(ROOT, "project3-pathtracer_b200", "pt_render"), "scene=%s" % p,
                                   "depth=5", "seed=3", "out=%s" % (tmp_path / "a.png"), "json=1"], text=True)
    infos = [json.loads(x) for x in out.strip().splitlines()]
    assert [i["frame"] for i in infos] == [0, 1, 2]
    s = pt.Scene(p)
    assert s.n_frames == 3
    imgs = []
    for k, info in enumerate(infos):
        assert info["file"].endswith("a.%d.png" % k)
        g, m, cam, lens = s.frame(k)
        want_sum, live, _ = oracle.render(oracle.make_scene(g, m, cam, lens), 0, 2, 5, 3)
        assert info["segments"] == int(live.sum())
        want8 = pt.image_to_rgb8(want_sum / np.float32(2), 96, 64)
        got8 = np.asarray(Image.open(info["file"]).convert("RGB"))
        assert (got8 == want8).all()  # two samples per pixel: the sum is order-independent, so bits match
        imgs.append(got8)
    assert (imgs[0] != imgs[1]).any() and (imgs[1] != imgs[2]).any()


def test_headless_driver_direct_lighting_switch(pt, oracle, tmp_path):
    """pt_render ... direct=1: same segments as the oracle with direct light sampling, pixels within one LSB"""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "scenes"))
    import gen_scenes
    p = tmp_path / "small.txt"
    p.write_text(gen_scenes.sample_scene((96, 64), 2))
    out = subprocess.check_output([os.path.join(ROOT, "project3-pathtracer_b200", "pt_render"), "scene=%s" % p,
                                   "depth=4", "seed=8", "direct=1", "out=%s" % (tmp_path / "d.png"), "json=1"], text=True)
    info = json.loads(out.strip().splitlines()[-1])
    g, m, cam, lens = pt.Scene(p).frame(0)
    want_sum, live, _ = oracle.render(oracle.make_scene(g, m, cam, lens, direct_lighting=True), 0, 2, 4, 8)
    assert oracle.last_shadow_rays > 0 and info["segments"] == int(live.sum())
    want8 = pt.image_to_rgb8(want_sum / np.float32(2), 96, 64)
    got8 = np.asarray(Image.open(info["file"]).convert("RGB"))
    assert np.abs(got8.astype(int) - want8.astype(int)).max() <= 1 and (got8 != want8).mean() < 1e-2
    plain, _, _ = oracle.render(oracle.make_scene(g, m, cam, lens), 0, 2, 4, 8)
    assert (pt.image_to_rgb8(plain / np.float32(2), 96, 64) != want8).mean() > 0.2  # and it is a different estimator


def test_cudaRaytraceCore_with_direct_lighting(pt, compat, oracle, sample_scene):
    """pt_compat_set_direct_lighting(1): the reference entry point renders with direct light sampling"""
    cam = with_resolution(sample_scene["camera"], 64, 64)
    rs = compat.RefScene([(sample_scene["geoms"], cam)], sample_scene["materials"], iterations=2)
    compat.reset(); compat.set_trace_depth(3); compat.set_seed(5); compat.set_exit_on_error(False)
    compat.set_direct_lighting(True)
    try:
        compat.cudaRaytraceCore(None, rs.camera, 0, 1, rs.materials, len(rs.materials), rs.geoms, len(rs.geoms))
        assert compat.last_status() == 0
        scn = oracle.make_scene(sample_scene["geoms"], sample_scene["materials"], cam, direct_lighting=True)
        want, _, _ = oracle.render(scn, 0, 1, 3, 5)
        assert oracle.last_shadow_rays > 0
        assert np.allclose(rs.image, want, rtol=1e-5, atol=1e-6)
    finally:
        compat.set_direct_lighting(False)
        compat.reset()
