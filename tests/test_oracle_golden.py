"""The CPU oracle against the golden vectors (no GPU).

ref_vectors.json holds outputs of the REFERENCE'S OWN functions (generated in the build container from
oracle/_ref, see oracle/gen_golden.py): the restatement must reproduce them bit for bit.
spec_vectors.json freezes our specification of the stubbed pieces; its Philox entries are Random123's KATs.
sampling_vectors.json: the reference's getRandomPointOnCube (pinned) and the frozen sampling / absorption stubs."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, f32, same_bits


def _geoms(ref_gold, pt):
    return np.frombuffer(bytes.fromhex(ref_gold["scene"]["geoms_hex"]), dtype=pt.GEOM_DTYPE)


def test_hash(oracle, ref_gold):
    for a, h in ref_gold["hash"]:
        assert oracle.hash(a) == h


def test_multiplyMV_and_getPointOnRay(oracle, ref_gold):
    for k in ref_gold["multiplyMV"]:
        assert same_bits(oracle.multiplyMV(f32(k["m"]), f32(k["v"])), f32(k["out"]))
    for k in ref_gold["getPointOnRay"]:
        assert same_bits(oracle.getPointOnRay(f32(k["o"]), f32(k["d"]), f32([k["t"]])[0]), f32(k["out"]))


def test_sphereIntersectionTest_bit_exact(oracle, ref_gold, pt):
    g = _geoms(ref_gold, pt)
    n_hits = 0
    for k in ref_gold["sphereIntersectionTest"]:
        o, d = f32(k["o"]).reshape(-1, 3), f32(k["d"]).reshape(-1, 3)
        t, p, n = oracle.intersect_one(g[k["geom"]], 0, o, d)
        want_t = f32(k["t"])
        assert same_bits(t, want_t)
        hit = want_t > 0
        n_hits += int(hit.sum())
        assert same_bits(p[hit], f32(k["p"]).reshape(-1, 3)[hit]) and same_bits(n[hit], f32(k["n"]).reshape(-1, 3)[hit])
    assert n_hits > 300


def test_reference_stubs_are_stubs(ref_gold):
    """boxIntersectionTest returns -1 and calculateBSDF returns 1 in the reference"""
    assert (f32(ref_gold["boxIntersectionTest_stub"]) == -1).all()
    assert ref_gold["calculateBSDF_stub"] == 1


def test_getRadiuses(oracle, ref_gold, pt):
    g = _geoms(ref_gold, pt)
    for k in ref_gold["getRadiuses"]:
        assert same_bits(oracle.getRadiuses(g[k["geom"]]), f32(k["out"]))


def test_hemisphere_reference_formula_bit_exact(oracle, ref_gold):
    h = ref_gold["hemisphere"]
    n, x1, x2 = f32(h["n"]).reshape(-1, 3), f32(h["xi1"]), f32(h["xi2"])
    want = f32(h["out"]).reshape(-1, 3)
    assert same_bits(oracle.hemisphere(n, x1, x2, ref=True), want)
    # the sampler the path loop uses swaps libm sin/cos for a reproducible polynomial: same direction to ~3e-7
    assert np.abs(oracle.hemisphere(n, x1, x2) - want).max() < 1e-6


def test_struct_layout(ref_gold, pt):
    lay = ref_gold["layout"]
    assert lay[:7] == [24, 56, 172, 52, 96, 64, 64]
    assert lay[7:14] == [pt.GEOM_DTYPE.fields[k][1] for k in
                         ("type", "materialid", "translation", "rotation", "scale", "transform", "inverseTransform")]
    assert lay[14:19] == [pt.CAMERA_DTYPE.fields[k][1] for k in ("resolution", "position", "view", "up", "fov")]
    assert lay[19:29] == [pt.MATERIAL_DTYPE.fields[k][1] for k in
                          ("color", "specularExponent", "specularColor", "hasReflective", "hasRefractive",
                           "indexOfRefraction", "hasScatter", "absorptionCoefficient", "reducedScatterCoefficient",
                           "emittance")]


def test_philox_random123_kat(oracle, spec_gold):
    for k in spec_gold["philox"]:
        assert oracle.philox(k["ctr"], k["key"]).tolist() == k["out"]
    assert oracle.u01(0) == 0.0 and oracle.u01(0xFFFFFFFF) < 1.0


def test_sincos_2pi(oracle, spec_gold):
    for k in spec_gold["sincos_2pi"]:
        u = float(f32([k["u"]])[0])
        s, c = oracle.sincos_2pi(u)
        assert same_bits([s, c], f32([k["s"], k["c"]]))
        assert abs(s - np.sin(2 * np.pi * u)) < 3e-7 and abs(c - np.cos(2 * np.pi * u)) < 3e-7


def test_box_raygen_closest_hit_frozen(oracle, spec_gold, ref_gold, pt):
    g = _geoms(ref_gold, pt)
    for k in spec_gold["boxIntersectionTest"]:
        t, p, n = oracle.intersect_one(g[k["geom"]], 1, f32(k["o"]).reshape(-1, 3), f32(k["d"]).reshape(-1, 3))
        assert same_bits(t, f32(k["t"])) and same_bits(p.ravel(), f32(k["p"])) and same_bits(n.ravel(), f32(k["n"]))
    cam = np.frombuffer(bytes.fromhex(ref_gold["scene"]["camera_hex"]), dtype=pt.CAMERA_DTYPE)
    r = spec_gold["raygen"]
    o, d = oracle.raygen(cam, (0.0, 0.0), r["seed"], r["pixel"], r["sample"])
    assert same_bits(o.ravel(), f32(r["o"])) and same_bits(d.ravel(), f32(r["d"]))
    lo, ld = oracle.raygen(cam, tuple(r["lens"]), r["seed"], r["pixel"], r["sample"])
    assert same_bits(lo.ravel(), f32(r["lens_o"])) and same_bits(ld.ravel(), f32(r["lens_d"]))
    gid, t, p, n = oracle.intersect_rays(g, o, d)
    ch = spec_gold["closest_hit"]
    assert gid.tolist() == ch["id"] and same_bits(t, f32(ch["t"]))


def test_box_test_agrees_with_analytic_cube(oracle):
    """unit cube at the origin, identity transform: hits land on the faces, normals are the face normals"""
    import importlib
    pt = importlib.import_module("project3-pathtracer_b200")
    g = np.zeros(1, pt.GEOM_DTYPE)
    g[0]["type"] = 1
    g[0]["transform"] = np.eye(4, dtype=np.float32).ravel()
    g[0]["inverseTransform"] = np.eye(4, dtype=np.float32).ravel()
    o = np.array([[0, 0, 3], [3, 0.2, 0.1], [0, 0, 0], [0, 2, 0], [0.2, 0.1, 0]], np.float32)
    d = np.array([[0, 0, -1], [-1, 0, 0], [0, 1, 0], [1, 0, 0], [0, 0, 2]], np.float32)
    t, p, n = oracle.intersect_one(g[0], 1, o, d)
    # distances are shortened by the 1e-4 pull-back (intersections.h:47)
    assert np.allclose(t, [2.4999, 2.4999, 0.4999, -1, 0.4999], atol=2e-6)
    assert np.allclose(n[0], [0, 0, 1]) and np.allclose(n[1], [1, 0, 0]) and np.allclose(n[2], [0, 1, 0])
    assert np.allclose(n[4], [0, 0, 1])  # from inside: still the outward face normal


def test_optics_frozen(oracle, spec_gold):
    for k in spec_gold["optics"]:
        n, i = f32(k["n"]), f32(k["i"])
        a, b = (float(x) for x in f32(k["ior"]))
        assert same_bits(oracle.reflect(n, i), f32(k["refl"]))
        tir, tr = oracle.refract(n, i, a, b)
        assert tir == k["tir"] and same_bits(tr, f32(k["trans"]))
        R, T = oracle.fresnel(n, i, a, b, tr, tir)
        assert same_bits([R, T], f32([k["R"], k["T"]]))
        assert 0.0 <= R <= 1.0 and abs(R + T - 1) < 1e-6
        if not tir:  # Snell: sin_t = eta * sin_i
            si = np.linalg.norm(np.cross(n, i)); st = np.linalg.norm(np.cross(n, tr))
            assert abs(st - (a / b) * si) < 1e-5


def test_render_fixture_frozen(oracle, spec_gold, ref_gold, pt):
    g = _geoms(ref_gold, pt)
    m = np.frombuffer(bytes.fromhex(ref_gold["scene"]["materials_hex"]), dtype=pt.MATERIAL_DTYPE)
    cam = np.frombuffer(bytes.fromhex(ref_gold["scene"]["camera_hex"]), dtype=pt.CAMERA_DTYPE).copy()
    cam["resolution"][0] = [40, 40]
    k = spec_gold["render_40x40_2spp_d8_seed11"]
    img, live, _ = oracle.render(oracle.make_scene(g, m, cam), 0, 2, 8, 11, threads=1)
    assert live.tolist() == k["live"] and same_bits(img.ravel(), f32(k["sum_rgb"]))
    # threads and pixel ranges do not change the result
    img2, live2, _ = oracle.render(oracle.make_scene(g, m, cam), 0, 2, 8, 11, threads=4)
    assert same_bits(img, img2) and live.tolist() == live2.tolist()
    a, la, _ = oracle.render(oracle.make_scene(g, m, cam), 0, 2, 8, 11, pix_begin=0, pix_end=700)
    b, lb, _ = oracle.render(oracle.make_scene(g, m, cam), 0, 2, 8, 11, pix_begin=700, pix_end=1600, sum_rgb=a)
    assert same_bits(b, img) and (la + lb).tolist() == live.tolist()


# ---- surface-point / direction sampling and absorption (a12, a13, a15, a20) ----
@pytest.fixture(scope="module")
def samp_gold():
    with open(os.path.join(GOLD, "sampling_vectors.json")) as f:
        return json.load(f)


def _as_cube(g):
    c = np.array(g).reshape(1).copy()
    c["type"] = 1
    return c


def test_getRandomPointOnCube_bit_exact(oracle, ref_gold, samp_gold, pt):
    """the restatement (thrust minstd + hash + the host build's right-to-left draws) against the reference's own
    getRandomPointOnCube, src/intersections.h:133-175: 9 transforms x 256 seeds"""
    g = _geoms(ref_gold, pt)
    seeds = f32(samp_gold["seeds"])
    assert len(samp_gold["ref_cube_points"]) == len(g)
    for k in samp_gold["ref_cube_points"]:
        got = oracle.random_points(_as_cube(g[k["geom"]]), seeds)
        assert same_bits(got.ravel(), f32(k["p"]))
    assert (f32(samp_gold["ref_stubs"]) == 0).all()  # what the reference's three stubs return today


def test_sampling_spec_frozen(oracle, ref_gold, samp_gold, pt):
    g = _geoms(ref_gold, pt)
    seeds, u = f32(samp_gold["seeds"]), f32(samp_gold["u"]).reshape(-1, 3)
    for k in samp_gold["sphere_points"]:
        assert same_bits(oracle.random_points(g[k["geom"]], seeds).ravel(), f32(k["p"]))
    for k in samp_gold["points_u"]:
        assert same_bits(oracle.points_u(g[k["geom"]], u).ravel(), f32(k["p"]))
    assert same_bits(oracle.sphere_dirs(u[:, 0], u[:, 1]).ravel(), f32(samp_gold["sphere_dirs"]))
    t = samp_gold["transmission"]
    ab, dist = f32(t["absorption"]).reshape(-1, 3), f32(t["distance"])
    assert same_bits(oracle.transmission(ab, dist).ravel(), f32(t["T"]))


def test_sampling_properties(oracle, ref_gold, pt):
    g = _geoms(ref_gold, pt)
    rng = np.random.default_rng(9)
    u = rng.random((4000, 3), dtype=np.float32)
    d = oracle.sphere_dirs(u[:, 0], u[:, 1]).astype(np.float64)
    assert np.abs(np.linalg.norm(d, axis=1) - 1).max() < 1e-6 and np.abs(d.mean(axis=0)).max() < 0.05
    for gi in range(len(g)):
        inv = np.array(g[gi]["inverseTransform"], np.float64).reshape(4, 4)
        p = oracle.points_u(g[gi], u).astype(np.float64)
        po = p @ inv[:3, :3].T + inv[:3, 3]  # object space
        if int(g[gi]["type"]) == 0:
            assert np.abs(np.linalg.norm(po, axis=1) - 0.5).max() < 1e-3  # on the sphere of radius .5
        else:
            assert np.abs(np.abs(po).max(axis=1) - 0.5).max() < 1e-3  # on a face (walls are scaled by .01: 100x rounding)
    # Beer-Lambert against exp() in binary64; zero absorption transmits everything, huge absorption nothing
    ab = (rng.random((2000, 3)) * 5).astype(np.float32)
    dist = (rng.random(2000) * 8).astype(np.float32)
    T = oracle.transmission(ab, dist)
    want = np.exp(-(ab * dist[:, None]).astype(np.float64))
    assert np.abs(T - want).max() <= 2e-7 * 1 + 0 and (np.abs(T - want) / want).max() < 2e-7
    assert (oracle.transmission([[0, 0, 0]], [3.0]) == 1).all() and (oracle.transmission([[200, 300, 1e30]], [1.0]) == 0).all()


def test_direct_light_sampling_is_unbiased(oracle, pt):
    """oracle only: a grey floor under a sphere light seen through one tiny pixel.  Radiance leaving the floor point
    below a sphere light of radius r at height D is albedo * Le * (r/D)^2; plain path tracing and direct light
    sampling must both converge to it."""
    from scenes_for_tests import build_geom
    m = np.zeros(2, pt.MATERIAL_DTYPE)
    m[0]["color"] = [0.5, 0.5, 0.5]
    m[1]["color"] = [1, 1, 1]
    m[1]["emittance"] = 2.0
    g = np.zeros(2, pt.GEOM_DTYPE)
    g[0] = build_geom(pt, 1, 0, (0, -0.5, 0), (0, 0, 0), (100, 1, 100))
    g[1] = build_geom(pt, 0, 1, (0, 5, 0), (0, 0, 0), (2, 2, 2))
    cam = np.zeros(1, pt.CAMERA_DTYPE)
    v = np.array([0, -1, -3.0]) / np.sqrt(10.0)
    cam["resolution"][0], cam["position"][0], cam["view"][0] = [2, 2], [0, 1, 3], v
    cam["up"][0], cam["fov"][0] = [0, 1, 0], [0.01, 0.01]
    spp, want = 40000, 0.5 * 2.0 * (1 / 5) ** 2
    for nee, tol in ((False, 0.03), (True, 0.01)):
        img, live, _ = oracle.render(oracle.make_scene(g, m, cam, direct_lighting=nee), 0, spp, 2, 3)
        assert abs(float(img.mean()) / spp - want) < tol * want
        assert (oracle.last_shadow_rays > 0) == nee
    g[1] = build_geom(pt, 1, 1, (0, 5, 0), (10, 20, 30), (2, 1, 3))  # a rotated box light: both estimators agree
    a = oracle.render(oracle.make_scene(g, m, cam, direct_lighting=False), 0, spp, 2, 3)[0].mean() / spp
    b = oracle.render(oracle.make_scene(g, m, cam, direct_lighting=True), 0, spp, 2, 3)[0].mean() / spp
    assert abs(a - b) < 0.03 * a


def test_generateRandomNumberFromThread_bit_exact(oracle, samp_gold):
    """the noise the reference's stub renderer writes (src/raytraceKernel.cu:29-36, 93-104), host build's draw order"""
    for k in samp_gold["ref_noise_host"]:
        got = oracle.noise_image(k["W"], k["H"], k["time"], True)
        assert same_bits(got.ravel(), f32(k["rgb"]))
        fwd = oracle.noise_image(k["W"], k["H"], k["time"], False)
        assert same_bits(fwd[:, ::-1], got)  # the other order is the same three numbers reversed


def test_small_helpers_bit_exact(oracle, samp_gold):
    """epsilonCheck, getInverseDirectionOfRay, getSignOfRay (src/intersections.h:37-43, 62-70) against the reference"""
    for k in samp_gold["epsilonCheck"]:
        assert int(oracle.epsilonCheck(float(f32([k["a"]])[0]), float(f32([k["b"]])[0]))) == k["r"]
    with np.errstate(divide="ignore"):
        for k in samp_gold["ray_helpers"]:
            inv, sign = oracle.ray_helpers(f32(k["d"]))
            assert same_bits(inv, f32(k["inv"])) and same_bits(sign, f32(k["sign"]))
