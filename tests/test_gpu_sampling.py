"""GPU parity: surface-point / direction sampling and Beer-Lambert absorption (8f
rank 4) through the C ABI against the oracle and against the reference's own getRandomPointOnCube (golden vectors).
Bar: identical bits (same arithmetic contract as the rest of the path)."""
import json
import os

import numpy as np
import pytest

from conftest import GOLD, f32, optics_scene, same_bits, with_resolution

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def samp_gold():
    with open(os.path.join(GOLD, "sampling_vectors.json")) as f:
        return json.load(f)


def _as_cube(g):
    c = np.array(g).reshape(1).copy()
    c["type"] = 1
    return c


def test_getRandomPointOnCube_matches_the_reference(pt, sample_scene, samp_gold):
    """the CUDA sampler against vectors produced by the reference's own code (src/intersections.h:133-175)"""
    g = sample_scene["geoms"]
    seeds = f32(samp_gold["seeds"])
    for k in samp_gold["ref_cube_points"]:
        got = pt.random_points_on_geom(_as_cube(g[k["geom"]]), seeds)
        assert same_bits(got.ravel(), f32(k["p"]))


def test_samplers_match_oracle_bitwise(pt, oracle, sample_scene):
    g = sample_scene["geoms"]
    rng = np.random.default_rng(21)
    seeds = np.concatenate([rng.integers(0, 2 ** 31, 20000), np.arange(64)]).astype(np.float32)
    u = rng.random((20000, 3), dtype=np.float32)
    u[:8] = [[0, 0, 0], [0.99999994, 0.99999994, 0.99999994], [0.5, 0, 0.99999994], [0.25, 0.25, 0.25],
             [0.75, 0.5, 0.5], [0, 0.99999994, 0], [0.125, 0.375, 0.625], [0.99999994, 0, 0]]
    for gi in range(len(g)):
        assert same_bits(pt.random_points_on_geom(g[gi:gi + 1], seeds), oracle.random_points(g[gi:gi + 1], seeds))
        assert same_bits(pt.points_on_geom_u(g[gi:gi + 1], u), oracle.points_u(g[gi:gi + 1], u))
    assert same_bits(pt.random_directions_in_sphere(u[:, 0], u[:, 1]), oracle.sphere_dirs(u[:, 0], u[:, 1]))


def test_transmission_matches_oracle_bitwise(pt, oracle, samp_gold):
    rng = np.random.default_rng(22)
    ab = np.concatenate([rng.random((50000, 3)) * 6, rng.random((2000, 3)) * 500, np.zeros((10, 3))]).astype(np.float32)
    dist = np.concatenate([rng.random(50000) * 12, rng.random(2000) * 3, rng.random(10)]).astype(np.float32)
    ab[:3] = [[np.inf, 0, -1], [np.nan, 1e-30, 87.0], [1e38, 1e-38, 1]]
    dist[:3] = [1.0, 1.0, 1.0]
    got, want = pt.calculate_transmission(ab, dist), oracle.transmission(ab, dist)
    assert same_bits(got, want)
    t = samp_gold["transmission"]
    assert same_bits(pt.calculate_transmission(f32(t["absorption"]).reshape(-1, 3), f32(t["distance"])).ravel(), f32(t["T"]))
    ok = np.isfinite(ab).all(axis=1) & (ab >= 0).all(axis=1)
    ref = np.exp(-(ab[ok].astype(np.float64) * dist[ok, None]))
    assert np.abs(got[ok] - ref).max() < 2e-7  # and close to the real exponential


def test_empty_and_invalid(pt, sample_scene):
    g = sample_scene["geoms"]
    assert pt.random_points_on_geom(g[0:1], []).shape == (0, 3)
    assert pt.points_on_geom_u(g[0:1], np.zeros((0, 3))).shape == (0, 3)
    assert pt.random_directions_in_sphere([], []).shape == (0, 3)
    assert pt.calculate_transmission(np.zeros((0, 3)), []).shape == (0, 3)
    mesh = g[0:1].copy()
    mesh["type"] = 2
    with pytest.raises(pt.PtError):
        pt.random_points_on_geom(mesh, [1.0])
    for bad in (-1.0, float("nan"), 4294967296.0, float("inf")):  # float -> unsigned is defined on [0, 2^32) only
        with pytest.raises(pt.PtError):
            pt.random_points_on_geom(g[0:1], [3.0, bad])


def test_absorbing_glass_paths_match_oracle(pt, oracle, sample_scene):
    """glass sphere and glass cube with ABSCOEFF > 0: every segment that runs inside them is attenuated by
    calculateTransmission; image and live counts identical to the oracle, and darker than without absorption"""
    g, m = optics_scene(pt, sample_scene)
    cam = with_resolution(sample_scene["camera"], 128, 128)
    m0 = m.copy()
    m[4]["absorptionCoefficient"] = [0.4, 1.5, 0.05]
    m[6]["absorptionCoefficient"] = [0.0, 0.3, 2.0]
    want_sum, want_live, _ = oracle.render(oracle.make_scene(g, m, cam), 0, 2, 12, 31)
    with pt.Context(g, m, cam) as c:
        c.render(0, 2, 12, 31)
        got = c.download_sum()
        _, _, live = c.counters()
    assert live[:12].tolist() == want_live.tolist()
    assert same_bits(got, want_sum)
    clear_sum, clear_live, _ = oracle.render(oracle.make_scene(g, m0, cam), 0, 2, 12, 31)
    assert clear_live.tolist() == want_live.tolist()  # absorption changes no path, only its weight
    assert want_sum.sum() < clear_sum.sum() and (want_sum <= clear_sum + 1e-6).all()
