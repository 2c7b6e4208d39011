"""Scenes and ray populations for the closest-hit filter tests (shared by tests/test_gpu_filter.py and
tools/filter_margin.py).  Transforms are built the way the reference does (src/utilities.cpp:74-90:
translate * rotate(x) * rotate(y) * rotate(z) * scale in binary32, inverse in binary32), so transform and
inverseTransform carry the same kind of rounding a loaded scene has."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def build_geom(pt, gtype, material, translation, rotation_deg, scale):
    g = np.zeros(1, pt.GEOM_DTYPE)
    t = np.asarray(translation, np.float32)
    r = np.asarray(rotation_deg, np.float32)
    s = np.asarray(scale, np.float32)
    T = np.eye(4, dtype=np.float32); T[:3, 3] = t
    ang = (r * np.float32(np.pi / 180)).astype(np.float32)
    cx, cy, cz = np.cos(ang).astype(np.float32)
    sx, sy, sz = np.sin(ang).astype(np.float32)
    Rx = np.array([[1, 0, 0, 0], [0, cx, -sx, 0], [0, sx, cx, 0], [0, 0, 0, 1]], np.float32)
    Ry = np.array([[cy, 0, sy, 0], [0, 1, 0, 0], [-sy, 0, cy, 0], [0, 0, 0, 1]], np.float32)
    Rz = np.array([[cz, -sz, 0, 0], [sz, cz, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    S = np.diag(np.append(s, np.float32(1))).astype(np.float32)
    M = (T @ Rx @ Ry @ Rz @ S).astype(np.float32)
    A = np.linalg.inv(M.astype(np.float64)).astype(np.float32)  # binary32 inverse of the binary32 transform
    A[3] = [0, 0, 0, 1]
    g["type"], g["materialid"] = gtype, material
    g["translation"], g["rotation"], g["scale"] = t, r, s
    g["transform"] = M.ravel()
    g["inverseTransform"] = A.ravel()
    return g[0]


def _sample(pt):
    with open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")) as f:
        s = json.load(f)["scene"]
    return (np.frombuffer(bytes.fromhex(s["geoms_hex"]), dtype=pt.GEOM_DTYPE).copy(),
            np.frombuffer(bytes.fromhex(s["materials_hex"]), dtype=pt.MATERIAL_DTYPE).copy(),
            np.frombuffer(bytes.fromhex(s["camera_hex"]), dtype=pt.CAMERA_DTYPE).copy())


def random_scene(pt, n, seed, extent=10.0, smin=0.05, smax=3.0, aniso=1.0, offset=(0, 0, 0)):
    """n spheres / cubes with random rotations; aniso > 1 stretches one axis by up to that factor"""
    rng = np.random.default_rng(seed)
    g = np.zeros(n, pt.GEOM_DTYPE)
    for i in range(n):
        s = np.exp(rng.uniform(np.log(smin), np.log(smax))) * np.ones(3)
        s[rng.integers(3)] *= np.exp(rng.uniform(0, np.log(aniso)))
        g[i] = build_geom(pt, int(rng.integers(2)), 0, rng.uniform(-extent, extent, 3) + np.asarray(offset),
                          rng.uniform(0, 360, 3), s)
    return g


def all_scenes(pt):
    g, m, cam = _sample(pt)
    out = {"sample": (g, m, cam)}
    out["random64"] = (random_scene(pt, 64, 1), m, cam)
    out["tiny_far"] = (random_scene(pt, 48, 2, extent=200.0, smin=0.02, smax=0.5), m, cam)
    out["aniso100"] = (random_scene(pt, 32, 3, aniso=100.0), m, cam)
    out["offset1e4"] = (random_scene(pt, 32, 4, offset=(1e4, -2e4, 5e3)), m, cam)
    out["touching"] = (touching_scene(pt), m, cam)
    return out


def touching_scene(pt):
    """surfaces that coincide or touch: stacked cubes, a sphere resting on a cube, concentric spheres, a duplicate"""
    gs = [build_geom(pt, 1, 0, (0, 0, 0), (0, 0, 0), (4, 1, 4)),
          build_geom(pt, 1, 0, (0, 1, 0), (0, 0, 0), (2, 1, 2)),       # shares the plane y = 0.5
          build_geom(pt, 0, 0, (0, 2.5, 0), (0, 0, 0), (2, 2, 2)),     # rests on the upper cube
          build_geom(pt, 0, 0, (0, 2.5, 0), (0, 30, 0), (2, 2, 2)),    # the same sphere, rotated frame
          build_geom(pt, 0, 0, (0, 2.5, 0), (0, 0, 0), (1, 1, 1)),     # concentric, inside
          build_geom(pt, 1, 0, (0, 0, 0), (0, 0, 0), (4, 1, 4)),       # exact duplicate of geom 0
          build_geom(pt, 1, 0, (3, 0, 0), (0, 45, 0), (2, 1, 2))]      # overlaps geom 0
    g = np.zeros(len(gs), pt.GEOM_DTYPE)
    for i, x in enumerate(gs):
        g[i] = x
    return g


def ray_sets(pt, ctx, geoms, n):
    """ray populations: uniform random, aimed at geoms (near-silhouette), bounce-like (start on surfaces)"""
    rng = np.random.default_rng(99)
    c = np.array([g["transform"].reshape(4, 4)[:3, 3] for g in geoms], np.float64)
    lo, hi = c.min(0) - 5, c.max(0) + 5
    o = rng.uniform(lo, hi, (n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True).astype(np.float32)
    sets = {"uniform": (o, d)}
    # aimed: toward a point on/near a random geom's bounding sphere -> many grazing rays
    k = rng.integers(len(geoms), size=n)
    rad = np.array([np.linalg.norm(g["transform"].reshape(4, 4)[:3, :3], 2) for g in geoms])[k]
    u = rng.normal(size=(n, 3)); u /= np.linalg.norm(u, axis=1, keepdims=True)
    target = c[k] + u * (rad * rng.uniform(0.3, 0.75, n))[:, None]
    d2 = (target - o).astype(np.float32)
    sets["aimed"] = (o, d2)  # unnormalised on purpose
    # bounce-like: origins = exact hit points pushed off the surface along the normal, cosine-ish directions
    gid, t, p, nr = ctx.intersect(o, d2, mode=pt.HIT_EXACT_SCAN)
    hit = gid >= 0
    if hit.sum() > 0:
        ob = (p[hit] + nr[hit] * np.float32(2e-4)).astype(np.float32)
        db = rng.normal(size=ob.shape).astype(np.float32)
        db /= np.linalg.norm(db, axis=1, keepdims=True).astype(np.float32)
        flip = (db * nr[hit]).sum(1) < 0
        db[flip] = -db[flip]
        sets["bounce"] = (ob, db)
        # through: continue the ray into the geom (refraction-like, starts just inside)
        oi = (p[hit] - nr[hit] * np.float32(3e-4)).astype(np.float32)
        sets["inside"] = (oi, d2[hit] / np.linalg.norm(d2[hit], axis=1, keepdims=True).astype(np.float32))
    return sets
