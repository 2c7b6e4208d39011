"""Shared fixtures.  GPU tests are marked @pytest.mark.gpu and call the product through its C ABI; the oracle
(oracle/) is loaded here only as the checker."""
import importlib
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def pt():
    """the product package (directory name has a hyphen)"""
    return importlib.import_module("project3-pathtracer_b200")


@pytest.fixture(scope="session")
def oracle():
    from oracle_py import Oracle
    return Oracle()


@pytest.fixture(scope="session")
def ref_gold():
    with open(os.path.join(GOLD, "ref_vectors.json")) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def spec_gold():
    with open(os.path.join(GOLD, "spec_vectors.json")) as f:
        return json.load(f)


def f32(bits):
    return np.array(bits, dtype=np.uint32).view(np.float32)


def same_bits(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.uint32)
    return a.shape == b.shape and bool((a == b).all())


@pytest.fixture(scope="session")
def sample_scene(ref_gold, pt):
    """The reference loader's parse of its own data/scenes/sampleScene.txt (golden, bit-exact)."""
    s = ref_gold["scene"]
    return dict(
        geoms=np.frombuffer(bytes.fromhex(s["geoms_hex"]), dtype=pt.GEOM_DTYPE).copy(),
        materials=np.frombuffer(bytes.fromhex(s["materials_hex"]), dtype=pt.MATERIAL_DTYPE).copy(),
        camera=np.frombuffer(bytes.fromhex(s["camera_hex"]), dtype=pt.CAMERA_DTYPE).copy(),
        width=s["width"], height=s["height"], iterations=s["iterations"], image_name=s["image_name"],
    )


def with_resolution(camera, w, h):
    """same view, smaller frame: fov.x follows the aspect like src/scene.cpp:203-207 (square frames keep fov)"""
    cam = camera.copy()
    yscaled = np.tan(np.float32(cam["fov"][0][1]) * (np.pi / 180))
    fovx = np.arctan(np.float32(yscaled * w) / h) * 180 / np.pi
    cam["resolution"][0] = [w, h]
    cam["fov"][0][0] = np.float32(fovx)
    return cam


def optics_scene(pt, sample_scene):
    """sample-scene geometry with a mirror sphere, a glass sphere and a glass cube (materials re-pointed)"""
    g = sample_scene["geoms"].copy()
    m = sample_scene["materials"].copy()
    m[3]["hasReflective"] = 1.0
    m[3]["specularColor"] = [0.9, 0.9, 0.9]
    m[4]["hasRefractive"] = 1.0
    m[4]["indexOfRefraction"] = 1.5
    m[4]["specularColor"] = [1, 1, 1]
    m[4]["color"] = [0.95, 0.95, 0.95]
    m[6]["hasRefractive"] = 1.0
    m[6]["indexOfRefraction"] = 2.2
    m[6]["specularColor"] = [1, 1, 1]
    extra = np.zeros(1, pt.GEOM_DTYPE)
    extra[0] = g[6]
    return g, m
