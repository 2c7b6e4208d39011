import importlib, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
pt = importlib.import_module("project3-pathtracer_b200")
from scenes_for_tests import _sample, random_scene
_, m, cam = _sample(pt)
n = 33
rng = np.random.default_rng(n)
g = random_scene(pt, n, 100 + n, extent=4.0, smin=0.5, smax=2.0)
o = rng.uniform(-8, 8, (50_000, 3)).astype(np.float32)
d = rng.normal(size=(50_000, 3)).astype(np.float32)
with pt.Context(g, m, cam) as ctx:
    i = 11
    want = ctx.intersect(o[i:i+1], d[i:i+1], mode=pt.HIT_EXACT_SCAN)
    print("want", want[0], want[1], flush=True)
    got = ctx.intersect(o[i:i+1], d[i:i+1])
    print("got", got[0], got[1])
