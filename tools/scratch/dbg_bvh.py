import importlib, sys, numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
pt = importlib.import_module("project3-pathtracer_b200")
from scenes_for_tests import _sample, random_scene
_, m, cam = _sample(pt)
for n in (33, 64):
    rng = np.random.default_rng(n)
    g = random_scene(pt, n, 100 + n, extent=4.0, smin=0.5, smax=2.0)
    o = rng.uniform(-8, 8, (50_000, 3)).astype(np.float32)
    d = rng.normal(size=(50_000, 3)).astype(np.float32)
    with pt.Context(g, m, cam) as ctx:
        want = ctx.intersect(o, d, mode=pt.HIT_EXACT_SCAN)
        got = ctx.intersect(o, d, with_stats=True)
    bad = np.nonzero(got[0] != want[0])[0]
    print(n, "mismatch", len(bad), "fallbacks", got[4], "want hits", (want[0] >= 0).sum(), "got hits", (got[0] >= 0).sum())
    for i in bad[:8]:
        print("  ray", i, "want", want[0][i], want[1][i], "got", got[0][i], got[1][i])
