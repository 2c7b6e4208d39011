"""The N>1 host path on CPU: world_size-2 gloo.  Each rank renders its block of sample indices (with the CPU oracle
standing in for the device), the float images are combined with the same reduce the GPU path uses, and rank 0
compares with a single-process render of all samples."""
import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sample_ranges_tile_exactly():
    sh = importlib.import_module("project3-pathtracer_b200.sharding")
    for world in (1, 2, 3, 4, 8):
        for spp in (0, 1, 7, 8, 5000, 16384):
            seen = []
            for r in range(world):
                b, n = sh.sample_range(r, world, spp, first_sample=11)
                seen += list(range(b, b + n))
            assert seen == list(range(11, 11 + spp))
    with pytest.raises(ValueError):
        sh.sample_range(2, 2, 10)


def _worker(rank, world, port, spp, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import json
    import torch
    import torch.distributed as dist
    from oracle_py import Oracle
    pt = importlib.import_module("project3-pathtracer_b200")
    sh = importlib.import_module("project3-pathtracer_b200.sharding")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = json.load(open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")))["scene"]
    g = np.frombuffer(bytes.fromhex(s["geoms_hex"]), dtype=pt.GEOM_DTYPE).copy()
    m = np.frombuffer(bytes.fromhex(s["materials_hex"]), dtype=pt.MATERIAL_DTYPE).copy()
    cam = np.frombuffer(bytes.fromhex(s["camera_hex"]), dtype=pt.CAMERA_DTYPE).copy()
    cam["resolution"][0] = [48, 48]
    orc = Oracle()
    scn = orc.make_scene(g, m, cam)
    b, n = sh.sample_range(rank, world, spp)
    img, live, _ = orc.render(scn, b, n, 8, 5, threads=1)
    t = torch.from_numpy(img)
    lv = torch.from_numpy(live.astype(np.int64))
    sh.reduce_image(t, dst=0)
    dist.reduce(lv, dst=0)
    if rank == 0:
        full, full_live, _ = orc.render(scn, 0, spp, 8, 5, threads=1)
        q.put((float(np.abs(t.numpy() - full).max()), lv.numpy().tolist() == full_live.astype(np.int64).tolist()))
    dist.destroy_process_group()


def test_two_rank_gloo_reduce_matches_single_process():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 9, q)) for r in range(2)]
    for p in procs:
        p.start()
    err, live_ok = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert live_ok, "segment counts of the shards must add up exactly"
    assert err <= 1e-5, err  # only float summation order differs
