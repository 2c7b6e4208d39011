#!/usr/bin/env python
"""bench.py -- headline benchmark: path segments per second on the reference's sample scene.

Workload (BASELINE.json configs[1]): the shipped example scene (9 objects, diffuse + emissive), 800x800,
5000 samples per pixel, 8 bounces, stream compaction on.  One "step" = one whole render of that frame
(3.2e9 paths) on every rank.  With N ranks (torchrun, one per GPU) rank r renders samples [r*5000, (r+1)*5000)
of every pixel (sharding by sample index, weak scaling) and the per-GPU accumulation buffers are combined with
one NCCL reduce per step, inside the timed region.

  value      device-timed Mseg/s with everything resident in HBM (CUDA events on the launching stream)
  e2e        the same render through the public host API with HOST buffers: scene upload (H2D) + render +
             download of the float image (D2H) every step, wall clock around synchronised calls
  roofline   HBM roofline of the dominant kernel k_bounce (all launches of a step together)
  cpu_baseline  the CPU oracle (oracle/pt_oracle.c, "port") on configs[0] = same scene, 800x800, 1 spp, 8 bounces

`--impl reference` times the CPU implementation of the path on the host cores: the reference's own kernels are
TODO stubs, so this is the oracle port, which calls restatements of the reference's implemented
functions and our specification of the stubs.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "path segments per second, reference sample scene 800x800, 5000 spp, 8 bounces"
UNIT = "Mseg/s"
RES, SPP, DEPTH, SEED = 800, 5000, 8, 565
CPU_SPP = 32  # bounded CPU sample: 32 spp of the 800x800 frame (about 20 s of CPU work over the host cores)
WF_SPP = 50  # samples of the frame per wavefront: 32 M paths, 3.07 GB of path state (>> 126 MB L2); 100 wavefronts per step


def load_sample_scene(pt):
    """The reference loader's parse of its own data/scenes/sampleScene.txt (tests/golden/ref_vectors.json)."""
    with open(os.path.join(ROOT, "tests", "golden", "ref_vectors.json")) as f:
        s = json.load(f)["scene"]
    g = np.frombuffer(bytes.fromhex(s["geoms_hex"]), dtype=pt.GEOM_DTYPE).copy()
    m = np.frombuffer(bytes.fromhex(s["materials_hex"]), dtype=pt.MATERIAL_DTYPE).copy()
    c = np.frombuffer(bytes.fromhex(s["camera_hex"]), dtype=pt.CAMERA_DTYPE).copy()
    return g, m, c


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_oracle_rate(threads=0, spp=CPU_SPP):
    """configs[0] scaled to a ~20 CPU-second sample: sample scene, 800x800, CPU_SPP spp, 8 bounces on the host
    cores through the oracle port."""
    from oracle_py import Oracle
    pt_dtypes = importlib.import_module("project3-pathtracer_b200")
    g, m, c = load_sample_scene(pt_dtypes)
    orc = Oracle()
    scn = orc.make_scene(g, m, c)
    nthreads = threads if threads > 0 else orc.max_threads()
    _, live, secs = orc.render(scn, 0, spp, DEPTH, SEED, threads=nthreads)
    segs = int(live.sum())
    return segs, secs, nthreads


_JSON_FD = 1  # the process's original stdout (main() moves everything else to stderr)


def _emit(line):
    os.write(_JSON_FD, (line + "\n").encode())


def run_reference(args, rank):
    """--impl reference: the CPU path on the host cores (rank 0 only)."""
    if rank != 0:
        return
    rates, last = [], None
    for i in range(args.warmup + args.steps):
        segs, secs, nthreads = cpu_oracle_rate()
        if i >= args.warmup:
            rates.append((segs, secs))
        last = nthreads
    tot_s = sum(s for s, _ in rates)
    tot_t = sum(t for _, t in rates)
    val = tot_s / tot_t / 1e6
    sample = "sample scene 800x800, %d spp, 8 bounces per step (BASELINE configs[0] x %d); %d segments per step" % (CPU_SPP, CPU_SPP, rates[0][0])
    _emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "reference sample scene 800x800, 8 bounces; bounded sample: %d spp per step on host cores" % CPU_SPP},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": last, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--spp", type=int, default=SPP)
    ap.add_argument("--wf-spp", type=int, default=WF_SPP)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE line, the JSON: libraries that write to file descriptor 1 on their own (NCCL prints
    # "NCCL version ..." there when NCCL_DEBUG is set) are sent to stderr instead
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the path tracer has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    pt = importlib.import_module("project3-pathtracer_b200")
    geoms, mats, cam = load_sample_scene(pt)
    npix = RES * RES
    ctx = pt.Context(geoms, mats, cam, device=local)
    ctx.set_wavefront_paths(npix * args.wf_spp)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    sh = importlib.import_module("project3-pathtracer_b200.sharding")
    accum = sh.accum_tensor(ctx)  # zero-copy view of the float4 accumulation image in HBM
    first_sample = rank * args.spp

    def step():
        ctx.clear()
        ctx.render(first_sample, args.spp, DEPTH, SEED)
        if world > 1:
            with torch.cuda.stream(stream):
                sh.reduce_image(accum, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed: K whole renders, scene and buffers resident ----
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record(stream)
    for _ in range(args.steps):
        step()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    paths, segs, live = ctx.counters()  # of the last step (clear() resets them)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    tot = torch.tensor([float(segs), float(paths), float(launches)], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    segs_all, paths_all, launches_all = (float(x) for x in tot.tolist())
    ms_per_step = ms / args.steps
    value = segs_all / (ms_per_step * 1e-3) / 1e6

    # ---- end to end through the host API: H2D scene + render + D2H image, every step ----
    host_img = torch.empty((npix, 3), dtype=torch.float32).pin_memory()
    host_np = host_img.numpy()
    hg = torch.from_numpy(geoms.view(np.uint8)).pin_memory().numpy().view(pt.GEOM_DTYPE)
    hm = torch.from_numpy(mats.view(np.uint8)).pin_memory().numpy().view(pt.MATERIAL_DTYPE)
    hc = torch.from_numpy(cam.view(np.uint8)).pin_memory().numpy().view(pt.CAMERA_DTYPE)
    h2d = int(hg.nbytes + hm.nbytes + hc.nbytes)
    d2h = int(host_np.nbytes)

    def e2e_step():
        ctx.update_scene(hg, hm, hc)
        step()
        ctx.download_mean(args.spp * world if rank == 0 else args.spp, out=host_np)

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = segs_all * args.steps / e2e_s / 1e6

    if rank == 0:
        # roofline of k_bounce: algorithmic path-state bytes over all its launches of one step
        P1, S1 = float(paths), float(segs)
        alg_bytes = 96.0 * (S1 - P1) + 32.0 * P1
        peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(peaks_path):
            peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = alg_bytes / (ms_per_step * 1e-3) / 1e9
        n_bounce = max(1, (launches // args.steps) * DEPTH // (DEPTH + 1))  # launches per step = wavefronts * (DEPTH k_bounce + 1 k_accum_counts)
        traffic, traffic_note, pipes = None, "no ncu capture on file", None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tj = json.load(open(tpath))
            # DRAM bytes of one average k_bounce launch = measured DRAM/algorithmic ratio of the captured launch x the
            # algorithmic bytes of an average launch of this run
            traffic = tj["dram_over_algorithmic"] * alg_bytes / n_bounce
            traffic_note = "ncu dram read+write / algorithmic = %.3f on the captured launch (%s)" % (tj["dram_over_algorithmic"], tj["source"])
            pipes = tj.get("pipes_pct_of_peak")  # FP32 pipe utilisation next to the HBM fraction (same capture)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "reference sample scene (9 objects) %dx%d, %d spp per GPU, %d bounces, compaction on"
                                   % (RES, RES, args.spp, DEPTH),
                       "sharding": "by sample index; one NCCL reduce of the float4 image per step" if world > 1 else "single GPU",
                       "wavefront_paths": npix * args.wf_spp,
                       "l2": "path state per wavefront %.0f MB > 126 MB L2 (no flush needed)" % (npix * args.wf_spp * 96 / 1e6),
                       "seed": SEED},
            "spp_per_s": args.spp * world / (ms_per_step * 1e-3),
            "segments_per_step": segs_all, "paths_per_step": paths_all,
            "live_per_depth": [int(x) for x in live[:DEPTH]],
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches_all),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_nominal_8000": achieved / 8000.0,
                         "traffic": traffic, "traffic_note": traffic_note, "peak_source": peak_src, "kernel": "k_bounce",
                         "ncu_pipes_pct_of_peak": pipes,
                         "algorithmic_bytes_per_launch": alg_bytes / n_bounce, "launches_per_step": n_bounce,
                         "avg_launch_us": 1e3 * ms_per_step / n_bounce,
                         "note": "achieved = algorithmic bytes 96*(S-P) + 32*P of a step / CUDA-event time of the step; the step is "
                                 "%d k_bounce launches (99.9 %% of its GPU time, profiles/r01_launches_v14.csv), two wavefronts in "
                                 "flight on two streams, so this is bytes per average launch / (step time / launches)" % n_bounce},
        }
        if world == 1 and not args.no_cpu_baseline:
            csegs, csecs, cthreads = cpu_oracle_rate()
            out["cpu_baseline"] = {"value": csegs / csecs / 1e6, "unit": UNIT, "cores": cthreads, "kind": "port",
                                   "sample": "BASELINE configs[0] x %d: sample scene 800x800, %d spp, 8 bounces (%d segments, %.2f s)"
                                             % (CPU_SPP, CPU_SPP, csegs, csecs)}
            s1, t1, _ = cpu_oracle_rate(threads=1, spp=2)  # the same oracle on ONE core
            out["cpu_baseline_1core"] = {"value": s1 / t1 / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
                                         "sample": "sample scene 800x800, 2 spp, 8 bounces (%d segments, %.2f s)" % (s1, t1)}
            # the other kernel the north star asks a roofline for: the stable stream-compaction primitive (an HBM-bound
            # kernel by nature), device-timed on its own, outside the timed region above
            try:
                n_c, keep = 1 << 27, 0.7
                rng = np.random.default_rng(1)
                vals = rng.integers(0, 2 ** 32, n_c, dtype=np.uint32)
                flg = (rng.random(n_c, dtype=np.float32) < keep).astype(np.uint8)
                kept, cms = pt.compact_u32_timed(vals, flg, iters=10)
                cbytes = 5.0 * n_c + 4.0 * len(kept)
                out["compaction_primitive"] = {
                    "kernels": "k_compact_count + k_compact_scan + k_compact_scatter (pt_compact_u32)", "elements": n_c,
                    "keep": keep, "ms": cms, "Gelem_per_s": n_c / cms / 1e6, "bound": "hbm", "achieved": cbytes / cms / 1e6,
                    "peak": peak, "unit": "GB/s", "frac": cbytes / cms / 1e6 / peak,
                    "algorithmic_bytes": "5 B read per element + 4 B written per kept element"}
                del vals, flg, kept
            except Exception as e:  # never let the side measurement take the headline line down
                out["compaction_primitive"] = {"error": str(e)}
        _emit(json.dumps(out))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
