#!/usr/bin/env python3
"""Benchmark of the photon-mapping hot path (BASELINE.json metric: photon-bounces/s).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA arm
    python bench.py --impl reference [--steps K] [--warmup W]       # the reference's CPU path

One step = one bake pass over the workload: every emitter's photons traced to `depth` bounces and
deposited.  Default workload = BASELINE.json configs[1]: example.png, 1e8 photons x 3 bounces on
one B200.  For N > 1 (torchrun, one rank per GPU) every rank traces the same per-GPU budget as a
disjoint Philox sub-range of an N-times larger job (weak scaling) and the atlases are summed with
one NCCL reduce inside the timed region.

Printed keys follow the driver contract; `value` is device-timed whole-job bounces/s with scene
and atlas resident in HBM, `e2e` is the same metric through the host-buffer C-ABI call
(fmgi_bake = what performGlobalIlluminationCl runs), copies included.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT / "flatmatch-global-illumination_b200"))

WORKLOADS = {
    # name: (scene fixture, total photons per GPU, depth, lightmap texels per m^2 (0 = as parsed, 200))
    "example_1e8x3": ("example_scene.npz", 1.0e8, 3, 0),          # BASELINE.json configs[1]
    "example_default_x8": ("example_scene.npz", 1.538e9, 8, 0),   # configs[0]: reference default density
    "example_1e9x4": ("example_scene.npz", 1.0e9, 4, 0),          # north_star target
    "synth4000_1e9x4": ("synth4000_scene.npz", 1.0e9, 4, 0),      # configs[2]: ~21.5k rectangles, 0.46 GB atlas
    "synth4000_hires_1e9x4": ("synth4000_scene.npz", 1.0e9, 4, 800),  # configs[3]: 4x texel density, 1.8 GB atlas
    "synth800_1e8x4": ("synth800_scene.npz", 1.0e8, 4, 0),
}
def _baseline_metric():
    """BASELINE.json's metric string; `value` is its first half (device-timed photon-bounces/s), the
    second half (example.png bake wall time) is reported as `example_bake_wall_s`."""
    try:
        return json.loads((ROOT / "BASELINE.json").read_text())["metric"]
    except Exception:
        return "photon-bounces/sec (device-timed) at 1/2/4/8 B200; example.png bake wall time"


METRIC = _baseline_metric()
UNIT = "photon-bounces/s"


def load_scene(name, tile_size=0):
    """Rect tables of the fixture as plain numpy (no oracle code involved).  tile_size > 0 rebuilds
    the lightmap layout at that texel density (fmgi.layout.retile)."""
    import fmgi
    from fmgi import layout

    z = np.load(ROOT / "tests" / "golden" / name)
    walls = z["walls"].view(fmgi.RECT_DTYPE)
    num_texels = int(z["num_texels"])
    if tile_size:
        walls, num_texels = layout.retile(walls, float(tile_size))
    walls = fmgi.aligned_rects(walls)
    windows = fmgi.aligned_rects(z["windows"].view(fmgi.RECT_DTYPE))
    lights = fmgi.aligned_rects(z["lights"].view(fmgi.RECT_DTYPE))
    return walls, windows, lights, num_texels


def emitter_area(windows, lights):
    a = 0.0
    for r in list(windows) + list(lights):
        w, h = r["width"][:3].astype(np.float64), r["height"][:3].astype(np.float64)
        a += float(np.linalg.norm(w) * np.linalg.norm(h))
    return a


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons during the timed region (NVML; nvidia-smi semantics)."""

    def __init__(self, index=0, period=0.1):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self._nv = None
        self.period = period

    def _run(self):
        nv = self._nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self._nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ---------------------------------------------------------------------------------------------------
# CPU legs (the only places that execute anything under oracle/)
# ---------------------------------------------------------------------------------------------------

def _cpu_worker(job):
    """One process: the compiled reference (oracle/_ref) timed around performPhotonMappingNative
    (photonmap.c:408); deposits are counted exactly by replaying the same rand() stream through the
    bit-identical restatement (not timed)."""
    fixture, spa, depth, seed, use_ref = job
    sys.path.insert(0, str(ROOT / "oracle"))
    import refbind

    scene = refbind.Scene.load(ROOT / "tests" / "golden" / fixture)
    orc = refbind.OracleLib()
    if use_ref:
        ref = refbind.RefLib(runtime_depth=(depth != 8))
        _, secs = ref.photonmap_native(scene, spa, seed, depth)
        _, st = orc.bake(scene, spa, depth, orc.ACCEL_BSP, orc.RNG_LIBC, seed)
    else:
        t0 = time.perf_counter()
        _, st = orc.bake(scene, spa, depth, orc.ACCEL_BSP, orc.RNG_LIBC, seed)
        secs = time.perf_counter() - t0
    return secs, st["deposits"], st["photons"]


def cpu_reference_rate(fixture, depth, spa, procs, seed0=1):
    """Aggregate bounces/s of `procs` independent seeded processes (the reference itself is
    single-threaded, photonmap.c:267-272).  Returns (rate, kind, deposits, max seconds)."""
    import multiprocessing as mp

    use_ref = (ROOT / "oracle" / "_ref" / "libfmgi_ref.so").exists()
    jobs = [(fixture, spa, depth, seed0 + i, use_ref) for i in range(procs)]
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    secs = max(r[0] for r in res)
    deposits = sum(r[1] for r in res)
    return deposits / secs, ("reference" if use_ref else "port"), deposits, secs


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fixture, _photons, depth, _tile = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    _w, windows, lights, _n = load_scene(fixture)
    spa = max(int(args.cpu_photons / emitter_area(windows, lights)), 1)
    for _ in range(args.warmup):
        cpu_reference_rate(fixture, depth, max(spa // 10, 1000), cores)
    t_total, d_total, kind = 0.0, 0, "port"
    for k in range(args.steps):
        _, kind, dep, secs = cpu_reference_rate(fixture, depth, spa, cores, seed0=100 + 1000 * k)
        t_total += secs
        d_total += dep
    value = d_total / t_total
    sample = (f"{cores} processes x spa={spa} ({args.cpu_photons:.3g} photons each) per step, depth {depth}, "
              f"distinct srand seeds, BSP build included")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "scene": fixture, "depth": depth,
                   "note": "reference CPU path performPhotonMappingNative on host cores, bounded sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------

def ncu_facts(workload):
    """Counters of the dominant kernel for this workload from the committed `ncu --set full` capture
    (profiles/ncu_facts.json, written by profiles/summarize_ncu.py --facts); None if not captured."""
    p = ROOT / "profiles" / "ncu_facts.json"
    if not p.exists():
        return None
    return json.loads(p.read_text()).get(workload)


def reference_app_wall_time():
    """The other half of the BASELINE metric: wall time of the reference's untouched main.c linked
    against libfmgi_cuda.so (process start -> last tile PNG written) on example.png with its default
    1e8 photons/m^2 and 8 bounces.  None if the prebuilt binary or the layout is missing."""
    import subprocess
    import tempfile

    app = ROOT / "flatmatch-global-illumination_b200" / "build" / "globalIllumination"
    png = ROOT / "tests" / "golden" / "example.png"
    if not app.exists() or not png.exists():
        return None
    with tempfile.TemporaryDirectory() as tmp:
        os.mkdir(os.path.join(tmp, "tiles"))
        best = None
        for _ in range(2):
            t0 = time.perf_counter()
            # one visible GPU: driver initialisation time grows with the number of GPUs it enumerates
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0])
            r = subprocess.run([str(app), str(png)], cwd=tmp, capture_output=True, env=env)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return None
            best = dt if best is None else min(best, dt)
        return best


def run_ours(args):
    import torch
    import torch.distributed as dist

    import fmgi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    fixture, photons, depth, tile_size = WORKLOADS[args.workload]
    if args.photons:
        photons = args.photons
    strong = args.total_photons > 0          # configs[4]: fixed total budget split over the GPUs
    if strong:
        photons = args.total_photons / world
    if args.depth:
        depth = args.depth
    walls, windows, lights, num_texels = load_scene(fixture, tile_size)
    area = emitter_area(windows, lights)
    spa_gpu = int(photons / area)                 # per-GPU density
    spa_job = spa_gpu * world                     # weak scaling: the job grows with N
    scene = fmgi.DeviceScene(walls, windows, lights, num_texels, device=local)
    atlas = torch.zeros((num_texels, 4), dtype=torch.float32, device="cuda")
    flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2
    stream = torch.cuda.current_stream()

    from fmgi.distributed import bake_sharded

    def trace(buf, spa, shard, num_shards):
        scene.trace(buf.data_ptr(), spa, stream=stream.cuda_stream, max_depth=depth, seed=args.seed,
                    shard=shard, num_shards=num_shards, deposit=args.deposit)

    def step():
        atlas.zero_()
        flush.fill_(1.0)                           # L2 flush between timed iterations
        bake_sharded(trace, atlas, spa_job, rank, world, dist=dist)   # trace + one NCCL reduce onto rank 0

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # rectangle tests per ray (the T of the algorithmic flop count): a scene property, measured once by a
    # short untimed trace with the counting kernel variant (the timed kernel does not count, two
    # instructions less per walk step)
    scene.trace(atlas.data_ptr(), max(1, int(2.0e6 / area)), stream=stream.cuda_stream, max_depth=depth, seed=args.seed,
                count_tests=1)
    st = scene.sync()
    tests_per_ray = st["rect_tests"] / max(st["rays"], 1)

    for _ in range(args.warmup):
        step()
    barrier()
    kernel_ms, deposits, rays, photons_done = [], 0, 0, 0
    launches0 = scene.sync()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        barrier()
        ev0.record(stream)
        for _ in range(args.steps):
            step()
            st = scene.sync()                       # counters + CUDA-event time of the trace kernel
            kernel_ms.append(st["trace_ms"])
            deposits += st["deposits"]; rays += st["rays"]; photons_done += st["photons"]
        ev1.record(stream)
        barrier()
    ms = ev0.elapsed_time(ev1)
    launches = st["kernel_launches"] - launches0          # our kernels inside the timed region (this rank)

    tot = torch.tensor([float(deposits), float(rays), float(photons_done), ms, float(np.mean(kernel_ms))],
                       dtype=torch.float64, device="cuda")
    if world > 1:
        mx = tot.clone()
        dist.all_reduce(tot, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ms, kms = float(mx[3]), float(mx[4])
    else:
        kms = float(tot[4])
    deposits_all, rays_all, photons_all = float(tot[0]), float(tot[1]), float(tot[2])
    value = deposits_all / (ms * 1e-3)

    # ---- e2e: host buffers through the C-ABI call performGlobalIlluminationCl wraps ----------------
    tex = torch.zeros((num_texels, 4), dtype=torch.float32).pin_memory().numpy()    # pinned host atlas
    geo = fmgi.make_geometry(walls, windows, lights, tex)
    e2e_opts = dict(max_depth=depth, seed=args.seed, shard=rank, num_shards=world, device=local, deposit=args.deposit)
    fmgi.bake(geo, spa_job, **e2e_opts)            # warm-up (context, module load)
    barrier()
    t0 = time.perf_counter()
    e2e_dep = 0
    for _ in range(args.e2e_steps):
        tex[...] = 0
        e2e_dep += fmgi.bake(geo, spa_job, **e2e_opts)["deposits"]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    e2e_t = torch.tensor([float(e2e_dep), e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        s = e2e_t.clone()
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_value = float(s[0]) / float(e2e_t[1])
    else:
        e2e_value = e2e_dep / e2e_s
    rect_bytes = 80 * (len(walls) + len(windows) + len(lights))
    atlas_bytes = 16 * num_texels

    if rank == 0:
        pk, pk_kind = peaks()
        sms = st["num_sms"]
        f_hz = pk.get("sm_max_mhz", 1965.0) * 1e6
        fp32_peak = sms * 128 * 2 * f_hz / 1e12                      # TFLOP/s, FMA = 2
        flops_per_ray = 13.0 * tests_per_ray + 150.0
        rays_per_s_kernel = (rays_all / world) / (args.steps * kms * 1e-3)
        achieved = rays_per_s_kernel * flops_per_ray / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if strong else "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "scene": fixture, "rectangles": int(len(walls)),
                       "emitters": int(len(windows) + len(lights)), "atlas_texels": num_texels,
                       "photons_per_gpu_per_step": photons_all / world / args.steps, "depth": depth,
                       "samples_per_area_per_gpu": spa_gpu, "parallelism": f"photon-range shards x{world}",
                       "atlas_bytes": atlas_bytes, "texels_per_m2": tile_size or 200,
                       "tier": ["auto", "soup", "grid"][st["tier"]],
                       "l2": "flushed between steps (192 MiB fill)",
                       "deposit": ["vec4", "scalar", "warp_agg"][args.deposit]},
            "rays_per_s": rays_all / (ms * 1e-3), "photons_per_s": photons_all / (ms * 1e-3),
            "kernel_ms_per_step": kms,
            "roofline": {"bound": "fp32_issue", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                         "frac": achieved / fp32_peak, "traffic": None,
                         "peak_source": f"{sms} SMs x 128 lanes x 2 x {pk.get('sm_max_mhz', 1965.0):.0f} MHz ({pk_kind})",
                         "flops_per_ray": flops_per_ray, "rect_tests_per_ray": tests_per_ray,
                         "note": "algorithmic FP32 work per SURVEY.md 8(d): 13*T+150 flops per ray"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": atlas_bytes + rect_bytes,
                    "d2h_bytes_per_step": atlas_bytes, "steps": args.e2e_steps,
                    "api": "fmgi_bake (host Geometry + pinned host atlas in, host atlas out; what "
                           "performGlobalIlluminationCl runs: table build, H2D, trace, D2H every call)"},
            "gpu_launches": launches,
            "clocks": clk.summary(),
        }
        facts = ncu_facts(args.workload) if not (args.photons or args.depth or strong) else None
        if facts and facts.get("tier") == line["config"]["tier"]:
            line["roofline"]["traffic"] = facts.get("dram_bytes_per_launch")
            line["roofline"]["ncu"] = facts
            if facts.get("warp_inst_per_ray"):
                # the resource that actually binds this divergent traversal: instruction issue slots.
                # warp instructions per ray come from the committed ncu capture of this kernel and scene,
                # the ray rate is live; peak = one warp instruction per cycle and SM sub-partition.
                issue_peak = sms * 4 * f_hz
                issue = rays_per_s_kernel * facts["warp_inst_per_ray"]
                line["roofline"]["issue"] = {
                    "achieved": issue / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s",
                    "frac": issue / issue_peak, "warp_inst_per_ray": facts["warp_inst_per_ray"],
                    "lanes_per_inst": facts.get("threads_per_instruction")}
        # the deposit instruction's own rate, measured in this run - at uniform-random texels
        # of an atlas-sized scratch buffer (HBM-bound once the atlas exceeds L2) and of an L2-resident one (what
        # a bake reaches when its deposits are local: photons are handed out emitter by emitter)
        try:
            dep_rate = (deposits_all / world) / (args.steps * kms * 1e-3)
            peak_atlas = fmgi.deposit_peak(num_texels, 300_000_000, device=local)
            peak_l2 = fmgi.deposit_peak(min(num_texels, 1 << 20), 300_000_000, device=local)
            line["roofline"]["deposit"] = {
                "achieved": dep_rate, "peak": peak_l2, "frac": dep_rate / peak_l2,
                "peak_uniform_over_atlas": peak_atlas, "unit": "deposits/s per GPU (RED.E.ADD.F32x4)",
                "note": "peak = bare deposit instruction at uniform-random texels of an L2-resident footprint; "
                        "peak_uniform_over_atlas = the same over a scratch buffer of the atlas size"}
        except Exception as e:                       # a probe must never cost the bench line
            line["roofline"]["deposit"] = {"error": str(e)}
        if world == 1 and args.workload.startswith("example") and not args.no_app:
            wall = reference_app_wall_time()
            if wall is not None:
                line["example_bake_wall_s"] = wall
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            cpu_spa = max(int(args.cpu_photons / area), 1)
            rate, kind, dep, secs = cpu_reference_rate(fixture, depth, cpu_spa, cores)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": kind,
                "sample": f"{cores} processes x {args.cpu_photons:.3g} photons (spa={cpu_spa}) of the same "
                          f"scene/depth, BSP build included ({dep:.3g} bounces, {secs:.1f} s)"}
        print(json.dumps(line))
    scene.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="example_1e8x3", choices=sorted(WORKLOADS))
    ap.add_argument("--photons", type=float, default=0.0, help="override photons per GPU per step")
    ap.add_argument("--total-photons", type=float, default=0.0,
                    help="strong scaling: total photons per step over all GPUs (photon-count sweeps)")
    ap.add_argument("--depth", type=int, default=0)
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--deposit", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-photons", type=float, default=1.5e6, help="CPU legs: photons per process per step")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-app", action="store_true", help="skip the example.png end-to-end wall-time run")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
