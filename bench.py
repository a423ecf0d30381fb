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
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "rectangle tables of the reference's parseLayout (tests/golden fixture); photons from libc rand()",
        "config": {"workload": args.workload, "scene": fixture, "depth": depth,
                   "note": "reference CPU path performPhotonMappingNative on host cores, bounded sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "per_core": value / cores,
                         "sample": sample},
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
    1e8 photons/m^2 and 8 bounces.  None if the prebuilt binary or the layout is missing.  The library's own
    account of where the call's time went (FMGI_STATS=2) rides along."""
    import re
    import subprocess
    import tempfile

    app = ROOT / "flatmatch-global-illumination_b200" / "build" / "globalIllumination"
    png = ROOT / "tests" / "golden" / "example.png"
    if not app.exists() or not png.exists():
        return None
    with tempfile.TemporaryDirectory() as tmp:
        os.mkdir(os.path.join(tmp, "tiles"))
        best, out = None, b""
        for _ in range(3):
            t0 = time.perf_counter()
            # one visible GPU: driver initialisation time grows with the number of GPUs it enumerates
            env = dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0],
                       FMGI_STATS="2")
            r = subprocess.run([str(app), str(png)], cwd=tmp, capture_output=True, env=env)
            dt = time.perf_counter() - t0
            if r.returncode != 0:
                return None
            if best is None or dt < best:
                best, out = dt, r.stdout
        res = {"wall_s": best}
        m2 = re.search(rb"\[INF\] fmgi exit_begin ([0-9.]+) ms", out)
        if m2:
            res["exit_begin_ms"] = float(m2.group(1))
        m = re.search(rb"\[INF\] fmgi breakdown: (.*)", out)
        if m:
            for part in m.group(1).decode().split(","):
                k, _, v = part.strip().partition(" ")
                try:
                    res[k + "_ms"] = float(v.split()[0])
                except (ValueError, IndexError):
                    pass
        return res


class Harness:
    """torch / torch.distributed plumbing shared by the measurements of one bench run."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
        torch.cuda.set_device(self.local)
        self.cpu_group = None
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            # host-side barrier for the phases in which rank 0 drives every GPU itself: a NCCL barrier would leave
            # a spinning kernel on the other ranks' GPUs, and two processes on one GPU are time-sliced
            self.cpu_group = dist.new_group(backend="gloo")
        self.flush = torch.empty(192 * 1024 * 1024 // 4, dtype=torch.float32, device="cuda")   # > 126 MB L2
        self.stream = torch.cuda.current_stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def barrier_idle(self):
        """Barrier that leaves the GPUs idle while waiting (gloo)."""
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier(group=self.cpu_group)

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def measure(h, workload, steps, warmup, e2e_steps, photons=0.0, total_photons=0.0, depth=0, clocks=False,
            deposit_peak=False):
    """One workload on all ranks: device-timed bake steps (trace + the NCCL atlas reduce for N > 1) and the
    host-buffer C-ABI call.  Returns the result dict on rank 0, None elsewhere."""
    import fmgi
    from fmgi.distributed import bake_sharded

    torch, dist, args = h.torch, h.dist, h.args
    world, rank, local = h.world, h.rank, h.local
    fixture, wl_photons, wl_depth, tile_size = WORKLOADS[workload]
    photons = photons or wl_photons
    strong = total_photons > 0               # configs[4]: fixed total budget split over the GPUs
    if strong:
        photons = total_photons / world
    depth = depth or wl_depth
    walls, windows, lights, num_texels = load_scene(fixture, tile_size)
    area = emitter_area(windows, lights)
    spa_gpu = int(photons / area)                 # per-GPU density
    spa_job = spa_gpu * world                     # weak scaling: the job grows with N
    t0 = time.perf_counter()
    scene = fmgi.DeviceScene(walls, windows, lights, num_texels, device=local)
    scene_create_ms = 1e3 * (time.perf_counter() - t0)
    atlas = torch.zeros((num_texels, 4), dtype=torch.float32, device="cuda")
    stream = h.stream

    def trace(buf, spa, shard, num_shards):
        scene.trace(buf.data_ptr(), spa, stream=stream.cuda_stream, max_depth=depth, seed=args.seed,
                    shard=shard, num_shards=num_shards, deposit=args.deposit)

    def step():
        atlas.zero_()
        h.flush.fill_(1.0)                         # L2 flush between timed iterations
        bake_sharded(trace, atlas, spa_job, rank, world, dist=dist)   # trace + one NCCL reduce onto rank 0

    # rectangle tests per ray (the T of the algorithmic flop count): a scene property, measured once by a
    # short untimed trace with the counting kernel variant (the timed kernel does not count, two
    # instructions less per walk step)
    scene.trace(atlas.data_ptr(), max(1, int(2.0e6 / area)), stream=stream.cuda_stream, max_depth=depth, seed=args.seed,
                count_tests=1)
    st = scene.sync()
    tests_per_ray = st["rect_tests"] / max(st["rays"], 1)

    for _ in range(warmup):
        step()
    h.barrier()
    kernel_ms, deposits, rays, photons_done = [], 0, 0, 0
    launches0 = scene.sync()["kernel_launches"]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        h.barrier()
        ev0.record(stream)
        for _ in range(steps):
            step()
            st = scene.sync()                       # counters + CUDA-event time of the trace kernel
            kernel_ms.append(st["trace_ms"])
            deposits += st["deposits"]; rays += st["rays"]; photons_done += st["photons"]
        ev1.record(stream)
        h.barrier()
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
    scene.close()
    del atlas

    # ---- e2e: host buffers through the C-ABI call performGlobalIlluminationCl wraps ----------------
    # ONE fmgi_bake call per step on rank 0 with num_gpus = N: the library shards the photons over the N GPUs
    # itself (host threads), folds the atlases over peer memory and returns the slices over every GPU's PCIe
    # link - what the reference-facing entry point does with FMGI_GPUS=N.  The other ranks wait at the barrier.
    e2e = None
    if e2e_steps > 0:
        h.barrier_idle()
        if rank == 0:
            tex = torch.zeros((num_texels, 4), dtype=torch.float32).pin_memory().numpy()    # pinned host atlas
            geo = fmgi.make_geometry(walls, windows, lights, tex)
            e2e_opts = dict(max_depth=depth, seed=args.seed, num_gpus=world, device=local, deposit=args.deposit)
            first = fmgi.bake(geo, spa_job, **e2e_opts)       # warm-up (contexts, streams, peer mappings, module load)
            e2e_dep, parts, e2e_s = 0, [], 0.0
            for _ in range(e2e_steps):
                tex[...] = 0                                   # the caller's zeroed atlas (parseLayout.c:526-533): not timed
                t0 = time.perf_counter()
                r = fmgi.bake(geo, spa_job, **e2e_opts)        # blocking: returns with the host atlas written
                e2e_s += time.perf_counter() - t0
                e2e_dep += r["deposits"]
                parts.append(r)
            mean = lambda k: float(np.mean([q[k] for q in parts]))
            e2e = {"value": e2e_dep / e2e_s, "unit": UNIT,
                   "h2d_bytes_per_step": 16 * num_texels + 80 * (len(walls) + len(windows) + len(lights)),
                   "d2h_bytes_per_step": 16 * num_texels, "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps,
                   "api": "fmgi_bake(num_gpus=N) on rank 0: host Geometry + pinned host atlas in, host atlas out; "
                          "what performGlobalIlluminationCl runs - table build, H2D (under the trace), trace, "
                          "peer-memory fold, D2H every call",
                   # (room tier, 2048 rectangles and more: the box decomposition runs beside the rectangle tables)
                   "breakdown_ms": {"host_tables": (max(mean("prepare_ms"), mean("grid_build_ms"))
                                                    if parts[0]["tier"] == 4 and len(walls) >= 2048
                                                    else mean("prepare_ms") + mean("grid_build_ms")),
                                    "grid_build": mean("grid_build_ms"), "table_upload": mean("upload_ms"),
                                    "atlas_h2d_overlapped": mean("h2d_ms"), "trace_device": mean("trace_ms"),
                                    "fold": mean("reduce_ms"), "atlas_d2h": mean("d2h_ms"), "total": mean("total_ms")},
                   "first_call_ms": first["total_ms"], "first_call_init_ms": first["init_ms"]}
            del tex, geo
        h.barrier_idle()

    if rank != 0:
        return None
    pk, pk_kind = peaks()
    sms = st["num_sms"]
    f_hz = pk.get("sm_max_mhz", 1965.0) * 1e6
    fp32_peak = sms * 128 * 2 * f_hz / 1e12                      # TFLOP/s, FMA = 2
    flops_per_ray = 13.0 * tests_per_ray + 150.0
    rays_per_s_kernel = (rays_all / world) / (steps * kms * 1e-3)
    achieved = rays_per_s_kernel * flops_per_ray / 1e12
    res = {
        "value": value, "ms_per_step": ms / steps, "steps": steps, "warmup": warmup,
        "scaling": "strong" if strong else "weak",
        "config": {"workload": workload, "scene": fixture, "rectangles": int(len(walls)),
                   "emitters": int(len(windows) + len(lights)), "atlas_texels": num_texels,
                   "photons_per_gpu_per_step": photons_all / world / steps, "depth": depth,
                   "samples_per_area_per_gpu": spa_gpu, "parallelism": f"photon-range shards x{world}",
                   "atlas_bytes": 16 * num_texels, "texels_per_m2": tile_size or 200,
                   "tier": {0: "auto", 1: "soup", 2: "grid", 4: "rooms"}[st["tier"]],
                   "l2": "flushed between steps (192 MiB fill)",
                   "deposit": ["vec4", "scalar", "warp_agg"][args.deposit]},
        "rays_per_s": rays_all / (ms * 1e-3), "photons_per_s": photons_all / (ms * 1e-3),
        "kernel_ms_per_step": kms,
        # step - kernel: atlas zero + L2 flush fill + (N > 1) the NCCL reduce of the atlas
        "non_kernel_ms_per_step": ms / steps - kms,
        "scene_create_ms": scene_create_ms,
        "roofline": {"bound": "fp32_issue", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                     "frac": achieved / fp32_peak, "traffic": None,
                     "peak_source": f"{sms} SMs x 128 lanes x 2 x {pk.get('sm_max_mhz', 1965.0):.0f} MHz ({pk_kind})",
                     "flops_per_ray": flops_per_ray, "rect_tests_per_ray": tests_per_ray,
                     "note": "algorithmic FP32 work per SURVEY.md 8(d): 13*T+150 flops per ray; T = rectangle tests per ray "
                             "(room tier: boxes crossed, a 14-flop slab test each, + face grids looked up)"},
        "gpu_launches": launches,
    }
    if e2e:
        res["e2e"] = e2e
    if clocks:
        res["clocks"] = clk.summary()
    src_hash = fmgi.lib().fmgi_source_hash().decode()
    facts = ncu_facts(workload) if not (photons != wl_photons or depth != wl_depth or strong) else None
    if facts and facts.get("tier") == res["config"]["tier"]:
        if facts.get("src_hash") != src_hash:
            # counters of another build of the kernel: kept for reference, never quoted as this kernel's
            res["roofline"]["ncu_stale"] = {"src_hash": facts.get("src_hash"), "library_src_hash": src_hash,
                                            "note": "profiles/ncu_facts.json was captured on another build"}
        else:
            res["roofline"]["traffic"] = facts.get("dram_bytes_per_launch")
            res["roofline"]["ncu"] = facts
            if facts.get("warp_inst_per_ray"):
                # the resource that actually binds this divergent traversal: instruction issue slots.
                # warp instructions per ray come from the committed ncu capture of THIS build (src_hash) and
                # scene, the ray rate is live; peak = one warp instruction per cycle and SM sub-partition.
                issue_peak = sms * 4 * f_hz
                issue = rays_per_s_kernel * facts["warp_inst_per_ray"]
                res["roofline"]["issue"] = {
                    "achieved": issue / 1e9, "peak": issue_peak / 1e9, "unit": "G warp-inst/s",
                    "frac": issue / issue_peak, "warp_inst_per_ray": facts["warp_inst_per_ray"],
                    "lanes_per_inst": facts.get("threads_per_instruction")}
    if deposit_peak:
        # the deposit instruction's own rate, measured in this run - at uniform-random texels
        # of an atlas-sized scratch buffer (HBM-bound once the atlas exceeds L2) and of an L2-resident one (what
        # a bake reaches when its deposits are local: photons are handed out emitter by emitter)
        try:
            dep_rate = (deposits_all / world) / (steps * kms * 1e-3)
            peak_atlas = fmgi.deposit_peak(num_texels, 300_000_000, device=local)
            peak_l2 = fmgi.deposit_peak(min(num_texels, 1 << 20), 300_000_000, device=local)
            res["roofline"]["deposit"] = {
                "achieved": dep_rate, "peak": peak_l2, "frac": dep_rate / peak_l2,
                "peak_uniform_over_atlas": peak_atlas, "unit": "deposits/s per GPU (RED.E.ADD.F32x4)",
                "note": "peak = bare deposit instruction at uniform-random texels of an L2-resident footprint; "
                        "peak_uniform_over_atlas = the same over a scratch buffer of the atlas size"}
        except Exception as e:                       # a probe must never cost the bench line
            res["roofline"]["deposit"] = {"error": str(e)}
    return res


def run_ours(args):
    import fmgi

    h = Harness(args)
    strong = args.total_photons > 0
    main = measure(h, args.workload, args.steps, args.warmup, args.e2e_steps if args.e2e_steps >= 0 else args.steps,
                   photons=args.photons, total_photons=args.total_photons, depth=args.depth, clocks=True,
                   deposit_peak=True)
    # BASELINE.json configs[2..4] ride along on the default line so that the driver-run record carries them:
    # the 21.5k-rectangle scene with its 0.46 GB atlas, the same at 4x texel density (1.83 GB atlas: the reduce
    # and the PCIe read-back at size), and the fixed-total photon sweep on example.png (strong scaling).
    secondary = {}
    default_run = not (args.photons or args.depth or strong or args.no_secondary) and args.workload == "example_1e8x3"
    if default_run:
        for name in ("synth4000_1e9x4", "synth4000_hires_1e9x4"):
            r = measure(h, name, 3, 2, 2)
            if r:
                secondary[name] = {k: r[k] for k in ("value", "ms_per_step", "kernel_ms_per_step", "non_kernel_ms_per_step",
                                                     "scene_create_ms", "e2e", "roofline", "config", "steps")}
        sweep = []
        for total in (1e6, 1e7, 1e8, 1e9):
            r = measure(h, "example_default_x8", 5, 3, 0, total_photons=total, depth=8)
            if r:
                sweep.append({"total_photons": total, "value": r["value"], "ms_per_step": r["ms_per_step"],
                              "kernel_ms_per_step": r["kernel_ms_per_step"], "steps": r["steps"]})
        if sweep:
            secondary["sweep_example_d8"] = sweep
    if h.rank == 0:
        fixture = WORKLOADS[args.workload][0]
        area = emitter_area(*load_scene(fixture)[1:3])
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": h.world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": main["scaling"], "vs_baseline": None, "dtype": "f32",
            "data": ("rectangle tables of the reference's parseLayout on its own example.png (tests/golden fixture); "
                     "photons are generated on the device (Philox)" if fixture.startswith("example") else
                     "synthetic layout (fmgi/synth.py) parsed by the reference's parseLayout (tests/golden fixture); "
                     "photons are generated on the device (Philox)"),
            "config": main["config"],
        }
        for k in ("rays_per_s", "photons_per_s", "kernel_ms_per_step", "non_kernel_ms_per_step", "roofline", "e2e",
                  "gpu_launches", "clocks"):
            if k in main:
                line[k] = main[k]
        line["src_hash"] = fmgi.lib().fmgi_source_hash().decode()
        if secondary:
            line["secondary"] = secondary
        if h.world == 1 and args.workload.startswith("example") and not args.no_app:
            app = reference_app_wall_time()
            if app is not None:
                line["example_bake_wall_s"] = app["wall_s"]
                line["example_bake"] = app
        if h.world == 1 and not args.no_cpu:
            depth = args.depth or WORKLOADS[args.workload][2]
            cores = os.cpu_count() or 1
            cpu_spa = max(int(args.cpu_photons / area), 1)
            rate, kind, dep, secs = cpu_reference_rate(fixture, depth, cpu_spa, cores)
            line["cpu_baseline"] = {
                "value": rate, "unit": UNIT, "cores": cores, "kind": kind, "per_core": rate / cores,
                "sample": f"{cores} processes x {args.cpu_photons:.3g} photons (spa={cpu_spa}) of the same "
                          f"scene/depth, BSP build included ({dep:.3g} bounces, {secs:.1f} s)"}
        print(json.dumps(line))
    h.close()


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
    ap.add_argument("--e2e-steps", type=int, default=-1, help="host-buffer fmgi_bake steps (default: --steps)")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the secondary workloads (BASELINE configs 2-4) of the default run")
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
