#!/usr/bin/env python3
"""profiles/<tag>_scale_{1,2,4,8}.json (tools/gpu_scale_r2.sh) -> the scaling tables in markdown.

    python profiles/make_scaling_md.py r2h > profiles/r2h_scaling.md
"""
import json
import sys
from pathlib import Path

HERE = Path(__file__).resolve().parent
tag = sys.argv[1]
lines = {n: json.loads((HERE / f"{tag}_scale_{n}.json").read_text()) for n in (1, 2, 4, 8)}


def sci(x):
    m, e = f"{x:.3e}".split("e")
    return f"{m}×10^{int(e)}"


print(f"# Multi-GPU measurements (one box, 8× B200), kernels of build {lines[1]['src_hash']} ({tag})\n")
print("`gpurun --gpus 8 -- bash tools/gpu_scale_r2.sh`: the driver-shaped `bench.py` line at N = 1, 2, 4, 8 (torchrun, one rank per\n"
      "GPU, NCCL atlas reduce inside the timed region, L2 flushed between steps); `secondary` carries BASELINE configs[2], [3], [4].\n"
      f"Raw lines: `{tag}_scale_{{1,2,4,8}}.json`. \"value\" = whole-job photon-bounces/s, device-timed, max over ranks. \"e2e\" = ONE\n"
      "`fmgi_bake(num_gpus=N)` call on rank 0 with a pinned host atlas (table build, sharding over host threads, reduce-scatter fold\n"
      "over NVLink peer mappings, every GPU returns its slice over its own PCIe link).\n")
print("## Weak scaling (per-GPU budget fixed) — BASELINE configs[1], [2], [3]\n")
print("| workload | GPUs | bounces/s | ms/step | trace kernel ms | efficiency | e2e bounces/s (in-library, host buffers) | e2e fold ms | e2e read-back ms |")
print("|---|---|---|---|---|---|---|---|---|")
names = {"example_1e8x3": "example_1e8x3 (172 rect., 1.8 MB atlas, 10⁸ photons × 3 per GPU)",
         "synth4000_1e9x4": "synth4000_1e9x4 (21.5k rect., 0.46 GB atlas, 10⁹ photons × 4 per GPU)",
         "synth4000_hires_1e9x4": "synth4000_hires_1e9x4 (4× texel density, 1.83 GB atlas)"}
for wl, label in names.items():
    base = None
    for n in (1, 2, 4, 8):
        d = lines[n] if wl == "example_1e8x3" else lines[n]["secondary"][wl]
        base = base or d["value"]
        bd = d["e2e"]["breakdown_ms"]
        print(f"| {label if n == 1 else ''} | {n} | {sci(d['value'])} | {d['ms_per_step']:.2f} | {d['kernel_ms_per_step']:.2f} | "
              f"{d['value'] / (n * base):.3f} | {sci(d['e2e']['value'])} | {bd['fold']:.2f} | {bd['atlas_d2h']:.2f} |")
print()
d8 = lines[8]
print(f"First `fmgi_bake` call of a process: {d8['e2e']['first_call_ms']:.0f} ms at 8 GPUs (driver + primary contexts "
      f"{d8['e2e']['first_call_init_ms']:.0f} ms, peer mappings), {lines[1]['e2e']['first_call_ms']:.0f} ms at 1 GPU; every later call is "
      "the steady state above.\n")
print("## Fixed total budget (strong scaling) — BASELINE configs[4]: example.png, 8 bounces, total photons over all GPUs\n")
steady = lines[1]["secondary"]["sweep_example_d8"][-1]["value"]
print(f"| total photons | GPUs | bounces/s | ms/step | trace kernel ms | efficiency vs 1 GPU at the same total | efficiency vs N × the 1-GPU steady rate ({sci(steady)}) |")
print("|---|---|---|---|---|---|---|")
for i, q1 in enumerate(lines[1]["secondary"]["sweep_example_d8"]):
    for n in (1, 2, 4, 8):
        q = lines[n]["secondary"]["sweep_example_d8"][i]
        tp = q["total_photons"]
        print(f"| {'10^%d' % round(__import__('math').log10(tp)) if n == 1 else ''} | {n} | {sci(q['value'])} | {q['ms_per_step']:.3f} | "
              f"{q['kernel_ms_per_step']:.3f} | {q['value'] / (n * q1['value']):.2f} | {q['value'] / (n * steady):.2f} |")
