set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -x -k "device_grid or synth4000_grid or more_z or random_soups or outside" 2>&1 | tail -8
FMGI_DEBUG_TIMING=1 python - 2>&1 <<'PY' | grep -E "fmgi\]|bounces" | cut -c1-700
import sys, time, numpy as np, torch
sys.path.insert(0, "flatmatch-global-illumination_b200"); sys.path.insert(0, ".")
import bench, fmgi
fixture, photons, depth, tile = bench.WORKLOADS["synth4000_1e9x4"]
walls, windows, lights, n = bench.load_scene(fixture, tile)
spa = int(photons / bench.emitter_area(windows, lights))
tex = torch.zeros((n, 4), dtype=torch.float32).pin_memory().numpy()
geo = fmgi.make_geometry(walls, windows, lights, tex)
for i in range(3):
    t0 = time.perf_counter(); r = fmgi.bake(geo, spa, max_depth=depth); dt = time.perf_counter() - t0
    print(i, f"{1e3*dt:.1f} ms", f"{r['deposits']/dt:.4g} bounces/s", "grid_build_ms", r["grid_build_ms"], "prepare", r["prepare_ms"], "upload", r["upload_ms"], flush=True)
PY
