set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
tail -5 gpurun_out/r2b_tests.log
FMGI_DEBUG_TIMING=1 python - > gpurun_out/r2b_e2e_debug.log 2>&1 <<'PY'
import sys, time, numpy as np, torch
sys.path.insert(0, "flatmatch-global-illumination_b200"); sys.path.insert(0, ".")
import bench, fmgi
for wl in ("synth4000_1e9x4", "synth4000_hires_1e9x4"):
    fixture, photons, depth, tile = bench.WORKLOADS[wl]
    walls, windows, lights, n = bench.load_scene(fixture, tile)
    spa = int(photons / bench.emitter_area(windows, lights))
    for pinned in (True, False):
        tex = torch.zeros((n, 4), dtype=torch.float32)
        if pinned: tex = tex.pin_memory()
        tex = tex.numpy()
        geo = fmgi.make_geometry(walls, windows, lights, tex)
        for i in range(3):
            t0 = time.perf_counter(); r = fmgi.bake(geo, spa, max_depth=depth); dt = time.perf_counter() - t0
            print(wl, "pinned" if pinned else "pageable", i, f"{1e3*dt:.1f} ms", f"{r['deposits']/dt:.4g} bounces/s", flush=True)
PY
cat gpurun_out/r2b_e2e_debug.log
