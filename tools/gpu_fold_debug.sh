set -u
mkdir -p gpurun_out
FMGI_DEBUG_TIMING=1 python - > gpurun_out/fold_debug.log 2>&1 <<'PY'
import sys, numpy as np
sys.path.insert(0, "flatmatch-global-illumination_b200"); sys.path.insert(0, "oracle")
import fmgi, refbind
scene = refbind.Scene.load("tests/golden/example_scene.npz")
spa, depth = 200000, 5
tex1 = fmgi.aligned_texels(scene.num_texels)
st1 = fmgi.bake(fmgi.make_geometry(scene.walls, scene.windows, scene.lights, tex1), spa, max_depth=depth, seed=5)
for g in (2,):
    for rep in range(2):
        texg = fmgi.aligned_texels(scene.num_texels)
        stg = fmgi.bake(fmgi.make_geometry(scene.walls, scene.windows, scene.lights, texg), spa, max_depth=depth, seed=5, num_gpus=g)
        d = np.abs(texg.astype(np.float64) - tex1)
        bad = np.nonzero(d.max(axis=1) > 0.05 + 1e-5 * np.abs(tex1).max(axis=1))[0]
        print("gpus", g, "rep", rep, "max abs diff", d.max(), "bad texels", len(bad), "of", scene.num_texels, "first/last bad", (bad[:3], bad[-3:]) if len(bad) else None,
              "sum ratio", texg[:, :3].sum(dtype=np.float64) / tex1[:, :3].sum(dtype=np.float64))
        half = scene.num_texels // 2
        print("  sum ratio first half", texg[:half, :3].sum(dtype=np.float64) / tex1[:half, :3].sum(dtype=np.float64), "second half", texg[half:, :3].sum(dtype=np.float64) / tex1[half:, :3].sum(dtype=np.float64))
        if len(bad): print("  example", bad[0], texg[bad[0]], tex1[bad[0]])
PY
cat gpurun_out/fold_debug.log | cut -c1-500
