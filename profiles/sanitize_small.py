"""Small end-to-end exercise of every kernel for `compute-sanitizer --tool memcheck` (tiny budgets)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "flatmatch-global-illumination_b200"))
import fmgi  # noqa: E402

for fixture, spa in (("example_scene.npz", 3000), ("synth800_scene.npz", 300), ("synth4000_scene.npz", 20)):
    z = np.load(ROOT / "tests" / "golden" / fixture)
    walls = fmgi.aligned_rects(z["walls"].view(fmgi.RECT_DTYPE))
    windows = fmgi.aligned_rects(z["windows"].view(fmgi.RECT_DTYPE))
    lights = fmgi.aligned_rects(z["lights"].view(fmgi.RECT_DTYPE))
    n = int(z["num_texels"])
    for tier in (fmgi.TIER_SOUP, fmgi.TIER_GRID, fmgi.TIER_ROOMS):
        if tier == fmgi.TIER_SOUP and len(walls) > 1000:
            continue
        tex = fmgi.aligned_texels(n)
        geo = fmgi.make_geometry(walls, windows, lights, tex)
        st = fmgi.bake(geo, spa, max_depth=8, tier=tier)
        rgb, _ = fmgi.bake_tiles(geo, walls, spa, tier=tier)
        print(fixture, "tier", tier, "photons", st["photons"], "deposits", st["deposits"], "rgb bytes", rgb.size)
    if len(walls) < 1000:
        for tier in (fmgi.TIER_GRID, fmgi.TIER_ROOMS):
            s = fmgi.DeviceScene(walls, windows, lights, n, tier=tier)
            rng = np.random.default_rng(0)
            o = rng.uniform(-5, 30, (20000, 3)).astype(np.float32)    # many rays start outside the grid / the root box
            d = rng.normal(size=(20000, 3)).astype(np.float32)
            s.closest_hit(o, d)
            s.paths(0, 8, 1, 0, 2000)
            s.close()
            if fixture.startswith("example"):
                small = fmgi.aligned_texels(n)
                fmgi.ambient_occlusion(fmgi.make_geometry(walls[:40], windows, lights, small), tier=tier)
print("sanitize run complete")
