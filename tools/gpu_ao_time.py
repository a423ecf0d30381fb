"""Ambient occlusion (fmgi_ambient_occlusion) per tier on the example flat and synth800: wall time of the second call."""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "flatmatch-global-illumination_b200")); sys.path.insert(0, str(ROOT / "oracle"))
import numpy as np
import fmgi, refbind
for name in ("example_scene", "synth800_scene"):
    sc = refbind.Scene.load(ROOT / "tests" / "golden" / f"{name}.npz")
    res = {}
    for tier, label in ((fmgi.TIER_GRID, "grid"), (fmgi.TIER_ROOMS, "rooms")):
        for rep in range(3):
            tex = fmgi.aligned_texels(sc.num_texels)
            geo = fmgi.make_geometry(sc.walls, sc.windows, sc.lights, tex)
            t0 = time.perf_counter()
            fmgi.ambient_occlusion(geo, tier=tier)
            dt = time.perf_counter() - t0
        res[label] = (dt * 1e3, tex[:, 0].copy())
    a, b = res["grid"][1], res["rooms"][1]
    m = a > 0
    print(name, "grid %.2f ms rooms %.2f ms" % (res["grid"][0], res["rooms"][0]), "texels differing > 1e-5 rel:",
          float(np.mean(np.abs(a[m] - b[m]) > 1e-5 * np.maximum(a[m], 1e-3))), "nan:", int(np.isnan(b).sum()))
