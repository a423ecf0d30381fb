#!/usr/bin/env python3
"""tests/golden/<name>_scene.npz -> the flat binary tools/simt_walk_sim.cpp reads."""
import sys
import numpy as np

z = np.load(sys.argv[1])
with open(sys.argv[2], "wb") as f:
    np.array([z["walls"].size // 80, z["windows"].size // 80, z["lights"].size // 80], dtype="<i4").tofile(f)
    for k in ("walls", "windows", "lights"):
        z[k].tofile(f)
