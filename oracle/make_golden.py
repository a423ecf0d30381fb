#!/usr/bin/env python3
"""TEST INFRASTRUCTURE — generates the committed fixtures under tests/golden/ by running the
UNMODIFIED reference (oracle/_ref/libfmgi_ref*.so, built from /root/reference by oracle/Makefile).

Run here (where /root/reference exists); the GPU box only reads the fixtures.

    python oracle/make_golden.py scene            # example_scene.npz + example_facts.json
    python oracle/make_golden.py synth            # synth800_scene.npz, synth4000_scene.npz
    python oracle/make_golden.py atlas --depth 8 --spa 5000000 --procs 8
    python oracle/make_golden.py atlas --depth 3 --spa 5000000 --procs 8

`atlas` runs `procs` independent processes of performPhotonMappingNative (photonmap.c:408) with
srand(seed0 + i), sums the raw atlases of the first and second half of the processes in float64
and stores the normalised luminance (main.c:68-79 scaling, rectangle.c:277 weights) of the
base-level texels for both halves, plus the raw RGB energy totals.  Two halves give the tests a
measured Monte-Carlo noise floor.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import multiprocessing as mp
import sys
import time
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE))
import refbind as rb  # noqa: E402

GOLDEN = HERE.parent / "tests" / "golden"
EXAMPLE_PNG = Path("/root/reference/example.png")
LUMA = np.array([0.2126, 0.7152, 0.0722])  # rectangle.c:277


def cmd_scene(_args):
    ref = rb.RefLib()
    px = rb.load_layout_png(EXAMPLE_PNG)
    scene, gjson = ref.parse_rgba(px, 30.0, 200.0)
    cmap = ref.collision_map_json(px)
    GOLDEN.mkdir(parents=True, exist_ok=True)
    scene.save(GOLDEN / "example_scene.npz")
    facts = {
        "source": "parseLayout (parseLayout.c:359) on example.png, scale 30 px/m, TILE_SIZE 200",
        "num_walls": len(scene.walls), "num_windows": len(scene.windows), "num_lights": len(scene.lights),
        "num_box_walls": len(scene.box_walls), "num_texels": scene.num_texels,
        "base_texels": int(scene.base_texel_mask().sum()),
        "sha256_geometry_json": hashlib.sha256(gjson).hexdigest(), "len_geometry_json": len(gjson),
        "sha256_collision_map_json": hashlib.sha256(cmap).hexdigest(), "len_collision_map_json": len(cmap),
        "sha256_walls": hashlib.sha256(scene.walls.tobytes()).hexdigest(),
        "sha256_windows": hashlib.sha256(scene.windows.tobytes()).hexdigest(),
        "sha256_lights": hashlib.sha256(scene.lights.tobytes()).hexdigest(),
        "photon_counts_spa_100000": scene.photon_counts(100000),
        "start": [scene.meta["startX"], scene.meta["startY"]],
        "native_atlas_sha256": {},
    }
    # small seeded runs of the compiled reference: the restatement must reproduce these bit for bit
    refd = rb.RefLib(runtime_depth=True)
    for spa, seed, depth in [(3000, 1, 8), (3000, 7, 8), (2000, 3, 3), (2000, 5, 4), (1000, 9, 1)]:
        lib = ref if depth == 8 else refd
        tex, _ = lib.photonmap_native(scene, spa, seed, depth)
        facts["native_atlas_sha256"][f"spa{spa}_seed{seed}_depth{depth}"] = hashlib.sha256(tex.tobytes()).hexdigest()
    (GOLDEN / "example_facts.json").write_text(json.dumps(facts, indent=1) + "\n")
    print(json.dumps(facts, indent=1))


def _worker(job):
    depth, spa, seed, fixture = job
    ref = rb.RefLib(runtime_depth=(depth != 8))
    scene = rb.Scene.load(GOLDEN / f"{fixture}_scene.npz")
    tex, secs = ref.photonmap_native(scene, spa, seed, depth)
    return tex.astype(np.float64), secs


def cmd_atlas(args):
    scene = rb.Scene.load(GOLDEN / f"{args.fixture}_scene.npz")
    jobs = [(args.depth, args.spa, args.seed0 + i, args.fixture) for i in range(args.procs)]
    t0 = time.time()
    with mp.Pool(args.procs) as pool:
        res = pool.map(_worker, jobs)
    wall = time.time() - t0
    half = args.procs // 2
    mask = scene.base_texel_mask()
    out = {}
    for name, part in (("a", res[:half]), ("b", res[half:])):
        raw = np.sum([r[0] for r in part], axis=0)
        spa_total = args.spa * len(part)
        norm = scene.normalisation(spa_total)
        lum = (raw[:, :3] @ LUMA) * norm
        out[f"lum_{name}"] = lum[mask][:: args.stride].astype(np.float32)
        # per-wall mean luminance: almost noise-free, catches any per-surface bias
        out[f"wall_lum_{name}"] = np.array([lum[int(w["lightmapSetup"][0]): int(w["lightmapSetup"][0]) +
                                               int(w["lightmapSetup"][1]) * int(w["lightmapSetup"][2])].mean()
                                           for w in scene.walls], dtype=np.float64)
        out[f"rgb_total_{name}"] = raw[:, :3].sum(axis=0)
        out[f"spa_{name}"] = np.int64(spa_total)
    photons = sum(scene.photon_counts(args.spa))
    out.update(stride=np.int64(args.stride), depth=np.int64(args.depth), photons_per_half=np.int64(photons * half),
               cpu_seconds=np.array([r[1] for r in res]), wall_seconds=np.float64(wall),
               seeds=np.array([j[2] for j in jobs]))
    path = GOLDEN / f"{args.fixture}_native_depth{args.depth}.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path}: {photons * args.procs:.3e} photons in {wall:.0f} s wall, "
          f"per-process {np.mean(out['cpu_seconds']):.0f} s")


def cmd_synth(_args):
    """Synthetic layouts (BASELINE.json configs[2]): pixels from our generator, rectangles from the
    reference's parseLayout (parseLayout.c:359), scale 30 px/m, TILE_SIZE 200."""
    sys.path.insert(0, str(HERE.parent / "flatmatch-global-illumination_b200"))
    from fmgi import synth

    ref = rb.RefLib()
    for name, kw in (("synth800", dict(size_px=800, seed=1)),
                     ("synth4000", dict(size_px=4000, seed=1, room_min_m=2.4, room_max_m=5.0))):
        img = synth.make_layout(**kw)
        scene, _ = ref.parse_rgba(img, 30.0, 200.0)
        scene.meta["generator"] = json.dumps(kw, sort_keys=True)
        scene.save(GOLDEN / f"{name}_scene.npz")
        print(name, len(scene.walls), "walls", len(scene.windows), "windows", len(scene.lights), "lights",
              scene.num_texels, "texels", hashlib.sha256(scene.walls.tobytes()).hexdigest()[:16])


def cmd_ao(_args):
    """performAmbientOcclusionNative (photonmap.c:480) on example.png: base-level texel values."""
    ref = rb.RefLib()
    scene = rb.Scene.load(GOLDEN / "example_scene.npz")
    t0 = time.time()
    tex = ref.ambient_occlusion_native(scene)
    mask = scene.base_texel_mask()
    assert np.array_equal(tex[:, 0], tex[:, 1]) and np.array_equal(tex[:, 0], tex[:, 2]) and not tex[~mask].any()
    np.savez_compressed(GOLDEN / "example_ao_native.npz", ao=tex[mask, 0], seconds=np.float64(time.time() - t0))
    print("wrote example_ao_native.npz", tex[mask, 0].mean(), time.time() - t0)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    sub.add_parser("scene")
    sub.add_parser("synth")
    sub.add_parser("ao")
    a = sub.add_parser("atlas")
    a.add_argument("--depth", type=int, default=8)
    a.add_argument("--spa", type=int, default=5_000_000)
    a.add_argument("--procs", type=int, default=8)
    a.add_argument("--seed0", type=int, default=1000)
    a.add_argument("--fixture", default="example")
    a.add_argument("--stride", type=int, default=1, help="keep every stride-th base texel (fixture size)")
    args = ap.parse_args()
    {"scene": cmd_scene, "atlas": cmd_atlas, "synth": cmd_synth, "ao": cmd_ao}[args.cmd](args)
