"""Test wiring.  `-m "not gpu"` runs on a CPU-only box; `-m gpu` needs a B200 and calls the CUDA
path through the C ABI (ctypes over lib/libfmgi_cuda.so).  /root/reference is never read here:
the reference enters only through the prebuilt oracle/_ref/*.so and the fixtures in tests/golden/.
"""
import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "flatmatch-global-illumination_b200"))
sys.path.insert(0, str(ROOT / "oracle"))
sys.path.insert(0, str(ROOT))
GOLDEN = ROOT / "tests" / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def facts():
    return json.loads((GOLDEN / "example_facts.json").read_text())


@pytest.fixture(scope="session")
def scene():
    import refbind

    return refbind.Scene.load(GOLDEN / "example_scene.npz")


@pytest.fixture(scope="session")
def oracle():
    import subprocess

    import refbind

    subprocess.run(["make", "-s", "-C", str(ROOT / "oracle"), "oracle"], check=True)
    return refbind.OracleLib()


@pytest.fixture(scope="session")
def reflib():
    import refbind

    try:
        return refbind.RefLib(False), refbind.RefLib(True)
    except (FileNotFoundError, OSError) as e:
        pytest.skip(f"compiled reference not available: {e}")


@pytest.fixture(scope="session")
def fmgi():
    import fmgi as m

    return m


@pytest.fixture(scope="session")
def dev_scene(fmgi, scene):
    if fmgi.lib().fmgi_device_count() < 1:
        pytest.fail("no CUDA device visible: the product path has no CPU fallback")
    s = fmgi.DeviceScene(scene.walls, scene.windows, scene.lights, scene.num_texels)
    yield s
    s.close()


@pytest.fixture(scope="session")
def dev_scene_grid(fmgi, scene):
    """The same flat traced through the floor-plan grid instead of the shared-memory soup."""
    s = fmgi.DeviceScene(scene.walls, scene.windows, scene.lights, scene.num_texels, tier=fmgi.TIER_GRID)
    yield s
    s.close()


@pytest.fixture(scope="session")
def dev_scene_rooms(fmgi, scene):
    """The same flat through the room tier (box decomposition; what AUTO picks for axis-parallel scenes)."""
    s = fmgi.DeviceScene(scene.walls, scene.windows, scene.lights, scene.num_texels, tier=fmgi.TIER_ROOMS)
    yield s
    s.close()


@pytest.fixture(scope="session")
def dev_scene_soup(fmgi, scene):
    """The same flat through the brute-force soup tier (explicit: AUTO picks the grid above 64 colliders);
    its horizontal rectangles go through the grid's plane tables (kernel variant soup + planes)."""
    s = fmgi.DeviceScene(scene.walls, scene.windows, scene.lights, scene.num_texels, tier=fmgi.TIER_SOUP)
    yield s
    s.close()


@pytest.fixture(scope="session")
def dev_scene_soup_plain(fmgi, scene):
    """Soup tier with the plane tables switched off: every collider in the shared-memory scan."""
    import os

    os.environ["FMGI_SOUP_PLANES"] = "0"
    try:
        s = fmgi.DeviceScene(scene.walls, scene.windows, scene.lights, scene.num_texels, tier=fmgi.TIER_SOUP)
    finally:
        del os.environ["FMGI_SOUP_PLANES"]
    yield s
    s.close()


_RECORD = {}


@pytest.fixture(scope="session")
def record():
    """Measured parity figures (mismatch counts, identical-path shares, radiance statistics) are collected here and
    written to gpurun_out/parity_r2.json at the end of the session; the committed copy is profiles/parity_r2.json."""
    def put(name, **values):
        _RECORD.setdefault(name, {}).update({k: (float(v) if isinstance(v, (np.floating, float)) else
                                                 int(v) if isinstance(v, (np.integer, int)) else v)
                                             for k, v in values.items()})
    return put


def pytest_sessionfinish(session, exitstatus):
    if not _RECORD:
        return
    out = ROOT / "gpurun_out"
    try:
        out.mkdir(exist_ok=True)
        path = out / "parity_r2.json"
        old = json.loads(path.read_text()) if path.exists() else {}
        old.update(_RECORD)
        path.write_text(json.dumps(old, indent=1, sort_keys=True) + "\n")
    except OSError:
        pass


@pytest.fixture(scope="session")
def synth800():
    import refbind

    return refbind.Scene.load(GOLDEN / "synth800_scene.npz")


@pytest.fixture(scope="session")
def synth4000():
    import refbind

    return refbind.Scene.load(GOLDEN / "synth4000_scene.npz")


def outside_rays(scene, n, seed):
    """Rays whose origins lie OUTSIDE the walls' bounding box in x, y or z (up to 8 m away, a few much
    farther), with uniform directions: towards the flat, past it and away from it."""
    rng = np.random.default_rng(seed)
    w = scene.walls
    corners = np.concatenate([w["pos"][:, :3], w["pos"][:, :3] + w["width"][:, :3] + w["height"][:, :3]])
    lo, hi = corners.min(axis=0), corners.max(axis=0)
    o = rng.uniform(lo - 8, hi + 8, size=(n, 3))
    axis = rng.integers(0, 3, n)
    side = rng.integers(0, 2, n)
    dist = rng.uniform(1e-4, 8, n)
    dist[: n // 50] = rng.uniform(100, 1e5, n // 50)          # far away: the walk must not step off the grid
    o[np.arange(n), axis] = np.where(side == 1, hi[axis] + dist, lo[axis] - dist)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    return o.astype(np.float32), d.astype(np.float32)


def random_rays(scene, n, seed):
    """Rays as the path produces them: half start inside the flat's bounding box with uniform
    directions, half start on a wall (offset 1e-5 along the new direction, photonmap.c:254)."""
    rng = np.random.default_rng(seed)
    w = scene.walls
    corners = np.concatenate([w["pos"][:, :3], w["pos"][:, :3] + w["width"][:, :3] + w["height"][:, :3]])
    lo, hi = corners.min(axis=0), corners.max(axis=0)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    k = n // 2
    wi = rng.integers(0, len(w), size=k)
    a, b = rng.random((k, 1), dtype=np.float32), rng.random((k, 1), dtype=np.float32)
    on = w["pos"][wi, :3] + a * w["width"][wi, :3] + b * w["height"][wi, :3]
    nrm = w["n"][wi, :3]
    flip = np.sum(d[:k] * nrm, axis=1) < 0
    d[:k][flip] -= 2 * np.sum(d[:k][flip] * nrm[flip], axis=1, keepdims=True) * nrm[flip]
    o[:k] = on + d[:k] * np.float32(1e-5)
    return o, d
