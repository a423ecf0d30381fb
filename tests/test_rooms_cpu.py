"""CPU checks of the room tier's box decomposition (csrc/rooms_build.cpp): a host replay of the device traversal against a
brute-force scan with the reference's intersects() semantics (rectangle.c:67-95) on random rays.
tests/cpu/rooms_check.cpp does the work; no GPU, no CUDA library involved."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "flatmatch-global-illumination_b200" / "csrc"


@pytest.fixture(scope="session")
def rooms_checker(tmp_path_factory):
    exe = tmp_path_factory.mktemp("rooms") / "rooms_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", str(ROOT / "include"),
                    str(ROOT / "tests" / "cpu" / "rooms_check.cpp"), str(CSRC / "rooms_build.cpp"), "-o", str(exe), "-lpthread"], check=True)
    return exe


def run_check(checker, tmp_path, walls, windows, lights, rays, expect=0, env=None):
    path = tmp_path / "scene.bin"
    with open(path, "wb") as f:
        np.array([len(walls), len(windows), len(lights)], dtype="<i4").tofile(f)
        for t in (walls, windows, lights):
            np.ascontiguousarray(t).tofile(f)
    import os

    r = subprocess.run([str(checker), str(path), str(rays)], capture_output=True, text=True,
                       env=dict(os.environ, **(env or {})))
    assert r.returncode == expect, r.stdout[-3000:] + r.stderr[-1000:]
    return r.stdout


@pytest.mark.parametrize("fixture,rays,max_steps", [("scene", 100000, 2.2), ("synth800", 100000, 1.6), ("synth4000", 20000, 1.6)])
def test_rooms_of_the_fixture_layouts(rooms_checker, tmp_path, request, fixture, rays, max_steps):
    sc = request.getfixturevalue(fixture)
    out = run_check(rooms_checker, tmp_path, sc.walls, sc.windows, sc.lights, rays)
    steps = float(re.search(r"steps/ray ([0-9.]+)", out).group(1))
    assert steps < max_steps, out           # a ray in a room leaves it through ONE face: one or two boxes per ray


def test_rooms_with_many_z_planes(rooms_checker, tmp_path, fmgi):
    from test_gpu_parity import staircase_scene

    walls, windows, lights, _ = staircase_scene(fmgi)
    run_check(rooms_checker, tmp_path, walls, windows, lights, 100000)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_rooms_of_random_axis_parallel_soups(rooms_checker, tmp_path, fmgi, seed):
    from test_gpu_parity import random_scene

    walls, windows, lights, _ = random_scene(fmgi, seed, n_axis=160, n_general=0)
    run_check(rooms_checker, tmp_path, walls, windows, lights, 100000)


def test_rooms_refuse_arbitrarily_oriented_colliders(rooms_checker, tmp_path, fmgi):
    from test_gpu_parity import random_scene

    walls, windows, lights, _ = random_scene(fmgi, 1)
    out = run_check(rooms_checker, tmp_path, walls, windows, lights, 100, expect=3)
    assert "refused" in out


def test_rooms_refuse_big_random_soups(rooms_checker, tmp_path, fmgi):
    """1000 rectangles floating in space cut each other into 20 boxes per rectangle (a flat: one box per two): the
    builder gives up early and AUTO stays on the grid."""
    from test_gpu_parity import random_scene

    walls, windows, lights, _ = random_scene(fmgi, 1, n_axis=1000, n_general=0)
    out = run_check(rooms_checker, tmp_path, walls, windows, lights, 100, expect=3)
    assert "too many boxes" in out


def test_rooms_do_not_depend_on_the_number_of_builder_threads(rooms_checker, tmp_path, synth4000):
    """The kd subtrees and the face grids are built by a pool of threads and concatenated in a fixed order: every table
    the device reads is byte-identical for 1, 3 and 8 builder threads (21.5k rectangles: the threaded path)."""
    sums = set()
    for threads in ("1", "3", "8"):
        out = run_check(rooms_checker, tmp_path, synth4000.walls, synth4000.windows, synth4000.lights, 2000,
                        env={"FMGI_BUILD_THREADS": threads})
        sums.add(re.search(r"tables checksum ([0-9a-f]+)", out).group(1))
    assert len(sums) == 1, sums


def test_warp_replay_tool_builds_and_prefers_interleaved_steps(tmp_path, scene):
    """tools/rooms_warp_sim.cpp (the analysis behind profiles/r2_rooms_lanes.md) still compiles against the tables and
    reproduces the shape of the measurement: one box step per iteration costs clearly more warp instructions per ray than
    two or three, and walking every ray to its end is not better than three."""
    exe = tmp_path / "rooms_warp_sim"
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", str(ROOT / "include"), str(ROOT / "tools" / "rooms_warp_sim.cpp"),
                    str(CSRC / "rooms_build.cpp"), "-o", str(exe), "-lpthread"], check=True)
    path = tmp_path / "scene.bin"
    with open(path, "wb") as f:
        np.array([len(scene.walls), len(scene.windows), len(scene.lights)], dtype="<i4").tofile(f)
        for t in (scene.walls, scene.windows, scene.lights):
            np.ascontiguousarray(t).tofile(f)
    out = subprocess.run([str(exe), str(path), "3", "60", "120"], capture_output=True, text=True, check=True).stdout
    rows = {int(m.group(1)): float(m.group(2)) for m in re.finditer(r"^(\d+)\s+([0-9.]+)\s+[0-9.]+\s+[0-9.]+\s*$", out, re.M)}
    assert {1, 2, 3, 64} <= set(rows), out
    assert rows[1] > 1.15 * rows[3] and rows[64] >= rows[3] and abs(rows[2] / rows[3] - 1) < 0.1, out
