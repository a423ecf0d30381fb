"""CPU checks of the host-built scene tables of the grid tier (csrc/scene_prep.cpp): structure of the head table
and a host replay of the device walk against a brute-force scan with the reference's intersects() semantics
(rectangle.c:67-95).  tests/cpu/scene_tables_check.cpp does the work; no GPU, no CUDA library involved."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "flatmatch-global-illumination_b200" / "csrc"


@pytest.fixture(scope="session")
def checker(tmp_path_factory):
    exe = tmp_path_factory.mktemp("scene_tables") / "scene_tables_check"
    subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-I", str(ROOT / "include"),
                    str(ROOT / "tests" / "cpu" / "scene_tables_check.cpp"), str(CSRC / "scene_prep.cpp"),
                    str(CSRC / "geosphere.cpp"), "-o", str(exe)], check=True)
    return exe


def run_check(checker, tmp_path, walls, windows, lights, rays, cell=0.0):
    path = tmp_path / "scene.bin"
    with open(path, "wb") as f:
        np.array([len(walls), len(windows), len(lights)], dtype="<i4").tofile(f)
        for t in (walls, windows, lights):
            np.ascontiguousarray(t).tofile(f)
    r = subprocess.run([str(checker), str(path), str(rays), str(cell)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    return r.stdout


@pytest.mark.parametrize("fixture,rays", [("scene", 40000), ("synth800", 40000), ("synth4000", 15000)])
def test_grid_tables_of_the_fixture_layouts(checker, tmp_path, request, fixture, rays):
    sc = request.getfixturevalue(fixture)
    out = run_check(checker, tmp_path, sc.walls, sc.windows, sc.lights, rays)
    assert "misc 0" in out          # parseLayout output: nothing takes the slow path


@pytest.mark.parametrize("cell", [0.35, 0.9, 2.7, 50.0])
def test_grid_tables_do_not_depend_on_the_cell_size(checker, tmp_path, scene, cell):
    run_check(checker, tmp_path, scene.walls, scene.windows, scene.lights, 15000, cell)


def test_more_z_planes_than_the_plane_table_ride_in_the_walk_lists(checker, tmp_path, fmgi):
    from test_gpu_parity import staircase_scene

    walls, windows, lights, _ = staircase_scene(fmgi)
    out = run_check(checker, tmp_path, walls, windows, lights, 40000)
    assert "planes 8/8" in out and "misc 0" not in out


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_soups_with_arbitrarily_oriented_rectangles(checker, tmp_path, fmgi, seed):
    from test_gpu_parity import random_scene

    walls, windows, lights, _ = random_scene(fmgi, seed)
    out = run_check(checker, tmp_path, walls, windows, lights, 40000)
    assert "misc 0" not in out
