"""CPU suite: pins the oracle (oracle/photon_oracle.c) against outputs of the compiled reference
and checks the host-side logic that does not need a GPU."""
import hashlib
import re
from pathlib import Path

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, random_rays


def test_scene_fixture_matches_recorded_hashes(scene, facts):
    # SURVEY.md section 4 golden facts for example.png (scale 30, TILE_SIZE 200)
    assert (len(scene.walls), len(scene.windows), len(scene.lights)) == (172, 7, 3)
    assert scene.num_texels == 113964 and int(scene.base_texel_mask().sum()) == 85056
    assert hashlib.sha256(scene.walls.tobytes()).hexdigest() == facts["sha256_walls"]
    assert facts["sha256_walls"] == "c695e700f49b590b0bb4d46bdd1db94c48f740d5d4018177d551819deca827e3"
    assert hashlib.sha256(scene.windows.tobytes()).hexdigest() == facts["sha256_windows"]
    assert hashlib.sha256(scene.lights.tobytes()).hexdigest() == facts["sha256_lights"]
    assert facts["sha256_geometry_json"] == "12ea9a756f7e520ff3bf3ba9475f40cc819c4f3ac035b32b8c1cd76c4f27b268"
    assert facts["sha256_collision_map_json"] == "5ef7d1df3a40817938d61f27bce5c6c46674c35c16c07a2944d9adcd1d97e731"
    assert scene.photon_counts(100000) == facts["photon_counts_spa_100000"]
    assert sum(scene.photon_counts(100000)) == 1538348


@pytest.mark.parametrize("spa,seed,depth", [(3000, 1, 8), (3000, 7, 8), (2000, 3, 3), (2000, 5, 4), (1000, 9, 1)])
def test_oracle_reproduces_reference_atlas_bit_for_bit(oracle, scene, facts, spa, seed, depth):
    """libc rand() + BSP: same srand seed -> same raw atlas bytes as performPhotonMappingNative
    (photonmap.c:408) compiled from the reference (hashes recorded by oracle/make_golden.py)."""
    tex, st = oracle.bake(scene, spa, depth, oracle.ACCEL_BSP, oracle.RNG_LIBC, seed)
    want = facts["native_atlas_sha256"][f"spa{spa}_seed{seed}_depth{depth}"]
    assert hashlib.sha256(tex.tobytes()).hexdigest() == want
    assert st["photons"] == sum(scene.photon_counts(spa))


def test_oracle_matches_live_reference(oracle, reflib, scene):
    ref8, refd = reflib
    for spa, seed, depth in [(1500, 11, 8), (1500, 12, 2)]:
        a, _ = (ref8 if depth == 8 else refd).photonmap_native(scene, spa, seed, depth)
        b, _ = oracle.bake(scene, spa, depth, oracle.ACCEL_BSP, oracle.RNG_LIBC, seed)
        assert np.array_equal(a, b)


def test_oracle_probes_match_live_reference(oracle, reflib, scene):
    ref8, _ = reflib
    o, d = random_rays(scene, 20000, 3)
    ri, rt = ref8.closest_hit_linear(scene.walls, o, d)
    oi, ot = oracle.closest_hit(scene.walls, o, d, oracle.ACCEL_LINEAR)
    assert np.array_equal(ri, oi) and np.array_equal(rt, ot)
    rb_, rbt = ref8.closest_hit_bsp(scene.walls, o, d)
    bi, bt = oracle.closest_hit(scene.walls, o, d, oracle.ACCEL_BSP)
    base = np.where(bi >= 0, scene.walls["lightmapSetup"][np.maximum(bi, 0), 0], -1)
    assert np.array_equal(rb_, base) and np.array_equal(rbt, bt)
    rng = np.random.default_rng(5)
    for wi in (0, 17, 60, 171):
        w = scene.walls[wi]
        uv = rng.random((2000, 2), dtype=np.float32)
        pts = w["pos"][:3] + uv[:, :1] * w["width"][:3] + uv[:, 1:] * w["height"][:3]
        assert np.array_equal(ref8.tile_ids(w, pts), oracle.tile_ids(w, pts))


def test_bsp_and_linear_scan_agree(oracle, scene):
    # BSP == brute force (0 target mismatches, distance within 1e-4 relative)
    o, d = random_rays(scene, 200000, 7)
    li, lt = oracle.closest_hit(scene.walls, o, d, oracle.ACCEL_LINEAR)
    bi, bt = oracle.closest_hit(scene.walls, o, d, oracle.ACCEL_BSP)
    hit = li >= 0
    assert 0.05 < 1 - hit.mean() < 0.6
    assert (li != bi).mean() < 1e-5
    same = hit & (li == bi)
    assert np.max(np.abs(lt[same] - bt[same]) / np.maximum(lt[same], 1e-3)) < 1e-3


def test_philox_known_answers(oracle):
    # Random123 kat_vectors for philox4x32-10
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kats:
        assert oracle.philox(ctr, key).tolist() == want
    # ... and for philox2x32-10, the generator the CUDA tracer draws from (csrc/philox.cuh)
    for ctr, key, want in PHILOX2X32_KATS:
        assert oracle.philox2x32(ctr, key).tolist() == want


PHILOX2X32_KATS = [
    ([0, 0], 0, [0xFF1DAE59, 0x6CD10DF2]),
    ([0xFFFFFFFF, 0xFFFFFFFF], 0xFFFFFFFF, [0x2C3F628B, 0xAB4FD7AD]),
    ([0x243F6A88, 0x85A308D3], 0x13198A2E, [0xDD7CE038, 0xF62A4C12]),
]


def test_philox_mode_statistics_match_libc_mode(oracle, scene):
    """The Philox stream changes the samples, not the distribution: deposits/photon, mirror share
    and total energy agree with the reference's libc stream within Monte-Carlo noise."""
    spa = 20000
    a, sa = oracle.bake(scene, spa, 8, oracle.ACCEL_BSP, oracle.RNG_LIBC, 3)
    b, sb = oracle.bake(scene, spa, 8, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 3)
    assert sa["photons"] == sb["photons"]
    assert abs(sa["deposits"] / sb["deposits"] - 1) < 0.01
    assert abs(sa["mirror_bounces"] / sb["mirror_bounces"] - 1) < 0.02
    assert abs(a[:, :3].sum() / b[:, :3].sum() - 1) < 0.01
    assert 4.8 < sb["deposits"] / sb["photons"] < 5.1          # 4.968 at depth 8


def test_philox_shards_partition_the_photon_set(oracle, scene):
    """Photon-range sharding (multi-GPU): the shards' counters add up to the unsharded run and the
    summed atlases agree up to float summation order."""
    spa, depth = 4000, 4
    whole, sw = oracle.bake(scene, spa, depth, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 9)
    parts = [oracle.bake(scene, spa, depth, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 9, g, 3) for g in range(3)]
    for k in ("photons", "rays", "deposits", "mirror_bounces"):
        assert sum(p[1][k] for p in parts) == sw[k]
    total = np.sum([p[0].astype(np.float64) for p in parts], axis=0)
    assert np.allclose(total, whole, rtol=1e-5, atol=1e-3)


def test_c_abi_library_exports_every_declared_symbol(fmgi):
    """include/fmgi.h is the boundary: every function it declares must be exported by
    lib/libfmgi_cuda.so (no compute is attempted here)."""
    hdr = (ROOT / "include" / "fmgi.h").read_text()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(performGlobalIlluminationCl|fmgi_[a-z_0-9]+)\s*\(", hdr))
    declared -= {"fmgi_rect", "fmgi_geometry", "fmgi_options", "fmgi_stats", "fmgi_scene", "fmgi_status"}
    assert declared == set(fmgi.EXPORTS)
    L = fmgi.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert b"sm_100a" in L.fmgi_version()
    import ctypes

    assert ctypes.sizeof(fmgi.Geometry) == 80 and fmgi.RECT_DTYPE.itemsize == 80


def test_no_gpu_means_loud_failure_not_fallback(fmgi, scene):
    if fmgi.lib().fmgi_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(fmgi.FmgiError):
        fmgi.DeviceScene(scene.walls, scene.windows, scene.lights, scene.num_texels)
    tex = fmgi.aligned_texels(scene.num_texels)
    geo = fmgi.make_geometry(scene.walls, scene.windows, scene.lights, tex)
    with pytest.raises(fmgi.FmgiError):
        fmgi.bake(geo, 1000)


def test_product_never_touches_the_oracle():
    pkg = ROOT / "flatmatch-global-illumination_b200"
    for f in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu*")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")):
        txt = f.read_text()
        assert "refbind" not in txt and "photon_oracle" not in txt and "libfmgi_ref" not in txt, f


@pytest.mark.parametrize("fixture", ["example_scene.npz", "synth800_scene.npz", "synth4000_scene.npz"])
def test_retile_reproduces_the_reference_layout(fixture):
    """fmgi.layout.retile at TILE_SIZE 200 must give exactly the tile grids and atlas offsets
    parseLayout/createRectangleV produced (rectangle.c:24-42, parseLayout.c:512-517)."""
    import refbind
    from fmgi import layout

    sc = refbind.Scene.load(GOLDEN / fixture)
    walls, num_texels = layout.retile(sc.walls, 200.0)
    assert num_texels == sc.num_texels
    assert np.array_equal(walls["lightmapSetup"], sc.walls["lightmapSetup"])
    hi, n_hi = layout.retile(sc.walls[:200], 800.0)
    assert n_hi > 3.5 * layout.retile(sc.walls[:200], 200.0)[1]


def test_oracle_tonemap_equals_reference_saveas(oracle, reflib, scene):
    """orc_tonemap_tiles (main.c:68-79 + rectangle.c:263-336 restated) gives byte for byte the pixel
    buffer the reference's saveAs hands to write_png_file, with and without the extra floor tint."""
    import refbind

    ref8, _ = reflib
    spa = 8000
    raw, _ = oracle.bake(scene, spa, 8, oracle.ACCEL_BSP, oracle.RNG_LIBC, 5)
    norm = refbind.aligned_texels(scene.num_texels)
    norm[...] = raw
    for w in scene.walls:                      # main.c:68-79, float/double promotions as in the reference
        b, tw, th = (int(x) for x in w["lightmapSetup"][:3])
        lw = np.sqrt(np.float32((w["width"][:3].astype(np.float32) ** 2).sum(dtype=np.float32)), dtype=np.float32)
        lh = np.sqrt(np.float32((w["height"][:3].astype(np.float32) ** 2).sum(dtype=np.float32)), dtype=np.float32)
        tps = np.float32(np.float32(tw * th) / np.float32(np.float32(lw * lh) * np.float32(spa)))
        norm[b:b + tw * th, :3] = raw[b:b + tw * th, :3] * np.float32(0.35 * float(tps))
    for tint in (0, 1):
        mine = oracle.tonemap_tiles(scene, raw, spa, tint)
        off = 0
        for w in scene.walls:
            tw, th = int(w["lightmapSetup"][1]), int(w["lightmapSetup"][2])
            want = ref8.save_tile(w, norm, tint)
            assert np.array_equal(mine[off:off + 3 * tw * th], want)
            off += 3 * tw * th
        assert off == mine.size and (mine > 0).mean() > 0.9


def test_geosphere_regenerates_the_reference_direction_set(fmgi, reflib):
    """geosphere.cpp rebuilds geoSphere4 (geoSphere.c:148, the AO direction table) from the generator's
    construction (geoSphere.py): same 481 float triples as a set, whatever the order."""
    ref8, _ = reflib
    want = ref8.geosphere4()
    got = fmgi.geosphere(4)
    assert got.shape == want.shape == (481, 3)
    assert sorted(map(tuple, got.tolist())) == sorted(map(tuple, want.tolist()))
    assert (got[:, 2] > 0).all() and np.allclose(np.linalg.norm(got.astype(np.float64), axis=1), 1, atol=1e-6)


def test_oracle_ambient_occlusion_equals_reference(oracle, reflib, fmgi):
    """photonmap.c:436-491 restated: bit-equal to performAmbientOcclusionNative on a small closed room
    (with the reference's own table order, since the sums are float)."""
    import refbind
    from test_gpu_parity import staircase_scene

    ref8, _ = reflib
    walls, windows, lights, num_texels = staircase_scene(fmgi)
    sc = refbind.Scene(walls, windows, lights, num_texels)
    want = ref8.ambient_occlusion_native(sc)
    got = oracle.ambient_occlusion(sc, ref8.geosphere4(), oracle.ACCEL_BSP)
    assert np.array_equal(want, got)
    assert 0.1 < want[sc.base_texel_mask(), 0].mean() < 10 and not want[~sc.base_texel_mask()].any()
