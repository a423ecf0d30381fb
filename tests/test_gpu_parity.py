"""GPU suite (-m gpu): the CUDA path, called through the C ABI, against the oracle and against
fixtures produced by the compiled reference.

Bars (BASELINE.json north_star):
  * index/integer work — closest-hit wall index, texel index, Philox words, photon budgets,
    counters under sharding: bit-exact;
  * radiance: per-texel relative RMS of the normalised luminance below 2 % and total deposited
    energy within 0.1 % of the reference's native CPU path on the same layout/depth.
"""
import ctypes

import numpy as np
import pytest

from conftest import GOLDEN, outside_rays, random_rays

pytestmark = pytest.mark.gpu

LUMA = np.array([0.2126, 0.7152, 0.0722])  # rectangle.c:277


def device_atlas(num_texels):
    import torch

    return torch.zeros((num_texels, 4), dtype=torch.float32, device="cuda")


def edge_distance(walls, idx, p):
    """Distance (m) of point p from the nearest edge of walls[idx], measured in the rectangle's own (u, v) frame."""
    w = walls[idx]
    wd, ht = w["width"][:3].astype(np.float64), w["height"][:3].astype(np.float64)
    e = p - w["pos"][:3].astype(np.float64)
    lw, lh = np.linalg.norm(wd), np.linalg.norm(ht)
    u, v = e @ wd / lw, e @ ht / lh
    return min(abs(u), abs(lw - u), abs(v), abs(lh - v))


def split_mismatches(walls, o, d, gi, gt, ci, ct, eps=2e-5):
    """Closest-hit index mismatches are legitimate only where the ray passes through a rectangle EDGE: one side's
    rounding counts the point as inside (edges are inclusive, rectangle.c:93), the other's as outside, or two
    rectangles share the edge and tie (the reference's strict `<` keeps the lowest index, photonmap.cl:199).
    Returns (mismatches at an edge, mismatches anywhere else): the second number must be zero."""
    at_edge = elsewhere = 0
    for r in np.nonzero(gi != ci)[0]:
        near = []
        for idx, t in ((gi[r], gt[r]), (ci[r], ct[r])):
            if idx >= 0:
                near.append(edge_distance(walls, idx, o[r].astype(np.float64) + float(t) * d[r].astype(np.float64)))
        if near and min(near) < eps:
            at_edge += 1
        else:
            elsewhere += 1
    return at_edge, elsewhere


def gpu_bake(dev_scene, spa, **opts):
    import torch

    atlas = device_atlas(dev_scene.num_texels)
    dev_scene.trace(atlas.data_ptr(), spa, stream=torch.cuda.current_stream().cuda_stream, **opts)
    st = dev_scene.sync()
    return atlas.cpu().numpy(), st


# ---- bit-exact pieces -----------------------------------------------------------------------------


def test_philox_on_device(fmgi, oracle):
    kats = [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ]
    for ctr, key, want in kats:
        assert fmgi.philox(ctr, key).tolist() == want
    rng = np.random.default_rng(1)
    for _ in range(20):
        ctr, key = rng.integers(0, 2**32, 4, dtype=np.uint64), rng.integers(0, 2**32, 2, dtype=np.uint64)
        assert np.array_equal(fmgi.philox(ctr, key), oracle.philox(ctr, key))
    # the tracer's own generator: Philox2x32-10 (Random123 known answers, then random blocks vs the oracle)
    from test_oracle import PHILOX2X32_KATS

    for ctr, key, want in PHILOX2X32_KATS:
        assert fmgi.philox2x32(ctr, key).tolist() == want
    for _ in range(20):
        ctr, key = rng.integers(0, 2**32, 2, dtype=np.uint64), int(rng.integers(0, 2**32))
        assert np.array_equal(fmgi.philox2x32(ctr, key), oracle.philox2x32(ctr, key))


TIERS = {"soup": 1, "grid": 2, "rooms": 4}          # fmgi.TIER_*


@pytest.fixture(params=["soup_planes", "soup", "grid", "rooms"])
def any_tier_scene(request, dev_scene_soup, dev_scene_soup_plain, dev_scene_grid, dev_scene_rooms):
    """The four closest-hit kernels on the same flat: brute-force soup with the horizontal rectangles in the
    plane tables, plain brute-force soup, floor-plan grid, box decomposition (room tier: what AUTO picks here)."""
    s = {"soup_planes": dev_scene_soup, "soup": dev_scene_soup_plain, "grid": dev_scene_grid,
         "rooms": dev_scene_rooms}[request.param]
    s.tier_name = request.param
    return s


def test_tier_fixtures_run_the_kernels_they_name(dev_scene, dev_scene_soup, dev_scene_soup_plain, dev_scene_grid,
                                                 dev_scene_rooms):
    spa = 2000
    assert gpu_bake(dev_scene_soup, spa, count_tests=1)[1]["tier"] == 1
    assert gpu_bake(dev_scene_soup_plain, spa)[1]["tier"] == 1
    assert gpu_bake(dev_scene_grid, spa)[1]["tier"] == 2
    assert gpu_bake(dev_scene_rooms, spa)[1]["tier"] == 4
    assert gpu_bake(dev_scene, spa)[1]["tier"] == 4           # AUTO: 172 axis-parallel colliders > 64 -> room tier
    t_rooms = gpu_bake(dev_scene_rooms, spa, count_tests=1)[1]
    assert 1.0 < t_rooms["rect_tests"] / t_rooms["rays"] < 8.0     # boxes crossed + face grids looked up per ray
    # rectangle tests per ray tell the kernels apart: plain soup scans every pair block of the ray's sign
    # (94 tests), soup + planes only the x / y lists plus the counted plane lookups, the grid a handful
    t_plain = gpu_bake(dev_scene_soup_plain, spa, count_tests=1)[1]
    t_planes = gpu_bake(dev_scene_soup, spa, count_tests=1)[1]
    t_grid = gpu_bake(dev_scene_grid, spa, count_tests=1)[1]
    per_ray = [t["rect_tests"] / t["rays"] for t in (t_plain, t_planes, t_grid)]
    assert per_ray[0] > per_ray[1] > 20 > per_ray[2] > 2, per_ray


def test_closest_hit_matches_oracle(any_tier_scene, oracle, scene, record):
    """>= 1e6 rays: same wall index as the reference's linear scan with intersects()
    (rectangle.c:67, photonmap.cl:194-206); distance within 1e-4 relative (expect
    < 1e-6 index mismatches, ties excepted)."""
    o, d = random_rays(scene, 1_000_000, 11)
    gi, gt = any_tier_scene.closest_hit(o, d)
    ci, ct = oracle.closest_hit(scene.walls, o, d, oracle.ACCEL_LINEAR)
    mism = gi != ci
    both = (gi >= 0) & ~mism
    rel = np.abs(gt[both] - ct[both]) / np.maximum(ct[both], 1e-4)
    at_edge, elsewhere = split_mismatches(scene.walls, o, d, gi, gt, ci, ct)
    record(f"closest_hit_example_{any_tier_scene.tier_name}", rays=len(o), index_mismatches=int(mism.sum()),
           at_rectangle_edge=at_edge, elsewhere=elsewhere, max_rel_distance_error=float(rel.max()),
           hit_share=float(both.mean()))
    # measured on a B200 (profiles/parity_r2.json): 0-3 mismatches per 1e6 rays, every one of them a ray through
    # a rectangle edge (half of these rays start ON a wall, so edges are met far more often than by chance)
    assert elsewhere == 0 and at_edge <= 8, f"{elsewhere} mismatches away from any edge, {at_edge} at edges"
    assert both.mean() > 0.5
    assert rel.max() < 1e-4
    assert np.all(np.isinf(gt[gi < 0]))


@pytest.mark.parametrize("which", ["example", "synth800"])
def test_closest_hit_from_outside_the_bounding_box(any_tier_scene, fmgi, oracle, scene, synth800, record, which):
    """Origins outside the walls' bounding box in x, y or z - an emitter above the wall z range, a probe ray from
    the garden: rays that enter the flat find the reference's hit, rays that start outside and move away hit
    nothing, and the grid walk ends at once instead of stepping off its table (t_exit clamped to 0)."""
    if which == "example":
        sc, dev = scene, any_tier_scene
    else:
        if any_tier_scene.tier_name not in ("grid", "rooms"):
            pytest.skip("synth800 runs through the grid and the room tier")
        sc, dev = synth800, fmgi.DeviceScene(synth800.walls, synth800.windows, synth800.lights, synth800.num_texels,
                                             tier=TIERS[any_tier_scene.tier_name])
    o, d = outside_rays(sc, 400_000, 31)
    gi, gt = dev.closest_hit(o, d)
    ci, ct = oracle.closest_hit(sc.walls, o, d, oracle.ACCEL_LINEAR)
    mism = gi != ci
    both = (gi >= 0) & ~mism
    rel = np.abs(gt[both] - ct[both]) / np.maximum(ct[both], 1e-4)
    at_edge, elsewhere = split_mismatches(sc.walls, o, d, gi, gt, ci, ct)
    record(f"closest_hit_outside_{which}_{any_tier_scene.tier_name}", rays=len(o), index_mismatches=int(mism.sum()),
           at_rectangle_edge=at_edge, elsewhere=elsewhere, max_rel_distance_error=float(rel.max()),
           hit_share=float(both.mean()))
    assert elsewhere == 0 and at_edge <= 4, f"{elsewhere} mismatches away from any edge, {at_edge} at edges"
    assert 0.02 < both.mean() < 0.9                 # some rays enter the flat, many miss it
    assert rel.max() < 1e-4
    assert np.all(np.isinf(gt[gi < 0]))
    if which != "example":
        dev.close()


def test_texel_index_is_bit_exact(dev_scene, oracle, scene):
    """getTileIdAt (rectangle.c:205): identical index for random points on every wall, including
    points exactly on and slightly outside the edges (clamping)."""
    rng = np.random.default_rng(2)
    total = 0
    for wi in range(len(scene.walls)):
        w = scene.walls[wi]
        uv = rng.random((6000, 2), dtype=np.float32) * np.float32(1.02) - np.float32(0.01)
        uv[:50] = rng.integers(0, 2, (50, 2)).astype(np.float32)          # corners
        tw, th = int(w["lightmapSetup"][1]), int(w["lightmapSetup"][2])
        uv[50:150, 0] = rng.integers(0, tw + 1, 100) / np.float32(tw)     # texel borders
        uv[150:250, 1] = rng.integers(0, th + 1, 100) / np.float32(th)
        pts = (w["pos"][:3] + uv[:, :1] * w["width"][:3] + uv[:, 1:] * w["height"][:3]).astype(np.float32)
        got = dev_scene.tile_ids(np.full(len(pts), wi, dtype=np.int32), pts)
        want = oracle.tile_ids(w, pts)
        assert np.array_equal(got, want), f"wall {wi}"
        total += len(pts)
    assert total >= 1_000_000


def test_sampler_moments_and_sky_fold(fmgi):
    """vector3_cl.c:102-149: cosine-weighted hemisphere, E[n.d] = 2/3, zero mean tangentially;
    the window sampler folds onto +U, which is -z for any horizontal normal (light only goes down)."""
    n = 400_000
    for normal in ([0, 0, 1], [0, 0, -1], [1, 0, 0], [0, -0.99999994, 0], [0.6, 0.8, 0]):
        nn = np.array(normal, dtype=np.float32)
        d = fmgi.sample_dirs(nn, False, 5, n).astype(np.float64)
        assert np.allclose(np.linalg.norm(d, axis=1), 1, atol=2e-6)
        cosn = d @ (nn / np.linalg.norm(nn))
        assert cosn.min() >= -1e-6
        assert abs(cosn.mean() - 2 / 3) < 2e-3
        tang = d - np.outer(cosn, nn / np.linalg.norm(nn))
        assert np.all(np.abs(tang.mean(axis=0)) < 3e-3)
    for normal in ([1, 0, 0], [0, -0.99999994, 0], [0.6, 0.8, 0]):
        d = fmgi.sample_dirs(np.array(normal, dtype=np.float32), True, 6, n)
        assert d[:, 2].max() <= 1e-6          # never upwards
        assert abs((d.astype(np.float64) @ np.array(normal) / np.linalg.norm(normal)).mean() - 2 / 3) < 2e-3


def test_photon_paths_match_oracle(any_tier_scene, oracle, scene, record):
    dev_scene = any_tier_scene
    """Same Philox sub-streams -> the same sequence of deposited texels, photon by photon.  The
    oracle evaluates sqrt/sin/cos in double like the reference, the kernel in float with SFU
    sin/cos, so a tiny share of paths may part ways at a texel border."""
    depth, seed, count = 8, 77, 40000
    for e in (0, 3, 8):
        got = dev_scene.paths(e, depth, seed, 5, count)
        want = oracle.trace_paths(scene, e, depth, seed, 5, count)
        same = np.all(got == want, axis=1)
        record(f"paths_example_{dev_scene.tier_name}_emitter{e}", photons=count, depth=depth,
               identical_paths=float(same.mean()), identical_first_bounce=float((got[:, 0] == want[:, 0]).mean()))
        assert same.mean() > 0.995, f"emitter {e}: {same.mean():.5f} identical paths"
        # first bounce depends only on emission: stricter
        assert (got[:, 0] == want[:, 0]).mean() > 0.9995


def test_budgets_and_counters_are_exact(any_tier_scene, oracle, scene):
    dev_scene = any_tier_scene
    """Photon budget per emitter follows photonmap.c:414-418; counters agree with the oracle run
    on the same Philox streams (rays/deposits may differ by the few border paths above)."""
    spa, depth = 20000, 8
    assert dev_scene.photon_count(spa) == sum(scene.photon_counts(spa))
    atlas, st = gpu_bake(dev_scene, spa, max_depth=depth, seed=5)
    _, so = oracle.bake(scene, spa, depth, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 5)
    assert st["photons"] == so["photons"] == sum(scene.photon_counts(spa))
    assert abs(st["deposits"] - so["deposits"]) <= 2e-4 * so["deposits"]
    assert abs(st["rays"] - so["rays"]) <= 2e-4 * so["rays"]
    assert abs(st["mirror_bounces"] - so["mirror_bounces"]) <= 1e-3 * so["mirror_bounces"]
    assert st["rays"] >= st["deposits"] >= st["mirror_bounces"]
    assert np.all(atlas[:, 3] == 0)
    assert np.all(atlas[~scene.base_texel_mask()] == 0)      # mip slots are never written


def test_rectangle_test_counter_is_opt_in_and_changes_nothing(dev_scene, dev_scene_grid):
    """fmgi_options.count_tests selects the counting instantiation of the trace kernel (two more instructions per
    walk step): same photons, rays, deposits and atlas; rect_tests only then.  The grid does a handful of tests
    per ray where the brute-force scan of this scene did 67."""
    spa = 40000
    plain, sp = gpu_bake(dev_scene_grid, spa, max_depth=5, seed=21)
    counted, sc = gpu_bake(dev_scene_grid, spa, max_depth=5, seed=21, count_tests=1)
    for k in ("photons", "rays", "deposits", "mirror_bounces"):
        assert sp[k] == sc[k]
    assert sp["rect_tests"] == 0
    assert 2.0 < sc["rect_tests"] / sc["rays"] < 10.0
    assert np.allclose(plain, counted, rtol=1e-5, atol=1e-2)
    auto, sa = gpu_bake(dev_scene, spa, max_depth=5, seed=21)          # AUTO picks the room tier for this flat
    # same photons; the handful of rays through rectangle edges may end differently in the two structures
    assert sa["tier"] == 4 and sa["photons"] == sp["photons"] and abs(sa["rays"] - sp["rays"]) <= 1e-5 * sp["rays"]


def test_room_tier_deposits_the_grid_tiers_energy(dev_scene_rooms, dev_scene_grid, record):
    """Same photons, same streams: the two closest-hit structures must find the same hits - the room tier neither loses
    photons at corners or beside window niches (it relocates an origin that lies behind a wall, rooms_walk) nor
    invents any.  Total energy to 2 ppm, ray counts to 1e-5 (rays through rectangle edges may end differently)."""
    spa = 2_000_000                                   # 3e7 photons
    a_r, s_r = gpu_bake(dev_scene_rooms, spa, max_depth=8, seed=77)
    a_g, s_g = gpu_bake(dev_scene_grid, spa, max_depth=8, seed=77)
    assert s_r["tier"] == 4 and s_g["tier"] == 2 and s_r["photons"] == s_g["photons"]
    e_r, e_g = a_r[:, :3].sum(dtype=np.float64), a_g[:, :3].sum(dtype=np.float64)
    record("energy_rooms_vs_grid_depth8", photons=int(s_r["photons"]), energy_rel_diff=float(e_r / e_g - 1),
           rays_rel_diff=float(s_r["rays"] / s_g["rays"] - 1), deposits_rel_diff=float(s_r["deposits"] / s_g["deposits"] - 1))
    assert abs(e_r / e_g - 1) < 2e-6
    assert abs(s_r["rays"] / s_g["rays"] - 1) < 1e-5 and abs(s_r["deposits"] / s_g["deposits"] - 1) < 1e-5


def test_deposit_peak_probe(fmgi, scene):
    """fmgi_probe_deposit_peak (the deposit roofline of SURVEY.md 8d-ii): the bare RED.E.ADD.F32x4 at uniform-random
    texels.  An L2-resident footprint sustains an order of magnitude more than the bake deposits (measured
    1.9e11/s on a B200); a footprint far beyond L2 is bound by random 32-byte sector traffic (2.6e10/s at 457 MB)."""
    small = fmgi.deposit_peak(scene.num_texels, 100_000_000)
    large = fmgi.deposit_peak(28_591_084, 100_000_000)
    assert small > 2e10 and large > 2e9
    assert large < small


def test_small_bake_matches_oracle_texel_by_texel(dev_scene, oracle, scene):
    """Same streams, small budget: atlases agree texel by texel except where a border path moved
    one deposit to the neighbouring texel; energy agrees to 1e-4."""
    spa, depth = 30000, 4
    atlas, _ = gpu_bake(dev_scene, spa, max_depth=depth, seed=21)
    want, _ = oracle.bake(scene, spa, depth, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 21)
    assert abs(atlas[:, :3].sum(dtype=np.float64) / want[:, :3].sum(dtype=np.float64) - 1) < 1e-4
    differing = np.any(np.abs(atlas[:, :3] - want[:, :3]) > 1e-3 * np.maximum(want[:, :3], 1), axis=1)
    assert differing.mean() < 2e-3


# ---- statistical parity against the reference's native CPU path -----------------------------------------


def parity_stats(lum_gpu, lum_ref, floor_frac=0.05):
    """S  = RMS(L_gpu - L_ref) / mean(L_ref) over base texels;
    S' = sqrt(mean(((L_gpu - L_ref)/L_ref)^2)) over texels brighter than floor_frac * mean."""
    s = np.sqrt(np.mean((lum_gpu - lum_ref) ** 2)) / lum_ref.mean()
    lit = lum_ref > floor_frac * lum_ref.mean()
    sp = np.sqrt(np.mean(((lum_gpu[lit] - lum_ref[lit]) / lum_ref[lit]) ** 2))
    return s, sp, lit.mean()


@pytest.mark.parametrize("depth,photons,tier", [(8, 1.0e9, "rooms"), (3, 1.0e9, "rooms"), (8, 1.0e9, "grid"),
                                                 (3, 1.0e9, "grid"), (8, 1.0e9, "soup_planes"), (3, 1.0e9, "soup")])
def test_radiance_parity_with_native_reference(dev_scene_rooms, dev_scene_grid, dev_scene_soup, dev_scene_soup_plain, scene,
                                               record, depth, photons, tier):
    """BASELINE.json: per-texel relative RMS < 2 % after the reference's normalisation and total
    deposited energy within 0.1 %, against performPhotonMappingNative (photonmap.c:408) at the
    same depth.  The reference side is the committed fixture (6.15e8 photons over 8 seeded
    processes of the compiled reference, two independent halves)."""
    z = np.load(GOLDEN / f"example_native_depth{depth}.npz")
    assert int(z["depth"]) == depth
    lum_a, lum_b = z["lum_a"].astype(np.float64), z["lum_b"].astype(np.float64)
    lum_ref = 0.5 * (lum_a + lum_b)
    area = sum(scene.photon_counts(1_000_000)) / 1e6
    spa = int(photons / area)
    dev_scene = {"rooms": dev_scene_rooms, "grid": dev_scene_grid, "soup_planes": dev_scene_soup,
                 "soup": dev_scene_soup_plain}[tier]
    atlas, st = gpu_bake(dev_scene, spa, max_depth=depth, seed=2024)
    mask = scene.base_texel_mask()
    lum_gpu = ((atlas[:, :3].astype(np.float64) @ LUMA) * scene.normalisation(spa))[mask]

    # noise floor measured on the reference itself: two independent halves
    s_ab, sp_ab, _ = parity_stats(lum_a, lum_b)
    s, sp, lit = parity_stats(lum_gpu, lum_ref)
    print(f"depth {depth}: S={s:.4%} S'={sp:.4%} (lit share {lit:.3f}); reference half-vs-half "
          f"S={s_ab:.4%} S'={sp_ab:.4%}; gpu photons {st['photons']:.3e}")
    e_ref = 0.5 * (z["rgb_total_a"] / float(z["spa_a"]) + z["rgb_total_b"] / float(z["spa_b"]))
    e_gpu = atlas[:, :3].sum(axis=0, dtype=np.float64) / spa
    rel = np.abs(e_gpu / e_ref - 1)
    print(f"energy per unit density rel. diff {rel}")
    record(f"radiance_example_depth{depth}_{tier}", gpu_photons=int(st["photons"]), S=float(s), S_per_texel=float(sp),
           reference_half_vs_half_S_per_texel=float(sp_ab), energy_rel_diff=[float(x) for x in e_gpu / e_ref - 1])
    assert sp < 0.02, f"per-texel relative RMS {sp:.4%}"
    assert s < 0.02
    # a systematic error would not shrink with photon count: GPU-vs-reference must not exceed the
    # reference's own half-vs-half scatter (each half has half the reference's photons)
    assert sp < sp_ab * 1.05
    assert np.all(rel < 1e-3)


# ---- the drop-in boundary ---------------------------------------------------------------------------------


def test_perform_global_illumination_cl_accumulates_into_host_atlas(fmgi, scene, oracle):
    """global_illumination_cl.h:10 semantics: texels = initial contents + raw deposits
    (CL_MEM_COPY_HOST_PTR, global_illumination_cl.c:295,312); lane 3 and mip slots untouched."""
    tex = fmgi.aligned_texels(scene.num_texels)
    rng = np.random.default_rng(0)
    tex[...] = rng.random(tex.shape, dtype=np.float32)
    before = tex.copy()
    geo = fmgi.make_geometry(scene.walls, scene.windows, scene.lights, tex)
    spa = 10000
    fmgi.perform_global_illumination_cl(geo, spa)
    delta = tex.astype(np.float64) - before
    mask = scene.base_texel_mask()
    assert np.array_equal(tex[:, 3], before[:, 3])
    assert np.array_equal(tex[~mask], before[~mask])
    want, so = oracle.bake(scene, spa, 8, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 1)
    assert abs(delta[:, :3].sum() / want[:, :3].sum(dtype=np.float64) - 1) < 2e-4


def test_bake_is_deterministic_and_shards_add_up(fmgi, dev_scene, scene):
    spa, depth = 50000, 5
    a, sa = gpu_bake(dev_scene, spa, max_depth=depth, seed=3)
    b, sb = gpu_bake(dev_scene, spa, max_depth=depth, seed=3)
    for k in ("photons", "rays", "deposits", "mirror_bounces"):
        assert sa[k] == sb[k]
    assert np.allclose(a, b, rtol=1e-5, atol=1e-2)          # float atomics commute up to rounding
    parts = [gpu_bake(dev_scene, spa, max_depth=depth, seed=3, shard=g, num_shards=4) for g in range(4)]
    for k in ("photons", "rays", "deposits", "mirror_bounces"):
        assert sum(p[1][k] for p in parts) == sa[k]
    total = np.sum([p[0].astype(np.float64) for p in parts], axis=0)
    assert np.allclose(total, a, rtol=1e-5, atol=1e-2)
    c, _ = gpu_bake(dev_scene, spa, max_depth=depth, seed=4)
    assert not np.allclose(a, c, rtol=1e-3, atol=1.0)


@pytest.mark.parametrize("deposit", [0, 1, 2])
def test_deposit_variants_agree(dev_scene, deposit):
    spa = 40000
    ref, sr = gpu_bake(dev_scene, spa, max_depth=6, seed=9, deposit=0)
    got, sg = gpu_bake(dev_scene, spa, max_depth=6, seed=9, deposit=deposit)
    assert sr["deposits"] == sg["deposits"]
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-2)


def test_edge_cases(fmgi, scene):
    empty = np.zeros(0, dtype=fmgi.RECT_DTYPE)
    # no emitters: nothing happens
    s = fmgi.DeviceScene(scene.walls, empty, empty, scene.num_texels)
    atlas, st = gpu_bake(s, 100000)
    assert st["photons"] == 0 and not atlas.any()
    s.close()
    # no colliders: every photon leaves (photonmap.c:200-201)
    s = fmgi.DeviceScene(empty, scene.windows, scene.lights, scene.num_texels)
    atlas, st = gpu_bake(s, 1000)
    assert st["photons"] == sum(scene.photon_counts(1000)) and st["rays"] == st["photons"]
    assert st["deposits"] == 0 and not atlas.any()
    s.close()
    # zero density
    s = fmgi.DeviceScene(scene.walls, scene.windows, scene.lights, scene.num_texels)
    atlas, st = gpu_bake(s, 0)
    assert st["photons"] == 0
    # depth 1: one deposit at most per photon
    atlas, st = gpu_bake(s, 5000, max_depth=1)
    assert st["rays"] == st["photons"] and st["deposits"] <= st["photons"]
    s.close()
    # a wall whose tile range leaves the atlas is rejected, not traced
    bad = scene.walls.copy()
    bad["lightmapSetup"][0, 0] = scene.num_texels
    with pytest.raises(fmgi.FmgiError):
        fmgi.DeviceScene(bad, scene.windows, scene.lights, scene.num_texels)


def rotated_scene(scene, angle_deg=27.0):
    """The example flat rotated about the z axis: no wall is axis parallel any more, so every
    collider takes the general-rectangle path."""
    a = np.deg2rad(angle_deg)
    rot = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], dtype=np.float64)

    def turn(rects):
        out = rects.copy()
        for f in ("pos", "width", "height"):
            out[f][:, :3] = (rects[f][:, :3].astype(np.float64) @ rot.T).astype(np.float32)
        # normal as createRectangleV builds it: normalized(cross(height, width)) (rectangle.c:22)
        c = np.cross(out["height"][:, :3].astype(np.float32), out["width"][:, :3].astype(np.float32))
        out["n"][:, :3] = (c / np.linalg.norm(c, axis=1, keepdims=True)).astype(np.float32)
        return out

    return turn(scene.walls), turn(scene.windows), turn(scene.lights)


@pytest.mark.parametrize("tier", ["soup", "grid"])
def test_general_rectangles(fmgi, oracle, scene, tier):
    import refbind

    walls, windows, lights = rotated_scene(scene)
    rs = refbind.Scene(walls, windows, lights, scene.num_texels)
    # the room tier needs axis-parallel colliders: asked for explicitly it refuses, AUTO falls back to the grid
    with pytest.raises(fmgi.FmgiError, match="room tier"):
        fmgi.DeviceScene(rs.walls, rs.windows, rs.lights, rs.num_texels, tier=fmgi.TIER_ROOMS)
    auto = fmgi.DeviceScene(rs.walls, rs.windows, rs.lights, rs.num_texels)
    assert gpu_bake(auto, 500)[1]["tier"] == fmgi.TIER_GRID
    auto.close()
    s = fmgi.DeviceScene(rs.walls, rs.windows, rs.lights, rs.num_texels,
                         tier=fmgi.TIER_SOUP if tier == "soup" else fmgi.TIER_GRID)
    o, d = random_rays(rs, 200_000, 4)
    gi, gt = s.closest_hit(o, d)
    ci, ct = oracle.closest_hit(rs.walls, o, d, oracle.ACCEL_LINEAR)
    assert (gi != ci).mean() < 2e-4
    both = (gi >= 0) & (gi == ci)
    assert np.max(np.abs(gt[both] - ct[both]) / np.maximum(ct[both], 1e-3)) < 1e-3
    spa = 20000
    atlas, st = gpu_bake(s, spa, max_depth=6, seed=8)
    want, so = oracle.bake(rs, spa, 6, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 8)
    assert st["photons"] == so["photons"]
    assert abs(st["deposits"] / so["deposits"] - 1) < 2e-3
    assert abs(atlas[:, :3].sum(dtype=np.float64) / want[:, :3].sum(dtype=np.float64) - 1) < 2e-3
    s.close()


def test_in_library_multi_gpu_bake(fmgi, scene):
    """fmgi_bake with num_gpus > 1: one host thread per GPU, disjoint photon ranges; GPU g folds slice g of every
    GPU's atlas (reading the peers' copies over NVLink peer mappings) onto the caller's values of that slice and
    writes it back over its own PCIe link."""
    n = fmgi.lib().fmgi_device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    spa, depth = 200000, 5
    tex1 = fmgi.aligned_texels(scene.num_texels)
    st1 = fmgi.bake(fmgi.make_geometry(scene.walls, scene.windows, scene.lights, tex1), spa, max_depth=depth, seed=5)
    for g in sorted({2, min(n, 4), n}):
        texg = fmgi.aligned_texels(scene.num_texels)
        stg = fmgi.bake(fmgi.make_geometry(scene.walls, scene.windows, scene.lights, texg), spa, max_depth=depth,
                        seed=5, num_gpus=g)
        assert stg["num_gpus"] == g
        for k in ("photons", "rays", "deposits", "mirror_bounces"):
            assert stg[k] == st1[k], (g, k)
        # fp32 sums in another order: the hottest texel (2.2e4, ~1400 deposits) moves by 1.4e-5 relative
        assert np.allclose(texg, tex1, rtol=5e-5, atol=5e-2)
    # the caller's current device survives the call, whichever device the bake starts on, and a bake that
    # starts on GPU 1 uses distinct GPUs (1, 2, ... wrapping to 0), adds onto the caller's values and leaves
    # lane 3 / the mip slots alone
    import torch

    torch.cuda.set_device(n - 1)
    rng = np.random.default_rng(3)
    texd = fmgi.aligned_texels(scene.num_texels)
    texd[...] = rng.random(texd.shape, dtype=np.float32)
    before = texd.copy()
    std = fmgi.bake(fmgi.make_geometry(scene.walls, scene.windows, scene.lights, texd), spa, max_depth=depth, seed=5,
                    num_gpus=2, device=1)
    assert torch.cuda.current_device() == n - 1
    torch.cuda.set_device(0)
    assert std["num_gpus"] == 2 and std["deposits"] == st1["deposits"]
    assert np.allclose(texd[:, :3] - before[:, :3], tex1[:, :3], rtol=1e-4, atol=5e-2)
    assert np.array_equal(texd[:, 3], before[:, 3])
    mask = scene.base_texel_mask()
    assert np.array_equal(texd[~mask], before[~mask])


# ---- grid tier on the synthetic multi-room layouts (BASELINE.json configs[2]) -------------------------------


@pytest.mark.parametrize("tier", ["soup", "grid", "rooms"])
def test_synth800_closest_hit_and_bake(fmgi, oracle, synth800, tier):
    sc = synth800
    s = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels, tier=TIERS[tier])
    o, d = random_rays(sc, 300_000, 21)
    gi, gt = s.closest_hit(o, d)
    ci, ct = oracle.closest_hit(sc.walls, o, d, oracle.ACCEL_LINEAR)
    assert (gi != ci).mean() < 1e-5
    both = (gi >= 0) & (gi == ci)
    assert np.max(np.abs(gt[both] - ct[both]) / np.maximum(ct[both], 1e-4)) < 1e-4
    spa, depth = 3000, 5
    atlas, st = gpu_bake(s, spa, max_depth=depth, seed=3)
    want, so = oracle.bake(sc, spa, depth, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 3)
    assert st["photons"] == so["photons"]
    assert abs(st["deposits"] - so["deposits"]) <= 3e-4 * so["deposits"]
    assert abs(atlas[:, :3].sum(dtype=np.float64) / want[:, :3].sum(dtype=np.float64) - 1) < 3e-4
    e = len(sc.windows) + 2                     # a ceiling light
    got = s.paths(e, depth, 9, 0, 8000)
    ref = oracle.trace_paths(sc, e, depth, 9, 0, 8000)
    assert np.all(got == ref, axis=1).mean() > 0.995
    s.close()


def test_synth4000_grid_tier(fmgi, oracle, synth4000):
    """~21.5k rectangles, 1228 emitters, 0.46 GB atlas: AUTO picks the grid tier."""
    sc = synth4000
    s = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels)
    o, d = random_rays(sc, 40_000, 5)
    gi, gt = s.closest_hit(o, d)
    ci, ct = oracle.closest_hit(sc.walls, o, d, oracle.ACCEL_LINEAR)
    assert (gi != ci).mean() < 5e-5
    both = (gi >= 0) & (gi == ci)
    assert np.max(np.abs(gt[both] - ct[both]) / np.maximum(ct[both], 1e-4)) < 1e-4
    spa, depth = 25, 4                         # the oracle's linear scan costs 21.5k tests per ray
    atlas, st = gpu_bake(s, spa, max_depth=depth, seed=3)
    assert st["tier"] == fmgi.TIER_ROOMS
    assert st["photons"] == sum(sc.photon_counts(spa))
    want, so = oracle.bake(sc, spa, depth, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 3)
    assert abs(st["deposits"] - so["deposits"]) <= 1e-3 * so["deposits"] + 2
    assert abs(atlas[:, :3].sum(dtype=np.float64) / want[:, :3].sum(dtype=np.float64) - 1) < 1e-3
    s.close()


@pytest.mark.parametrize("k", [2, 3, 4])
def test_pooled_kernel_matches_k_trace(fmgi, scene, synth800, monkeypatch, k):
    """trace_pool.cuh (opt-in, FMGI_POOL_K rays per lane): the same photon streams walked from a per-warp pool in
    shared memory - identical counters, atlas equal up to the order of the float atomics, on example.png and on
    the ceiling-lit synth800 layout."""
    for sc, spa, depth in ((scene, 400_000, 8), (synth800, 150_000, 4)):
        monkeypatch.setenv("FMGI_POOL_K", "0")
        classic = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels, tier=fmgi.TIER_GRID)
        monkeypatch.setenv("FMGI_POOL_K", str(k))
        pooled = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels, tier=fmgi.TIER_GRID)
        a, sa = gpu_bake(classic, spa, max_depth=depth, seed=31)
        b, sb = gpu_bake(pooled, spa, max_depth=depth, seed=31)
        assert sa["pool_rays"] == 0 and sb["pool_rays"] == k
        for key in ("photons", "rays", "deposits", "mirror_bounces"):
            assert sa[key] == sb[key], key
        assert np.allclose(a, b, rtol=1e-4, atol=0.5)
        assert abs(b[:, :3].sum(dtype=np.float64) / a[:, :3].sum(dtype=np.float64) - 1) < 1e-6
        # a bake too small to keep the pools busy falls back to k_trace
        _, ss = gpu_bake(pooled, 1000, max_depth=depth, seed=31)
        assert ss["pool_rays"] == 0
        classic.close(); pooled.close()


def test_device_grid_table_equals_host(fmgi, scene, synth800, synth4000, monkeypatch):
    """csrc/grid_build.cuh assembles the floor-plan grid table on the GPU for scenes of 2048 colliders and more
    (count / scan / scatter / per-list ordering / layout).  It must be the table scene_prep.cpp builds on the
    host, bit for bit: flats, the 21.5k-rectangle layout, a staircase with more z planes than the plane table
    (misc records in the walk lists) and a random soup with arbitrarily oriented rectangles."""
    import refbind

    cases = [("example", scene), ("synth800", synth800), ("synth4000", synth4000)]
    w, wi, li, n = staircase_scene(fmgi)
    cases.append(("staircase", refbind.Scene(w, wi, li, n)))
    w, wi, li, n = random_scene(fmgi, 5)
    cases.append(("random", refbind.Scene(w, wi, li, n)))
    for name, sc in cases:
        tables = {}
        for how in ("host", "device"):
            monkeypatch.setenv("FMGI_GRID_BUILD", how)
            s = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels, tier=fmgi.TIER_GRID)
            tables[how] = s.grid_table()
            s.close()
        assert tables["host"].shape == tables["device"].shape, name
        assert len(tables["host"]) > 100 and np.array_equal(tables["host"], tables["device"]), name
    monkeypatch.delenv("FMGI_GRID_BUILD")
    # the default picks the device builder for the big layout and reports its time
    s = fmgi.DeviceScene(synth4000.walls, synth4000.windows, synth4000.lights, synth4000.num_texels)
    _, st = gpu_bake(s, 10, max_depth=2)
    assert 0 < st["grid_build_ms"] < 50
    s.close()


def test_bounds_checked_build_reports_violations(fmgi, scene, monkeypatch):
    """lib/libfmgi_cuda_checked.so (-DFMGI_CHECKED, the stand-in for compute-sanitizer memcheck, which is closed on the
    GPU pool): run the suite with FMGI_LIB pointing at it and every data-dependent index is compared with its table
    size.  This test checks the checker: with FMGI_CHECK_SELFTEST the atlas is declared half its size, so deposits
    into the upper half must be counted and refused; the regular build reports -1 (no checks compiled in)."""
    s = fmgi.DeviceScene(scene.walls, scene.windows, scene.lights, scene.num_texels)
    _, st = gpu_bake(s, 2000)
    if st["bounds_violations"] < 0:
        s.close()
        pytest.skip("regular build: bounds checks are compiled out (run with FMGI_LIB=.../libfmgi_cuda_checked.so)")
    assert st["bounds_violations"] == 0
    monkeypatch.setenv("FMGI_CHECK_SELFTEST", "1")
    with pytest.raises(fmgi.FmgiError, match="index violations"):
        gpu_bake(s, 2000)
    monkeypatch.delenv("FMGI_CHECK_SELFTEST")
    s.close()


def test_chunk_size_does_not_change_the_sample_set(dev_scene, monkeypatch):
    """Small bakes are handed out in smaller photon chunks (32..256 per warp claim) so that every SM gets work;
    the chunk size only changes who traces which photon."""
    spa, depth = 3000, 6                                    # 46k photons: chunks of 32
    ref, sr = gpu_bake(dev_scene, spa, max_depth=depth, seed=8)
    for chunk in ("1", "7", "256", "4096"):
        monkeypatch.setenv("FMGI_CHUNK", chunk)
        got, sg = gpu_bake(dev_scene, spa, max_depth=depth, seed=8)
        for key in ("photons", "rays", "deposits", "mirror_bounces"):
            assert sr[key] == sg[key], (chunk, key)
        assert np.allclose(got, ref, rtol=1e-5, atol=1e-2)
    monkeypatch.delenv("FMGI_CHUNK")


def wall_mean_luminance(atlas_dev, walls, spa):
    """Per-wall mean of the NORMALISED luminance (main.c:68-79 scaling, rectangle.c:277 weights) of a RAW device
    atlas, computed on the device: mean over the wall's base texels of lum * 0.35 * tiles / (area * spa)."""
    import torch

    lum = (atlas_dev[:, :3].double() @ torch.tensor(LUMA, dtype=torch.float64, device=atlas_dev.device))
    csum = torch.cat([torch.zeros(1, dtype=torch.float64, device=lum.device), torch.cumsum(lum, 0)])
    base = torch.tensor(walls["lightmapSetup"][:, 0].astype(np.int64), device=lum.device)
    tiles = torch.tensor((walls["lightmapSetup"][:, 1].astype(np.int64) * walls["lightmapSetup"][:, 2]), device=lum.device)
    sums = (csum[base + tiles] - csum[base]).cpu().numpy()
    tiles = tiles.cpu().numpy().astype(np.float64)
    area = (np.linalg.norm(walls["width"][:, :3].astype(np.float64), axis=1) *
            np.linalg.norm(walls["height"][:, :3].astype(np.float64), axis=1))
    return sums / tiles * 0.35 * tiles / (area * spa), tiles


@pytest.mark.parametrize("tile_size", [200, 800])
def test_radiance_parity_synth4000(fmgi, synth4000, record, tile_size):
    """BASELINE.json configs[2] and [3] against the compiled reference: the 21.5k-rectangle layout, 1e9 photons x 4
    bounces, per-wall mean of the normalised luminance and total energy vs 5.45e8 photons of
    performPhotonMappingNative at depth 4 (fixture: oracle/make_golden.py atlas --fixture synth4000 --depth 4, two
    independent halves).  tile_size 800 is the hi-res layout (fmgi.layout.retile, 1.83 GB atlas): the normalised
    radiance of a wall must not depend on its texel density, so the same fixture serves."""
    import torch
    from fmgi import layout

    sc = synth4000
    z = np.load(GOLDEN / "synth4000_native_depth4.npz")
    depth = int(z["depth"])
    assert depth == 4
    walls, num_texels = (sc.walls, sc.num_texels) if tile_size == 200 else layout.retile(sc.walls, float(tile_size))
    s = fmgi.DeviceScene(walls, sc.windows, sc.lights, num_texels)
    area = sum(sc.photon_counts(1_000_000)) / 1e6
    spa = int(1.0e9 / area)
    atlas = device_atlas(num_texels)
    s.trace(atlas.data_ptr(), spa, stream=torch.cuda.current_stream().cuda_stream, max_depth=depth, seed=4000)
    st = s.sync()
    assert st["tier"] == fmgi.TIER_ROOMS and abs(st["photons"] - 1e9) < 1e6
    wall_gpu, _ = wall_mean_luminance(atlas, s.walls, spa)
    e_gpu = atlas[:, :3].sum(dim=0, dtype=torch.float64).cpu().numpy() / spa
    if tile_size == 200:                        # the strided per-texel sample of the fixture
        mask = torch.tensor(sc.base_texel_mask(), device="cuda")
        lum_all = (atlas[:, :3].double() @ torch.tensor(LUMA, dtype=torch.float64, device="cuda"))
        norm = torch.tensor(sc.normalisation(spa), device="cuda")
        lum_gpu = (lum_all * norm)[mask][:: int(z["stride"])].cpu().numpy()
    del atlas
    s.close()

    wall_a, wall_b = z["wall_lum_a"], z["wall_lum_b"]
    wall_ref = 0.5 * (wall_a + wall_b)
    ref_tiles = (sc.walls["lightmapSetup"][:, 1].astype(np.int64) * sc.walls["lightmapSetup"][:, 2])
    bright = wall_ref > 0.05 * wall_ref.mean()
    big = bright & (ref_tiles >= 1024)          # walls of >= 5 m^2: 8121 of 21552, where the noise is small
    rel = (wall_gpu - wall_ref) / wall_ref
    rel_ab = (wall_a - wall_b) / wall_ref
    rms = lambda v: float(np.sqrt(np.mean(v ** 2)))
    wrms = lambda v, w: float(np.sqrt(np.sum(w * v ** 2) / np.sum(w)))
    e_ref = 0.5 * (z["rgb_total_a"] / float(z["spa_a"]) + z["rgb_total_b"] / float(z["spa_b"]))
    figures = dict(gpu_photons=int(st["photons"]), walls=int(bright.sum()), big_walls=int(big.sum()),
                   big_wall_rel_rms=rms(rel[big]), big_wall_rel_rms_reference_halves=rms(rel_ab[big]),
                   area_weighted_rel_rms=wrms(rel[bright], ref_tiles[bright]),
                   area_weighted_rel_rms_reference_halves=wrms(rel_ab[bright], ref_tiles[bright]),
                   all_wall_rel_rms=rms(rel[bright]), all_wall_rel_rms_reference_halves=rms(rel_ab[bright]),
                   mean_bias=float(np.average(rel[bright], weights=ref_tiles[bright])),
                   energy_rel_diff=[float(x) for x in e_gpu / e_ref - 1])
    if tile_size == 200:
        lum_a, lum_b = z["lum_a"].astype(np.float64), z["lum_b"].astype(np.float64)
        s_ab, sp_ab, _ = parity_stats(lum_a, lum_b)
        s_g, sp_g, _ = parity_stats(lum_gpu, 0.5 * (lum_a + lum_b))
        figures.update(S_per_texel=float(sp_g), S_per_texel_reference_halves=float(sp_ab))
        # ~80 deposits per texel on the reference side: per texel this is a noise check only
        assert sp_g < 1.05 * sp_ab
    print(f"synth4000 tile_size {tile_size}: {figures}")
    record(f"radiance_synth4000_depth4_tile{tile_size}", **figures)
    # per-wall bars: rel. RMS < 1 % where the reference's own noise allows it, energy within 0.1 %
    assert figures["big_wall_rel_rms"] < 0.01 and figures["area_weighted_rel_rms"] < 0.01
    assert figures["big_wall_rel_rms"] < 1.2 * figures["big_wall_rel_rms_reference_halves"]
    assert figures["all_wall_rel_rms"] < 1.2 * figures["all_wall_rel_rms_reference_halves"]
    assert abs(figures["mean_bias"]) < 1e-3
    assert np.all(np.abs(e_gpu / e_ref - 1) < 1e-3)


def test_accumulation_passes_keep_the_sample_set(dev_scene, monkeypatch):
    """Big bakes are traced in several fp32 accumulation passes (finer shards into a scratch atlas
    that is folded into the caller's atlas): same photons, same counters, same sums."""
    spa, depth = 60000, 6
    monkeypatch.setenv("FMGI_ACCUM_PASSES", "1")
    one, s1 = gpu_bake(dev_scene, spa, max_depth=depth, seed=17)
    monkeypatch.setenv("FMGI_ACCUM_PASSES", "5")
    five, s5 = gpu_bake(dev_scene, spa, max_depth=depth, seed=17)
    for k in ("photons", "rays", "deposits", "mirror_bounces"):
        assert s1[k] == s5[k]
    assert s5["kernel_launches"] - s1["kernel_launches"] >= 9
    assert np.allclose(one, five, rtol=1e-5, atol=1e-2)
    assert np.all(five[:, 3] == 0)


def staircase_scene(fmgi):
    """A closed 6 x 4 x 3 m room with a 12-step staircase of horizontal slabs, each at its own height, top and
    bottom faces: 12 + 1 distinct z planes per normal sign, more than the grid tier's 8 + 8 plane table, so some
    horizontal rectangles have to ride in the walk lists."""
    def rect(px, py, pz, w, h, n, base, tw, th):
        r = np.zeros(1, dtype=fmgi.RECT_DTYPE)[0]
        r["pos"][:3] = (px, py, pz); r["width"][:3] = w; r["height"][:3] = h; r["n"][:3] = n
        r["lightmapSetup"][:3] = (base, tw, th)
        return r

    walls, base = [], 0

    def add(px, py, pz, w, h):
        nonlocal base
        n = np.cross(np.array(h, dtype=np.float32), np.array(w, dtype=np.float32))     # rectangle.c:22
        n = (n / np.linalg.norm(n)).astype(np.float32)
        walls.append(rect(px, py, pz, w, h, n, base, 8, 8))
        base += 8 * 8 + 16 + 4 + 1                     # full mip chain of an 8 x 8 tile (rectangle.c:166-192)

    X, Y, Z = 6.0137, 4.0211, 3.0071
    add(X, 0, 0, (-X, 0, 0), (0, Y, 0))                # floor, normal +z (parseLayout.c:466 orientation)
    add(0, 0, Z, (X, 0, 0), (0, Y, 0))                 # ceiling, normal -z
    add(0, 0, 0, (X, 0, 0), (0, 0, Z))                 # wall y = 0, normal +y
    add(X, Y, 0, (-X, 0, 0), (0, 0, Z))                # wall y = Y, normal -y
    add(0, Y, 0, (0, -Y, 0), (0, 0, Z))                # wall x = 0, normal +x
    add(X, 0, 0, (0, Y, 0), (0, 0, Z))                 # wall x = X, normal -x
    for i in range(12):
        # off-lattice numbers on purpose: with round ones, rays from tile centres along the symmetric
        # direction set pass exactly through rectangle edges and hit/miss becomes a rounding matter
        x0, z = 0.4037 * i + 0.3119, 0.2011 * (i + 1) + 0.0113
        add(x0 + 0.3977, 1.0171, z, (-0.3977, 0, 0), (0, 1.9871, 0))          # tread, top face (normal +z)
        add(x0, 1.0171, z - 0.0503, (0.3977, 0, 0), (0, 1.9871, 0))          # underside (normal -z)
    light = rect(2.5, 1.5, Z - 0.001, (1.0, 0, 0), (0, 1.0, 0), (0, 0, -1), 0, 1, 1)
    walls = np.array(walls, dtype=fmgi.RECT_DTYPE)
    return walls, np.zeros(0, dtype=fmgi.RECT_DTYPE), np.array([light], dtype=fmgi.RECT_DTYPE), base


@pytest.mark.parametrize("tier", ["soup", "grid", "rooms"])
def test_more_z_planes_than_the_plane_table(fmgi, oracle, tier):
    import refbind

    walls, windows, lights, num_texels = staircase_scene(fmgi)
    sc = refbind.Scene(walls, windows, lights, num_texels)
    s = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels, tier=TIERS[tier])
    o, d = random_rays(sc, 200_000, 9)
    gi, gt = s.closest_hit(o, d)
    ci, ct = oracle.closest_hit(sc.walls, o, d, oracle.ACCEL_LINEAR)
    assert (gi != ci).mean() < 2e-5
    both = (gi >= 0) & (gi == ci)
    assert both.mean() > 0.9                            # closed room: almost every ray hits something
    assert np.max(np.abs(gt[both] - ct[both]) / np.maximum(ct[both], 1e-4)) < 1e-4
    spa, depth = 200000, 8
    atlas, st = gpu_bake(s, spa, max_depth=depth, seed=2)
    want, so = oracle.bake(sc, spa, depth, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, 2)
    assert st["photons"] == so["photons"] > 100000
    assert abs(st["deposits"] - so["deposits"]) <= 3e-4 * so["deposits"]
    assert abs(atlas[:, :3].sum(dtype=np.float64) / want[:, :3].sum(dtype=np.float64) - 1) < 3e-4
    got = s.paths(0, depth, 4, 0, 20000)
    ref = oracle.trace_paths(sc, 0, depth, 4, 0, 20000)
    assert np.all(got == ref, axis=1).mean() > 0.995
    s.close()


@pytest.mark.parametrize("tint", [0, 1])
def test_device_tonemap_is_byte_exact(fmgi, dev_scene, oracle, scene, tint):
    """fmgi_scene_tonemap = main.c:68-79 + saveAs_core (rectangle.c:293-336) on the device: the packed RGB
    tiles equal the oracle's (which equals the reference's saveAs output) byte for byte, including black
    texels (0/0 -> 0) and the floor tint."""
    import torch

    spa = 30000
    atlas = device_atlas(dev_scene.num_texels)
    dev_scene.trace(atlas.data_ptr(), spa, stream=torch.cuda.current_stream().cuda_stream, max_depth=8, seed=3)
    dev_scene.sync()
    rgb = torch.zeros(dev_scene.tile_bytes(), dtype=torch.uint8, device="cuda")
    dev_scene.tonemap(atlas.data_ptr(), spa, rgb.data_ptr(), tint_extra=tint,
                      stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    got = rgb.cpu().numpy()
    want = oracle.tonemap_tiles(scene, atlas.cpu().numpy(), spa, tint)
    assert got.size == want.size == 3 * int(scene.base_texel_mask().sum())
    diff = got != want
    assert diff.mean() < 1e-6, f"{diff.sum()} differing bytes"
    assert (got > 0).mean() > 0.9


def test_device_png_tiles_decode_to_the_packed_tiles(fmgi, dev_scene, scene, synth800):
    """fmgi_scene_tiles_png: every wall's tile as a complete PNG file assembled on the GPU (stored-block zlib stream,
    Adler-32 and CRC-32 computed on the device).  PIL - which checks both checksums - must decode each file to exactly
    the RGB bytes fmgi_scene_tonemap packs, i.e. to what the reference's saveAs hands to write_png_file
    (rectangle.c:338-346); a tile taller than one stored block (65535 raw bytes) is part of the case."""
    import io

    import torch
    from PIL import Image

    from fmgi import layout

    big_walls, big_texels = layout.retile(synth800.walls, 3000.0)       # some tiles above 65535 raw bytes
    big = fmgi.DeviceScene(big_walls, synth800.windows, synth800.lights, big_texels)
    for s, spa, tint in ((dev_scene, 30000, 0), (dev_scene, 30000, 1), (big, 3000, 0)):
        atlas = device_atlas(s.num_texels)
        stream = torch.cuda.current_stream().cuda_stream
        s.trace(atlas.data_ptr(), spa, stream=stream, max_depth=8, seed=3)
        s.sync()
        rgb = torch.zeros(s.tile_bytes(), dtype=torch.uint8, device="cuda")
        s.tonemap(atlas.data_ptr(), spa, rgb.data_ptr(), tint_extra=tint, stream=stream)
        total, off = fmgi.tile_png_layout(s.walls)
        png = torch.zeros(total, dtype=torch.uint8, device="cuda")
        s.tiles_png(atlas.data_ptr(), spa, png.data_ptr(), tint_extra=tint, stream=stream)
        torch.cuda.synchronize()
        want, files = rgb.cpu().numpy(), png.cpu().numpy()
        at, raw_max = 0, 0
        for i, w in enumerate(s.walls):
            tw, th = int(w["lightmapSetup"][1]), int(w["lightmapSetup"][2])
            raw_max = max(raw_max, th * (3 * tw + 1))
            img = Image.open(io.BytesIO(files[int(off[i]): int(off[i + 1])].tobytes()))
            img.load()
            assert img.size == (tw, th) and img.mode == "RGB", i
            assert np.array_equal(np.asarray(img).reshape(-1), want[at: at + 3 * tw * th]), i
            at += 3 * tw * th
        assert at == want.size
        if s is big:
            assert raw_max > 65535
    big.close()


def test_bake_tiles_png_through_the_host_call(fmgi, scene):
    import io

    from PIL import Image

    spa = 20000
    tex = fmgi.aligned_texels(scene.num_texels)
    geo = fmgi.make_geometry(scene.walls, scene.windows, scene.lights, tex)
    files, off, st = fmgi.bake_tiles_png(geo, scene.walls, spa, tint_extra=0, seed=6, max_depth=8)
    rgb, _ = fmgi.bake_tiles(geo, scene.walls, spa, tint_extra=0, seed=6, max_depth=8)
    assert st["deposits"] > 0 and not tex.any()                     # the float atlas is not written back
    at = 0
    for i, w in enumerate(scene.walls):
        tw, th = int(w["lightmapSetup"][1]), int(w["lightmapSetup"][2])
        img = np.asarray(Image.open(io.BytesIO(files[int(off[i]): int(off[i + 1])].tobytes()))).reshape(-1)
        d = np.abs(img.astype(np.int32) - rgb[at: at + 3 * tw * th].astype(np.int32))
        assert d.max() <= 1                                          # two bakes: float atomics order differs
        at += 3 * tw * th


def test_bake_tiles_returns_packed_tiles_only(fmgi, oracle, scene):
    spa = 20000
    tex = fmgi.aligned_texels(scene.num_texels)
    before = tex.copy()
    geo = fmgi.make_geometry(scene.walls, scene.windows, scene.lights, tex)
    rgb, st = fmgi.bake_tiles(geo, scene.walls, spa, tint_extra=0, seed=6, max_depth=8)
    assert np.array_equal(tex, before)                       # the float atlas is not written back
    tex2 = fmgi.aligned_texels(scene.num_texels)
    fmgi.bake(fmgi.make_geometry(scene.walls, scene.windows, scene.lights, tex2), spa, seed=6, max_depth=8)
    want = oracle.tonemap_tiles(scene, tex2, spa, 0)
    d = np.abs(rgb.astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3            # two bakes: float atomics order differs
    assert st["deposits"] > 0


def test_ambient_occlusion_matches_reference(fmgi, scene):
    """fmgi_ambient_occlusion vs performAmbientOcclusionNative (photonmap.c:480) on example.png (fixture made
    by the compiled reference).  Float result: the direction set is the same but summed in another order
    and the distances come from another exact closest-hit search, so texels agree to ~1e-6 relative; a ray
    grazing a rectangle edge may flip between hit and miss, which moves one texel by up to 1/481 of its range."""
    want = np.load(GOLDEN / "example_ao_native.npz")["ao"].astype(np.float64)
    tex = fmgi.aligned_texels(scene.num_texels)
    tex[...] = 7.0                                            # must be overwritten on the base level only
    fmgi.ambient_occlusion(fmgi.make_geometry(scene.walls, scene.windows, scene.lights, tex))
    mask = scene.base_texel_mask()
    assert np.all(tex[~mask] == 7.0)
    got = tex[mask]
    assert np.array_equal(got[:, 0], got[:, 1]) and np.array_equal(got[:, 0], got[:, 2]) and not got[:, 3].any()
    rel = np.abs(got[:, 0] - want) / np.maximum(want, 1e-3)
    # measured: 99.91 % of the 85 056 texels within 1e-5, the rest are single hit/miss flips (a miss counts
    # as distance 10, so one flipped ray can move a texel by a few per cent), mean ratio 1 + 3e-6
    assert np.mean(rel < 1e-5) > 0.998, np.mean(rel < 1e-5)
    assert rel.max() < 0.08
    assert abs(got[:, 0].mean() / want.mean() - 1) < 2e-5


@pytest.mark.parametrize("tier", ["soup", "grid", "rooms"])
def test_ambient_occlusion_matches_oracle_on_small_room(fmgi, oracle, tier):
    import refbind

    walls, windows, lights, num_texels = staircase_scene(fmgi)
    sc = refbind.Scene(walls, windows, lights, num_texels)
    want = oracle.ambient_occlusion(sc, fmgi.geosphere(4), oracle.ACCEL_LINEAR)
    tex = fmgi.aligned_texels(num_texels)
    fmgi.ambient_occlusion(fmgi.make_geometry(sc.walls, sc.windows, sc.lights, tex), tier=TIERS[tier])
    mask = sc.base_texel_mask()
    rel = np.abs(tex[mask, 0] - want[mask, 0]) / np.maximum(want[mask, 0], 1e-3)
    assert np.mean(rel < 1e-5) > 0.99 and rel.max() < 0.08
    assert abs(tex[mask, 0].mean(dtype=np.float64) / want[mask, 0].mean(dtype=np.float64) - 1) < 2e-3


def random_scene(fmgi, seed, n_axis=120, n_general=25):
    """Random rectangle soup in a 12 x 9 x 3 m box: axis-parallel rectangles of all six orientations at
    many different plane coordinates (more z planes than the plane table holds) plus arbitrarily
    oriented ones, with normals as createRectangleV builds them (rectangle.c:22)."""
    rng = np.random.default_rng(seed)
    walls = np.zeros(n_axis + n_general, dtype=fmgi.RECT_DTYPE)
    base = 0
    for i in range(len(walls)):
        pos = rng.uniform([0, 0, 0], [12, 9, 3]).astype(np.float32)
        if i < n_axis:
            k = rng.integers(0, 3)
            ai, aj = [a for a in range(3) if a != k]
            if rng.random() < 0.5:
                ai, aj = aj, ai
            w = np.zeros(3, np.float32); h = np.zeros(3, np.float32)
            w[ai] = rng.uniform(0.3, 4.0) * rng.choice([-1, 1]); h[aj] = rng.uniform(0.3, 3.0) * rng.choice([-1, 1])
        else:
            w = rng.normal(size=3).astype(np.float32); w *= rng.uniform(0.5, 3.0) / np.linalg.norm(w)
            h = np.cross(w, rng.normal(size=3)).astype(np.float32); h *= rng.uniform(0.5, 3.0) / np.linalg.norm(h)
        n = np.cross(h, w).astype(np.float32)
        n = (n / np.float32(np.linalg.norm(n))).astype(np.float32)
        walls[i]["pos"][:3] = pos; walls[i]["width"][:3] = w; walls[i]["height"][:3] = h; walls[i]["n"][:3] = n
        walls[i]["lightmapSetup"][:3] = (base, 4, 2)
        base += 8 + 2 + 1
    light = np.zeros(1, dtype=fmgi.RECT_DTYPE)
    light[0]["pos"][:3] = (5, 4, 2.9); light[0]["width"][:3] = (1, 0, 0); light[0]["height"][:3] = (0, 1, 0)
    light[0]["n"][:3] = (0, 0, -1); light[0]["lightmapSetup"][:3] = (0, 1, 1)
    return walls, np.zeros(0, dtype=fmgi.RECT_DTYPE), light, base


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("tier", ["soup", "grid"])
def test_random_soups_closest_hit_and_paths(fmgi, oracle, seed, tier):
    """Property test of the closest-hit search on scenes parseLayout would never produce."""
    import refbind

    walls, windows, lights, num_texels = random_scene(fmgi, seed)
    sc = refbind.Scene(walls, windows, lights, num_texels)
    s = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels,
                         tier=fmgi.TIER_SOUP if tier == "soup" else fmgi.TIER_GRID)
    o, d = random_rays(sc, 100_000, seed)
    gi, gt = s.closest_hit(o, d)
    ci, ct = oracle.closest_hit(sc.walls, o, d, oracle.ACCEL_LINEAR)
    assert (gi != ci).mean() < 2e-4, (gi != ci).sum()
    both = (gi >= 0) & (gi == ci)
    assert both.mean() > 0.15
    assert np.max(np.abs(gt[both] - ct[both]) / np.maximum(ct[both], 1e-3)) < 1e-3
    got = s.paths(0, 6, seed, 0, 20000)
    ref = oracle.trace_paths(sc, 0, 6, seed, 0, 20000)
    assert np.all(got == ref, axis=1).mean() > 0.99
    s.close()


@pytest.mark.parametrize("seed", [1, 2, 3])
@pytest.mark.parametrize("tier", ["grid", "rooms"])
def test_random_axis_parallel_soups(fmgi, oracle, seed, tier):
    """160 axis-parallel rectangles of all six orientations floating at random in a box - overlapping, crossing,
    seen from behind (back-face culling lets rays through): nothing like a flat, the hard case for the room tier's
    box decomposition, whose leaves must still put every rectangle on a face."""
    import refbind

    walls, windows, lights, num_texels = random_scene(fmgi, seed, n_axis=160, n_general=0)
    sc = refbind.Scene(walls, windows, lights, num_texels)
    s = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels, tier=TIERS[tier])
    o, d = random_rays(sc, 200_000, seed)
    gi, gt = s.closest_hit(o, d)
    ci, ct = oracle.closest_hit(sc.walls, o, d, oracle.ACCEL_LINEAR)
    at_edge, elsewhere = split_mismatches(sc.walls, o, d, gi, gt, ci, ct)
    assert elsewhere == 0 and at_edge <= 8, (at_edge, elsewhere)
    both = (gi >= 0) & (gi == ci)
    assert both.mean() > 0.15
    assert np.max(np.abs(gt[both] - ct[both]) / np.maximum(ct[both], 1e-3)) < 1e-4
    got = s.paths(0, 6, seed, 0, 20000)
    ref = oracle.trace_paths(sc, 0, 6, seed, 0, 20000)
    assert np.all(got == ref, axis=1).mean() > 0.99
    spa = 100000
    atlas, st = gpu_bake(s, spa, max_depth=6, seed=seed)
    want, so = oracle.bake(sc, spa, 6, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, seed)
    assert st["photons"] == so["photons"] and abs(st["deposits"] - so["deposits"]) <= 1e-3 * so["deposits"] + 2
    s.close()


def test_hires_retiled_layout_texel_index(fmgi, oracle, synth800):
    """4x texel density (BASELINE configs[3]): tile grids from fmgi.layout.retile, texel index still the
    reference's getTileIdAt bit for bit."""
    import refbind
    from fmgi import layout

    walls, num_texels = layout.retile(synth800.walls, 800.0)
    assert num_texels > 3.5 * synth800.num_texels
    sc = refbind.Scene(walls, synth800.windows, synth800.lights, num_texels)
    s = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels)
    rng = np.random.default_rng(4)
    for wi in rng.choice(len(sc.walls), 40, replace=False):
        w = sc.walls[wi]
        uv = rng.random((3000, 2), dtype=np.float32)
        pts = (w["pos"][:3] + uv[:, :1] * w["width"][:3] + uv[:, 1:] * w["height"][:3]).astype(np.float32)
        assert np.array_equal(s.tile_ids(np.full(len(pts), wi, dtype=np.int32), pts), oracle.tile_ids(w, pts))
    atlas, st = gpu_bake(s, 2000, max_depth=4, seed=1)
    assert st["deposits"] > 0 and np.all(atlas[~sc.base_texel_mask()] == 0)
    s.close()


def test_radiance_parity_grid_tier_synth800(fmgi, synth800):
    """The same radiance bars on a layout that goes through the GRID tier and is lit mostly by ceiling
    lights (cosine emitters): 547 rectangles, 31 emitters, against 6.0e8 photons of the compiled reference
    (fixture: every 8th base texel + per-wall mean luminance, two independent halves)."""
    sc = synth800
    z = np.load(GOLDEN / "synth800_native_depth8.npz")
    stride, depth = int(z["stride"]), int(z["depth"])
    s = fmgi.DeviceScene(sc.walls, sc.windows, sc.lights, sc.num_texels, tier=fmgi.TIER_GRID)
    area = sum(sc.photon_counts(1_000_000)) / 1e6
    spa = int(2.0e9 / area)
    atlas, st = gpu_bake(s, spa, max_depth=depth, seed=99)
    s.close()
    mask = sc.base_texel_mask()
    lum_all = (atlas[:, :3].astype(np.float64) @ LUMA) * sc.normalisation(spa)
    lum_gpu = lum_all[mask][::stride]
    lum_a, lum_b = z["lum_a"].astype(np.float64), z["lum_b"].astype(np.float64)
    s_ab, sp_ab, _ = parity_stats(lum_a, lum_b)
    s_g, sp_g, lit = parity_stats(lum_gpu, 0.5 * (lum_a + lum_b))
    print(f"synth800 depth {depth}: S={s_g:.4%} S'={sp_g:.4%} (lit {lit:.3f}); reference half-vs-half S={s_ab:.4%} "
          f"S'={sp_ab:.4%}; gpu photons {st['photons']:.3e}")
    assert sp_g < 0.02 and s_g < 0.02 and sp_g < 1.05 * sp_ab
    # per-wall mean luminance: noise is negligible at this level, so this is a bias check per surface
    wall_ref = 0.5 * (z["wall_lum_a"] + z["wall_lum_b"])
    wall_gpu = np.array([lum_all[int(w["lightmapSetup"][0]): int(w["lightmapSetup"][0]) +
                                 int(w["lightmapSetup"][1]) * int(w["lightmapSetup"][2])].mean() for w in sc.walls])
    bright = wall_ref > 0.05 * wall_ref.mean()
    wall_rel = (wall_gpu[bright] - wall_ref[bright]) / wall_ref[bright]
    ab_rel = (z["wall_lum_a"][bright] - z["wall_lum_b"][bright]) / wall_ref[bright]
    print(f"per-wall mean luminance: rel. RMS {np.sqrt(np.mean(wall_rel ** 2)):.4%} (reference half-vs-half "
          f"{np.sqrt(np.mean(ab_rel ** 2)):.4%}), mean bias {wall_rel.mean():+.4%}")
    assert np.sqrt(np.mean(wall_rel ** 2)) < 0.01 and abs(wall_rel.mean()) < 2e-3
    e_ref = 0.5 * (z["rgb_total_a"] / float(z["spa_a"]) + z["rgb_total_b"] / float(z["spa_b"]))
    e_gpu = atlas[:, :3].sum(axis=0, dtype=np.float64) / spa
    print(f"energy per unit density rel. diff {e_gpu / e_ref - 1}")
    assert np.all(np.abs(e_gpu / e_ref - 1) < 1e-3)


def test_full_size_properties(dev_scene, scene):
    """BASELINE.json configs[1] at full size (example.png, 1e8 photons x 3 bounces): size-independent
    properties instead of an oracle run - exact budgets, counter identities, determinism of the sample set,
    additivity over photon shards (what the multi-GPU reduce relies on) and energy per photon equal to a
    10x smaller bake within Monte-Carlo noise."""
    area = sum(scene.photon_counts(1_000_000)) / 1e6
    spa, depth = int(1.0e8 / area), 3
    whole, sw = gpu_bake(dev_scene, spa, max_depth=depth, seed=12)
    assert sw["photons"] == sum(scene.photon_counts(spa)) and abs(sw["photons"] - 1e8) < 100
    # every ray either deposits or ends its photon; a photon makes at most `depth` deposits
    assert sw["deposits"] <= sw["rays"] <= sw["deposits"] + sw["photons"]
    assert sw["deposits"] <= depth * sw["photons"] and sw["mirror_bounces"] < sw["deposits"]
    assert 2.2 < sw["deposits"] / sw["photons"] < 2.3           # 2.254 at depth 3
    again, sa = gpu_bake(dev_scene, spa, max_depth=depth, seed=12)
    for k in ("photons", "rays", "deposits", "mirror_bounces"):
        assert sw[k] == sa[k]
    # same deposits in another order: float atomics round differently (about 3e4 additions per texel)
    assert np.allclose(whole, again, rtol=1e-4, atol=1.0)
    parts = [gpu_bake(dev_scene, spa, max_depth=depth, seed=12, shard=g, num_shards=8) for g in range(8)]
    for k in ("photons", "rays", "deposits", "mirror_bounces"):
        assert sum(p[1][k] for p in parts) == sw[k]
    total = np.sum([p[0].astype(np.float64) for p in parts], axis=0)
    # The whole bake adds ~3e4 deposits of ~16 into each fp32 texel; once a texel passes 2^16 every add rounds
    # by up to 2^-8 in the same direction, so the single sum drifts from the sum of eight 8x smaller sums by a
    # few 1e-5 relative (measured 4.5e-5; the library re-zeroes its fp32 accumulator every 2^28 photons).
    assert np.allclose(total, whole, rtol=1e-3, atol=1.0)
    assert abs(total[:, :3].sum() / whole[:, :3].sum(dtype=np.float64) - 1) < 2e-4
    assert np.all(whole[:, 3] == 0) and np.all(whole[~scene.base_texel_mask()] == 0)
    # colour lanes: every deposit is (18 or 16, ., 18) x 0.9^k x tint -> R >= G >= B-ish ordering never inverts R < 0
    assert whole.min() >= 0
    small, ss = gpu_bake(dev_scene, spa // 10, max_depth=depth, seed=13)
    e_big = whole[:, :3].sum(dtype=np.float64) / sw["photons"]
    e_small = small[:, :3].sum(dtype=np.float64) / ss["photons"]
    assert abs(e_big / e_small - 1) < 1e-3
