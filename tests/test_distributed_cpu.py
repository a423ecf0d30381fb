"""world_size-2 gloo test of the N > 1 path on CPU: photon-range shards + one additive reduce.
The tracer here is the oracle in Philox mode (tests may use the oracle); on GPUs the same
bake_sharded() drives fmgi_scene_trace and NCCL (bench.py)."""
import os
import socket
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _rank_main(rank, world, port, spa, depth, seed, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, str(ROOT / "flatmatch-global-illumination_b200"))
    sys.path.insert(0, str(ROOT / "oracle"))
    import torch
    import torch.distributed as dist

    import refbind
    from fmgi.distributed import bake_sharded

    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = refbind.Scene.load(GOLDEN / "example_scene.npz")
    orc = refbind.OracleLib()
    counters = {}

    def trace(atlas, spa_job, shard, num_shards):
        tex = refbind.aligned_texels(scene.num_texels)
        _, st = orc.bake(scene, spa_job, depth, orc.ACCEL_LINEAR, orc.RNG_PHILOX, seed, shard, num_shards, texels=tex)
        counters.update(st)
        atlas += torch.from_numpy(tex)

    atlas = torch.zeros((scene.num_texels, 4), dtype=torch.float32)
    bake_sharded(trace, atlas, spa, rank, world, dist=dist)
    c = torch.tensor([counters["photons"], counters["deposits"]], dtype=torch.int64)
    dist.all_reduce(c)
    if rank == 0:
        np.savez(out_path, atlas=atlas.numpy(), photons=int(c[0]), deposits=int(c[1]))
    dist.destroy_process_group()


def test_two_rank_sharded_bake_equals_single_rank(tmp_path, oracle, scene):
    import torch.multiprocessing as mp

    spa, depth, seed, world = 3000, 4, 13, 2
    out = tmp_path / "rank0.npz"
    mp.spawn(_rank_main, args=(world, _free_port(), spa, depth, seed, str(out)), nprocs=world, join=True)
    got = np.load(out)
    want, st = oracle.bake(scene, spa, depth, oracle.ACCEL_LINEAR, oracle.RNG_PHILOX, seed)
    assert int(got["photons"]) == st["photons"] and int(got["deposits"]) == st["deposits"]
    assert np.allclose(got["atlas"], want, rtol=1e-5, atol=1e-3)


def test_photon_ranges_tile_the_budget():
    from fmgi.distributed import photon_range

    for n in (0, 1, 7, 468833, 10**10 + 3):
        for world in (1, 2, 3, 8):
            edges = [photon_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
