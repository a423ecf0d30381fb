"""One-process-per-GPU photon sharding.

Photons are independent, so the path shards with no data-path collective: rank r traces the
contiguous photon-index range [N_e * r / W, N_e * (r+1) / W) of every emitter e (the Philox
counter is the global photon index, so the sample set does not depend on W).  The one real
exchange step is additive: the per-rank lightmap atlases are summed onto rank 0 with a single
reduce (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations


def shard_of(rank: int, world: int) -> dict:
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    return {"shard": rank, "num_shards": world}


def photon_range(n: int, rank: int, world: int):
    """The photon indices of one emitter this rank owns (mirrors fill_jobs() in fmgi_api.cu)."""
    return n * rank // world, n * (rank + 1) // world


def bake_sharded(trace, atlas, spa_job: int, rank: int, world: int, dist=None, group=None, dst: int = 0):
    """trace(atlas, spa_job, shard=..., num_shards=...) accumulates this rank's sub-range into its
    own atlas tensor; afterwards rank `dst` holds the sum over ranks."""
    trace(atlas, spa_job, **shard_of(rank, world))
    if world > 1:
        dist.reduce(atlas, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return atlas
