"""Pair-sharded batches: independent DEM pairs across the GPUs of one node, no data-path collective.

The reference's only parallelism is GNU parallel fanning whole ``dem_align.py`` processes over files
(reference ``demcoreg/dem_align_parallel.pbs:122``, ``dem_coreg_all.sh:116,139``).  Here: one process per
GPU (``torch.distributed``), pair ``p`` goes to rank ``p % world``; the only communication is the
gather of three floats per pair at the end.
"""
import os


def shard(n_items, rank, world):
    """Indices of the items rank ``rank`` of ``world`` processes (round-robin, like ``parallel -j``)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return list(range(rank, n_items, world))


def align_batch(pairs, align_fn, group=None):
    """Run ``align_fn(pair) -> dict(dx, dy, dz, n_iter)`` on this rank's shard of ``pairs`` and gather.

    Returns the full, ordered result list on every rank (``all_gather_object`` of small dicts).
    Works without an initialised process group (single process).
    """
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        rank, world = dist.get_rank(group), dist.get_world_size(group)
    else:
        rank, world = 0, 1
    mine = shard(len(pairs), rank, world)
    local = []
    for idx in mine:
        r = align_fn(pairs[idx])
        local.append((idx, {k: r[k] for k in ("dx", "dy", "dz", "n_iter") if k in r}))
    if world == 1:
        gathered = [local]
    else:
        gathered = [None] * world
        dist.all_gather_object(gathered, local, group=group)
    out = [None] * len(pairs)
    for part in gathered:
        for idx, r in part:
            out[idx] = r
    return out


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
