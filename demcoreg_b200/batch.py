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


def stream_pairs(host_pairs, device=None):
    """Double-buffered upload of host-resident pairs: yields ``(ref, src, mask_source)`` device objects for pair ``k`` while
    pair ``k+1`` is being copied on a separate CUDA stream, so that a batch streaming through one GPU is bound by
    max(PCIe, kernels) instead of their sum.

    ``host_pairs``: sequence of dicts ``ref, src`` (float32 arrays or PINNED torch tensors), ``ref_gt, src_gt, ndv`` and
    optionally ``mask`` (uint8) + ``mask_gt``.  All pairs must share the raster shapes (two device buffer sets are reused).
    The consumer must finish its GPU work on the yielded rasters (``align`` / ``nuth_iteration`` synchronise on their
    result) before advancing the generator."""
    import numpy as np
    import torch
    from .raster import Raster, MaskSource

    def pinned(a, dtype):
        if isinstance(a, np.ndarray):
            return torch.from_numpy(np.ascontiguousarray(a, dtype=dtype)).pin_memory()
        return a

    host_pairs = list(host_pairs)
    if not host_pairs:
        return
    first = host_pairs[0]
    has_mask = first.get('mask') is not None
    sets = []
    for _ in range(2):
        r = Raster(torch.empty(first['ref'].shape, dtype=torch.float32, device=device or "cuda"), first['ref_gt'], first['ndv'])
        s_ = Raster(torch.empty(first['src'].shape, dtype=torch.float32, device=device or "cuda"), first['src_gt'], first['ndv'])
        m = MaskSource(torch.empty(first['mask'].shape, dtype=torch.uint8, device=device or "cuda"), first['mask_gt']) if has_mask else None
        sets.append((r, s_, m))
    cs = torch.cuda.Stream()
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    free = [torch.cuda.Event(), torch.cuda.Event()]
    for ev in free:
        ev.record()

    def upload(k):
        b = k & 1
        hp = host_pairs[k]
        r, s_, m = sets[b]
        with torch.cuda.stream(cs):
            cs.wait_event(free[b])
            r.data.copy_(pinned(hp['ref'], np.float32), non_blocking=True)
            s_.data.copy_(pinned(hp['src'], np.float32), non_blocking=True)
            if m is not None:
                m.data.copy_(pinned(hp['mask'], np.uint8), non_blocking=True)
            ready[b].record(cs)
        r.gt = tuple(float(g) for g in hp['ref_gt'])
        s_.gt = tuple(float(g) for g in hp['src_gt'])
        r.ndv = s_.ndv = hp['ndv']
        s_.pending_dz = []
        if m is not None:
            m.gt = tuple(float(g) for g in hp['mask_gt'])

    upload(0)
    for k in range(len(host_pairs)):
        b = k & 1
        torch.cuda.current_stream().wait_event(ready[b])
        # NB the next upload overwrites the OTHER buffer set, whose consumer finished before this generator was advanced
        if k + 1 < len(host_pairs):
            upload(k + 1)
        yield sets[b]
        free[b].record()



def align_stream(host_pairs, mode_kw=None, final_pass=True, device=None):
    """Align a stream of host-resident pairs on this rank's GPU: ``dem_align`` (mode nuth) to convergence per pair.

    ``stream_pairs`` uploads pair ``k+1`` (pinned host memory, copy stream) while the whole loop of pair ``k`` runs;
    ONE ``NuthAligner`` serves every pair (``reset``: iteration outputs, workspace and record array are allocated
    once).  The source rasters are shifted in place in the upload buffers (the reference works on a copy; here the buffer
    is overwritten by the next upload anyway).  ``final_pass``: one more front + filter pass on the aligned pair, the
    basis of the reference's after-alignment statistics (``dem_align.py:366-392``).
    Returns a list of dicts ``dx, dy, dz, n_iter`` (+ ``after_med, after_nmad, after_count`` with ``final_pass``).
    """
    from . import engine
    kw = dict(tol=0.02, max_offset=100, max_dz=100, slope_lim=(0.1, 40), max_iter=30)
    if mode_kw:
        kw.update(mode_kw)
    al = None
    bufs = None
    out = []
    for d_ref, d_src, d_ms in stream_pairs(host_pairs, device=device):
        if al is None:
            al = engine.NuthAligner(d_ref, d_src, d_ms, **kw)
        else:
            al.reset(d_ref, d_src, d_ms, **kw)
        recs = al.run()
        dx, dy, dz = al.totals
        r = dict(dx=dx, dy=dy, dz=dz, n_iter=len(recs), exit_code=int(al.st.exit_code),
                 select_fallbacks=sum(1 for it in recs if not it['select_engine']),
                 bin_fallbacks=sum(1 for it in recs if not it['bin_engine']))
        if final_pass and al.st.exit_code == 0:
            res, grid, bufs = engine.nuth_iteration(d_ref, d_src, d_ms, True, kw['max_dz'], (0.1, 40), bufs=bufs)
            r.update(after_med=float(res.med1), after_nmad=float(res.nmad1), after_count=int(res.count_mad))
        out.append(r)
    return out
