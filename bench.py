#!/usr/bin/env python
"""bench.py -- Nuth-Kaab co-registration iteration throughput (BASELINE.json metric).

A "step" is one full dem_align iteration on one DEM pair: compute_offset(mode='nuth')
(reference dem_align.py:62-172: warp both rasters to the common grid, dh, masks, outlier
filters, Horn slope/aspect, medians, aspect-bin statistics, fit) followed by apply_xy_shift /
apply_z_shift (coreglib.py:14-55), with device-resident inputs and (dx, dy, dz) on the host at
the end of every step.  Workload at N=1: BASELINE.json configs[1] (8192x8192 pair at 8 m with a
stable-surface mask).  N>1: one independent pair per GPU (pair-sharded, no collective), weak scaling.

The same JSON line carries three more measured objects (BASELINE.json configs[2..4]):
  "config3"  N=1: compute_offset_ncc / compute_offset_sad, +-16 px window on a 4096^2 pair (time, GFLOP/s, bytes);
  "batch"    every N: 256 DISTINCT host-resident 4096^2 pairs sharded p % N over the ranks, each streamed through
             batch.align_stream (pinned upload overlapped with the previous pair's loop, dem_align to convergence,
             final statistics pass): pairs/s WITH the uploads, next to a measured all-ranks H2D bandwidth;
  "band"     N>1: ONE 32768^2 pair at 2 m split into row bands (NCCL halo exchange + histogram all-reduces).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size 8192]
                    [--no-config3] [--no-batch] [--no-band] [--batch-pairs 256] [--band-size 32768]
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALG_BYTES_ITER = 90.0      # algorithmic bytes per pixel per iteration
# dram__bytes_read.sum + dram__bytes_write.sum of the front end (front_fast_kernel + front_kernel over the listed tiles) on
# the default 8192^2 workload, from the `ncu --set full` capture summarised in profiles/r2/front_r2_raw.csv:
# 591.41 + 43.73 MB read, 653.27 + 7.87 MB written
FRONT_KERNEL_DRAM_BYTES_8192 = 1296272640
# algorithmic bytes per pixel of each profiled stage
STAGE_INFO = [
    # R ref 4 + src 4 + mask 1, W dh 4 + slope 4 + aspect 4 = 21 B/px (the kernels really write dh 4 + tan(slope) 4
    # + mask bits 1 + bin code 2 = 11 B/px: the aspect raster is replaced by its 1-degree bin code)
    ("front end: front_fast_kernel (interior tiles) + front_kernel (listed tiles): TMA-staged warp + dh + masks + Horn + bin codes",
     21.0),
    ("msel_*: median + MAD of dh in ONE pass with two brackets (slope median on a side stream)", 10.0),
    ("mask_update_kernel (MAD filter + y + moments of y, fused)", 9.0 + 4.0),
    ("(side streams only: scan/compress/dz median, med_dh)", 0.0),
    ("nuth_moments_reduce/final", 0.0),
    ("fb_pass* (bin count + exact bin medians; shares bandwidth with the side streams)", 12.0),
    ("result copy", 0.0),
]
# front 2 (fast + listed tiles) + median & MAD in one pass 10 (2 sample, bracket, main, 2 refine, rekey, find, 2 refine) +
# 3 selections x (sample gather, bracket, main, 2 refine) 15 + mask update 1 + scan/stride/compress 3 + moments 2 + bins 6 = 39
# (+ one z-shift fold every 4th step)
KERNELS_PER_STEP = 2 + 10 + 3 * 5 + 1 + 3 + 2 + 6


def peaks():
    fn = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(fn):
        with open(fn) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_workload(size, res, rank=0, cfg_id=2):
    from demcoreg_b200 import synth
    seed = 20260000 + 10 * cfg_id + 1000 * rank
    period = min(size, 8192)      # an 8192^2 period; nothing repeats at the bench size
    return synth.make_pair(n=size, res=res, seed=seed, period=period, static_mask_frac=0.3, nodata_frac=0.01,
                           outlier_frac=0.001, mask_tile=period)


class ClockSampler(object):
    """Sample SM clocks / throttle reasons with nvidia-smi while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def oracle_iteration_seconds(p, crop, shift_gt):
    """One oracle iteration (numpy/scipy restatement, 1 thread) on a crop of the workload; returns (seconds, pixels)."""
    from oracle import pygeotools_restated as pg, dem_align as oda
    ref = p['ref'][:crop, :crop]
    src = p['src'][:crop, :crop]
    msk = p['mask'][:crop, :crop] if p['mask'] is not None else None
    gt = p['gt']
    gts = (gt[0] + shift_gt[0], gt[1], 0.0, gt[3] + shift_gt[1], 0.0, gt[5])
    ms = oda.MaskSource(msk, gt) if msk is not None else None
    t0 = time.perf_counter()
    oda.compute_offset(pg.MemDataset(ref, gt, p['ndv']), pg.MemDataset(src, gts, p['ndv']), mask_source=ms)
    return time.perf_counter() - t0, crop * crop


def _ref_worker(args):
    """One worker = one single-threaded process (the reference's execution model) on ONE crop of the bench workload: the
    size x size pair of rank 0 is cut into g x g crops, worker (ci, cj) takes its crop of both rasters and of the mask."""
    size, res, gy, gx, ci, cj, nrep = args
    p = make_workload(size, res, 0)
    ch, cw = size // gy, size // gx
    q = dict(p)
    for k in ('ref', 'src', 'mask'):
        q[k] = np.ascontiguousarray(p[k][ci * ch:(ci + 1) * ch, cj * cw:(cj + 1) * cw]) if p[k] is not None else None
    gt = p['gt']
    q['gt'] = (gt[0] + cj * cw * gt[1], gt[1], 0.0, gt[3] + ci * ch * gt[5], 0.0, gt[5])
    del p
    out = []
    for _ in range(nrep):
        sec, _px = oracle_iteration_seconds(q, max(ch, cw), (3.4, -1.8))   # (the crop slices of that helper are no-ops here)
        out.append((sec, ch * cw))
    return out


def run_reference(args):
    """--impl reference: the reference's CPU path.  The reference itself cannot be imported (pygeotools / GDAL are
    absent), so this times the oracle restatement on the SAME workload as the GPU arm: the size x size pair is cut
    into g x g crops and every crop is one single-threaded process (the reference's only parallelism is one process per
    raster, `parallel -j`, dem_align_parallel.pbs:122) -- all host cores busy, a step = one iteration over every pixel of
    the pair."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = min(len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1), 128)
    nrep = args.warmup + args.steps
    g = max(1, min(int(np.floor(np.sqrt(cores))), args.size // 1024))
    while args.size % g:
        g -= 1
    gy = gx = g
    while gy * gx * 2 <= cores and args.size % (gx * 2) == 0 and args.size // (gx * 2) >= 512:
        gx *= 2                     # more cores than square crops: split the crops' columns once more
    crop = args.size // gy
    cropw = args.size // gx
    nw = gy * gx
    with mp.get_context("spawn").Pool(nw) as pool:
        out = pool.map(_ref_worker, [(args.size, args.res, gy, gx, k // gx, k % gx, nrep) for k in range(nw)])
    # per step: all workers run concurrently; the step time is the slowest worker's
    step_s = [max(out[w][k][0] for w in range(nw)) for k in range(args.warmup, nrep)]
    px = nw * crop * cropw
    val = px * len(step_s) / sum(step_s) / 1e9
    line = {"metric": "nuth_kaab_iteration_throughput", "value": val, "unit": "Gpix/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(step_s) / len(step_s),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": "dem_align -mode nuth, one iteration per step on a synthetic %dx%d fractal-terrain pair "
                                   "at %g m with a 30%% stable-surface mask, 1%% nodata (the GPU arm's pair), processed as "
                                   "%d crops of %dx%d by %d concurrent single-thread processes"
                                   % (args.size, args.size, args.res, nw, crop, cropw, nw),
                       "pixels_per_step": px},
            "cpu_baseline": {"value": val, "unit": "Gpix/s", "cores": nw, "kind": "port",
                             "sample": "every pixel of the %dx%d workload per step: %d concurrent single-thread oracle "
                                       "iterations (numpy/scipy restatement of the reference, plot path off) on its %dx%d "
                                       "crops; %d host cores available" % (args.size, args.size, nw, crop, cropw, cores)},
            "e2e": {"value": val, "unit": "Gpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_band(args):
    """Row-band mode (BASELINE.json configs[4]): one size x size pair, rows split over the ranks; NCCL send/recv for the
    halos, all-reduce of the selection / bin histograms between the phases of dcr_band_phase.  Strong scaling."""
    import torch
    import torch.distributed as dist
    from demcoreg_b200 import band, synth, _lib
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, res = args.size, args.res
    period = min(n, 2048)
    # every rank builds ONLY its own rows from the periodic tile (the halos come from the neighbours over NCCL)
    p = synth.make_pair(n=period, res=res, seed=20260050, static_mask_frac=0.3, nodata_frac=0.01)
    r0, r1 = band.band_rows(n, rank, world)
    rows = np.arange(r0, r1) % period
    reps = n // period

    def band_of(a):
        return torch.from_numpy(np.ascontiguousarray(np.tile(a[rows], (1, reps)))).cuda()
    gt = p['gt']
    al = band.BandAligner(band_of(p['ref']), band_of(p['src']), band_of(p['mask']), n, gt, p['ndv'], rank, world, halo=16)
    if world > 1:
        al.exchange_halos()
    else:
        pass  # single band: no neighbours

    def reduce_fn(t):
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)

    def step():
        nph = al.begin_iteration()
        for ph in range(nph):
            for t in al.run_phase(ph):
                reduce_fn(t)
        r = al.result()
        al.apply_shift(r.dx, r.dy, r.dz)
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    for _ in range(args.warmup):
        r = step()
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    pix = 0
    for _ in range(args.steps):
        r = step()
        pix += al.grid.height * al.grid.width
    e1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    if rank == 0:
        peak, peak_src = peaks()
        ms_step = ms_total / args.steps
        value = pix / (ms_total * 1e-3) / 1e9
        ach = ALG_BYTES_ITER * (pix / args.steps) / (ms_step * 1e-3) / 1e9
        line = {"metric": "nuth_kaab_iteration_throughput", "value": value, "unit": "Gpix/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": "dem_align -mode nuth, one iteration per step on ONE synthetic %dx%d pair at %g m (30%% "
                                       "stable-surface mask, 1%% nodata) split into %d row bands, NCCL halo exchange + "
                                       "histogram all-reduce (radix selects)" % (n, n, res, world),
                           "collectives_per_step": 22, "halo_rows": 16,
                           "l2": "inputs larger than the 126 MB L2", "last_step_shift_m": [r.dx, r.dy, r.dz]},
                "clocks": clocks, "gpu_launches": (1 + 5 * 6 + 6 + 9 + 1) * args.steps,
                "roofline_iteration": {"bound": "hbm", "achieved": ach, "peak": peak * world, "unit": "GB/s",
                                       "frac": ach / (peak * world), "algorithmic_bytes_per_px": ALG_BYTES_ITER,
                                       "peak_source": peak_src + " x n_gpus"},
                "e2e": None, "roofline": None, "cpu_baseline": None}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()



def bench_config3(args, torch):
    """BASELINE configs[2]: compute_offset_ncc and compute_offset_sad with a +-16 px window on a 4096^2 pair
    (device-resident masked rasters in, (2p+1)^2 float64 map + sub-pixel offset on the host out).  N = 1, rank 0."""
    from demcoreg_b200 import coreglib, engine, synth
    n, res, p = 4096, 8.0, 16
    pr = synth.make_pair_torch(n=n, res=res, seed=20260030, shift=(5.3 * res, -3.6 * res, 0.5), noise=0.3, outlier_frac=0.0,
                               static_mask_frac=0.05, nodata_frac=0.0)
    m = pr['mask']
    d1, d2 = pr['ref'], pr['src']
    J = 2 * p + 1
    kh = n - 2 * p
    g = torch.Generator(device="cuda")
    g.manual_seed(9)
    rn = torch.randn((n, n), generator=g, device="cuda")
    kn = torch.randn((kh, kh), generator=g, device="cuda")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(reps):
            out = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps, out
    ncc_ms, (mn, _) = timed(lambda: engine.ncc_corr(d1, m, d2, m, p, rn, kn), 3)
    info = {}
    sad_ms, ms_ = timed(lambda: engine.sad_search(d1, m, d2, m, p, info=info), 3)
    peak, _src = peaks()
    mneg = -ms_
    so = np.array(coreglib.find_subpixel_peak_position(mneg, 'parabolic')) - p
    sn = np.array(coreglib.find_subpixel_peak_position(mn, 'parabolic')) - p
    flops = 2.0 * J * J * kh * kh
    ndiff = float(J * J) * kh * kh
    return {"workload": "compute_offset_ncc / compute_offset_sad, pad=(16,16), synthetic 4096x4096 pair at 8 m displaced by "
                        "(5.3, -3.6) px, 5 % masked, device-resident",
            "ncc": {"ms": ncc_ms, "gflops": flops / (ncc_ms * 1e-3) / 1e9, "flop": flops,
                    "fp32_fma_peak_frac": flops / (ncc_ms * 1e-3) / (148 * 128 * 2 * 1.965e9),
                    "hbm_frac_at_8_B_per_px": 8.0 * n * n / (ncc_ms * 1e-3) / 1e9 / peak,
                    "sp_offset_px": [float(sn[0]), float(sn[1])],
                    "note": "direct shared-memory correlation, FP32 FMA over one tile row, FP64 beyond; includes the z-score passes"},
            "sad": {"ms": sad_ms, "gdiff_per_s": ndiff / (sad_ms * 1e-3) / 1e9,
                    "hbm_frac_at_24_B_per_px": 24.0 * n * n / (sad_ms * 1e-3) / 1e9 / peak,
                    "stages_ms": info.get('stage_ms'), "offsets_per_offset_path": info.get('n_per_offset'),
                    "sp_offset_px": [float(so[0]), float(so[1])],
                    "note": "batched over all 1089 offsets: sample brackets, 2 passes of 1.8e10 differences out of shared "
                            "memory, exact P25/P75 per offset"}}


def bench_h2d(torch, world, dist, nbytes=1 << 30, reps=6):
    """All ranks copy a pinned 1 GiB buffer host->device at the same time: the measured ceiling of the upload path."""
    h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return nbytes * reps / (float(t.item()) * 1e-3) / 1e9   # GB/s per GPU with all GPUs copying


def bench_batch(args, torch, rank, world, local, dist):
    """BASELINE configs[3]: a batch of DISTINCT host-resident 4096^2 pairs, pair p on rank p % world (no collective on the
    data path), every pair uploaded from pinned host memory and aligned to convergence through batch.align_stream."""
    from demcoreg_b200 import batch as _batch, synth
    n, res = args.batch_size, 8.0
    mine = _batch.shard(args.batch_pairs, rank, world)
    # each rank keeps to its own slice of the host cores (the per-pair host work is one thread)
    try:
        cpus = sorted(os.sched_getaffinity(0))
        per = max(2, len(cpus) // max(1, world))
        sl = cpus[(local * per) % len(cpus):(local * per) % len(cpus) + per]
        if len(sl) >= 2:
            os.sched_setaffinity(0, sl)
    except Exception:
        cpus = None
    rng = np.random.default_rng(20260040)
    shifts = rng.uniform(-8.0, 8.0, size=(args.batch_pairs, 2))
    dzs = rng.uniform(-1.0, 1.0, size=args.batch_pairs)
    nloc = len(mine)
    h_ref = torch.empty((nloc, n, n), dtype=torch.float32, pin_memory=True)
    h_src = torch.empty((nloc, n, n), dtype=torch.float32, pin_memory=True)
    h_msk = torch.empty((nloc, n, n), dtype=torch.uint8, pin_memory=True)
    scale = synth.terrain_scale_torch(torch, n, res, 20260040, "cuda")
    gt = None
    for k, pidx in enumerate(mine):
        pr = synth.make_pair_torch(n=n, res=res, seed=20260040 + 7 * pidx, shift=(float(shifts[pidx, 0]), float(shifts[pidx, 1]),
                                   float(dzs[pidx])), scale=scale, static_mask_frac=0.3, nodata_frac=0.01)
        h_ref[k].copy_(pr['ref'])
        h_src[k].copy_(pr['src'])
        h_msk[k].copy_(pr['mask'])
        gt = pr['gt']
    torch.cuda.synchronize()
    del pr
    torch.cuda.empty_cache()
    pairs = [dict(ref=h_ref[k], src=h_src[k], mask=h_msk[k], ref_gt=gt, src_gt=gt, mask_gt=gt, ndv=synth.NDV) for k in range(nloc)]
    h2d_gbs = bench_h2d(torch, world, dist)
    _batch.align_stream(pairs[:2])          # warm-up: buffers, workspace, first-use initialisation
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    res_list = _batch.align_stream(pairs)
    e1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([e0.elapsed_time(e1) * 1e-3, wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    err = 0.0
    iters = 0
    fb = 0
    bad = 0
    for k, pidx in enumerate(mine):
        r = res_list[k]
        if r['exit_code']:
            bad += 1
            continue
        err = max(err, abs(r['dx'] - shifts[pidx, 0]), abs(r['dy'] - shifts[pidx, 1]), abs(r['dz'] + dzs[pidx]))
        iters += r['n_iter']
        fb += r['select_fallbacks'] + r['bin_fallbacks']
    agg = torch.tensor([err, float(iters), float(fb), float(bad)], dtype=torch.float64, device="cuda")
    if world > 1:
        mx = agg[:1].clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = agg[1:].clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        agg = torch.cat([mx, sm])
    agg = agg.cpu().numpy()
    secs = float(t[0].item())
    bytes_pair = n * n * 9
    return {"workload": "%d distinct synthetic %dx%d pairs at %g m (30 %% stable-surface mask, 1 %% nodata, shifts U(-8,8) m), "
                        "host-resident (pinned), pair p on rank p %% %d, dem_align -mode nuth to tol 0.02 m + final statistics pass"
                        % (args.batch_pairs, n, n, res, world),
            "pairs": args.batch_pairs, "pairs_per_s": args.batch_pairs / secs, "seconds": secs, "wall_seconds": float(t[1].item()),
            "h2d_bytes_per_pair": bytes_pair, "upload_gbs_per_gpu_needed": bytes_pair * (args.batch_pairs / world) / secs / 1e9,
            "h2d_gbs_per_gpu_measured_all_ranks_copying": h2d_gbs,
            "upload_bound_pairs_per_s": world * h2d_gbs * 1e9 / bytes_pair,
            "mean_iterations": float(agg[1]) / max(1, args.batch_pairs - int(agg[3])),
            "max_abs_error_vs_injected_m": float(agg[0]), "engine_fallbacks": int(agg[2]), "pairs_failed": int(agg[3])}


def bench_band(args, torch, rank, world, local, dist):
    """BASELINE configs[4]: ONE band_size^2 pair at 2 m split into row bands over the ranks (NCCL halo exchange, all-reduced
    histograms).  Every rank builds only its own rows, on the device: an 8192^2-periodic clean terrain pair tiled across
    the raster + white noise that is unique to every pixel of the big raster (seeded per 1024-row block, so the raster
    does not depend on the number of ranks)."""
    from demcoreg_b200 import band, synth
    n, res, per = args.band_size, 2.0, min(args.band_size, 8192)
    base = synth.make_pair_torch(n=per, res=res, seed=20260050, noise=0.0, outlier_frac=0.0, static_mask_frac=0.3, nodata_frac=0.01)
    r0, r1 = band.band_rows(n, rank, world)
    reps = n // per
    rows = torch.arange(r0, r1, device="cuda") % per

    def band_of(a):
        return a[rows].repeat(1, reps)
    ref_b = band_of(base['ref'])
    msk_b = band_of(base['mask'])
    src_b = band_of(base['src'])
    hole = src_b == synth.NDV
    blk = 1024
    for b0 in range((r0 // blk) * blk, r1, blk):
        g = torch.Generator(device="cuda")
        g.manual_seed(20260051 + b0 // blk)
        nz = torch.randn((blk, n), generator=g, device="cuda", dtype=torch.float32) * 0.3
        lo, hi = max(b0, r0), min(b0 + blk, r1)
        src_b[lo - r0:hi - r0] += nz[lo - b0:hi - b0]
    src_b[hole] = synth.NDV
    del base, hole, nz
    torch.cuda.empty_cache()
    al = band.BandAligner(ref_b, src_b, msk_b, n, (500000.0, res, 0.0, 4100000.0, 0.0, -res), synth.NDV, rank, world, halo=64,
                          keep_aspect=False)
    del ref_b, src_b, msk_b
    torch.cuda.empty_cache()
    comm = band.Comm()
    band.exchange_halos_native(al, comm)

    def step(native=True):
        r = band.iteration_native(al, comm) if native else band.iteration_distributed(al)
        al.apply_shift(r.dx, r.dy, r.dz)
        return r

    def timed(native, nst):
        for _ in range(2):
            r = step(native)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        pix = 0
        for _ in range(nst):
            r = step(native)
            pix += al.grid.height * al.grid.width
            engines.append((int(r.select_engine), int(r.bin_engine)))
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()) / nst, pix, r
    nst = 4
    engines = []
    ms_py, _, _ = timed(False, 2)     # interpreter-driven phases over torch.distributed (round 1's path), for comparison
    ms_step, pix, r = timed(True, nst)
    peak, _ = peaks()
    ach = ALG_BYTES_ITER * (pix / nst) / (ms_step * 1e-3) / 1e9
    torch.cuda.synchronize()
    comm.close()
    return {"workload": "dem_align -mode nuth, one iteration per step on ONE synthetic %dx%d pair at %g m (30 %% stable-surface "
                        "mask, 1 %% nodata) split into %d row bands; NCCL halo exchange + histogram all-reduces" % (n, n, res, world),
            "ms_per_iteration": ms_step, "gpix_per_s": pix / nst / (ms_step * 1e-3) / 1e9, "steps": nst,
            "roofline_frac_of_n_x_peak": ach / (peak * world), "halo_rows": 64,
            "path": "dcr_band_iteration: 33 phases + their ncclAllReduce enqueued natively on one stream (library-owned "
                    "communicator); fast front kernel on the band, bins from the front kernel, one-pass sample-bracket selects "
                    "(sample keys / counts / histograms all-reduced), deferred z shifts",
            "ms_per_iteration_interpreter_phases": ms_py,
            "last_step_shift_m": [r.dx, r.dy, r.dz], "select_engine": int(r.select_engine), "bin_engine": int(r.bin_engine),
            "radix_reruns": sum(1 for e in engines if e != (1, 1)), "steps_counted": len(engines)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", type=int, default=8192)
    ap.add_argument("--res", type=float, default=8.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config3", action="store_true", help="skip the NCC / SAD object (BASELINE configs[2])")
    ap.add_argument("--no-batch", action="store_true", help="skip the 256-pair batch object (BASELINE configs[3])")
    ap.add_argument("--no-band", action="store_true", help="skip the row-band object at N > 1 (BASELINE configs[4])")
    ap.add_argument("--batch-pairs", type=int, default=256)
    ap.add_argument("--batch-size", type=int, default=4096)
    ap.add_argument("--band-size", type=int, default=32768)
    ap.add_argument("--mode", default="pairs", choices=["pairs", "band"],
                    help="pairs: one pair per GPU, no collective (default; configs[1]/[3]); band: ONE size x size pair "
                         "split into row bands over the GPUs with NCCL halo exchange + all-reduces (configs[4])")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.mode == "band":
        return run_band(args)

    import torch
    import torch.distributed as dist
    from demcoreg_b200 import _lib, engine, coreglib
    from demcoreg_b200.raster import Raster, MaskSource
    import contextlib
    import io

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    size, res = args.size, args.res
    p = make_workload(size, res, rank)
    ref = Raster(p['ref'], p['gt'], p['ndv'])
    src0 = Raster(p['src'], p['gt'], p['ndv'])
    ms = MaskSource(p['mask'], p['gt'])

    quiet = contextlib.redirect_stdout(io.StringIO())

    class Grid2(object):
        pass

    def make_stepper(src):
        # dem_align's loop body (compute_offset + apply_xy_shift + apply_z_shift, reference dem_align.py:306-343), one
        # native call per iteration; tol = 0 and a huge max_iter so that the loop never "converges" during the bench
        return engine.NuthAligner(ref, src, ms, tol=0.0, max_offset=1e9, max_iter=1 << 30)

    def step(al, profile=False):
        r = al.step(profile=profile)
        g = Grid2()
        g.height, g.width = al.grid_hw
        return r, g

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident arm -------------------------------------------------------------
    src = src0.copy()
    al = make_stepper(src)
    for _ in range(args.warmup):
        r, grid = step(al)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    # EXACTLY args.steps iterations in one native call (dcr_align_nuth, max_steps = K): what dem_align.align runs, with no
    # interpreter work between iterations
    recs = al.run(max_steps=args.steps)
    e1.record()
    barrier()
    assert len(recs) == args.steps and all(rc['status'] == 0 for rc in recs), "bench iterations did not all run"
    pix = sum(rc['h'] * rc['w'] for rc in recs)
    engines = {"select_engine": "one-pass sample-bracket" if all(rc['select_engine'] for rc in recs) else "radix fallback seen",
               "radix_reruns": sum(1 for rc in recs if not rc['select_engine']),
               "bin_engine": "two-pass shared-memory bins" if all(rc['bin_engine'] for rc in recs) else "radix bins seen",
               "fast_fallback": sum(1 for rc in recs if not rc['bin_engine'])}
    r = al.st.last
    clocks = sampler.stop() if rank == 0 else None
    ms_total = e0.elapsed_time(e1)
    t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = world * pix / (ms_total * 1e-3) / 1e9

    # ---- per-stage CUDA-event timing (same stream, same workload) -> roofline of the dominant stage
    stage = np.zeros(7)
    nprof = max(3, min(args.steps, 5))
    for _ in range(nprof):
        r, grid = step(al, profile=True)
        stage += np.array(r.stage_ms[:7])
    stage /= nprof
    npx = grid.height * grid.width
    peak, peak_src = peaks()
    dom = int(np.argmax(stage[:6]))
    dom_bytes = STAGE_INFO[dom][1] * npx
    ach = dom_bytes / (stage[dom] * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": STAGE_INFO[dom][0], "achieved": ach, "peak": peak, "unit": "GB/s",
                "frac": ach / peak, "traffic": (FRONT_KERNEL_DRAM_BYTES_8192 if (dom == 0 and size == 8192) else None),
                "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom_bytes, "ms_per_launch": float(stage[dom]),
                "stages_ms": {STAGE_INFO[k][0]: float(stage[k]) for k in range(7)}}
    iter_ach = ALG_BYTES_ITER * npx / (ms_step * 1e-3) / 1e9
    roofline_iter = {"bound": "hbm", "achieved": iter_ach, "peak": peak, "unit": "GB/s", "frac": iter_ach / peak,
                     "algorithmic_bytes_per_px": ALG_BYTES_ITER}

    # ---- DEM pairs/s (second half of the BASELINE metric): the whole dem_align loop to convergence on the
    #      device-resident pair (reference dem_align.py:306-357), per GPU, aggregated over the ranks
    from demcoreg_b200 import dem_align as _da
    _da.align(ref, src0, mask_source=ms, verbose=False)   # untimed: first-use allocations of the loop's own buffers
    barrier()
    npairs = 3
    e0.record()
    for _ in range(npairs):
        res_loop = _da.align(ref, src0, mask_source=ms, verbose=False)
    e1.record()
    barrier()
    tp = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    pairs_per_s = world * npairs / (float(tp.item()) * 1e-3)

    # ---- end-to-end arm: host buffers in, (dx,dy,dz) out, through the public API -------------
    h_ref = torch.from_numpy(p['ref']).pin_memory()
    h_src = torch.from_numpy(p['src']).pin_memory()
    h_msk = torch.from_numpy(p['mask']).pin_memory()
    al.finish()
    gts = src.gt  # a realistic sub-pixel state (after the timed steps)
    # demcoreg_b200.batch.stream_pairs: two device buffer sets, the upload of step k+1 (copy stream, pinned host memory)
    # overlaps the iteration of step k, as a batch of pairs streams through one GPU; every step's 604 MB of H2D copies
    # lie inside the timed region
    from demcoreg_b200 import batch as _batch
    hp = dict(ref=h_ref, src=h_src, mask=h_msk, ref_gt=ref.gt, src_gt=gts, mask_gt=p['gt'], ndv=p['ndv'])

    def e2e_run(nsteps):
        pix = 0
        for d_ref, d_src, d_ms in _batch.stream_pairs([hp] * nsteps):
            rr, g2, _ = engine.nuth_iteration(d_ref, d_src, d_ms)   # host gets (dx, dy, dz): D2H inside
            pix += g2.height * g2.width
        return pix

    e2e_run(2)
    barrier()
    e0.record()
    n_e2e = max(3, min(args.steps, 5))
    pix2 = e2e_run(n_e2e)
    e1.record()
    barrier()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_val = world * pix2 / (float(t.item()) * 1e-3) / 1e9
    h2d = int(h_ref.numel() * 4 + h_src.numel() * 4 + h_msk.numel())
    import ctypes
    d2h = int(ctypes.sizeof(_lib.IterResult))

    # free the 8192^2 working set before the other configurations
    del al, src, src0, ref, ms, h_ref, h_src, h_msk, hp
    engine._ws_cache.clear()
    torch.cuda.empty_cache()
    batch_obj = band_obj = config3_obj = None
    if not args.no_batch:
        batch_obj = bench_batch(args, torch, rank, world, local, dist)
        engine._ws_cache.clear()
        torch.cuda.empty_cache()
    if world > 1 and not args.no_band:
        band_obj = bench_band(args, torch, rank, world, local, dist)
        engine._ws_cache.clear()
        torch.cuda.empty_cache()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    if world == 1 and not args.no_config3:
        config3_obj = bench_config3(args, torch)

    # ---- CPU baseline: the oracle on a bounded sample of the same workload (rank 0, N=1 only) ----
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        crop = min(size, 4096)
        sec, px = oracle_iteration_seconds(p, crop, (gts[0] - p['gt'][0], gts[3] - p['gt'][3]))
        cpu = {"value": px / sec / 1e9, "unit": "Gpix/s", "cores": 1, "kind": "port",
               "sample": "one oracle iteration (numpy/scipy restatement, plot path off) on the %dx%d top-left crop of "
                         "the workload, %.1f s" % (crop, crop, sec)}

    line = {"metric": "nuth_kaab_iteration_throughput", "value": value, "unit": "Gpix/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "dem_align -mode nuth, one iteration per step on a synthetic %dx%d fractal-terrain pair "
                                   "at %g m with a 30%% stable-surface mask, 1%% nodata, injected shift (+3.2,-1.7,+0.5) m; "
                                   "one pair per GPU" % (size, size, res),
                       "pixels_per_step_per_gpu": npx, "l2": "inputs (2 x %d MB + mask) larger than the 126 MB L2" % (npx * 4 >> 20),
                       "last_step_shift_m": [r.dx, r.dy, r.dz], "engines": engines,
                       "period": "terrain and masks generated at the full %d^2 size (no tiling)" % min(size, 8192),
                       "pairs_per_s": pairs_per_s, "pairs_loop": {"n_iter": res_loop['n_iter'], "dx": res_loop['dx'],
                                                                  "dy": res_loop['dy'], "dz": res_loop['dz'],
                                                                  "note": "dem_align.align to convergence (tol 0.02 m), device-resident pair, incl. the src copy"}},
            "clocks": clocks, "gpu_launches": KERNELS_PER_STEP * args.steps,
            "e2e": {"value": e2e_val, "unit": "Gpix/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    # the ceiling the upload path allows: measured H2D rate per GPU with all ranks copying at once
                    "h2d_gbs_per_gpu_all_ranks_copying": (batch_obj or {}).get("h2d_gbs_per_gpu_measured_all_ranks_copying"),
                    "upload_bound_value": (world * (batch_obj["h2d_gbs_per_gpu_measured_all_ranks_copying"] * 1e9 / h2d) * npx / 1e9
                                           if batch_obj else None),
                    "note": "engine.nuth_iteration on host-resident pinned rasters through batch.stream_pairs: uploads double-"
                            "buffered on a copy stream (step k+1's H2D overlaps step k's kernels); bound by the host->device "
                            "copies (PCIe at N=1, the host side of the box when all ranks copy at once)"},
            "roofline": roofline, "roofline_iteration": roofline_iter, "cpu_baseline": cpu,
            "config3": config3_obj, "batch": batch_obj, "band": band_obj}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
