"""Row-band mode under real NCCL: run with
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        -m demcoreg_b200.tools.band_check [--size 2048] [--iters 3]
Every rank builds only its own rows of a synthetic pair, halos are exchanged with NCCL send/recv, the iteration
runs phase by phase with NCCL all-reduces, and rank 0 compares every iteration with the single-GPU path.
Prints one JSON line; exit code 1 on mismatch."""
import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=2048)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--halo", type=int, default=16)
    ap.add_argument("--native", action="store_true",
                    help="the library's own NCCL communicator: dcr_band_exchange_halos + dcr_band_iteration (one native call "
                         "per iteration) instead of interpreter-driven phases over torch.distributed")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    from demcoreg_b200 import band, engine, coreglib, synth
    from demcoreg_b200.raster import Raster, MaskSource
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = args.size
    p = synth.make_pair(n=n, res=8.0, seed=20260050, period=min(n, 1024), static_mask_frac=0.3, nodata_frac=0.01)
    r0, r1 = band.band_rows(n, rank, world)
    al = band.BandAligner(torch.from_numpy(p['ref'][r0:r1]).cuda(), torch.from_numpy(p['src'][r0:r1]).cuda(),
                          torch.from_numpy(p['mask'][r0:r1]).cuda(), n, p['gt'], p['ndv'], rank, world, halo=args.halo)
    comm = None
    if args.native:
        comm = band.Comm()
        band.exchange_halos_native(al, comm)
    else:
        al.exchange_halos()
    ok = True
    msgs = []
    if rank == 0:
        ref = Raster(p['ref'], p['gt'], p['ndv'])
        src = Raster(p['src'], p['gt'], p['ndv'])
        ms = MaskSource(p['mask'], p['gt'])
    for it in range(args.iters):
        res = band.iteration_native(al, comm) if args.native else band.iteration_distributed(al)
        if rank == 0:
            r1g, _, _ = engine.nuth_iteration(ref, src, ms, radix_only=True)
            same = (list(res.nuth.bin_count) == list(r1g.nuth.bin_count) and res.count_final == r1g.count_final
                    and res.dz_med == r1g.dz_med and res.med_dh == r1g.med_dh
                    and abs(res.dx - r1g.dx) < 1e-9 and abs(res.dy - r1g.dy) < 1e-9)
            ok = ok and same
            msgs.append(dict(iter=it, dx=res.dx, dy=res.dy, dz=res.dz, same_as_single_gpu=bool(same)))
            with contextlib.redirect_stdout(io.StringIO()):
                coreglib.apply_xy_shift(src, r1g.dx, r1g.dy, createcopy=False)
                coreglib.apply_z_shift(src, np.float32(r1g.dz), createcopy=False)
        al.apply_shift(res.dx, res.dy, res.dz)
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        print(json.dumps(dict(world=world, size=n, native=bool(args.native), ok=bool(ok), iterations=msgs)))
    if comm is not None:
        torch.cuda.synchronize()
        comm.close()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
