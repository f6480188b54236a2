"""Small invocations of every kernel family of the hot path, for compute-sanitizer (profiles/sanitize.sh):
front kernels (fast interior tiles + listed tiles, manual staging), one-pass selects incl. heavy ties, radix selects, mask
update, fast and radix bins, the native loop with deferred z shifts, row bands (2 virtual ranks), NCC / SAD, tilt fit."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from demcoreg_b200 import band, coreglib, dem_align, engine, synth, tiltcorr  # noqa: F401
from demcoreg_b200.raster import MaskSource, Raster

which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = 384
p = synth.make_pair(n=n, res=30.0, seed=20260001, static_mask_frac=0.2, nodata_frac=0.01)


def gds(key, gt=None):
    return Raster(p[key], gt or p['gt'], p['ndv'])


ms = MaskSource(p['mask'], p['gt'])
g = list(p['gt']); g[0] += 13.7; g[3] -= 21.3
if which in ("all", "iter"):
    for tan in (False, True):
        r, _, _ = engine.nuth_iteration(gds('ref'), gds('src', tuple(g)), ms, tan_slope=tan)
        assert r.status == 0
    r, _, _ = engine.nuth_iteration(gds('ref'), gds('src', tuple(g)), ms, radix_only=True)
    os.environ["DCR_NO_TMA"] = "1"
    r, _, _ = engine.nuth_iteration(gds('ref'), gds('src', tuple(g)), ms)
    del os.environ["DCR_NO_TMA"]
    print("iteration ok", r.dx, r.dy, r.dz)
if which in ("all", "loop"):
    out = dem_align.align(gds('ref'), gds('src'), mask_source=ms, verbose=False, tol=0.0, max_iter=6)
    print("loop ok", out['n_iter'], out['dx'], out['dy'], out['dz'])
if which in ("all", "ties"):
    # heavy ties: quantised values -> bracket verification / radix fallbacks of the selects and the bins
    q = dict(p)
    q['src'] = np.where(p['src'] == p['ndv'], p['ndv'], np.round(p['src'] * 2) / 2).astype(np.float32)
    q['ref'] = np.where(p['ref'] == p['ndv'], p['ndv'], np.round(p['ref'] * 2) / 2).astype(np.float32)
    r, _, _ = engine.nuth_iteration(Raster(q['ref'], p['gt'], p['ndv']), Raster(q['src'], p['gt'], p['ndv']), ms)
    print("ties ok", r.status, r.select_engine, r.bin_engine)
if which in ("all", "band"):
    vg = band.VirtualGroup(p['ref'], p['src'], p['mask'], p['gt'], p['ndv'], 2, halo=8)
    res = vg.iteration()
    print("band ok", res[0].dx, res[0].dy)
if which in ("all", "corr"):
    m = p['mask'].astype(bool)
    a = np.ma.array(p['ref'][:200, :200], mask=m[:200, :200])
    b = np.ma.array(p['src'][:200, :200], mask=m[:200, :200])
    mm, io, so, _ = coreglib.compute_offset_ncc(a, b, pad=(4, 4))
    mm, io, so = coreglib.compute_offset_sad(a, b, pad=(3, 3))
    print("corr ok", io, so)
if which in ("all", "tilt"):
    out = dem_align.align(gds('ref'), gds('src'), verbose=False, tiltcorr=True, max_iter=3)
    print("tilt ok", out['n_iter'])
torch.cuda.synchronize()
print("driver done")
