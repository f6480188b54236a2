#! /usr/bin/env python
"""Drop-in for reference ``demcoreg/robust_stats.py`` and for ``malib.get_stats_dict`` as used by
``dem_align.py:384, 392`` -- robust statistics of a dz raster, computed on the device.

Order statistics come from the exact selection kernels (``dcr_select_quantiles``), moments from the FP64
reduction kernel (``dcr_moments``).  The reference's quirks are kept: ``Count`` prints ``n - 1``
(``robust_stats.py:55``) and the spread percentiles are the asymmetric (15.9, 84.2) (``robust_stats.py:70``).
"""
import argparse
import os
import sys

import numpy as np

from . import _lib, engine


def _as_device(dz, mask=None):
    torch = _lib.require_cuda()
    if isinstance(dz, np.ma.MaskedArray):
        mask = np.ma.getmaskarray(dz)
        dz = dz.filled(0)
    if isinstance(dz, np.ndarray):
        dz = torch.from_numpy(np.ascontiguousarray(dz, dtype=np.float32)).cuda()
    if isinstance(mask, np.ndarray):
        mask = torch.from_numpy(np.ascontiguousarray(mask).view(np.uint8)).cuda()
    if mask is not None:
        mask = (mask != 0).to(torch.uint8).contiguous()
    return dz.contiguous(), mask


def robust_stats(dz, mask=None):
    """The statistics ``robust_stats.py:55-86`` prints, as a dict (device computation)."""
    x, m = _as_device(dz, mask)
    mom = engine.moments(x, m, 1)
    n = mom['count']
    out = {'count': n - 1}
    out['rmse'] = float(np.sqrt(mom['sumsq'] / n)) if n else float('nan')
    out['mean'] = mom['mean']
    out['std'] = mom['std']
    (med, p16, p84), _ = engine.select_quantiles(x, [50.0, 15.9, 84.2], m, 1)
    out['med'] = float(med)
    out['p16'], out['p84'] = float(p16), float(p84)
    out['spread'] = float(np.float32(p84) - np.float32(p16))
    (absmed, p68, p95), _ = engine.select_quantiles(x, [50.0, 68.3, 95.0], m, 1, transform=1, center=0.0)
    out['absmed'] = float(absmed)
    (mad,), _ = engine.select_quantiles(x, [50.0], m, 1, transform=1, center=float(med))
    out['nmad'] = 1.4826 * float(mad)
    out['p68'], out['p95'] = float(p68), float(p95)
    return out


def get_stats_dict(dz, mask=None, full=True):
    """``malib.get_stats_dict(a, full=True)`` (reference ``dem_align.py:384, 392, 522, 530``):
    count, min, max, ptp, mean, std, nmad, med, median, p16, p84, spread, mode."""
    torch = _lib.require_cuda()
    x, m = _as_device(dz, mask)
    mom = engine.moments(x, m, 1)
    d = {}
    n = mom['count']
    if n == 0:
        return d
    stride = 1
    if not full and n >= 4e6:
        stride = int(np.around(n / 4e6))
    xs, ms = x, m
    if stride > 1:
        xs = engine.compress_strided(x.reshape(-1), m.reshape(-1) if m is not None else None, 1, stride)
        ms = None
    (med, p16, p84), _ = engine.select_quantiles(xs, [50.0, 15.9, 84.1], ms, 1)
    (mad,), _ = engine.select_quantiles(xs, [50.0], ms, 1, transform=1, center=float(med))
    d['count'] = int(n)
    d['min'] = float(mom['min'])
    d['max'] = float(mom['max'])
    d['ptp'] = float(np.float32(mom['max']) - np.float32(mom['min']))
    d['mean'] = float(mom['mean'])
    d['std'] = float(mom['std'])
    d['nmad'] = float(np.float32(mad) * np.float32(1.4826))
    d['med'] = float(med)
    d['median'] = float(med)
    d['p16'] = float(p16)
    d['p84'] = float(p84)
    d['spread'] = float(abs(np.float32(p84) - np.float32(p16)))
    # scipy.stats.mstats.mode: smallest most-frequent value.  Not on the hot path (its result is unused by
    # dem_align); torch.unique is used as a library call here.
    v = xs.reshape(-1) if ms is None else xs.reshape(-1)[ms.reshape(-1) == 0]
    vals, counts = torch.unique(v, return_counts=True)
    d['mode'] = float(vals[int(torch.argmax(counts))])
    return d


def getparser():
    parser = argparse.ArgumentParser(description="Compute robust stats from dz or pc_align sample.csv")
    parser.add_argument('fn', type=str, help='Input filename')
    parser.add_argument('-outdir', default=None, help='Output directory')
    parser.add_argument('-col', type=int, default=4, help='Column containing values for stats')
    return parser


def main(argv=None):
    parser = getparser()
    args = parser.parse_args(argv)
    fn = args.fn
    ext = os.path.splitext(fn)[1]
    if 'csv' in ext:
        a = np.genfromtxt(fn, delimiter=',', skip_header=1)
        dz_m = a[:, args.col]
        dz_m = dz_m[~np.isnan(dz_m)]
        st = robust_stats(dz_m.astype(np.float32))
    elif 'tif' in ext or 'npz' in ext:
        from . import rasterio_lite as rio
        r = rio.read_array(fn)
        a = np.ma.masked_values(r['array'], r['ndv'], shrink=False) if r['ndv'] is not None else np.ma.array(r['array'])
        st = robust_stats(a)
    else:
        sys.exit('Unsupported input type')
    print("Count: %i" % st['count'])
    print("RMSE: %0.3f" % st['rmse'])
    print("Mean Error: %0.3f" % st['mean'])
    print("Standard Deviation: %0.3f" % st['std'])
    print("Median Error: %0.3f" % st['med'])
    print("16th Percentile: %0.3f" % st['p16'])
    print("84th Percentile: %0.3f" % st['p84'])
    print("Spread: %0.3f" % st['spread'])
    print("Absolute Median Error: %0.3f" % st['absmed'])
    print("NMAD: %0.3f" % st['nmad'])
    print("68.3th Percentile: %0.3f" % st['p68'])
    print("95th Percentile: %0.3f" % st['p95'])
    return st


if __name__ == "__main__":
    main()
