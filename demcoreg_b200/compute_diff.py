#! /usr/bin/env python
"""Drop-in for reference ``demcoreg/compute_diff.py:28-140``: warp two rasters to a common grid
(``memwarp_multi``: extent ``-te`` intersection / union / first / last, resolution ``-tr`` min / max / mean / a number -- the
general GDAL cubic resample handles inputs of different pixel size) and write ``r2 - r1``; ``-rate`` adds the change rate in
m/yr from the dates in the two file names (the reference's ``*_ts.tif`` timestamp-raster variant is not built).

Same primitive as the iteration's front end (``dcr_warp_diff_horn``).
"""
import argparse
import datetime
import os
import re
import sys

import numpy as np

from . import _lib, engine
from . import rasterio_lite as rio


def compute_diff(r1, r2, extent='intersection', res='max'):
    """``r2 - r1`` on the common grid (reference ``compute_diff.py:97-105``) -> (masked array, geotransform)."""
    if abs(r1.gt[1] - r2.gt[1]) >= 1e-3 or abs(engine.target_res([r1, r2], res) - r1.gt[1]) >= 1e-3:
        # different pixel sizes (or a target resolution of its own): resample first, then the two share a lattice
        r1, r2 = engine.memwarp_multi([r1, r2], extent=extent, res=res)
        extent = 'intersection'
    grid, bufs = engine.front_only(r1, r2, None, remove_outliers=False, keep_warped=False, extent=extent)
    base = ((bufs.maskbits & _lib.M_BASE) != 0).cpu().numpy()
    return np.ma.array(bufs.dh.cpu().numpy(), mask=base), tuple(grid.gt)


def fn_getdatetime(fn):
    """Date in a file name (``pygeotools.timelib.fn_getdatetime`` <- ``compute_diff.py:81-82``, its common patterns):
    the first ``YYYYMMDD`` (1950-2099) in the base name, optionally followed by ``[_T]HHMM[SS]``; ``None`` without one."""
    base = os.path.basename(fn)
    for m in re.finditer(r'(?<!\d)((?:19[5-9]\d|20\d\d)(?:0[1-9]|1[0-2])(?:0[1-9]|[12]\d|3[01]))(?:[_T]?(\d{6}|\d{4})(?!\d))?', base):
        try:
            d = datetime.datetime.strptime(m.group(1), '%Y%m%d')
        except ValueError:
            continue
        t = m.group(2)
        if t:
            try:
                d = d.replace(hour=int(t[0:2]), minute=int(t[2:4]), second=int(t[4:6]) if len(t) == 6 else 0)
            except ValueError:
                pass
        return d
    return None


def getparser():
    parser = argparse.ArgumentParser(description="Compute difference between two rasters")
    parser.add_argument('-outdir', default=None, help='Output directory')
    parser.add_argument('-tr', default='max', help='Output resolution (default: %(default)s)')
    parser.add_argument('-te', default='intersection', help='Output extent (default: %(default)s)')
    parser.add_argument('-t_srs', default='first', help='Output projection (default: %(default)s)')
    parser.add_argument('-rate', action='store_true',
                        help='Attempt to generate change rate products (dz/dt) in m/yr (default: %(default)s)')
    parser.add_argument('fn1', type=str, help='Raster filename 1')
    parser.add_argument('fn2', type=str, help='Raster filename 2')
    return parser


def main(argv=None):
    args = getparser().parse_args(argv)
    if args.fn1 == args.fn2:
        sys.exit('Input filenames are identical')                          # compute_diff.py:37-38
    if args.te not in ('intersection', 'union', 'first', 'last'):
        sys.exit("-te %s: extents intersection / union / first / last are built" % args.te)
    t_factor = None
    if args.rate:                                                           # compute_diff.py:53-93
        t1, t2 = fn_getdatetime(args.fn1), fn_getdatetime(args.fn2)
        if t1 is not None and t2 is not None and t1 != t2:
            dt = t2 - t1
            t_factor = abs(dt.total_seconds() / datetime.timedelta(days=365.25).total_seconds())
            print("Time differences is %s, dh/%0.3f" % (dt, t_factor))
        else:
            print("Unable to extract timestamps for input images")
            args.rate = False
    r1, r2 = rio.read_raster(args.fn1), rio.read_raster(args.fn2)
    print("Warping rasters to same res/extent/proj")
    print("Computing raster difference")
    diff, gt = compute_diff(r1, r2, extent=args.te, res=args.tr)
    if diff.count() == 0:                                   # compute_diff.py:107-109
        sys.exit("No valid overlap between input rasters")
    print("Raster difference stats:")                       # compute_diff.py:123-125 (malib.print_stats)
    from . import robust_stats
    st = robust_stats.get_stats_dict(diff, full=False)
    print("count: %i min: %0.2f max: %0.2f mean: %0.2f std: %0.2f med: %0.2f mad: %0.2f" % (
        st['count'], st['min'], st['max'], st['mean'], st['std'], st['med'], st['nmad']))
    outdir = args.outdir or os.path.split(os.path.abspath(args.fn1))[0]
    if not os.path.exists(outdir):
        os.makedirs(outdir)
    outprefix = '%s_%s' % (os.path.splitext(os.path.basename(args.fn1))[0], os.path.splitext(os.path.basename(args.fn2))[0])
    dst = os.path.join(outdir, outprefix + '_diff.tif')
    print("Writing raster difference map")
    rio.write_raster(dst, diff.filled(-9999.0), gt, -9999.0, r1.proj)
    print(dst)
    if args.rate:                                           # compute_diff.py:132-136
        print("Writing rate map")
        rate_fn = os.path.join(outdir, outprefix + '_diff_rate.tif')
        rio.write_raster(rate_fn, (diff / np.float32(t_factor)).astype(np.float32).filled(-9999.0), gt, -9999.0, r1.proj)
        print(rate_fn)
    return dst


if __name__ == "__main__":
    main()
