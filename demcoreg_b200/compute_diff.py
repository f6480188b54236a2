#! /usr/bin/env python
"""Drop-in for the core of reference ``demcoreg/compute_diff.py:97-133``: warp two rasters to their common grid
(``memwarp_multi``, equal resolution, extent ``-te`` intersection / union / first / last) and write ``r2 - r1``.

Same primitive as the iteration's front end (``dcr_warp_diff_horn``).  ``-tr`` is accepted for rasters of equal
resolution (min / max / mean coincide); unequal resolutions need the down-sampling GDAL warp, which is out of scope
(SURVEY 8f-3), as are the ``-rate`` products (time stamps parsed from file names by pygeotools.timelib).
"""
import argparse
import os

import numpy as np

from . import _lib, engine
from . import rasterio_lite as rio


def compute_diff(r1, r2, extent='intersection'):
    """``r2 - r1`` on the common grid (reference ``compute_diff.py:97-105``) -> (masked array, geotransform)."""
    grid, bufs = engine.front_only(r1, r2, None, remove_outliers=False, keep_warped=False, extent=extent)
    base = ((bufs.maskbits & _lib.M_BASE) != 0).cpu().numpy()
    return np.ma.array(bufs.dh.cpu().numpy(), mask=base), tuple(grid.gt)


def getparser():
    parser = argparse.ArgumentParser(description="Compute difference between two rasters")
    parser.add_argument('fn1', type=str, help='Raster filename 1')
    parser.add_argument('fn2', type=str, help='Raster filename 2')
    parser.add_argument('-tr', default='max', help='Output resolution (default: %(default)s)')
    parser.add_argument('-te', default='intersection', choices=['intersection', 'union', 'first', 'last'],
                        help='Output extent (default: %(default)s)')
    parser.add_argument('-outdir', default=None, help='Output directory')
    return parser


def main(argv=None):
    args = getparser().parse_args(argv)
    r1, r2 = rio.read_raster(args.fn1), rio.read_raster(args.fn2)
    print("Raster1 - Raster2: computing %s - %s" % (os.path.basename(args.fn2), os.path.basename(args.fn1)))
    diff, gt = compute_diff(r1, r2, extent=args.te)
    if diff.count() == 0:                                   # compute_diff.py:107-109
        import sys
        sys.exit("No valid overlap between input rasters")
    print("Raster difference stats:")                       # compute_diff.py:123-125 (malib.print_stats)
    from . import robust_stats
    st = robust_stats.get_stats_dict(diff, full=False)
    print("count: %i min: %0.2f max: %0.2f mean: %0.2f std: %0.2f med: %0.2f mad: %0.2f" % (
        st['count'], st['min'], st['max'], st['mean'], st['std'], st['med'], st['nmad']))
    outdir = args.outdir or os.path.split(os.path.abspath(args.fn1))[0]
    dst = os.path.join(outdir, '%s_%s_diff.tif' % (os.path.splitext(os.path.basename(args.fn1))[0],
                                                    os.path.splitext(os.path.basename(args.fn2))[0]))
    rio.write_raster(dst, diff.filled(-9999.0), gt, -9999.0)
    print("Writing out: %s" % dst)
    return dst


if __name__ == "__main__":
    main()
