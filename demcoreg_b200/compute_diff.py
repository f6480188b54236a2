#! /usr/bin/env python
"""Drop-in for the core of reference ``demcoreg/compute_diff.py:97-133``: warp two rasters to their common grid
(``memwarp_multi``, equal resolution, extent 'intersection') and write ``r2 - r1``.

Same primitive as the iteration's front end (``dcr_warp_diff_horn``); the ``-tr/-te`` resampling options of the
reference need the down-sampling GDAL warp and are out of scope.
"""
import argparse
import os

import numpy as np

from . import _lib, engine
from . import rasterio_lite as rio


def compute_diff(r1, r2):
    """``r2 - r1`` on the common grid -> (masked array, geotransform)."""
    grid, bufs = engine.front_only(r1, r2, None, remove_outliers=False, keep_warped=False)
    base = ((bufs.maskbits & _lib.M_BASE) != 0).cpu().numpy()
    return np.ma.array(bufs.dh.cpu().numpy(), mask=base), tuple(grid.gt)


def getparser():
    parser = argparse.ArgumentParser(description="Compute difference between two rasters")
    parser.add_argument('fn1', type=str, help='Raster filename 1')
    parser.add_argument('fn2', type=str, help='Raster filename 2')
    parser.add_argument('-outdir', default=None, help='Output directory')
    return parser


def main(argv=None):
    args = getparser().parse_args(argv)
    r1, r2 = rio.read_raster(args.fn1), rio.read_raster(args.fn2)
    print("Raster1 - Raster2: computing %s - %s" % (os.path.basename(args.fn2), os.path.basename(args.fn1)))
    diff, gt = compute_diff(r1, r2)
    outdir = args.outdir or os.path.split(os.path.abspath(args.fn1))[0]
    dst = os.path.join(outdir, '%s_%s_diff.tif' % (os.path.splitext(os.path.basename(args.fn1))[0],
                                                    os.path.splitext(os.path.basename(args.fn2))[0]))
    rio.write_raster(dst, diff.filled(-9999.0), gt, -9999.0)
    print("Writing out: %s" % dst)
    return dst


if __name__ == "__main__":
    main()
