"""``-mode ncc`` / ``-mode sad`` branch of ``dem_align.compute_offset`` (reference ``dem_align.py:126-143``)."""
import sys

import numpy as np

from . import _lib, coreglib, engine


def compute_offset_corr(ref_dem_ds, src_dem_ds, mode, remove_outliers, max_offset, max_dz, slope_lim, mask_source, plot,
                        ncc_noise=None, _details=None):
    """Front end identical to the Nuth path (warp, dh, outlier / slope masks, dz = median), then the NCC or SAD
    search on the warped rasters re-masked with the final ``static_mask`` -> ``(dx, dy, dz, static_mask, fig)``."""
    # (masks, counts and dz only: the aspect-bin statistics and the fit of the Nuth branch are skipped)
    res, grid, bufs = engine.nuth_iteration(ref_dem_ds, src_dem_ds, mask_source, remove_outliers, max_dz, slope_lim,
                                            keep_warped=True, no_bins=True)
    if res.status == 1:
        sys.exit("No overlapping, unmasked pixels shared between input DEMs")
    r = float(grid.gt[1])
    max_offset_px = (max_offset / r) + 1
    pad = (int(max_offset_px), int(max_offset_px))
    static = (bufs.maskbits & _lib.M_DIFF) != 0
    dz = np.float32(res.dz_med)
    fig = None
    if mode == 'sad':
        m, int_offset, sp_offset = coreglib.compute_offset_sad((bufs.ref_w, static), (bufs.src_w, static), pad=pad)
    else:
        m, int_offset, sp_offset, fig = coreglib.compute_offset_ncc((bufs.ref_w, static), (bufs.src_w, static), pad=pad,
                                                                    prefilter=False, plot=plot, noise=ncc_noise)
    # geotransform has a negative y resolution (dem_align.py:131-135)
    dx = sp_offset[1] * grid.gt[1]
    dy = sp_offset[0] * grid.gt[5]
    if _details is not None:
        _details.update(m=m, int_offset=int_offset, sp_offset=sp_offset, result=res, grid=grid)
    return -dx, -dy, -dz, static.cpu().numpy(), fig
