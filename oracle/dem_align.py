"""CPU restatement of reference ``demcoreg/dem_align.py:26-172`` and the loop ``:306-392``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED at the
pygeotools/GDAL boundary.

``dem_mask.get_mask`` (reference ``dem_mask.py:420-567``) needs multi-GB external
land-cover / glacier datasets and is out of scope; the hot path only consumes its
boolean result (``dem_align.py:85-86``).  Here a stable-surface mask is a uint8
raster fixed in world coordinates (``MaskSource``), sampled nearest-neighbour onto
the clipped grid each iteration, which is what ``get_mask`` does with its inputs.
"""
import math
import sys
import time

import numpy as np

from . import pygeotools_restated as pg
from . import coreglib


class MaskSource(object):
    """A world-fixed uint8 raster (1 = masked) standing in for ``dem_mask`` inputs."""

    def __init__(self, array, gt):
        self.array = np.ascontiguousarray(array, dtype=np.uint8)
        self.gt = tuple(float(g) for g in gt)

    def sample(self, ds):
        """Nearest-neighbour sample onto the grid of ``ds``; outside the mask raster = masked."""
        gt = ds.GetGeoTransform()
        res = gt[1]
        H, W = ds.RasterYSize, ds.RasterXSize
        ox = (gt[0] - self.gt[0]) / res
        oy = (self.gt[3] - gt[3]) / res
        nx0 = int(math.floor(ox + 0.5))
        ny0 = int(math.floor(oy + 0.5))
        out = np.ones((H, W), dtype=bool)
        Hm, Wm = self.array.shape
        i0, i1 = max(0, -ny0), min(H, Hm - ny0)
        j0, j1 = max(0, -nx0), min(W, Wm - nx0)
        if i1 > i0 and j1 > j0:
            out[i0:i1, j0:j1] = self.array[i0 + ny0:i1 + ny0, j0 + nx0:j1 + nx0] != 0
        return out


def get_mask(ds, mask_list, dem_fn=None, mask_source=None):
    """Reference ``dem_align.py:26-30`` -> ``dem_mask.get_mask``: scalar ``False`` when there is no mask."""
    if mask_source is None:
        return False
    return mask_source.sample(ds)


def outlier_filter(diff, f=3, perc=None, max_dz=100):
    """Reference ``dem_align.py:32-49``."""
    diff = np.ma.masked_outside(diff, -max_dz, max_dz)
    if perc is not None:
        diff = pg.perc_fltr(diff, perc)
    else:
        diff = pg.mad_fltr(diff, f)
    return diff


def get_filtered_slope(ds, slope_lim=(0.1, 40)):
    """Reference ``dem_align.py:51-60``."""
    slope = pg.gdaldem_mem_ds(ds, processing='slope', returnma=True, computeEdges=False)
    slope = pg.range_fltr(slope, slope_lim)
    return slope


def compute_offset(ref_dem_ds, src_dem_ds, src_dem_fn=None, mode='nuth', remove_outliers=True, max_offset=100,
                   max_dz=100, slope_lim=(0.1, 40), mask_list=[], plot=True, mask_source=None,
                   ncc_noise=None, details=None, time_plot_path=False, with_mode=False):
    """Reference ``dem_align.py:62-172``.  Returns ``(dx, dy, dz, static_mask, fig)``.

    ``details`` (a dict, optional) receives the intermediates used by the parity tests.
    ``with_mode=False`` skips ``scipy.stats.mstats.mode`` inside ``print_stats`` (its result
    is unused by ``dem_align``).
    """
    ref_dem_clip_ds, src_dem_clip_ds = pg.memwarp_multi([ref_dem_ds, src_dem_ds],
                                                        res='max', extent='intersection', t_srs=src_dem_ds, r='cubic')
    res = float(pg.get_res(ref_dem_clip_ds, square=True)[0])
    max_offset_px = (max_offset / res) + 1
    pad = (int(max_offset_px), int(max_offset_px))
    src_dem_gt = np.array(src_dem_clip_ds.GetGeoTransform())

    ref_dem = pg.ds_getma(ref_dem_clip_ds, 1)
    src_dem = pg.ds_getma(src_dem_clip_ds, 1)

    diff = src_dem - ref_dem

    static_mask = get_mask(src_dem_clip_ds, mask_list, src_dem_fn, mask_source=mask_source)
    diff = np.ma.array(diff, mask=static_mask)

    if diff.count() == 0:
        sys.exit("No overlapping, unmasked pixels shared between input DEMs")

    if details is not None:
        details['ref_w'] = ref_dem_clip_ds.array
        details['src_w'] = src_dem_clip_ds.array
        details['gt'] = tuple(src_dem_gt)
        details['count_base'] = diff.count()

    if remove_outliers:
        diff = outlier_filter(diff, f=3, max_dz=max_dz)

    slope = get_filtered_slope(src_dem_clip_ds, slope_lim=slope_lim)
    aspect = pg.gdaldem_mem_ds(src_dem_clip_ds, processing='aspect', returnma=True, computeEdges=False)

    ref_dem_clip_ds = None
    src_dem_clip_ds = None

    diff = np.ma.array(diff, mask=np.ma.getmaskarray(slope))
    static_mask = np.ma.getmaskarray(diff)

    diff_stats = pg.print_stats(diff, with_mode=with_mode)
    dz = diff_stats[5]

    if details is not None:
        details['dh'] = diff
        details['slope'] = slope
        details['aspect'] = aspect
        details['count_final'] = diff.count()
        details['dz_med'] = dz

    fig = None
    dx = 0
    dy = 0

    if mode == "sad":
        ref_dem = np.ma.array(ref_dem, mask=static_mask)
        src_dem = np.ma.array(src_dem, mask=static_mask)
        m, int_offset, sp_offset = coreglib.compute_offset_sad(ref_dem, src_dem, pad=pad)
        dx = sp_offset[1] * src_dem_gt[1]
        dy = sp_offset[0] * src_dem_gt[5]
        if details is not None:
            details['m'] = m
    elif mode == "ncc":
        ref_dem = np.ma.array(ref_dem, mask=static_mask)
        src_dem = np.ma.array(src_dem, mask=static_mask)
        m, int_offset, sp_offset, fig = coreglib.compute_offset_ncc(ref_dem, src_dem,
                                                                    pad=pad, prefilter=False, plot=plot, noise=ncc_noise)
        dx = sp_offset[1] * src_dem_gt[1]
        dy = sp_offset[0] * src_dem_gt[5]
        if details is not None:
            details['m'] = m
    elif mode == "nuth":
        out = coreglib.compute_offset_nuth(diff, slope, aspect, plot=plot, return_details=details is not None,
                                           time_plot_path=time_plot_path)
        fit_param, fig = out[0], out[1]
        if details is not None:
            details['nuth'] = out[2]
            details['fit'] = fit_param
        if fit_param is None:
            print("Failed to calculate horizontal shift")
        else:
            dx = fit_param[0] * np.sin(np.deg2rad(fit_param[1]))
            dy = fit_param[0] * np.cos(np.deg2rad(fit_param[1]))
            med_slope = pg.fast_median(slope)
            nuth_dz = fit_param[2] * np.tan(np.deg2rad(med_slope))
    elif mode == "all":
        print("Not yet implemented")
    elif mode == "none":
        raise NameError("reference dem_align.py:166-170 references undefined names in mode 'none'")
    return -dx, -dy, -dz, static_mask, fig


def _final_diff(ref_dem_ds, src_dem_ds_align, res, mask_source, max_dz):
    """Reference ``dem_align.py:366-392``: final difference on the common grid, its statistics and the filtered version."""
    ref_c, src_c = pg.memwarp_multi([ref_dem_ds, src_dem_ds_align], res=res, extent='intersection', r='cubic')
    ref_a = pg.ds_getma(ref_c, 1)
    src_a = pg.ds_getma(src_c, 1)
    diff_align = src_a - ref_a
    smf = get_mask(src_c, [], None, mask_source=mask_source)
    smf = np.logical_or(np.ma.getmaskarray(diff_align), smf)
    out = dict(gt=src_c.GetGeoTransform())
    out['diff_align_stats'] = pg.get_stats_dict(diff_align[~smf], full=True)
    filt = np.ma.array(diff_align, mask=smf)
    filt = outlier_filter(filt, f=3, max_dz=max_dz)
    slope = get_filtered_slope(src_c)
    filt = np.ma.array(filt, mask=np.ma.getmaskarray(slope))
    out['diff_align_filt_stats'] = pg.get_stats_dict(filt, full=True)
    out['diff_align'] = diff_align
    out['diff_align_filt'] = filt
    return out


def align(ref_dem_ds, src_dem_ds, mode='nuth', mask_source=None, tol=0.02, max_offset=100, max_dz=100,
          slope_lim=(0.1, 40), max_iter=30, res='mean', verbose=False, time_plot_path=False, records=None,
          final_stats=True, tiltcorr=False, polyorder=1):
    """The iteration loop of reference ``dem_align.py:291-392`` (no file I/O, no plots).

    Returns a dict with the cumulative shift, iteration count, per-iteration wall
    times and (``final_stats``) the ``get_stats_dict`` of the final difference
    before/after filtering (``dem_align.py:366-392``).
    """
    src_dem_ds_align = src_dem_ds.copy()
    ref_dem_ds, src_dem_ds_align = pg.memwarp_multi([ref_dem_ds, src_dem_ds_align],
                                                    extent='intersection', res=res, r='cubic')
    res = float(pg.get_res(src_dem_ds_align, square=True)[0])
    n = 1
    dx_total = 0
    dy_total = 0
    dz_total = 0
    iters = []
    tiltcorr_done = False
    tilt = None
    fin = None
    max_iter0 = max_iter
    while True:
        t0 = time.perf_counter()
        det = {} if records is not None else None
        dx, dy, dz, static_mask, fig = compute_offset(ref_dem_ds, src_dem_ds_align, None, mode, max_offset=max_offset,
                                                      mask_list=[], max_dz=max_dz, slope_lim=slope_lim, plot=True,
                                                      mask_source=mask_source, details=det,
                                                      time_plot_path=time_plot_path)
        dx_total += dx
        dy_total += dy
        dz_total += dz
        src_dem_ds_align = coreglib.apply_xy_shift(src_dem_ds_align, dx, dy, createcopy=False)
        src_dem_ds_align = coreglib.apply_z_shift(src_dem_ds_align, dz, createcopy=False)
        t1 = time.perf_counter()
        iters.append(dict(n=n, dx=float(dx), dy=float(dy), dz=float(dz), seconds=t1 - t0,
                          shape=static_mask.shape))
        if records is not None:
            records.append(det)
        if verbose:
            print("iter %i: dx=%+0.4f dy=%+0.4f dz=%+0.4f (%.2fs)" % (n, dx, dy, dz, t1 - t0))
        n += 1
        dm = np.sqrt(dx ** 2 + dy ** 2 + dz ** 2)
        dm_total = np.sqrt(dx_total ** 2 + dy_total ** 2 + dz_total ** 2)
        if dm_total > max_offset:
            sys.exit("Total offset exceeded specified max_offset (%0.2f m). Consider increasing -max_offset argument" % max_offset)
        if n > max_iter or dm < tol:
            # final difference (dem_align.py:366-392); with -tiltcorr: ONE polynomial fit of the filtered residuals,
            # removed from the source, then the loop goes on with max_iter more iterations (dem_align.py:396-447)
            fin = _final_diff(ref_dem_ds, src_dem_ds_align, res, mask_source, max_dz) if (final_stats or tiltcorr) else None
            if tiltcorr and not tiltcorr_done:
                gt = fin['gt']
                vals, resid, coeff = pg.ma_fitpoly(fin['diff_align_filt'], order=polyorder, gt=gt, perc=(0, 100), origmask=False)
                tilt = dict(coeff=coeff, vals=vals, val_stats=pg.get_stats_dict(vals))
                xgrid, ygrid = pg.get_xy_grids(src_dem_ds_align)
                valgrid = coeff(xgrid, ygrid)   # geolib.polyval2d(xgrid, ygrid, coeff) on the model (see pg.ma_fitpoly)
                src_dem_ds_align = coreglib.apply_z_shift(src_dem_ds_align, -valgrid, createcopy=False)
                tiltcorr_done = True
                max_iter = n + max_iter0
            else:
                break

    out = dict(dx=float(dx_total), dy=float(dy_total), dz=float(dz_total), n_iter=n - 1, iters=iters,
               src_align=src_dem_ds_align, ref=ref_dem_ds, res=res)
    if tilt is not None:
        out['tiltcorr'] = tilt
    if fin is not None:
        out.update({k: fin[k] for k in ('diff_align_stats', 'diff_align_filt_stats', 'diff_align', 'diff_align_filt')})
    return out
