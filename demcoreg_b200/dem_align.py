#! /usr/bin/env python
"""Drop-in for the iteration path of reference ``demcoreg/dem_align.py`` (``:26-172``, ``:203-392``).

``compute_offset`` runs ONE chained sequence of sm_100a kernels per call
(``dcr_nuth_iteration``); the command line (``getparser``/``main``) keeps the reference's
options and defaults.  Datasets are ``demcoreg_b200.raster.Raster`` objects (device memory).
"""
import argparse
import json
import os
import sys

import numpy as np

from . import _lib, coreglib, engine
from .raster import Raster, MaskSource

mask_choices = ['toa', 'snodas', 'modscag', 'bareground', 'glaciers', 'nlcd', 'none']


def get_mask(ds, mask_list, dem_fn=None, mask_source=None):
    """Reference ``dem_align.py:26-30``.  ``dem_mask.get_mask`` needs external land-cover data and is
    out of scope: a ``MaskSource`` (world-fixed uint8 raster) takes its place; no mask -> scalar ``False``."""
    if mask_source is None:
        if mask_list and mask_list != ['none']:
            raise NotImplementedError("dem_mask layers %s need external datasets (RGI/NLCD/...); pass a MaskSource" % mask_list)
        return False
    return mask_source


def outlier_filter(diff, f=3, perc=None, max_dz=100):
    """Reference ``dem_align.py:32-49`` on a masked array or ``(tensor, mask)`` pair -> same kind of object.

    Absolute ``max_dz`` filter, then ``perc`` percentile filter or (default) median +- ``f`` * NMAD.
    """
    torch = _lib.require_cuda()
    is_ma = not (isinstance(diff, (tuple, list)) and hasattr(diff[0], "is_cuda"))
    x, m = coreglib._to_device_masked(diff)
    print("Removing outliers")
    print("Initial pixel count:")
    m = (m != 0)
    print(int((~m).sum()))
    print("Absolute dz filter: %0.2f" % max_dz)
    lim = np.float32(max_dz)
    m = m | (x < -float(lim)) | (x > float(lim))
    print(int((~m).sum()))
    bits = m.to(torch.uint8).contiguous()
    if perc is not None:
        (lo, hi), _ = engine.select_quantiles(x, [perc[0], perc[1]], bits, 1)
    else:
        (med,), _ = engine.select_quantiles(x, [50.0], bits, 1)
        (madv,), _ = engine.select_quantiles(x, [50.0], bits, 1, transform=1, center=float(med))
        nmad = np.float32(madv) * np.float32(1.4826)
        lo = np.float32(med) - np.float32(f) * nmad
        hi = np.float32(med) + np.float32(f) * nmad
    m = m | (x < float(lo)) | (x > float(hi))
    print(int((~m).sum()))
    if is_ma:
        return np.ma.array(x.cpu().numpy(), mask=m.cpu().numpy())
    return x, m


def get_filtered_slope(ds, slope_lim=(0.1, 40)):
    """Reference ``dem_align.py:51-60``: Horn slope (degrees) of ``ds`` masked outside ``slope_lim``."""
    print("Computing slope")
    s, _ = engine.horn(ds, True, False)
    print("Slope filter: %0.2f - %0.2f" % slope_lim)
    a = s.cpu().numpy()
    slope = np.ma.masked_values(a, -9999.0, shrink=False)
    print("Initial count: %i" % slope.count())
    slope = np.ma.masked_outside(slope, *slope_lim)
    print(slope.count())
    return slope


_EXIT_MSG = {
    1: "No overlapping, unmasked pixels shared between input DEMs",
    2: "Not enough dh samples",
    3: "Not enough slope/aspect samples",
}


def compute_offset(ref_dem_ds, src_dem_ds, src_dem_fn, mode='nuth', remove_outliers=True, max_offset=100,
                   max_dz=100, slope_lim=(0.1, 40), mask_list=['glaciers', ], plot=True, mask_source=None,
                   return_static_mask=True, _details=None, _apply_z=False):
    """Reference ``dem_align.py:62-172`` -> ``(dx, dy, dz, static_mask, fig)``.

    ``mask_source`` (extension): a ``MaskSource`` standing in for ``dem_mask.get_mask``; with it
    ``mask_list`` is ignored.  ``return_static_mask=False`` (extension) skips the device->host copy of
    the final mask, which the Nuth loop only uses for the closing plot.
    """
    if mask_source is None and mask_list and mask_list != ['none']:
        get_mask(None, mask_list)
    fig = None
    if mode == 'nuth':
        # (the slope raster stays on the device and is not returned: tan(slope) mode, like the native loop)
        res, grid, bufs = engine.nuth_iteration(ref_dem_ds, src_dem_ds, mask_source, remove_outliers, max_dz, slope_lim,
                                                apply_z=_apply_z, tan_slope=True)
        if res.status in _EXIT_MSG:
            sys.exit(_EXIT_MSG[res.status])
        if _details is not None:
            _details['result'] = res
            _details['grid'] = grid
            _details['bufs'] = bufs
        dz = np.float32(res.dz_med)
        if res.status == 4:
            print("Failed to calculate horizontal shift")
            dx, dy = 0, 0
        else:
            fit_param = res.fit
            dx = fit_param[0] * np.sin(np.deg2rad(fit_param[1]))
            dy = fit_param[0] * np.cos(np.deg2rad(fit_param[1]))
            nuth_dz = fit_param[2] * np.tan(np.deg2rad(res.med_slope))
            print('Median dz: %0.2f\nNuth dz: %0.2f' % (dz, nuth_dz))
        static_mask = None
        if return_static_mask:
            static_mask = ((bufs.maskbits & _lib.M_DIFF) != 0).cpu().numpy()
        return -dx, -dy, -dz, static_mask, fig
    elif mode in ('ncc', 'sad'):
        from . import corr
        return corr.compute_offset_corr(ref_dem_ds, src_dem_ds, mode, remove_outliers, max_offset, max_dz, slope_lim,
                                        mask_source, plot)
    elif mode == 'all':
        print("Not yet implemented")
        return 0, 0, 0, None, None
    elif mode == 'none':
        raise NameError("reference dem_align.py:166-170 uses undefined names (outprefix, src_dem_orig) in mode 'none'")
    raise ValueError(mode)


def _looks_geographic(ds):
    """True when the dataset is evidently in geographic coordinates (degrees): GEOGCS / longlat projection string, or
    pixels smaller than 0.05 units on an origin inside +-360 / +-90."""
    pr = (ds.proj or '').strip().upper()
    if pr.startswith('GEOGCS') or pr.startswith('GEOGCRS') or '+PROJ=LONGLAT' in pr:
        return True
    if pr.startswith('PROJCS') or pr.startswith('PROJCRS') or pr.startswith('LOCAL_CS'):
        return False
    gt = ds.gt
    return abs(gt[1]) < 0.05 and abs(gt[0]) <= 360.0 and abs(gt[3]) <= 90.0


def getparser():
    """Reference ``dem_align.py:174-201``: same options, names and defaults."""
    parser = argparse.ArgumentParser(description="Perform DEM co-registration using multiple algorithms",
                                     formatter_class=argparse.ArgumentDefaultsHelpFormatter)
    parser.add_argument('ref_fn', type=str, help='Reference DEM filename')
    parser.add_argument('src_fn', type=str, help='Source DEM filename to be shifted')
    parser.add_argument('-mode', type=str, default='nuth', choices=['ncc', 'sad', 'nuth', 'none'],
                        help='Type of co-registration to use')
    parser.add_argument('-mask_list', nargs='+', type=str, default=[], choices=mask_choices,
                        help='Define masks to use to limit reference surfaces for co-registration')
    parser.add_argument('-tiltcorr', action='store_true',
                        help='After preliminary translation, fit polynomial to residual elevation offsets and remove')
    parser.add_argument('-polyorder', type=int, default=1, help='Specify order of polynomial fit')
    parser.add_argument('-tol', type=float, default=0.02,
                        help='When iterative translation magnitude is below this tolerance (meters), break and write out corrected DEM')
    parser.add_argument('-max_offset', type=float, default=100,
                        help='Maximum expected horizontal offset in meters, used to set search range for ncc and sad modes')
    parser.add_argument('-max_dz', type=float, default=100,
                        help='Maximum expected vertical offset in meters, used to filter outliers')
    res_choices = ['min', 'max', 'mean', 'common_scale_factor']
    parser.add_argument('-res', type=str, default='mean', choices=res_choices, help='Warp intputs to this resolution')
    parser.add_argument('-slope_lim', type=float, nargs=2, default=(0.1, 40),
                        help='Minimum and maximum surface slope limits to consider')
    parser.add_argument('-max_iter', type=int, default=30, help='Maximum number of iterations, if tol is not reached')
    parser.add_argument('-outdir', default=None, help='Output directory')
    # extension: a world-fixed uint8 stable-surface mask raster (1 = masked) replacing dem_mask layers
    parser.add_argument('-mask_fn', default=None, help='(extension) uint8 mask raster, 1 = exclude from co-registration')
    return parser


def align(ref_dem_ds, src_dem_ds, mode='nuth', mask_source=None, tol=0.02, max_offset=100, max_dz=100,
          slope_lim=(0.1, 40), max_iter=30, verbose=True, records=None, tiltcorr=False, polyorder=1, res='max'):
    """The iteration loop of reference ``dem_align.py:291-357`` on device-resident rasters.  ``res``: the resolution of the
    one-time ``memwarp_multi`` before the loop (``-res``; only matters for inputs of different pixel size).

    ``src_dem_ds`` is copied; the copy is shifted in place each iteration (geotransform + band) exactly
    as the reference does.  Returns ``dict(dx, dy, dz, n_iter, iters, src_align)``.
    """
    import contextlib
    import io
    src_dem_ds_align = src_dem_ds.copy()
    # dem_align.py:294-299: both rasters are clipped / resampled ONCE to their intersection before the loop (rasters that
    # already share extent and lattice pass through untouched); the loop then works on these
    ref_dem_ds, src_dem_ds_align = engine.memwarp_multi([ref_dem_ds, src_dem_ds_align], res=res)
    if mode == 'nuth' and records is None:
        return _align_nuth_native(ref_dem_ds, src_dem_ds_align, mask_source, tol, max_offset, max_dz, slope_lim, max_iter,
                                  verbose, tiltcorr, polyorder)
    if tiltcorr:
        raise NotImplementedError("-tiltcorr is wired to the native nuth loop (mode='nuth', records=None)")
    n = 1
    dx_total = 0
    dy_total = 0
    dz_total = 0
    iters = []
    sink = contextlib.nullcontext() if verbose else contextlib.redirect_stdout(io.StringIO())
    with sink:
        while True:
            print("*** Iteration %i ***" % n)
            det = {}
            # Nuth mode: the z shift of this iteration is applied on the device inside the same kernel chain
            dx, dy, dz, static_mask, fig = compute_offset(ref_dem_ds, src_dem_ds_align, None, mode, max_offset=max_offset,
                                                          mask_list=[], max_dz=max_dz, slope_lim=slope_lim, plot=True,
                                                          mask_source=mask_source, return_static_mask=False, _details=det,
                                                          _apply_z=(mode == 'nuth'))
            xyz_shift_str_iter = "dx=%+0.2fm, dy=%+0.2fm, dz=%+0.2fm" % (dx, dy, dz)
            print("Incremental offset: %s" % xyz_shift_str_iter)
            dx_total += dx
            dy_total += dy
            dz_total += dz
            xyz_shift_str_cum = "dx=%+0.2fm, dy=%+0.2fm, dz=%+0.2fm" % (dx_total, dy_total, dz_total)
            print("Cumulative offset: %s" % xyz_shift_str_cum)
            src_dem_ds_align = coreglib.apply_xy_shift(src_dem_ds_align, dx, dy, createcopy=False)
            if 'result' in det and det['result'].z_applied:
                print("Z shift offset: ", dz)
            else:
                src_dem_ds_align = coreglib.apply_z_shift(src_dem_ds_align, dz, createcopy=False)
            iters.append(dict(n=n, dx=float(dx), dy=float(dy), dz=float(dz)))
            if records is not None:
                records.append(det)
            n += 1
            print("\n")
            dm = np.sqrt(dx ** 2 + dy ** 2 + dz ** 2)
            dm_total = np.sqrt(dx_total ** 2 + dy_total ** 2 + dz_total ** 2)
            if dm_total > max_offset:
                sys.exit("Total offset exceeded specified max_offset (%0.2f m). Consider increasing -max_offset argument" % max_offset)
            if n > max_iter or dm < tol:
                break
    return dict(dx=float(dx_total), dy=float(dy_total), dz=float(dz_total), n_iter=n - 1, iters=iters,
                src_align=src_dem_ds_align, ref=ref_dem_ds)


def _align_nuth_native(ref_dem_ds, src_dem_ds_align, mask_source, tol, max_offset, max_dz, slope_lim, max_iter, verbose,
                       tiltcorr=False, polyorder=1):
    """The same loop through ``dcr_align_nuth``: every iteration is chained natively (no interpreter work between the
    kernels of consecutive iterations); the reference's per-iteration messages are replayed afterwards.

    ``tiltcorr`` (reference ``dem_align.py:396-447``): when the loop first stops, ONE polynomial of order ``polyorder`` is
    fitted to the filtered final difference and removed from the source; the loop then goes on for up to ``max_iter``
    more iterations."""
    al = engine.NuthAligner(ref_dem_ds, src_dem_ds_align, mask_source, tol, max_offset, max_dz, slope_lim, max_iter)
    al.run()
    tilt = None
    if tiltcorr and al.st.exit_code == 0:
        from . import tiltcorr as tc, robust_stats
        al.finish()
        if verbose:
            print("\n************")
            print("Calculating 'tiltcorr' 2D polynomial fit to residuals with order %i" % polyorder)
            print("************\n")
        # final difference on the common grid + its filtered version (dem_align.py:366-392): the masks of one more
        # front/filter pass with the DEFAULT slope limits, exactly what diff_stats() reports
        res, grid, bufs = engine.nuth_iteration(ref_dem_ds, src_dem_ds_align, mask_source, True, max_dz, (0.1, 40))
        if res.status in _EXIT_MSG:
            sys.exit(_EXIT_MSG[res.status])
        gt_clip = tuple(grid.gt)
        model, cnt = tc.polyfit2d(bufs.dh, bufs.maskbits, _lib.M_DIFF, gt_clip, polyorder)
        vals = tc.polyval2d(model, gt_clip, grid.height, grid.width)
        tilt = dict(coeff=tc.coeff_list(model), model=model, count=cnt, gt=gt_clip, vals=vals,
                    val_stats=robust_stats.get_stats_dict(vals, None, full=True))
        if verbose:
            print(np.array(tilt['coeff']))
        tc.apply_polyval(src_dem_ds_align, model, -1.0)          # apply_z_shift(src_dem_ds_align, -valgrid)
        # "Now use original tolerance, and number of iterations": max_iter = n + args.max_iter with n the NEXT iteration
        al.st.finished = 0
        al.st.max_iter = al.st.n_done + 1 + int(max_iter)
        al.run()
    iters = []
    dx_total = dy_total = 0.0
    dz_total = np.float32(0)
    for it in al.iters:
        if it['status'] in _EXIT_MSG:
            sys.exit(_EXIT_MSG[it['status']])
        dz = np.float32(it['dz'])
        dx_total += it['dx']
        dy_total += it['dy']
        dz_total = dz_total + dz
        if verbose:
            print("*** Iteration %i ***" % it['n'])
            if it['status'] == 4:
                print("Failed to calculate horizontal shift")
            else:   # dem_align.py:157-159
                nuth_dz = it['fit'][2] * np.tan(np.deg2rad(it['med_slope']))
                print('Median dz: %0.2f\nNuth dz: %0.2f' % (np.float32(it['dz_med']), nuth_dz))
            print("Incremental offset: dx=%+0.2fm, dy=%+0.2fm, dz=%+0.2fm" % (it['dx'], it['dy'], dz))
            print("Cumulative offset: dx=%+0.2fm, dy=%+0.2fm, dz=%+0.2fm" % (dx_total, dy_total, dz_total))
            print("\n")
        iters.append(dict(n=it['n'], dx=float(it['dx']), dy=float(it['dy']), dz=float(dz)))
    if al.st.exit_code == 5:
        sys.exit("Total offset exceeded specified max_offset (%0.2f m). Consider increasing -max_offset argument" % max_offset)
    al.finish()
    dxt, dyt, dzt = al.totals
    out = dict(dx=dxt, dy=dyt, dz=dzt, n_iter=len(iters), iters=iters, src_align=src_dem_ds_align, ref=ref_dem_ds)
    if tilt is not None:
        out['tiltcorr'] = tilt
    return out


def diff_stats(ref_dem_ds, src_dem_ds, mask_source=None, max_dz=100):
    """Final-difference products of reference ``dem_align.py:366-392`` (and ``:506-530`` for the original pair).

    Warp both rasters to their common grid, ``diff = src - ref``, then ``malib.get_stats_dict(full=True)`` of the
    difference over valid, un-masked pixels and of its filtered version (``outlier_filter`` + the slope mask with the
    DEFAULT ``slope_lim`` -- the reference ignores the user's limits here, ``dem_align.py:390, 528``).
    Returns ``dict(stats, stats_filt, diff, diff_filt, gt)`` with the rasters as masked arrays.
    """
    from . import robust_stats
    if abs(ref_dem_ds.gt[1] - src_dem_ds.gt[1]) >= 1e-3:
        # (dem_align.py:368-369, 506-507: memwarp_multi(res='max') -- rasters of different pixel size are resampled first)
        ref_dem_ds, src_dem_ds = engine.memwarp_multi([ref_dem_ds, src_dem_ds], res='max')
    res, grid, bufs = engine.nuth_iteration(ref_dem_ds, src_dem_ds, mask_source, True, max_dz, (0.1, 40))
    base = (bufs.maskbits & _lib.M_BASE) != 0
    filt = (bufs.maskbits & _lib.M_DIFF) != 0
    out = dict(stats=robust_stats.get_stats_dict(bufs.dh, base, full=True),
               stats_filt=robust_stats.get_stats_dict(bufs.dh, filt, full=True),
               gt=tuple(grid.gt))
    dh = bufs.dh.cpu().numpy()
    out['diff'] = np.ma.array(dh, mask=base.cpu().numpy())
    out['diff_filt'] = np.ma.array(dh, mask=filt.cpu().numpy())
    return out


def main(argv=None):
    """Reference ``dem_align.py:203-631`` (iteration loop + final shift products; plots are not produced)."""
    from . import rasterio_lite as rio
    parser = getparser()
    args = parser.parse_args(argv)
    ref_dem_fn = args.ref_fn
    src_dem_fn = args.src_fn
    mode = args.mode
    slope_lim = tuple(args.slope_lim)
    tiltcorr = args.tiltcorr
    polyorder = args.polyorder
    outdir = args.outdir
    if outdir is None:
        outdir = os.path.splitext(src_dem_fn)[0] + '_dem_align'
    if tiltcorr:
        outdir += '_tiltcorr'          # dem_align.py:233-234
    if not os.path.exists(outdir):
        os.makedirs(outdir)
    outprefix = '%s_%s' % (os.path.splitext(os.path.split(src_dem_fn)[-1])[0],
                           os.path.splitext(os.path.split(ref_dem_fn)[-1])[0])
    outprefix = os.path.join(outdir, outprefix)
    print("\nReference: %s" % ref_dem_fn)
    print("Source: %s" % src_dem_fn)
    print("Mode: %s" % mode)
    print("Output: %s\n" % outprefix)

    ref = rio.read_raster(ref_dem_fn)
    src = rio.read_raster(src_dem_fn)
    # dem_align.py:270-289: inputs must be in a projected CRS with linear units of metres.  There is no PROJ here; the
    # projection string and the geotransform decide (degrees-sized pixels inside +-360 / +-90)
    for fn, ds in ((ref_dem_fn, ref), (src_dem_fn, src)):
        if _looks_geographic(ds):
            print("%s CRS: '%s'" % (fn, ds.proj))
            print("Input DEMs must have projected CRS with linear units of meters (not WGS84 geographic, units of degrees)")
            print("For relatively limited DEM extents (<100-300 km), you can consider reprojecting using the appropriate UTM projection")
            print("Then rerun the dem_align.py command with the projected DEM(s)\n")
            sys.exit()
    mask_source = None
    if args.mask_fn:
        m = rio.read_array(args.mask_fn)
        mask_source = MaskSource(m['array'].astype(np.uint8), m['gt'])
    elif args.mask_list:
        get_mask(None, args.mask_list)

    if mode == 'none':
        # dem_align.py:166-170 (its names outprefix / src_dem_orig are undefined inside compute_offset, so the reference itself
        # dies with a NameError here): what it announces -- the source with the median bias over static surfaces removed
        res0, _, _ = engine.nuth_iteration(ref, src, mask_source, True, args.max_dz, slope_lim)
        if res0.status in _EXIT_MSG:
            sys.exit(_EXIT_MSG[res0.status])
        dz = np.float32(res0.dz_med)
        print("Skipping alignment, writing out DEM with median bias over static surfaces removed")
        dst_fn = outprefix + '_med%0.1f.tif' % dz
        sa = coreglib.apply_z_shift(src, -dz, createcopy=True)
        rio.write_raster(dst_fn, sa.numpy(), sa.gt, sa.ndv, sa.proj)
        sys.exit()
    out = align(ref, src, mode=mode, mask_source=mask_source, tol=args.tol, max_offset=args.max_offset,
                max_dz=args.max_dz, slope_lim=slope_lim, max_iter=args.max_iter, tiltcorr=tiltcorr, polyorder=polyorder,
                res=args.res)
    dx_total, dy_total, dz_total = out['dx'], out['dy'], out['dz']
    xyz_shift_str_cum_fn = '_%s_x%+0.2f_y%+0.2f_z%+0.2f' % (mode, dx_total, dy_total, dz_total)
    align_fn = outprefix + xyz_shift_str_cum_fn + '_align.tif'
    print("Writing out shifted src_dem with median vertical offset removed: %s" % align_fn)
    # dem_align.py:466-483: the ORIGINAL dataset with the TOTAL shifts applied once (not the loop's copy, whose band went
    # through one float32 add per iteration), then the tilt surface on its own grid
    sa = coreglib.apply_xy_shift(src.copy(), dx_total, dy_total, createcopy=False)
    sa = coreglib.apply_z_shift(sa, np.float32(dz_total), createcopy=False)
    if tiltcorr and 'tiltcorr' in out:
        from . import tiltcorr as tc
        tc.apply_polyval(sa, out['tiltcorr']['model'], -1.0)
    rio.write_raster(align_fn, sa.numpy(), sa.gt, sa.ndv, sa.proj)
    dm_total = float(np.sqrt(dx_total ** 2 + dy_total ** 2 + dz_total ** 2))
    # final elevation difference (dem_align.py:366-392, from the loop's dataset) and the original one (dem_align.py:506-530)
    after = diff_stats(out['ref'], out['src_align'], mask_source, args.max_dz)
    before = diff_stats(out['ref'], src, mask_source, args.max_dz)
    ndv_out = -9999.0
    rio.write_raster(outprefix + xyz_shift_str_cum_fn + '_align_diff.tif', after['diff'].filled(ndv_out), after['gt'], ndv_out)
    rio.write_raster(outprefix + xyz_shift_str_cum_fn + '_align_diff_filt.tif', after['diff_filt'].filled(ndv_out),
                     after['gt'], ndv_out)
    rio.write_raster(outprefix + '_orig_diff.tif', before['diff'].filled(ndv_out), before['gt'], ndv_out)
    # dem_align.py:488-499: the aligned source masked so that only the pixels of the filtered difference map survive -- the
    # filtered difference (nodata ndv_out) is cubic-warped onto the full source grid (band nodata 0) and lends its mask
    print("Applying filter to shiftec src_dem")
    dfr = Raster(after['diff_filt'].filled(ndv_out).astype(np.float32), after['gt'], ndv_out)
    g_full = _lib.Grid()
    for k in range(6):
        g_full.gt[k] = sa.gt[k]
    g_full.width, g_full.height = sa.RasterXSize, sa.RasterYSize
    geom = _lib.WarpGeom()
    import ctypes as _C
    dd = dfr.desc()
    _lib.check(_lib.lib().dcr_plan_onto_grid(_C.byref(dd), _C.byref(g_full), _C.byref(geom)))
    dfw = engine.warp_cubic(dfr, g_full, geom, 0.0).numpy()
    keep = ~np.ma.getmaskarray(np.ma.masked_values(dfw, 0.0))
    sa_arr = sa.numpy()
    fill = sa.ndv if sa.ndv is not None else ndv_out
    rio.write_raster(outprefix + xyz_shift_str_cum_fn + '_align_filt.tif', np.where(keep, sa_arr, np.float32(fill)), sa.gt,
                     fill, sa.proj)
    res_ref = float((ref.gt[1] + abs(ref.gt[5])) / 2.0)
    res_src = float((src.gt[1] + abs(src.gt[5])) / 2.0)
    g = after['gt']
    h, w = after['diff'].shape
    # json schema of dem_align.py:540-564 (consumed by dem_align_post.py:133-147); lon/lat need PROJ -> None
    info = {'src_fn': src_dem_fn, 'ref_fn': ref_dem_fn, 'align_fn': align_fn,
            'res': {'src': res_src, 'ref': res_ref, 'coreg': float(g[1])},
            'center_coord': {'lon': None, 'lat': None, 'x': g[0] + 0.5 * w * g[1], 'y': g[3] + 0.5 * h * g[5]},
            'shift': {'dx': dx_total, 'dy': dy_total, 'dz': float(dz_total), 'dm': dm_total},
            'before': before['stats'], 'before_filt': before['stats_filt'],
            'after': after['stats'], 'after_filt': after['stats_filt'],
            'n_iter': out['n_iter']}
    if tiltcorr and 'tiltcorr' in out:   # dem_align.py:553-556
        info['tiltcorr'] = {'coeff': out['tiltcorr']['coeff'], 'val_stats': out['tiltcorr']['val_stats']}
    json_fn = outprefix + xyz_shift_str_cum_fn + '_align_stats.json'
    with open(json_fn, 'w') as f:
        json.dump(info, f)
    return info


if __name__ == "__main__":
    main()
