"""CPU restatement of reference ``demcoreg/coreglib.py:14-485``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  PARITY UNPINNED at the
pygeotools/GDAL boundary (``oracle/pygeotools_restated.py``); numpy / scipy calls
are the real ones.  matplotlib is absent: every ``fig`` is ``None`` (the reference
already returns ``None`` figures, ``coreglib.py:196, 287-288``).
"""
import sys

import numpy as np

from . import pygeotools_restated as pg


def apply_xy_shift(ds, dx, dy, createcopy=True):
    """Reference ``coreglib.py:14-40``: ``gt[0] += dx; gt[3] += dy`` on a copy or in place."""
    gt_orig = ds.GetGeoTransform()
    gt_shift = np.copy(gt_orig)
    gt_shift[0] += dx
    gt_shift[3] += dy
    if createcopy:
        ds_align = ds.copy()
    else:
        ds_align = ds
    ds_align.SetGeoTransform(gt_shift)
    return ds_align


def apply_z_shift(ds, dz, createcopy=True):
    """Reference ``coreglib.py:42-55``: ``a = b_getma(band); a = a + dz; WriteArray(a.filled())``."""
    if createcopy:
        ds_shift = ds.copy()
    else:
        ds_shift = ds
    a = pg.b_getma(ds_shift)
    a = a + dz
    # band.WriteArray casts to the band type (float32)
    ds_shift.array = np.ascontiguousarray(a.filled(), dtype=np.float32)
    return ds_shift


def nuth_func(x, a, b, c):
    """Reference ``coreglib.py:58-63``."""
    y = a * np.cos(np.deg2rad(b - x)) + c
    return y


def find_first_peak(corr):
    """Reference ``coreglib.py:387-416``."""
    ind = corr.argmax()
    s = corr.shape[1]
    i = ind // s
    j = ind % s
    return i, j, corr.max()


def find_subpixel_peak_position(corr, subpixel_method='gaussian'):
    """Reference ``coreglib.py:419-485`` (IndexError -> map centre; index -1 wraps silently)."""
    default_peak_position = (corr.shape[0] / 2, corr.shape[1] / 2)
    peak1_i, peak1_j, dummy = find_first_peak(corr)
    try:
        c = corr[peak1_i, peak1_j]
        cl = corr[peak1_i - 1, peak1_j]
        cr = corr[peak1_i + 1, peak1_j]
        cd = corr[peak1_i, peak1_j - 1]
        cu = corr[peak1_i, peak1_j + 1]
        if np.any(np.array([c, cl, cr, cd, cu]) < 0) and subpixel_method == 'gaussian':
            subpixel_method = 'centroid'
        try:
            if subpixel_method == 'centroid':
                subp_peak_position = (((peak1_i - 1) * cl + peak1_i * c + (peak1_i + 1) * cr) / (cl + c + cr),
                                      ((peak1_j - 1) * cd + peak1_j * c + (peak1_j + 1) * cu) / (cd + c + cu))
            elif subpixel_method == 'gaussian':
                subp_peak_position = (peak1_i + ((np.log(cl) - np.log(cr)) / (2 * np.log(cl) - 4 * np.log(c) + 2 * np.log(cr))),
                                      peak1_j + ((np.log(cd) - np.log(cu)) / (2 * np.log(cd) - 4 * np.log(c) + 2 * np.log(cu))))
            elif subpixel_method == 'parabolic':
                subp_peak_position = (peak1_i + (cl - cr) / (2 * cl - 4 * c + 2 * cr),
                                      peak1_j + (cd - cu) / (2 * cd - 4 * c + 2 * cu))
        except Exception:
            subp_peak_position = default_peak_position
    except IndexError:
        subp_peak_position = default_peak_position
    return subp_peak_position[0], subp_peak_position[1]


def compute_offset_sad(dem1, dem2, pad=(9, 9), plot=False):
    """Reference ``coreglib.py:65-115`` (the per-offset prints are not reproduced)."""
    kernel = dem2[pad[0]:-pad[0], pad[1]:-pad[1]]
    m = np.zeros((pad[0] * 2 + 1, pad[1] * 2 + 1))
    for i in range(m.shape[0]):
        for j in range(m.shape[1]):
            ref = dem1[i:i + kernel.shape[0], j:j + kernel.shape[1]]
            diff = ref - kernel
            diff_iqr = pg.calcperc(diff, (25, 75))
            diff = np.ma.masked_outside(diff, *diff_iqr)
            m[i, j] = np.ma.abs(diff).sum() / diff.count()
    m = -m
    int_argmax = np.array(np.unravel_index(m.argmax(), m.shape))
    int_offset = int_argmax - pad
    sp_argmax = np.array(find_subpixel_peak_position(m, 'parabolic'))
    sp_offset = sp_argmax - pad
    return m, int_offset, sp_offset


def compute_offset_ncc(dem1, dem2, pad=(9, 9), prefilter=False, plot=False, noise=None):
    """Reference ``coreglib.py:118-198``.

    ``noise``: optional ``(ref_noise, kernel_noise)`` standard-normal arrays to
    replace the reference's unseeded ``np.random.randn`` (``coreglib.py:158-159``), so
    that two implementations can be compared.  ``None`` = reference behaviour.
    """
    if prefilter:
        raise NotImplementedError("prefilter=True needs malib.nanfill; dem_align always passes False")
    import scipy.signal
    stride = 1
    ref = dem1[::stride, ::stride]
    kernel = dem2[pad[0]:-pad[1]:stride, pad[0]:-pad[1]:stride]
    ref = (ref - ref.mean()) / ref.std()
    kernel = (kernel - kernel.mean()) / kernel.std()
    if noise is None:
        rn = np.random.randn(*ref.shape)
        kn = np.random.randn(*kernel.shape)
    else:
        rn, kn = noise
    ref_noise = np.ma.getmaskarray(ref) * rn
    kernel_noise = np.ma.getmaskarray(kernel) * kn
    ref = np.ma.asarray(ref).filled(0) + ref_noise
    kernel = np.ma.asarray(kernel).filled(0) + kernel_noise
    m = scipy.signal.correlate2d(ref, kernel, 'valid')
    int_argmax = np.array(np.unravel_index(m.argmax(), m.shape))
    int_offset = int_argmax * stride - pad
    sp_argmax = np.array(find_subpixel_peak_position(m, 'parabolic'))
    sp_offset = sp_argmax - pad
    fig = None
    return m, int_offset, sp_offset, fig


def nuth_bin_inputs(dh, slope, aspect, remove_outliers=True):
    """Reference ``coreglib.py:221-245``: common mask, x = aspect, y = dh/tan(slope), 3-sigma trim."""
    common_mask = ~(pg.common_mask([dh, aspect, slope]))
    xdata = aspect[common_mask].data
    ydata = (dh[common_mask] / np.tan(np.deg2rad(slope[common_mask]))).data
    n0 = ydata.size
    rmin = rmax = None
    if remove_outliers:
        f = 3
        sigma, u = (ydata.std(), ydata.mean())
        rmin = u - f * sigma
        rmax = u + f * sigma
        idx = (ydata >= rmin) & (ydata <= rmax)
        xdata = xdata[idx]
        ydata = ydata[idx]
    return xdata, ydata, dict(n_common=n0, rmin=rmin, rmax=rmax)


def compute_offset_nuth(dh, slope, aspect, min_count=100, remove_outliers=True, plot=True,
                        return_details=False, time_plot_path=False):
    """Reference ``coreglib.py:201-373``.

    ``return_details`` additionally returns the intermediate bin statistics (for
    parity tests); ``time_plot_path`` executes the numpy part of the always-on plot
    path (``coreglib.py:317-322``: ``np.digitize`` + 360 boolean selections) so its
    cost can be reported -- matplotlib itself is absent.
    """
    import scipy.optimize as optimization

    if dh.count() < min_count:
        sys.exit("Not enough dh samples")
    if slope.count() < min_count:
        sys.exit("Not enough slope/aspect samples")

    med_dh = pg.fast_median(dh)
    med_slope = pg.fast_median(slope)
    c_seed = (med_dh / np.tan(np.deg2rad(med_slope)))
    x0 = np.array([0.0, 0.0, c_seed])

    xdata, ydata, info = nuth_bin_inputs(dh, slope, aspect, remove_outliers)

    nbins = 360
    bin_range = (0., 360.)
    bin_count, bin_edges, bin_centers = pg.bin_stats(xdata, ydata, stat='count', nbins=nbins, bin_range=bin_range)
    bin_med, bin_edges, bin_centers = pg.bin_stats(xdata, ydata, stat='median', nbins=nbins, bin_range=bin_range)
    full_count = bin_count.filled(0).copy()
    full_med = np.ma.asarray(bin_med).filled(np.nan).copy()

    min_bin_sample_count = 9
    idx = (bin_count.filled(0) >= min_bin_sample_count)
    bin_count = bin_count[idx].data
    bin_med = bin_med[idx].data
    bin_centers = bin_centers[idx]

    fit = None
    fit_fig = None
    min_bin_count = 90
    bin_ptp = np.ptp(np.cos(np.radians(bin_centers))) if bin_centers.size else 0.0
    min_bin_ptp = 1.0

    if len(bin_med) >= min_bin_count and bin_ptp >= min_bin_ptp:
        fit = optimization.curve_fit(nuth_func, bin_centers, bin_med, x0)[0]
        if plot and time_plot_path:
            bin_idx = np.digitize(xdata, bin_edges)
            output = []
            for i in np.arange(1, len(bin_edges)):
                output.append(ydata[bin_idx == i])

    if return_details:
        details = dict(med_dh=med_dh, med_slope=med_slope, c_seed=c_seed, x0=x0,
                       bin_count=full_count, bin_med=full_med, n_final=ydata.size, **info)
        return fit, fit_fig, details
    return fit, fit_fig


def successive_med(a, first_axis=1, first_axis_only=False, sav_filter=False, sg_window=101, sg_poly=2, min_axes_count=350):
    """Reference ``coreglib.py:490-584`` with numpy / scipy.  (``:576`` uses ``med_second_smooth`` when ``sav_filter`` is
    False -- undefined there, the reference raises; restated with ``med_second``, the evident intent.)"""
    import scipy.signal
    second_axis = 0
    if first_axis == 0:
        second_axis = 1
    med_first = np.ma.median(a, axis=first_axis)
    count_first = np.ma.count(a, axis=first_axis)
    med_first[count_first < min_axes_count] = 0
    if sav_filter:
        med_first_smooth = scipy.signal.savgol_filter(med_first, window_length=sg_window, polyorder=sg_poly, mode='nearest')
        first_correction_surface = np.expand_dims(med_first_smooth, axis=first_axis)
    else:
        first_correction_surface = np.expand_dims(med_first, axis=first_axis)
    b = a - first_correction_surface
    if first_axis_only:
        med_second = np.zeros(a.shape[0 if second_axis == 1 else 1])
    else:
        med_second = np.ma.median(b, axis=second_axis)
        count_second = np.ma.count(a, axis=second_axis)
        med_second[count_second < min_axes_count] = 0
    if sav_filter:
        med_second_smooth = scipy.signal.savgol_filter(med_second, window_length=sg_window, polyorder=sg_poly, mode='nearest')
        second_correction_surface = np.expand_dims(med_second_smooth, axis=second_axis)
    else:
        second_correction_surface = np.expand_dims(med_second, axis=second_axis)
    b = b - second_correction_surface
    out = [b, med_first, med_second, first_correction_surface, second_correction_surface]
    if sav_filter:
        out.extend([med_first_smooth, med_second_smooth])
    return out
