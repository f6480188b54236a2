"""Drop-in for reference ``demcoreg/coreglib.py:14-485`` backed by ``libdemcoreg_b200.so``.

Same names, positional order, defaults and return arity as the reference
(SURVEY.md section 8b).  Rasters may be ``np.ma.MaskedArray`` (as in the reference; uploaded
once) or ``(tensor, mask_tensor)`` pairs already in device memory; datasets are
``demcoreg_b200.raster.Raster``.  matplotlib is not a dependency: every ``fig`` is ``None``
(the reference itself returns ``None`` figures, ``coreglib.py:196, 287-288``).
"""
import sys

import numpy as np

from . import _lib, engine
from .raster import Raster


def apply_xy_shift(ds, dx, dy, createcopy=True):
    """Apply horizontal shift to the dataset geotransform (reference ``coreglib.py:14-40``)."""
    print("X shift: ", dx)
    print("Y shift: ", dy)
    gt_orig = ds.GetGeoTransform()
    gt_shift = np.copy(gt_orig)
    gt_shift[0] += dx
    gt_shift[3] += dy
    print("Original geotransform:", gt_orig)
    print("Updated geotransform:", gt_shift)
    if createcopy:
        ds_align = ds.copy()
    else:
        ds_align = ds
    ds_align.SetGeoTransform(gt_shift)
    return ds_align


def apply_z_shift(ds, dz, createcopy=True):
    """Reference ``coreglib.py:42-55``: masked add of ``dz`` (scalar or array) to band 1, in float32."""
    if isinstance(dz, np.ndarray):
        print("Z shift offset array mean: ", dz.mean())
    else:
        print("Z shift offset: ", dz)
    if createcopy:
        ds_shift = ds.copy()
    else:
        ds_shift = ds
    engine.apply_z_shift_inplace(ds_shift, dz)
    return ds_shift


def nuth_func(x, a, b, c):
    """Reference ``coreglib.py:58-63``."""
    y = a * np.cos(np.deg2rad(b - x)) + c
    return y


def _to_device_masked(a):
    """masked array | (tensor, mask) -> (float32 CUDA tensor, bool/uint8 CUDA mask tensor)."""
    torch = _lib.require_cuda()
    if isinstance(a, (tuple, list)) and len(a) == 2 and hasattr(a[0], "is_cuda"):
        return a[0].to(torch.float32).contiguous(), a[1].to(torch.uint8).contiguous()
    a = np.ma.asarray(a)
    data = torch.from_numpy(np.ascontiguousarray(a.filled(0) if a.dtype == np.float32
                                                 else a.filled(0).astype(np.float32))).cuda()
    mask = torch.from_numpy(np.ascontiguousarray(np.ma.getmaskarray(a)).view(np.uint8)).cuda()
    return data, mask


def compute_offset_nuth(dh, slope, aspect, min_count=100, remove_outliers=True, plot=True):
    """Nuth and Kaab (2011) offset (reference ``coreglib.py:201-373``) -> ``(fit, fit_fig)``.

    ``fit = [a, b, c]`` of ``a*cos(deg2rad(b-x)) + c`` or ``None`` when there are too few bins.
    """
    torch = _lib.require_cuda()
    dh_t, dh_m = _to_device_masked(dh)
    sl_t, sl_m = _to_device_masked(slope)
    as_t, as_m = _to_device_masked(aspect)
    bits = (dh_m != 0).to(torch.uint8) * _lib.M_BASE + (sl_m != 0).to(torch.uint8) * _lib.M_SLOPE \
        + (as_m != 0).to(torch.uint8) * _lib.M_ASPECT
    bits = bits.contiguous()
    (med_dh,), n_dh = engine.select_quantiles(dh_t, [50.0], bits, _lib.M_BASE)
    (med_slope,), n_slope = engine.select_quantiles(sl_t, [50.0], bits, _lib.M_SLOPE)
    if n_dh < min_count:
        sys.exit("Not enough dh samples")
    if n_slope < min_count:
        sys.exit("Not enough slope/aspect samples")
    c_seed = (med_dh / np.tan(np.deg2rad(med_slope)))
    x0 = np.array([0.0, 0.0, c_seed])  # noqa: F841  (curve_fit seed; the closed-form fit does not need it)

    print("Computing common mask")
    st = engine.nuth_bin_stats(dh_t, sl_t, as_t, bits, _lib.M_BASE, _lib.M_SLOPE, _lib.M_ASPECT, remove_outliers)
    print("Initial sample count:")
    print(st.n_common)
    if remove_outliers:
        print("Removing outliers")
        print("3-sigma filter: %0.2f - %0.2f" % (st.rmin, st.rmax))
        print(st.n_final)

    bin_count = np.array(st.bin_count[:], dtype=np.float64)
    bin_med = np.array(st.bin_med[:], dtype=np.float64)
    bin_centers = np.arange(360, dtype=np.float64) + 0.5
    min_bin_sample_count = 9
    idx = bin_count >= min_bin_sample_count
    bin_med = bin_med[idx]
    bin_centers = bin_centers[idx]

    fit = None
    fit_fig = None
    min_bin_count = 90
    bin_ptp = np.ptp(np.cos(np.radians(bin_centers))) if bin_centers.size else 0.0
    min_bin_ptp = 1.0
    if len(bin_med) >= min_bin_count and bin_ptp >= min_bin_ptp:
        print("Computing fit")
        fit = engine.fit_nuth(bin_centers, bin_med)
        print(fit)
    return fit, fit_fig


def compute_offset_sad(dem1, dem2, pad=(9, 9), plot=False):
    """Sub-pixel offset by sum of absolute differences (reference ``coreglib.py:65-115``) -> ``(m, int_offset, sp_offset)``.

    Every offset of the (2*pad+1)^2 window: ``diff = dem1 window - dem2[pad:-pad]``, trimmed to its inter-quartile
    range, mean absolute value; all on the device (``dcr_sad_search``).  ``m`` is negated like the reference's.
    """
    if pad[0] != pad[1]:
        raise ValueError("square search windows only (the reference's slicing mixes pad[0]/pad[1], coreglib.py:139)")
    d1, m1 = _to_device_masked(dem1)
    d2, m2 = _to_device_masked(dem2)
    m = engine.sad_search(d1, m1, d2, m2, int(pad[0]))
    m = -m
    int_argmax = np.array(np.unravel_index(m.argmax(), m.shape))
    int_offset = int_argmax - pad
    sp_argmax = np.array(find_subpixel_peak_position(m, 'parabolic'))
    sp_offset = sp_argmax - pad
    return m, int_offset, sp_offset


def compute_offset_ncc(dem1, dem2, pad=(9, 9), prefilter=False, plot=False, noise=None):
    """Offset by normalized cross-correlation (reference ``coreglib.py:118-198``) -> ``(m, int_offset, sp_offset, fig)``.

    ``noise`` (extension): ``(ref_noise, kernel_noise)`` standard-normal arrays for the masked pixels; ``None``
    reproduces the reference, which draws them from the unseeded global ``np.random.randn`` (``coreglib.py:158-159``).
    """
    torch = _lib.require_cuda()
    if prefilter:
        raise NotImplementedError("prefilter=True (LoG edge filter via malib.nanfill) is outside the hot path; "
                                  "dem_align always passes prefilter=False (dem_align.py:140)")
    if pad[0] != pad[1]:
        raise ValueError("square search windows only (the reference's slicing mixes pad[0]/pad[1], coreglib.py:139)")
    p = int(pad[0])
    d1, m1 = _to_device_masked(dem1)
    d2, m2 = _to_device_masked(dem2)
    print("Adding random noise to masked regions")
    rn = kn = None
    kshape = (d2.shape[0] - 2 * p, d2.shape[1] - 2 * p)
    if noise is not None:
        rn = torch.from_numpy(np.ascontiguousarray(noise[0], dtype=np.float32)).cuda()
        kn = torch.from_numpy(np.ascontiguousarray(noise[1], dtype=np.float32)).cuda()
    elif bool(m1.any()) or bool(m2[p:-p, p:-p].any()):
        rn = torch.from_numpy(np.random.randn(*d1.shape).astype(np.float32)).cuda()
        kn = torch.from_numpy(np.random.randn(*kshape).astype(np.float32)).cuda()
    print("Running 2D correlation with search window (x,y): %i, %i" % (pad[1], pad[0]))
    m, _ = engine.ncc_corr(d1, m1, d2, m2, p, rn, kn)
    print("Computing sub-pixel peak")
    int_argmax = np.array(np.unravel_index(m.argmax(), m.shape))
    int_offset = int_argmax - pad
    sp_argmax = np.array(find_subpixel_peak_position(m, 'parabolic'))
    sp_offset = sp_argmax - pad
    fig = None
    return m, int_offset, sp_offset, fig


def find_first_peak(corr):
    """Reference ``coreglib.py:387-416``."""
    ind = corr.argmax()
    s = corr.shape[1]
    i = ind // s
    j = ind % s
    return i, j, corr.max()


def find_subpixel_peak_position(corr, subpixel_method='gaussian'):
    """Reference ``coreglib.py:419-485`` incl. IndexError -> map centre and the silent ``-1`` wrap."""
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


def successive_med(a, first_axis=1, first_axis_only=False, sav_filter=False, sg_window=101, sg_poly=2, min_axes_count=350):
    """Reference ``coreglib.py:490-584``: subtract the per-row / per-column medians of a difference map (along-track /
    cross-track correction).  Same arguments and return list ``[b, med_first, med_second, first_correction_surface,
    second_correction_surface(, med_first_smooth, med_second_smooth)]``.

    The line medians and counts (``np.ma.median`` / ``np.ma.count`` along an axis) and the two subtractions run on the
    device (``dcr_axis_median``, ``dcr_axis_subtract``); the Savitzky-Golay smoothing of the two 1-D profiles stays on the
    host (``scipy.signal.savgol_filter``).  Deviation: with ``sav_filter=False`` the reference dies at ``coreglib.py:576``
    (it uses ``med_second_smooth``, which only exists when ``sav_filter=True``); this function uses ``med_second`` there,
    which is what the line evidently means.
    """
    torch = _lib.require_cuda()
    second_axis = 0 if first_axis == 1 else 1
    am = np.ma.asarray(a)
    mask_np = np.ma.getmaskarray(am)
    d = torch.from_numpy(np.ascontiguousarray(am.filled(0).astype(np.float32))).cuda()
    m = torch.from_numpy(np.ascontiguousarray(mask_np.astype(np.uint8))).cuda()

    def profile(x, axis):
        med, cnt = engine.axis_median(x, m, axis)
        med = med.cpu().numpy()
        cnt = cnt.cpu().numpy()
        # np.ma.median masks a fully masked line; the reference then overwrites every line with too few samples by 0
        med = np.ma.array(med, mask=np.isnan(med))
        med[cnt < min_axes_count] = 0
        return med
    med_first = profile(d, first_axis)
    if sav_filter:
        import scipy.signal
        med_first_smooth = scipy.signal.savgol_filter(med_first, window_length=sg_window, polyorder=sg_poly, mode='nearest')
        corr1 = med_first_smooth
    else:
        corr1 = med_first
    first_correction_surface = np.expand_dims(corr1, axis=first_axis)
    engine.axis_subtract(d, torch.from_numpy(np.ascontiguousarray(np.ma.filled(corr1, 0).astype(np.float32))).cuda(),
                         1 if first_axis == 1 else 0)
    if first_axis_only:
        med_second = np.zeros(am.shape[0 if second_axis == 1 else 1])
    else:
        med_second = profile(d, second_axis)
    if sav_filter:
        import scipy.signal
        med_second_smooth = scipy.signal.savgol_filter(med_second, window_length=sg_window, polyorder=sg_poly, mode='nearest')
        corr2 = med_second_smooth
    else:
        corr2 = med_second
    second_correction_surface = np.expand_dims(corr2, axis=second_axis)
    engine.axis_subtract(d, torch.from_numpy(np.ascontiguousarray(np.ma.filled(corr2, 0).astype(np.float32))).cuda(),
                         1 if second_axis == 1 else 0)
    b = np.ma.array(d.cpu().numpy(), mask=mask_np)
    if first_axis_only:
        b = b.astype(np.float64)   # (the reference subtracts a float64 np.zeros profile there: its b becomes float64)
    out = [b, med_first, med_second, first_correction_surface, second_correction_surface]
    if sav_filter:
        out.extend([med_first_smooth, med_second_smooth])
    return out
