"""Synthetic fractal-terrain DEM pairs with a known injected shift.

Used by the tests and ``bench.py`` (``"data": "synthetic"``): there are no sample
rasters in the reference and no network.  Host-side numpy/scipy only; the arrays
are handed to the CUDA path (and, in tests, to the oracle) unchanged.
"""
import math

import numpy as np
import scipy.fft

NDV = -9999.0


def _spectrum(n, res, seed, beta, taper_px):
    """Complex Gaussian spectrum (rfft2 layout) * k^(-beta/2-0.5) * Gaussian band limit."""
    rng = np.random.default_rng(seed)
    ky = 2.0 * np.pi * np.fft.fftfreq(n, d=res)[:, None]
    kx = 2.0 * np.pi * np.fft.rfftfreq(n, d=res)[None, :]
    k = np.sqrt(kx * kx + ky * ky)
    amp = np.zeros_like(k)
    nz = k > 0
    amp[nz] = k[nz] ** (-beta / 2.0 - 0.5)
    if taper_px:
        kc = 2.0 * np.pi / (taper_px * res)
        amp *= np.exp(-(k / kc) ** 2)
    spec = (rng.standard_normal(k.shape) + 1j * rng.standard_normal(k.shape)) * amp
    return spec.astype(np.complex128), kx, ky


def _median_slope_deg(z, res):
    """Quick central-difference median slope (calibration only; not the Horn kernel)."""
    gy, gx = np.gradient(z.astype(np.float64), res)
    return float(np.median(np.degrees(np.arctan(np.hypot(gx, gy)))))


def blob_mask(n, seed, frac, beta=3.0, tile=None):
    """Threshold a low-frequency fBm so that ~``frac`` of the pixels are 1 (blobs)."""
    m = n if tile is None else tile
    spec, _, _ = _spectrum(m, 1.0, seed, beta, 8)
    f = scipy.fft.irfft2(spec, s=(m, m), workers=-1)
    thr = np.quantile(f[::4, ::4], 1.0 - frac)
    out = (f > thr).astype(np.uint8)
    if tile is not None and tile != n:
        reps = n // tile
        out = np.tile(out, (reps, reps))
    return out


def make_pair(n=2048, res=30.0, shift=(3.2, -1.7, 0.5), seed=20260010, noise=0.3, median_slope=12.0,
              taper_px=4, static_mask_frac=0.0, nodata_frac=0.0, outlier_frac=0.001, period=None,
              x0=500000.0, y0=4100000.0):
    """Return ``dict(ref, src, gt, ndv, mask)`` -- float32 rasters on one north-up grid.

    ``src[i, j] = T(x_j + sx, y_i - sy... )``: the source is the reference terrain evaluated at
    shifted coordinates via an exact Fourier phase shift, plus ``dz`` and white
    noise, so that the co-registration must recover ``(dx, dy, dz)_total ~= (+sx, +sy, -sz)``
    (sign derivation in SURVEY 8d).  ``period``: size of the periodic tile (<= n, n % period == 0);
    larger rasters tile it seamlessly.
    """
    p = n if period is None else period
    assert n % p == 0
    sx, sy, sz = shift
    spec, kx, ky = _spectrum(p, res, seed, 2.2, taper_px)
    ref = scipy.fft.irfft2(spec, s=(p, p), workers=-1)
    # calibrate the vertical scale so the median slope hits the target (bisection on a crop)
    crop = ref[:min(p, 1024), :min(p, 1024)]
    lo, hi = 1e-6, 1e6
    for _ in range(60):
        mid = math.sqrt(lo * hi)
        if _median_slope_deg(crop * mid, res) < median_slope:
            lo = mid
        else:
            hi = mid
    scale = math.sqrt(lo * hi)
    # source: T evaluated at (x + sx, y + sy) with y northwards = rows upwards.
    # Row index i grows southwards, so y_i = y0 - i*res and a shift of +sy in y is -sy in row space.
    phase = np.exp(1j * (kx * sx - ky * sy))
    src = scipy.fft.irfft2(spec * phase, s=(p, p), workers=-1)
    ref = (ref * scale + 2000.0)
    src = (src * scale + 2000.0 + sz)
    if p != n:
        reps = n // p
        ref = np.tile(ref, (reps, reps))
        src = np.tile(src, (reps, reps))
    if noise:
        rng = np.random.default_rng(seed + 1)
        src = src + rng.standard_normal(src.shape, dtype=np.float32) * np.float32(noise)
    ref = ref.astype(np.float32)
    src = src.astype(np.float32)
    if outlier_frac:
        rng = np.random.default_rng(seed + 4)
        k = int(outlier_frac * n * n)
        idx = rng.integers(0, n * n, size=k)
        sign = rng.choice(np.array([-200.0, 200.0], dtype=np.float32), size=k)
        src.ravel()[idx] += sign
    mask = None
    if static_mask_frac:
        mask = blob_mask(n, seed + 2, static_mask_frac, tile=min(n, 2048))
    if nodata_frac:
        holes = blob_mask(n, seed + 3, nodata_frac, tile=min(n, 2048)).astype(bool)
        src[holes] = np.float32(NDV)
        holes2 = blob_mask(n, seed + 5, nodata_frac * 0.5, tile=min(n, 2048)).astype(bool)
        ref[holes2] = np.float32(NDV)
    gt = (x0, res, 0.0, y0, 0.0, -res)
    return dict(ref=ref, src=src, gt=gt, ndv=NDV, mask=mask, shift=shift)
