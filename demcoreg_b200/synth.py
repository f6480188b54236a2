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
              x0=500000.0, y0=4100000.0, mask_tile=None):
    """Return ``dict(ref, src, gt, ndv, mask)`` -- float32 rasters on one north-up grid.

    ``src[i, j] = T(x_j + sx, y_i - sy... )``: the source is the reference terrain evaluated at
    shifted coordinates via an exact Fourier phase shift, plus ``dz`` and white
    noise, so that the co-registration must recover ``(dx, dy, dz)_total ~= (+sx, +sy, -sz)``
    (sign derivation in SURVEY 8d).  ``period``: size of the periodic tile (<= n, n % period == 0);
    larger rasters tile it seamlessly.  ``mask_tile``: period of the stable-surface / nodata blob masks (default
    ``min(n, 2048)``; ``n`` = masks without repetition).
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
    mt = min(n, 2048) if mask_tile is None else mask_tile
    mask = None
    if static_mask_frac:
        mask = blob_mask(n, seed + 2, static_mask_frac, tile=mt)
    if nodata_frac:
        holes = blob_mask(n, seed + 3, nodata_frac, tile=mt).astype(bool)
        src[holes] = np.float32(NDV)
        holes2 = blob_mask(n, seed + 5, nodata_frac * 0.5, tile=mt).astype(bool)
        ref[holes2] = np.float32(NDV)
    gt = (x0, res, 0.0, y0, 0.0, -res)
    return dict(ref=ref, src=src, gt=gt, ndv=NDV, mask=mask, shift=shift)


# ---------------------------------------------------------------------------------------------------------------
# The same generator on the device (torch.fft): batches of hundreds of distinct pairs and the rows of a 32768^2
# pair are built in milliseconds each instead of tens of seconds of host FFTs.  Synthetic-data plumbing only (bench.py,
# "data": "synthetic"); the rasters are handed to the CUDA path unchanged.
# ---------------------------------------------------------------------------------------------------------------
def _spectrum_torch(torch, n, res, seed, beta, taper_px, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    ky = (2.0 * math.pi) * torch.fft.fftfreq(n, d=res, device=device, dtype=torch.float64)[:, None]
    kx = (2.0 * math.pi) * torch.fft.rfftfreq(n, d=res, device=device, dtype=torch.float64)[None, :]
    k = torch.sqrt(kx * kx + ky * ky)
    amp = torch.where(k > 0, k.clamp_min(1e-300) ** (-beta / 2.0 - 0.5), torch.zeros_like(k))
    if taper_px:
        kc = 2.0 * math.pi / (taper_px * res)
        amp = amp * torch.exp(-(k / kc) ** 2)
    re = torch.randn(k.shape, generator=g, device=device, dtype=torch.float64)
    im = torch.randn(k.shape, generator=g, device=device, dtype=torch.float64)
    return torch.complex(re * amp, im * amp), kx, ky


def blob_mask_torch(torch, n, seed, frac, device, beta=3.0):
    spec, _, _ = _spectrum_torch(torch, n, 1.0, seed, beta, 8, device)
    f = torch.fft.irfft2(spec, s=(n, n))
    sub = f[::8, ::8].reshape(-1)
    thr = torch.quantile(sub[:4_000_000].float(), 1.0 - frac)
    return (f > thr).to(torch.uint8)


def terrain_scale_torch(torch, n, res, seed, device, median_slope=12.0, taper_px=4):
    """Vertical scale that gives the spectral law a median slope of ``median_slope`` degrees (bisection on a crop)."""
    spec, _, _ = _spectrum_torch(torch, n, res, seed, 2.2, taper_px, device)
    ref = torch.fft.irfft2(spec, s=(n, n))
    c = ref[:min(n, 1024), :min(n, 1024)]
    gy, gx = torch.gradient(c, spacing=res)
    g = torch.hypot(gx, gy).reshape(-1)
    med_g = float(torch.quantile(g[:4_000_000].float(), 0.5))
    return math.tan(math.radians(median_slope)) / med_g


def make_pair_torch(n=4096, res=8.0, shift=(3.2, -1.7, 0.5), seed=20260040, noise=0.3, scale=None, static_mask_frac=0.3,
                    nodata_frac=0.01, outlier_frac=0.001, device="cuda", x0=500000.0, y0=4100000.0):
    """Device twin of ``make_pair``: returns ``dict(ref, src, mask, gt, ndv, shift)`` with CUDA tensors (float32 / uint8).
    Same construction (spectral fBm, exact Fourier sub-pixel shift, white noise, +-200 m spikes, blob masks / nodata);
    the random streams are torch's, so the rasters differ from the numpy generator's for the same seed."""
    import torch
    sx, sy, sz = shift
    if scale is None:
        scale = terrain_scale_torch(torch, n, res, seed, device)
    spec, kx, ky = _spectrum_torch(torch, n, res, seed, 2.2, 4, device)
    ref = torch.fft.irfft2(spec, s=(n, n)) * scale + 2000.0
    ph = kx * sx - ky * sy
    src = torch.fft.irfft2(spec * torch.complex(torch.cos(ph), torch.sin(ph)), s=(n, n)) * scale + (2000.0 + sz)
    del spec, ph
    g = torch.Generator(device=device)
    g.manual_seed(int(seed) + 1)
    ref = ref.to(torch.float32)
    src = src.to(torch.float32)
    if noise:
        src = src + torch.randn((n, n), generator=g, device=device, dtype=torch.float32) * noise
    if outlier_frac:
        k = int(outlier_frac * n * n)
        idx = torch.randint(0, n * n, (k,), generator=g, device=device)
        sign = torch.where(torch.rand((k,), generator=g, device=device) < 0.5, -200.0, 200.0).to(torch.float32)
        src.view(-1).index_add_(0, idx, sign)
    mask = None
    if static_mask_frac:
        mask = blob_mask_torch(torch, n, seed + 2, static_mask_frac, device)
    if nodata_frac:
        src[blob_mask_torch(torch, n, seed + 3, nodata_frac, device).bool()] = NDV
        ref[blob_mask_torch(torch, n, seed + 5, nodata_frac * 0.5, device).bool()] = NDV
    gt = (x0, res, 0.0, y0, 0.0, -res)
    return dict(ref=ref.contiguous(), src=src.contiguous(), mask=mask, gt=gt, ndv=NDV, shift=shift, scale=scale)
