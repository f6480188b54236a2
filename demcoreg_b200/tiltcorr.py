"""``-tiltcorr``: 2-D polynomial fit to the filtered residuals and its removal (reference ``dem_align.py:396-447,
478-483``; ``geolib.ma_fitpoly`` / ``polyfit2d`` / ``polyval2d`` of pygeotools) on device rasters, through
``dcr_polyfit2d`` / ``dcr_polyval2d`` / ``dcr_polyval2d_apply``.
"""
import ctypes as C

import numpy as np

from . import _lib, engine


def _gt6(gt):
    return (C.c_double * 6)(*[float(g) for g in gt])


def polyfit2d(z, maskbits, reject, gt, order=1):
    """Least-squares polynomial of ``order`` over the pixels of the device raster ``z`` with ``(maskbits & reject) == 0``.

    Returns ``(model, count)``; ``model.coeff[: (order+1)**2]`` is ``geolib.polyfit2d``'s ``m`` (raw map coordinates)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    z = z.contiguous()
    h, w = int(z.shape[0]), int(z.shape[1])
    mp = None
    if maskbits is not None:
        maskbits = maskbits.contiguous()
        mp = C.c_void_p(maskbits.data_ptr())
    ws = engine.workspace(L.dcr_polyfit2d_workspace_bytes())
    model = _lib.Poly2d()
    cnt = C.c_int64(0)
    _lib.check(L.dcr_polyfit2d(C.c_void_p(z.data_ptr()), mp, int(reject), h, w, z.stride(0), _gt6(gt), int(order),
                               C.byref(model), C.byref(cnt), C.c_void_p(ws.data_ptr()), ws.numel(),
                               engine._stream_ptr(torch)))
    return model, int(cnt.value)


def polyval2d(model, gt, h, w):
    """``geolib.polyval2d`` on the pixel-centre grid of geotransform ``gt`` -> float32 device tensor (h, w)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    vals = torch.empty((int(h), int(w)), dtype=torch.float32, device="cuda")
    _lib.check(L.dcr_polyval2d(C.byref(model), _gt6(gt), int(h), int(w), C.c_void_p(vals.data_ptr()), vals.stride(0),
                               engine._stream_ptr(torch)))
    return vals


def apply_polyval(ds, model, sign=-1.0):
    """``coreglib.apply_z_shift(ds, sign * polyval2d(*get_xy_grids(ds), coeff), createcopy=False)`` on a ``Raster``."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    t = ds.data   # folds pending scalar shifts first
    _lib.check(L.dcr_polyval2d_apply(C.c_void_p(t.data_ptr()), ds.RasterYSize, ds.RasterXSize, t.stride(0),
                                     _gt6(ds.GetGeoTransform()), ds.ndv if ds.ndv is not None else 0.0,
                                     0 if ds.ndv is None else 1, C.byref(model), float(sign), engine._stream_ptr(torch)))
    return ds


def subtract_polyval(t, gt, model):
    """``diff_align -= vals`` on a device float32 tensor with geotransform ``gt``."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    _lib.check(L.dcr_polyval2d_apply(C.c_void_p(t.data_ptr()), int(t.shape[0]), int(t.shape[1]), t.stride(0), _gt6(gt), 0.0, 0,
                                     C.byref(model), -1.0, engine._stream_ptr(torch)))
    return t


def coeff_list(model):
    n = (model.order + 1) ** 2
    return [model.coeff[k] for k in range(n)]
