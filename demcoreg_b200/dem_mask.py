"""Raster-only pieces of reference ``demcoreg/dem_mask.py``: the class / threshold masks and the final
mask combination with its optional dilation.  Fetching and warping the external land-cover products (RGI, NLCD, bare
ground, SNODAS, MODSCAG, TOA: multi-GB downloads, ``dem_mask.py:32-101, 163-394``) is out of scope -- the caller hands in the
arrays already on the DEM grid.  Masks follow the reference convention of each function: True = valid surface."""
import ctypes as C

import numpy as np

from . import _lib, engine

# dem_mask.py:125-134: NLCD class filters (31 rock, 12 ice, 11 open water, 41-43 forest)
_NLCD = {
    'rock': (False, (31,)),
    'rock+ice': (False, (31, 12)),
    'rock+ice+water': (False, (31, 12, 11)),
    'not_forest': (True, (41, 42, 43)),
    'not_forest+not_water': (True, (41, 42, 43, 11)),
}


def _dev(a, dtype):
    torch = _lib.require_cuda()
    if hasattr(a, "is_cuda"):
        return a.to(dtype).contiguous().cuda()
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype).cuda()


def get_nlcd_mask(l, filter='not_forest'):
    """Reference ``dem_mask.py:111-139`` on the class array ``l`` (uint8 NLCD codes): bool mask, ``None`` for an unknown filter."""
    torch = _lib.require_cuda()
    print("Filtering NLCD LULC with: %s" % filter)
    if filter not in _NLCD:
        print("Invalid mask type")
        return None
    invert, classes = _NLCD[filter]
    lut = np.full(256, 1 if invert else 0, dtype=np.uint8)
    for c in classes:
        lut[c] = 0 if invert else 1
    d = _dev(l, torch.uint8)
    out = torch.empty_like(d)
    L = _lib.lib()
    _lib.check(L.dcr_mask_from_classes(C.c_void_p(d.data_ptr()), d.numel(), lut.ctypes.data_as(C.c_void_p),
                                       C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.bool().cpu().numpy()


def _threshold(x, thresh, greater):
    torch = _lib.require_cuda()
    d = _dev(x, torch.float32)
    out = torch.empty(d.shape, dtype=torch.uint8, device="cuda")
    L = _lib.lib()
    _lib.check(L.dcr_mask_threshold(C.c_void_p(d.data_ptr()), d.numel(), float(thresh), 1 if greater else 0,
                                    C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    return out.bool().cpu().numpy()


def get_bareground_mask(l, bareground_thresh=60):
    """Reference ``dem_mask.py:141-157``: True where the bare-ground percentage exceeds the threshold."""
    import sys
    print("Masking pixels with <%0.1f%% bare ground" % bareground_thresh)
    if bareground_thresh < 0.0 or bareground_thresh > 100.0:
        sys.exit("Invalid bare ground percentage")
    return _threshold(l, bareground_thresh, True)


def get_toa_mask(toa, toa_thresh=0.4):
    """Reference ``dem_mask.py:403-409``: True (valid) where the TOA reflectance is <= the threshold (masked input pixels
    stay valid, as ``~getmaskarray(masked_greater(toa, thresh))`` keeps them only if they were already masked)."""
    print("Applying TOA filter (masking values >= %0.2f)" % toa_thresh)
    t = np.ma.asarray(toa)
    valid = _threshold(t.filled(np.float32(0)), toa_thresh, False)
    return valid & ~np.ma.getmaskarray(t)


def combine_masks(shape, valid_masks, dilate=None):
    """Reference ``dem_mask.py:448-565`` (the raster part): AND of the per-layer valid masks, inverted to numpy's
    True = masked convention, optionally dilated by ``dilate`` iterations of ``scipy.ndimage.binary_dilation`` applied to the
    valid area's complement exactly as the reference writes it (``~binary_dilation(~newmask)``)."""
    torch = _lib.require_cuda()
    newmask = np.ones(shape, dtype=bool)
    for m in valid_masks:
        newmask = np.logical_and(m, newmask)
    print("Generating final mask to use for reference surfaces, and applying to input DEM")
    newmask = ~newmask                      # True = invalid
    if dilate is not None:
        print("Dilating mask with %i iterations" % dilate)
        grown = engine.binary_dilate(torch.from_numpy((~newmask).astype(np.uint8)).cuda(), dilate)
        newmask = ~(grown.bool().cpu().numpy())
    return newmask
