"""``Raster``: a device-resident stand-in for the ``gdal.Dataset`` (MEM) objects of the hot path.

The reference passes GDAL MEM datasets between ``dem_align.main``, ``compute_offset`` and
``coreglib.apply_xy_shift/apply_z_shift`` (reference ``dem_align.py:294-343``, ``coreglib.py:14-55``).
Here the single float32 band lives in HBM as a torch CUDA tensor; the handful of
``gdal.Dataset`` / ``gdal.Band`` methods those functions touch are duck-typed.
"""
import ctypes as C

import numpy as np

from . import _lib


class _Band(object):
    def __init__(self, ds):
        self._ds = ds

    def ReadAsArray(self):
        return self._ds.data.cpu().numpy()

    def WriteArray(self, a):
        torch = _lib.require_cuda()
        self._ds.data.copy_(torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)))

    def GetNoDataValue(self):
        return self._ds.ndv

    def SetNoDataValue(self, v):
        self._ds.ndv = float(v)


class Raster(object):
    """One float32 band in device memory + geotransform + nodata + projection string."""

    def __init__(self, data, gt, ndv=-9999.0, proj='LOCAL_CS["metres"]', device=None):
        torch = _lib.require_cuda()
        if isinstance(data, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(data, dtype=np.float32))
            data = t.to(device or "cuda", non_blocking=False)
        if data.dtype != torch.float32 or data.dim() != 2 or not data.is_cuda:
            raise ValueError("Raster needs a 2-D float32 CUDA tensor or numpy array")
        self._data = data.contiguous()
        # scalar z shifts not yet written to the band (coreglib.apply_z_shift defers them: the next iteration's
        # front kernel applies them while loading, bit-identically; any other reader goes through `.data`)
        self.pending_dz = []
        self.gt = tuple(float(g) for g in gt)
        self.ndv = None if ndv is None else float(ndv)
        self.proj = proj

    @property
    def data(self):
        """The band tensor with every deferred z shift applied."""
        self.materialize()
        return self._data

    @data.setter
    def data(self, t):
        self._data = t
        self.pending_dz = []

    def materialize(self):
        if self.pending_dz:
            import ctypes as C
            torch = _lib.require_cuda()
            L = _lib.lib()
            h, w = int(self._data.shape[0]), int(self._data.shape[1])
            while self.pending_dz:
                chunk, self.pending_dz = self.pending_dz[:16], self.pending_dz[16:]
                if len(chunk) == 1:   # the common case: vectorised single-shift kernel
                    _lib.check(L.dcr_apply_z_shift(C.c_void_p(self._data.data_ptr()), h, w, self._data.stride(0),
                                                   self.ndv if self.ndv is not None else 0.0, float(chunk[0]), None,
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
                    continue
                arr = (C.c_float * len(chunk))(*chunk)
                _lib.check(L.dcr_apply_z_shift_seq(C.c_void_p(self._data.data_ptr()), h, w, self._data.stride(0),
                                                   self.ndv if self.ndv is not None else 0.0, arr, len(chunk),
                                                   C.c_void_p(torch.cuda.current_stream().cuda_stream)))
            self.pending_dz = []

    # --- gdal.Dataset duck typing -------------------------------------------------------
    @property
    def RasterXSize(self):
        return int(self._data.shape[1])

    @property
    def RasterYSize(self):
        return int(self._data.shape[0])

    def GetGeoTransform(self):
        return self.gt

    def SetGeoTransform(self, gt):
        self.gt = tuple(float(g) for g in gt)

    def GetProjection(self):
        return self.proj

    def GetRasterBand(self, n=1):
        return _Band(self)

    def copy(self):
        """``iolib.mem_drv.CreateCopy('', ds, 0)``"""
        r = Raster(self._data.clone(), self.gt, self.ndv, self.proj)
        r.pending_dz = list(self.pending_dz)
        return r

    def numpy(self):
        return self.data.cpu().numpy()

    def desc(self):
        d = _lib.RasterDesc()
        for k in range(6):
            d.gt[k] = self.gt[k]
        d.width = self.RasterXSize
        d.height = self.RasterYSize
        d.ndv = self.ndv if self.ndv is not None else 0.0
        d.has_ndv = 0 if self.ndv is None else 1
        return d


class MaskSource(object):
    """World-fixed uint8 stable-surface mask (1 = masked) in device memory.

    Stands in for the output of ``dem_mask.get_mask`` (reference ``dem_align.py:26-30, 85``), whose
    external land-cover inputs are out of scope: it is sampled nearest-neighbour onto each
    iteration's common grid inside the fused kernel.
    """

    def __init__(self, array, gt, device=None):
        torch = _lib.require_cuda()
        if isinstance(array, np.ndarray):
            array = torch.from_numpy(np.ascontiguousarray(array, dtype=np.uint8)).to(device or "cuda")
        self.data = array.contiguous()
        self.gt = tuple(float(g) for g in gt)

    def desc(self):
        d = _lib.RasterDesc()
        for k in range(6):
            d.gt[k] = self.gt[k]
        d.width = int(self.data.shape[1])
        d.height = int(self.data.shape[0])
        d.ndv = 0.0
        d.has_ndv = 0
        return d


def plan_intersection(a, b, res=0.0):
    """``memwarp_multi([a, b], res, extent='intersection')`` bookkeeping -> (Grid, WarpGeom a, WarpGeom b)."""
    L = _lib.lib()
    grid, ga, gb = _lib.Grid(), _lib.WarpGeom(), _lib.WarpGeom()
    da, db = a.desc(), b.desc()
    _lib.check(L.dcr_plan_intersection(C.byref(da), C.byref(db), float(res), C.byref(grid), C.byref(ga), C.byref(gb)))
    return grid, ga, gb


def plan_extent(a, b, extent="union"):
    """``memwarp_multi([a, b], res='max', extent=...)`` bookkeeping for the extents other than 'intersection'
    (``warplib.parse_extent``: 'union' = bounding box, 'first' / 'last' = that raster's extent) -> (Grid, WarpGeom a, b)."""
    L = _lib.lib()
    if extent == "intersection":
        return plan_intersection(a, b, 0.0)

    def ext(r):
        gt = r.gt
        return [gt[0], gt[3] + gt[5] * r.RasterYSize, gt[0] + gt[1] * r.RasterXSize, gt[3]]
    ea, eb = ext(a), ext(b)
    ra, rb = (a.gt[1] + abs(a.gt[5])) / 2.0, (b.gt[1] + abs(b.gt[5])) / 2.0
    res = max(ra, rb)
    if abs(ra - res) > 1e-3 * res or abs(rb - res) > 1e-3 * res:
        raise _lib.DcrError("input resolutions differ (%.6f, %.6f): the down-sampling warp is out of scope" % (ra, rb))
    if extent == "union":
        te = [min(ea[0], eb[0]), min(ea[1], eb[1]), max(ea[2], eb[2]), max(ea[3], eb[3])]
    elif extent == "first":
        te = ea
    elif extent == "last":
        te = eb
    else:
        raise ValueError("extent must be one of intersection, union, first, last")
    grid = _lib.Grid()
    for k, v in enumerate((te[0], res, 0.0, te[3], 0.0, -res)):
        grid.gt[k] = v
    grid.width = int(round((te[2] - te[0]) / res))     # Python round(): half to even, like pygeotools
    grid.height = int(round((te[3] - te[1]) / res))
    ga, gb = _lib.WarpGeom(), _lib.WarpGeom()
    da, db = a.desc(), b.desc()
    _lib.check(L.dcr_plan_onto_grid(C.byref(da), C.byref(grid), C.byref(ga)))
    _lib.check(L.dcr_plan_onto_grid(C.byref(db), C.byref(grid), C.byref(gb)))
    return grid, ga, gb

