"""Host-side driver of the CUDA hot path: buffers, workspace and the C-ABI calls.

PyTorch is plumbing only (device memory, streams); all arithmetic happens in
``libdemcoreg_b200.so``.
"""
import ctypes as C

import numpy as np

from . import _lib
from .raster import Raster, MaskSource, plan_intersection

_ws_cache = {}
_res_cache = {}


def _stream_ptr(torch):
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def workspace(nbytes, device=None):
    """A cached uint8 device tensor of at least ``nbytes`` (grown on demand) for the C-ABI scratch."""
    import threading
    torch = _lib.require_cuda()
    dev = torch.cuda.current_device() if device is None else device
    key = (threading.get_ident(), dev)   # one scratch per (host thread, device): calls on it are stream-ordered
    t = _ws_cache.get(key)
    if t is None or t.numel() < nbytes:
        t = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda:%d" % dev)
        _ws_cache[key] = t
    return t


class IterBuffers(object):
    """Device outputs of one iteration's front end (dh, slope, aspect, mask bits) on the common grid."""

    def __init__(self, h, w, keep_warped=False):
        torch = _lib.require_cuda()
        self.h, self.w = h, w
        self.dh = torch.empty((h, w), dtype=torch.float32, device="cuda")
        self.slope = torch.empty((h, w), dtype=torch.float32, device="cuda")
        self.aspect = torch.empty((h, w), dtype=torch.float32, device="cuda")
        self.maskbits = torch.empty((h, w), dtype=torch.uint8, device="cuda")
        self.ref_w = torch.empty((h, w), dtype=torch.float32, device="cuda") if keep_warped else None
        self.src_w = torch.empty((h, w), dtype=torch.float32, device="cuda") if keep_warped else None


def make_front_args(ref, src, mask_source, max_dz, use_max_dz, slope_lim, keep_warped=False, dst_ndv=0.0, bufs=None,
                    extent='intersection'):
    """Plan the common grid (memwarp_multi, res='max', extent='intersection' by default) and fill ``dcr_front_args``."""
    L = _lib.lib()
    if extent == 'intersection':
        grid, gr, gs = plan_intersection(ref, src, 0.0)
    else:
        from .raster import plan_extent
        grid, gr, gs = plan_extent(ref, src, extent)
    h, w = grid.height, grid.width
    if bufs is None or bufs.h != h or bufs.w != w or (keep_warped and bufs.ref_w is None):
        bufs = IterBuffers(h, w, keep_warped)
    a = _lib.FrontArgs()
    a.ref = ref.data.data_ptr()
    a.ref_stride = ref.data.stride(0)
    a.ref_desc = ref.desc()
    a.ref_geom = gr
    if len(src.pending_dz) > 4:   # keep the per-load work small: fold the pending shifts in one RMW pass
        src.materialize()
    a.src = src._data.data_ptr()          # raw band: deferred z shifts are applied by the kernel while loading
    a.src_stride = src._data.stride(0)
    a.src_desc = src.desc()
    a.src_geom = gs
    a.src_ndz = len(src.pending_dz)
    for k, dzv in enumerate(src.pending_dz):
        a.src_dz[k] = float(dzv)
    if mask_source is not None:
        gm = _lib.WarpGeom()
        md = mask_source.desc()
        _lib.check(L.dcr_plan_onto_grid(C.byref(md), C.byref(grid), C.byref(gm)))
        a.smask = mask_source.data.data_ptr()
        a.smask_stride = mask_source.data.stride(0)
        a.smask_h, a.smask_w = md.height, md.width
        a.smask_nx0, a.smask_ny0 = gm.nx0, gm.ny0
    else:
        a.smask = None
    a.h, a.w = h, w
    a.ewres = grid.gt[1]
    a.nsres = grid.gt[5]
    a.dst_ndv = float(dst_ndv)
    a.max_dz = float(max_dz)
    a.use_max_dz = 1 if use_max_dz else 0
    a.slope_lo, a.slope_hi = float(slope_lim[0]), float(slope_lim[1])
    a.dh = bufs.dh.data_ptr()
    a.slope = bufs.slope.data_ptr()
    a.aspect = bufs.aspect.data_ptr()
    a.maskbits = bufs.maskbits.data_ptr()
    a.ref_w = bufs.ref_w.data_ptr() if bufs.ref_w is not None else None
    a.src_w = bufs.src_w.data_ptr() if bufs.src_w is not None else None
    return a, grid, bufs


def nuth_iteration(ref, src, mask_source=None, remove_outliers=True, max_dz=100, slope_lim=(0.1, 40),
                   keep_warped=False, bufs=None, profile=False, radix_only=False, apply_z=False, tan_slope=False, no_bins=False):
    """One full Nuth-Kaab iteration on the device (``dcr_nuth_iteration``).

    Returns ``(IterResult, Grid, IterBuffers)``; ``IterResult.dx/dy/dz`` is what
    ``dem_align.compute_offset`` returns (reference ``dem_align.py:172``).  ``apply_z=True`` additionally runs
    ``coreglib.apply_z_shift(src, dz)`` on the device before returning (``IterResult.z_applied``); the caller must
    then not apply ``dz`` again.  ``tan_slope=True`` (``DCR_ITER_TAN_SLOPE``): ``bufs.slope`` receives ``tan(slope)``
    instead of degrees -- what the ``dem_align`` loop runs with, since nothing there reads the slope raster back.
    """
    torch = _lib.require_cuda()
    L = _lib.lib()
    if apply_z:
        src.materialize()   # the device-side z shift writes the band itself: no deferred shifts may be pending
    a, grid, bufs = make_front_args(ref, src, mask_source, max_dz, remove_outliers, slope_lim, keep_warped, bufs=bufs)
    n = a.h * a.w
    nbytes = L.dcr_iteration_workspace_bytes(n)
    ws = workspace(nbytes)
    res = _lib.IterResult()
    flags = (_lib.ITER_REMOVE_OUTLIERS if remove_outliers else 0) | (_lib.ITER_PROFILE if profile else 0) \
        | (_lib.ITER_RADIX_ONLY if radix_only else 0) | (_lib.ITER_APPLY_Z if apply_z else 0) \
        | (_lib.ITER_TAN_SLOPE if tan_slope else 0) | (_lib.ITER_NO_BINS if no_bins else 0)
    _lib.check(L.dcr_nuth_iteration(C.byref(a), flags, C.byref(res),
                                    C.c_void_p(ws.data_ptr()), ws.numel(), _stream_ptr(torch)))
    return res, grid, bufs


class NuthAligner(object):
    """The ``dem_align`` iteration loop (reference ``dem_align.py:291-357``, mode ``nuth``) behind ONE native call
    per loop -- or per iteration (``step``) -- through ``dcr_align_nuth``.

    ``src`` is shifted IN PLACE: its geotransform after every iteration, its band through deferred z shifts that the
    front kernel applies while loading.  The native state (``dcr_align_state.src_dz``) is the ONLY owner of those pending
    shifts: ``run()`` / ``step()`` fold them into the band before they return, so that ``src`` is always consistent
    (band, geotransform, no ``Raster.pending_dz``) when the interpreter looks at it.

    One aligner can serve a whole batch of pairs (``reset``): the iteration outputs, the workspace and the record
    array are allocated once and reused.
    """

    def __init__(self, ref, src, mask_source=None, tol=0.02, max_offset=100, max_dz=100, slope_lim=(0.1, 40), max_iter=30,
                 remove_outliers=True):
        self.L = _lib.lib()
        self.cap = 0
        self._iters = (_lib.AlignIter * 64)()
        self._n = C.c_int(0)
        self.reset(ref, src, mask_source, tol, max_offset, max_dz, slope_lim, max_iter, remove_outliers)

    def reset(self, ref, src, mask_source=None, tol=0.02, max_offset=100, max_dz=100, slope_lim=(0.1, 40), max_iter=30,
              remove_outliers=True):
        """Point the loop at a new pair (fresh loop state); device buffers are kept when they are large enough."""
        torch = _lib.require_cuda()
        self.ref, self.src, self.mask_source = ref, src, mask_source
        st = self.st = _lib.AlignState()
        st.ref = ref.data.data_ptr()
        st.ref_stride = ref.data.stride(0)
        st.ref_desc = ref.desc()
        src.materialize()
        st.src = src._data.data_ptr()
        st.src_stride = src._data.stride(0)
        st.src_desc = src.desc()
        if mask_source is not None:
            st.smask = mask_source.data.data_ptr()
            st.smask_stride = mask_source.data.stride(0)
            st.smask_desc = mask_source.desc()
        h = min(ref.RasterYSize, src.RasterYSize) + 2
        w = min(ref.RasterXSize, src.RasterXSize) + 2
        if h * w > self.cap:
            self.cap = h * w
            self._dh = torch.empty(self.cap, dtype=torch.float32, device="cuda")
            self._slope = torch.empty(self.cap, dtype=torch.float32, device="cuda")
            self._maskbits = torch.empty(self.cap, dtype=torch.uint8, device="cuda")
            self.ws = workspace(self.L.dcr_iteration_workspace_bytes(self.cap))
        # (no aspect raster: nothing in the loop reads it -- the aspect-bin codes travel in the workspace)
        st.dh, st.slope, st.aspect, st.maskbits = self._dh.data_ptr(), self._slope.data_ptr(), None, self._maskbits.data_ptr()
        st.out_capacity = self.cap
        st.tol, st.max_offset = float(tol), float(max_offset)
        st.max_dz, st.slope_lo, st.slope_hi = float(max_dz), float(slope_lim[0]), float(slope_lim[1])
        st.max_iter = int(max_iter)
        st.remove_outliers = 1 if remove_outliers else 0
        self.grid_hw = (0, 0)
        self.iters = []
        return self

    def _publish(self):
        """Fold the pending z shifts into the band and publish the geotransform: ``src`` is consistent again."""
        torch = _lib.require_cuda()
        _lib.check(self.L.dcr_align_flush_z(C.byref(self.st), _stream_ptr(torch)))
        self.src.gt = tuple(self.st.src_desc.gt[k] for k in range(6))
        self.src.pending_dz = []

    def run(self, max_steps=0, profile=False, radix_only=False):
        """Run the loop to the end (``max_steps=0``) or exactly ``max_steps`` further iterations (fewer if it stops);
        returns the records of the iterations this call ran (dx, dy, dz, status, grid size, engines)."""
        torch = _lib.require_cuda()
        flags = (_lib.ITER_PROFILE if profile else 0) | (_lib.ITER_RADIX_ONLY if radix_only else 0) | _lib.ITER_TAN_SLOPE
        out = []
        left = int(max_steps)
        cap = len(self._iters)
        while True:
            chunk = cap if left <= 0 else min(left, cap)   # the record array bounds one native call, not the loop
            _lib.check(self.L.dcr_align_nuth(C.byref(self.st), chunk, flags, self._iters, cap, C.byref(self._n),
                                             C.c_void_p(self.ws.data_ptr()), self.ws.numel(), _stream_ptr(torch)))
            n = self._n.value
            for it in self._iters[:n]:
                out.append(dict(n=len(self.iters) + len(out) + 1, dx=it.dx, dy=it.dy, dz=it.dz, status=it.status, h=it.h,
                                w=it.w, select_engine=it.select_engine, bin_engine=it.bin_engine,
                                fit=(it.fit[0], it.fit[1], it.fit[2]), dz_med=it.dz_med, med_slope=it.med_slope))
            if left > 0:
                left -= n
                if left <= 0:
                    break
            if n < chunk or self.st.finished or self.st.exit_code:
                break
        if out:
            self.grid_hw = (out[-1]['h'], out[-1]['w'])
        self._publish()
        self.iters.extend(out)
        return out

    def step(self, profile=False):
        """Exactly one iteration (compute_offset + apply_xy_shift + apply_z_shift); returns the last ``IterResult``."""
        self.run(max_steps=1, profile=profile)
        return self.st.last

    def finish(self):
        """``src`` with every shift applied (``run`` / ``step`` already publish; kept for callers of the old protocol)."""
        self._publish()
        return self.src

    @property
    def totals(self):
        return float(self.st.dx_total), float(self.st.dy_total), float(self.st.dz_total)


def front_only(ref, src, mask_source=None, remove_outliers=True, max_dz=100, slope_lim=(0.1, 40), keep_warped=True,
               extent='intersection'):
    """Only the fused warp + dh + masks + Horn kernel (``dcr_warp_diff_horn``)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    a, grid, bufs = make_front_args(ref, src, mask_source, max_dz, remove_outliers, slope_lim, keep_warped, extent=extent)
    _lib.check(L.dcr_warp_diff_horn(C.byref(a), _stream_ptr(torch)))
    return grid, bufs


def select_quantiles(x, q_percent, mask=None, reject=0xFF, transform=0, center=0.0, radix_only=False):
    """Exact ``np.percentile`` (linear) of the unmasked elements of a device tensor -> (values f32, count).

    ``radix_only`` forces the 3-pass radix select (the fallback of the one-pass sample-bracket select)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    x = x.contiguous()
    nq = len(q_percent)
    q = (C.c_double * nq)(*[float(v) for v in q_percent])
    out = (C.c_float * nq)()
    cnt = C.c_int64(0)
    ws = workspace(L.dcr_select_workspace_bytes_n(x.numel()))
    mp = None
    if mask is not None:
        mask = mask.contiguous()
        mp = C.c_void_p(mask.data_ptr())
    fn = L.dcr_select_quantiles_radix if radix_only else L.dcr_select_quantiles
    _lib.check(fn(C.c_void_p(x.data_ptr()), mp, int(reject), x.numel(), int(transform), float(center),
                                      q, nq, out, C.byref(cnt), C.c_void_p(ws.data_ptr()), ws.numel(), _stream_ptr(torch)))
    return np.array(list(out), dtype=np.float32), int(cnt.value)


def moments(x, mask=None, reject=0xFF):
    """count, min, max, mean, std (float64 accumulation) of the unmasked elements."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    x = x.contiguous()
    out = (C.c_double * 6)()
    ws = workspace(1 << 20)
    mp = None
    if mask is not None:
        mask = mask.contiguous()
        mp = C.c_void_p(mask.data_ptr())
    _lib.check(L.dcr_moments(C.c_void_p(x.data_ptr()), mp, int(reject), x.numel(), out, C.c_void_p(ws.data_ptr()),
                             ws.numel(), _stream_ptr(torch)))
    return dict(count=int(out[0]), min=out[1], max=out[2], mean=out[3], std=out[4], sumsq=out[5])


def compress_strided(x, mask, reject, stride):
    """``compressed()[::stride]`` of the unmasked elements (row-major) as a device tensor."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    x = x.contiguous()
    n = x.numel()
    cap = n // max(1, stride) + 2
    dense = torch.empty(cap, dtype=torch.float32, device="cuda")
    nout = C.c_int64(0)
    ws = workspace((n // 2048 + 2) * 12 + 4096)
    mp = C.c_void_p(mask.contiguous().data_ptr()) if mask is not None else None
    _lib.check(L.dcr_compress_strided(C.c_void_p(x.data_ptr()), mp, int(reject), n, int(stride),
                                      C.c_void_p(dense.data_ptr()), cap, C.byref(nout), C.c_void_p(ws.data_ptr()),
                                      ws.numel(), _stream_ptr(torch)))
    return dense[:int(nout.value)]


def nuth_bin_stats(dh, slope, aspect, maskbits, reject_dh=_lib.M_DIFF, reject_slope=_lib.M_SLOPE,
                   reject_aspect=_lib.M_ASPECT, remove_outliers=True):
    """``dcr_nuth_bin_stats`` on device tensors -> NuthStats."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    n = dh.numel()
    ws = workspace(L.dcr_nuth_workspace_bytes(n))
    out = _lib.NuthStats()
    _lib.check(L.dcr_nuth_bin_stats(C.c_void_p(dh.data_ptr()), C.c_void_p(slope.data_ptr()), C.c_void_p(aspect.data_ptr()),
                                    C.c_void_p(maskbits.data_ptr()), reject_dh, reject_slope, reject_aspect, n,
                                    1 if remove_outliers else 0, C.byref(out), C.c_void_p(ws.data_ptr()), ws.numel(),
                                    _stream_ptr(torch)))
    return out


def fit_nuth(bin_centers, bin_med):
    """Closed-form least-squares fit of ``nuth_func`` (host, ``dcr_fit_nuth``) -> float64[3]."""
    L = _lib.lib()
    n = len(bin_centers)
    xc = (C.c_double * n)(*[float(v) for v in bin_centers])
    ym = (C.c_double * n)(*[float(v) for v in bin_med])
    fit = (C.c_double * 3)()
    _lib.check(L.dcr_fit_nuth(xc, ym, n, fit))
    return np.array([fit[0], fit[1], fit[2]], dtype=np.float64)


def warp_cubic(src, grid, geom, dst_ndv=0.0):
    """Standalone cubic warp of a Raster onto ``grid`` (``dcr_warp_cubic``) -> Raster."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    h, w = grid.height, grid.width
    dst = torch.empty((h, w), dtype=torch.float32, device="cuda")
    _lib.check(L.dcr_warp_cubic(C.c_void_p(src.data.data_ptr()), src.RasterYSize, src.RasterXSize, src.data.stride(0),
                                src.ndv if src.ndv is not None else 0.0, 0 if src.ndv is None else 1, C.byref(geom),
                                C.c_void_p(dst.data_ptr()), h, w, w, float(dst_ndv), _stream_ptr(torch)))
    return Raster(dst, tuple(grid.gt), dst_ndv, src.proj)


def resample_cubic(src, grid, dst_ndv=0.0):
    """GDAL cubic warp of a Raster onto a grid of ANOTHER resolution (``dcr_resample_cubic``) -> Raster."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    h, w = grid.height, grid.width
    dst = torch.empty((h, w), dtype=torch.float32, device="cuda")
    d = src.desc()
    _lib.check(L.dcr_resample_cubic(C.c_void_p(src.data.data_ptr()), C.byref(d), src.data.stride(0), C.byref(grid),
                                    C.c_void_p(dst.data_ptr()), w, float(dst_ndv), _stream_ptr(torch)))
    return Raster(dst, tuple(grid.gt), dst_ndv, src.proj)


def target_res(ds_list, res='max'):
    """``warplib.parse_res``: 'max' | 'min' | 'mean' of the inputs' (square) pixel sizes, or a number."""
    rl = [(ds.gt[1] + abs(ds.gt[5])) / 2.0 for ds in ds_list]
    if isinstance(res, str):
        if res == 'max':
            return max(rl)
        if res == 'min':
            return min(rl)
        if res == 'mean':
            return float(np.mean(rl))
        try:
            return float(res)
        except ValueError:
            raise NotImplementedError("res=%r: only min / max / mean / a number are built (reference dem_align.py:150 also "
                                      "lists common_scale_factor)" % res)
    return float(res)


def memwarp_multi(ds_list, extent='intersection', res='max'):
    """``warplib.memwarp_multi([a, b], res=res, extent=extent, r='cubic')`` (reference ``dem_align.py:294-299``: once before the
    loop, both rasters clipped / resampled to their intersection; ``compute_diff.py:97``).  A raster whose extent and
    resolution already equal the target grid is returned unchanged (same object); one of the target resolution is
    resampled by the translation-only GDAL cubic kernel; one of another resolution by the general kernel
    (``dcr_resample_cubic``: scaled footprint when down-sampling).  Band nodata 0 (pygeotools' ``dst_ndv``)."""
    from .raster import plan_extent
    a, b = ds_list
    tres = target_res(ds_list, res)
    rl = [(ds.gt[1] + abs(ds.gt[5])) / 2.0 for ds in ds_list]
    if all(abs(r - tres) < 1e-3 for r in rl):
        grid, ga, gb = plan_extent(a, b, extent)
        out = []
        for ds, g in ((a, ga), (b, gb)):
            out.append(ds if g.identity else warp_cubic(ds, grid, g, 0.0))
        return out
    # unequal resolutions: target extent in map coordinates, grid = int(round(extent / res)) (Python round, like pygeotools)
    L = _lib.lib()

    def ext(r):
        gt = r.gt
        return [gt[0], gt[3] + gt[5] * r.RasterYSize, gt[0] + gt[1] * r.RasterXSize, gt[3]]
    ea, eb = ext(a), ext(b)
    if extent == 'intersection':
        te = [max(ea[0], eb[0]), max(ea[1], eb[1]), min(ea[2], eb[2]), min(ea[3], eb[3])]
        if te[2] <= te[0] or te[3] <= te[1]:
            raise _lib.DcrError("no intersection between input extents")
    elif extent == 'union':
        te = [min(ea[0], eb[0]), min(ea[1], eb[1]), max(ea[2], eb[2]), max(ea[3], eb[3])]
    elif extent == 'first':
        te = ea
    elif extent == 'last':
        te = eb
    else:
        raise ValueError("extent must be one of intersection, union, first, last")
    grid = _lib.Grid()
    for k, v in enumerate((te[0], tres, 0.0, te[3], 0.0, -tres)):
        grid.gt[k] = v
    grid.width = int(round((te[2] - te[0]) / tres))
    grid.height = int(round((te[3] - te[1]) / tres))
    out = []
    for ds, e, r in ((a, ea, rl[0]), (b, eb, rl[1])):
        if abs(r - tres) < 1e-3 and all(abs(x - y) < 1e-3 for x, y in zip(e, te)):
            out.append(ds)
        elif abs(r - tres) < 1e-3:
            g = _lib.WarpGeom()
            d = ds.desc()
            _lib.check(L.dcr_plan_onto_grid(C.byref(d), C.byref(grid), C.byref(g)))
            out.append(warp_cubic(ds, grid, g, 0.0))
        else:
            out.append(resample_cubic(ds, grid, 0.0))
    return out


def horn(ds, want_slope=True, want_aspect=True):
    """Standalone Horn slope/aspect (``dcr_horn_slope_aspect``) -> (slope tensor|None, aspect tensor|None)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    h, w = ds.RasterYSize, ds.RasterXSize
    s = torch.empty((h, w), dtype=torch.float32, device="cuda") if want_slope else None
    a = torch.empty((h, w), dtype=torch.float32, device="cuda") if want_aspect else None
    gt = ds.GetGeoTransform()
    _lib.check(L.dcr_horn_slope_aspect(C.c_void_p(ds.data.data_ptr()), h, w, ds.data.stride(0),
                                       ds.ndv if ds.ndv is not None else 0.0, 0 if ds.ndv is None else 1,
                                       gt[1], gt[5], C.c_void_p(s.data_ptr()) if s is not None else None,
                                       C.c_void_p(a.data_ptr()) if a is not None else None, _stream_ptr(torch)))
    return s, a


def apply_z_shift_inplace(ds, dz):
    """``dcr_apply_z_shift`` on a Raster (scalar float32 ``dz`` or a device/numpy grid)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    h, w = ds.RasterYSize, ds.RasterXSize
    grid_ptr = None
    dzs = 0.0
    keep = None
    if isinstance(dz, np.ndarray) and dz.ndim == 2:
        keep = torch.from_numpy(np.ascontiguousarray(dz, dtype=np.float32)).cuda()
        grid_ptr = C.c_void_p(keep.data_ptr())
    elif hasattr(dz, "is_cuda"):
        keep = dz.to(torch.float32).contiguous()
        grid_ptr = C.c_void_p(keep.data_ptr())
    else:
        # scalar shift.  Raster.pending_dz can defer it to the next front kernel (dcr_front_args.src_dz), but the
        # per-load loop costs more than the 0.1 ms read-modify-write pass it saves (measured: front 0.88 -> 1.22 ms
        # at 8192^2), so it is written back at once.
        ds.pending_dz.append(float(np.float32(dz)))
        ds.materialize()
        return ds
    ds.materialize()
    _lib.check(L.dcr_apply_z_shift(C.c_void_p(ds._data.data_ptr()), h, w, ds._data.stride(0),
                                   ds.ndv if ds.ndv is not None else 0.0, dzs, grid_ptr, _stream_ptr(torch)))
    return ds


def ncc_corr(ref, ref_mask, src, src_mask, pad, ref_noise=None, kernel_noise=None):
    """``dcr_ncc_corr``: z-scored 'valid' cross-correlation map (float64, (2p+1)^2) of two device rasters."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    h, w = int(ref.shape[0]), int(ref.shape[1])
    ref, src = ref.contiguous(), src.contiguous()
    J = 2 * pad + 1
    m = (C.c_double * (J * J))()
    stats = (C.c_double * 4)()
    ws = workspace(L.dcr_ncc_workspace_bytes(h, w, pad))

    def ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None
    rm = ref_mask.to(torch.uint8).contiguous() if ref_mask is not None else None
    sm = src_mask.to(torch.uint8).contiguous() if src_mask is not None else None
    rn = ref_noise.to(torch.float32).contiguous() if ref_noise is not None else None
    kn = kernel_noise.to(torch.float32).contiguous() if kernel_noise is not None else None
    _lib.check(L.dcr_ncc_corr(ptr(ref), ptr(rm), ptr(src), ptr(sm), h, w, w, int(pad), ptr(rn), ptr(kn), m, stats,
                              C.c_void_p(ws.data_ptr()), ws.numel(), _stream_ptr(torch)))
    return np.array(m[:], dtype=np.float64).reshape(J, J), tuple(stats[:])


def sad_search(ref, ref_mask, src, src_mask, pad, per_offset=False, info=None):
    """``dcr_sad_search_ex``: IQR-trimmed mean absolute difference for every offset (float64, (2p+1)^2).

    ``per_offset=True`` forces the exact per-offset path (the fallback of the batched kernels) for every offset;
    ``info`` (dict) receives ``n_per_offset`` and the CUDA-event ``stage_ms`` of the batched stages."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    h, w = int(ref.shape[0]), int(ref.shape[1])
    ref, src = ref.contiguous(), src.contiguous()
    J = 2 * pad + 1
    m = (C.c_double * (J * J))()
    ws = workspace(L.dcr_sad_workspace_bytes(h, w, pad))
    rm = ref_mask.to(torch.uint8).contiguous() if ref_mask is not None else None
    sm = src_mask.to(torch.uint8).contiguous() if src_mask is not None else None
    nfb = C.c_int(0)
    stage = (C.c_float * 6)()

    def ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else None
    _lib.check(L.dcr_sad_search_ex(ptr(ref), ptr(rm), ptr(src), ptr(sm), h, w, w, int(pad),
                                   _lib.SAD_PER_OFFSET if per_offset else 0, m, C.byref(nfb),
                                   stage if (info is not None and not per_offset) else None,
                                   C.c_void_p(ws.data_ptr()), ws.numel(), _stream_ptr(torch)))
    if info is not None:
        info['n_per_offset'] = int(nfb.value)
        info['stage_ms'] = dict(zip(("prep", "sample", "pass1", "find", "pass2", "final"), [float(v) for v in stage]))
    return np.array(m[:], dtype=np.float64).reshape(J, J)


def axis_median(a, mask, axis):
    """``np.ma.median(a, axis)`` / ``np.ma.count(a, axis)`` of a device raster (``dcr_axis_median``): float32 tensor of the
    line medians (NaN for a fully masked line) and an int64 tensor of the unmasked counts."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    h, w = int(a.shape[0]), int(a.shape[1])
    nl = h if axis == 1 else w
    med = torch.empty(nl, dtype=torch.float32, device="cuda")
    cnt = torch.empty(nl, dtype=torch.int64, device="cuda")
    nb = L.dcr_axis_median_workspace_bytes(h, w, axis)
    ws = workspace(nb)
    m = mask.to(torch.uint8).contiguous() if mask is not None else None
    _lib.check(L.dcr_axis_median(C.c_void_p(a.data_ptr()), C.c_void_p(m.data_ptr()) if m is not None else None, h, w,
                                 a.stride(0), w, axis, C.c_void_p(med.data_ptr()), C.c_void_p(cnt.data_ptr()),
                                 C.c_void_p(ws.data_ptr()), ws.numel(), _stream_ptr(torch)))
    return med, cnt


def axis_subtract(a, corr, axis):
    """``a -= np.expand_dims(corr, axis)`` in place on the device (``dcr_axis_subtract``)."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    c = corr.to(torch.float32).contiguous()
    _lib.check(L.dcr_axis_subtract(C.c_void_p(a.data_ptr()), int(a.shape[0]), int(a.shape[1]), a.stride(0),
                                   C.c_void_p(c.data_ptr()), axis, _stream_ptr(torch)))
    return a


def binary_dilate(mask, iterations):
    """``scipy.ndimage.binary_dilation(mask, iterations=n)`` (cross structure) on a device uint8 0/1 mask -> new tensor."""
    torch = _lib.require_cuda()
    L = _lib.lib()
    m = mask.to(torch.uint8).contiguous().clone()
    scratch = torch.empty_like(m)
    _lib.check(L.dcr_binary_dilate(C.c_void_p(m.data_ptr()), int(m.shape[0]), int(m.shape[1]), int(iterations),
                                   C.c_void_p(scratch.data_ptr()), _stream_ptr(torch)))
    return m
