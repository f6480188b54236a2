"""Row-band mode: one very large DEM pair split into bands of rows, one band per GPU (BASELINE.json configs[4]).

The reference never tiles -- it loads whole arrays into RAM (reference ``dem_align.py:79-80``; memory note
``dem_align_parallel.pbs:3,121``).  Here every rank keeps rows ``[r0, r1)`` of both rasters (and of the
stable-surface mask) plus ``halo`` rows on each side, obtained from its neighbours with NCCL send/recv over
NVLink (``exchange_halos``).  An iteration runs the phases of ``dcr_band_phase``; between phases the integer
histograms / counters (and three FP64 sums) are all-reduced, so every rank ends up with the same
``dcr_iter_result`` -- bit-identical in its integer parts to the single-GPU path.

The same code runs G "virtual ranks" on one device (``VirtualGroup``) with an in-process sum standing in for
the collective; that is how the band logic is tested without a multi-GPU box.
"""
import ctypes as C

import numpy as np

from . import _lib
from .raster import Raster


def band_rows(height, rank, world):
    """Owned rows ``[r0, r1)`` of rank ``rank``: contiguous, near-equal bands."""
    r0 = (height * rank) // world
    r1 = (height * (rank + 1)) // world
    return r0, r1


class BandRaster(object):
    """Rows ``[row0, row0 + data.shape[0])`` of a raster of ``height`` rows (band + halos), in device memory."""

    def __init__(self, data, row0, height, gt, ndv):
        self.data = data.contiguous()
        self.row0 = int(row0)
        self.height = int(height)
        self.gt = tuple(float(g) for g in gt)
        self.ndv = None if ndv is None else float(ndv)

    @property
    def width(self):
        return int(self.data.shape[1])

    def desc(self):
        d = _lib.RasterDesc()
        for k in range(6):
            d.gt[k] = self.gt[k]
        d.width = self.width
        d.height = self.height
        d.ndv = self.ndv if self.ndv is not None else 0.0
        d.has_ndv = 0 if self.ndv is None else 1
        return d


def make_band(array_rows, r0, r1, height, halo, gt, ndv, torch, fill=None, device="cuda"):
    """Allocate band + halo storage and place the OWNED rows; halo rows are filled by ``exchange_halos``."""
    lo = max(0, r0 - halo)
    hi = min(height, r1 + halo)
    t = torch.empty((hi - lo, array_rows.shape[1]), dtype=array_rows.dtype, device=device)
    if fill is not None:
        t.fill_(fill)
    t[r0 - lo:r1 - lo] = array_rows
    b = BandRaster(t, lo, height, gt, ndv)
    b.halo = halo
    return b, (r0 - lo, r1 - lo)


def exchange_halos(band, owned, rank, world, group=None):
    """Fill the halo rows of ``band`` from the neighbouring ranks (NCCL send/recv, one batch per call).

    ``owned`` = (first, last) stored-row indices of the rows this rank owns.  Rank r sends its top owned rows to
    r-1 and its bottom owned rows to r+1 and receives its halos from them."""
    import torch.distributed as dist
    a, b = owned
    nrows = band.data.shape[0]
    top_halo, bot_halo = a, nrows - b
    ops = []
    keep = []
    if rank > 0 and top_halo > 0:
        ops.append(dist.P2POp(dist.irecv, band.data[0:a], rank - 1, group))
    if rank < world - 1 and bot_halo > 0:
        ops.append(dist.P2POp(dist.irecv, band.data[b:nrows], rank + 1, group))
    # what the neighbours need: the same halo depth (bands are symmetric), clipped by their raster edge
    if rank > 0:
        up = band.data[a:a + _halo_of_neighbour(band, rank - 1, world, below=True)].contiguous()
        keep.append(up)
        ops.append(dist.P2POp(dist.isend, up, rank - 1, group))
    if rank < world - 1:
        k = _halo_of_neighbour(band, rank + 1, world, below=False)
        dn = band.data[b - k:b].contiguous()
        keep.append(dn)
        ops.append(dist.P2POp(dist.isend, dn, rank + 1, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return band


def _halo_of_neighbour(band, nrank, world, below):
    """Rows the neighbour ``nrank`` stores beyond its owned band on the side facing us."""
    halo = band.halo
    r0, r1 = band_rows(band.height, nrank, world)
    if below:   # neighbour is above us; its bottom halo
        return min(band.height, r1 + halo) - r1
    return r0 - max(0, r0 - halo)


class BandAligner(object):
    """One rank's state of the row-band iteration."""

    def __init__(self, ref_rows, src_rows, mask_rows, height, gt, ndv, rank, world, halo=16, keep_aspect=True):
        torch = _lib.require_cuda()
        self.torch = torch
        self.rank, self.world, self.halo = rank, world, halo
        self.height = height
        self.r0, self.r1 = band_rows(height, rank, world)
        self.ref, self.owned = make_band(ref_rows, self.r0, self.r1, height, halo, gt, ndv, torch)
        self.src, _ = make_band(src_rows, self.r0, self.r1, height, halo, gt, ndv, torch)
        self.mask = None
        if mask_rows is not None:
            self.mask, _ = make_band(mask_rows, self.r0, self.r1, height, halo, gt, None, torch)
        for b in (self.ref, self.src, self.mask):
            if b is not None:
                b.halo = halo
        self.ws = None
        self.bufs = None
        self.pending_dz = []        # z shifts of the source not yet written into the band (applied by the front kernel)
        self.sticky_flags = 0       # engines that declined for this raster stay switched off (ITER_RADIX_BINS / _ONLY)
        self.keep_aspect = keep_aspect
        # how the halos were last filled: grow_halo() repeats it with the deeper halo.  Set by whoever exchanges
        # (exchange_halos -> torch.distributed, exchange_halos_native -> the library's communicator, VirtualGroup)
        self._exchange = None
        self.halo_regrown = 0

    def needed_halo(self, gr, gs):
        """Halo rows the current geometry needs: output row i reads source rows i + iy0 - 2 .. i + iy0 + 3 of each raster
        (4-tap cubic footprint + one Horn ring), and the bands are laid out on the raster row numbers -- the relative row
        offset of the two rasters (the cumulative y shift in pixels) comes on top.  Same value on every rank."""
        return abs(int(gs.iy0) - int(gr.iy0)) + 4

    def grow_halo(self, new_halo):
        """Re-lay the three band rasters with a deeper halo and fetch the new halo rows from the neighbours (collective: every
        rank sees the same geometry and grows in the same iteration).  The owned rows are kept as they are (with every z
        shift applied so far)."""
        torch = self.torch
        if self._exchange is None:
            raise _lib.DcrError("row-band halo of %d rows too small for the current shift and no exchange registered" % self.halo)
        own = self.r1 - self.r0
        if self.world > 1 and new_halo > min(band_rows(self.height, r, self.world)[1] - band_rows(self.height, r, self.world)[0]
                                             for r in range(self.world)):
            raise _lib.DcrError("shift of %d rows exceeds a whole band (%d rows): use fewer ranks" % (new_halo, own))
        self.flush_z()
        a, b = self.owned
        for name in ("ref", "src", "mask"):
            old = getattr(self, name)
            if old is None:
                continue
            nb, owned = make_band(old.data[a:b], self.r0, self.r1, self.height, new_halo, old.gt, old.ndv, torch)
            nb.halo = new_halo
            setattr(self, name, nb)
            new_owned = owned
        self.owned = new_owned
        self.halo = new_halo
        self.halo_regrown += 1
        self._exchange(self)

    def bands(self):
        return [b for b in (self.ref, self.src, self.mask) if b is not None]

    def exchange_halos(self, group=None):
        """NCCL halo exchange of all three rasters with the neighbouring ranks."""
        self.flush_z()
        for b in self.bands():
            exchange_halos(b, self.owned, self.rank, self.world, group)
        self._exchange = lambda al: al.exchange_halos(group)

    # ---- one iteration, phase by phase -------------------------------------------------------
    def begin_iteration(self, remove_outliers=True, max_dz=100, slope_lim=(0.1, 40)):
        torch = self.torch
        L = _lib.lib()
        grid, gr, gs = _lib.Grid(), _lib.WarpGeom(), _lib.WarpGeom()
        dr, dsrc = self.ref.desc(), self.src.desc()
        _lib.check(L.dcr_plan_intersection(C.byref(dr), C.byref(dsrc), 0.0, C.byref(grid), C.byref(gr), C.byref(gs)))
        need = self.needed_halo(gr, gs)
        if need > self.halo and self.world > 1:
            # the cumulative shift has outgrown the halo: deeper halo, rows re-exchanged with the neighbours
            self.grow_halo(max(2 * self.halo, need + 8))
            dr, dsrc = self.ref.desc(), self.src.desc()
        H, W = grid.height, grid.width
        rb = min(max(self.r0 - gr.iy0, 0), H)
        re = min(max(self.r1 - gr.iy0, 0), H)
        if self.rank == self.world - 1:
            re = H
        if self.rank == 0:
            rb = 0
        if re <= rb:
            raise _lib.DcrError("empty band on rank %d: use fewer ranks for this raster" % self.rank)
        nrows = re - rb
        if self.bufs is None or self.bufs[0].shape != (nrows, W):
            self.bufs = (torch.empty((nrows, W), dtype=torch.float32, device="cuda"),
                         torch.empty((nrows, W), dtype=torch.float32, device="cuda"),
                         torch.empty((nrows, W), dtype=torch.float32, device="cuda"),
                         torch.empty((nrows, W), dtype=torch.uint8, device="cuda"))
        a = _lib.FrontArgs()
        a.ref = self.ref.data.data_ptr()
        a.ref_stride = self.ref.data.stride(0)
        a.ref_desc = dr
        a.ref_geom = gr
        a.src = self.src.data.data_ptr()
        a.src_stride = self.src.data.stride(0)
        a.src_desc = dsrc
        a.src_geom = gs
        a.ref_row0, a.ref_nrows = self.ref.row0, int(self.ref.data.shape[0])
        a.src_row0, a.src_nrows = self.src.row0, int(self.src.data.shape[0])
        if self.mask is not None:
            gm = _lib.WarpGeom()
            md = self.mask.desc()
            _lib.check(L.dcr_plan_onto_grid(C.byref(md), C.byref(grid), C.byref(gm)))
            a.smask = self.mask.data.data_ptr()
            a.smask_stride = self.mask.data.stride(0)
            a.smask_h, a.smask_w = md.height, md.width
            a.smask_nx0, a.smask_ny0 = gm.nx0, gm.ny0
            a.smask_row0, a.smask_nrows = self.mask.row0, int(self.mask.data.shape[0])
        a.h, a.w = H, W
        a.row_begin, a.row_end = rb, re
        a.ewres, a.nsres = grid.gt[1], grid.gt[5]
        a.dst_ndv = 0.0
        a.max_dz = float(max_dz)
        a.use_max_dz = 1 if remove_outliers else 0
        a.slope_lo, a.slope_hi = float(slope_lim[0]), float(slope_lim[1])
        a.dh, a.slope, a.aspect, a.maskbits = (t.data_ptr() for t in self.bufs)
        if not self.keep_aspect:
            a.aspect = None             # the bin codes come out of the front kernel: the aspect raster is optional
        a.src_ndz = len(self.pending_dz)
        for k, z in enumerate(self.pending_dz):
            a.src_dz[k] = z
        self.args = a
        self.grid = grid
        self.flags = (_lib.ITER_REMOVE_OUTLIERS if remove_outliers else 0) | self.sticky_flags
        nbytes = L.dcr_iteration_workspace_bytes(nrows * W)
        if self.ws is None or self.ws.numel() < nbytes:
            self.ws = torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")
        return L.dcr_band_nphases()

    def run_phase(self, phase):
        """Run one phase; returns the tensors (views into the workspace) that must be sum-all-reduced."""
        torch = self.torch
        L = _lib.lib()
        ptrs = (C.c_void_p * 8)()
        counts = (C.c_int64 * 8)()
        kinds = (C.c_int * 8)()
        nred = C.c_int(0)
        _lib.check(L.dcr_band_phase(C.byref(self.args), self.flags, self.rank, self.world, phase,
                                    C.c_void_p(self.ws.data_ptr()), self.ws.numel(), ptrs, counts, kinds, C.byref(nred),
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        out = []
        dt = {0: (torch.int32, 4), 1: (torch.int64, 8), 2: (torch.float64, 8)}
        for k in range(nred.value):
            dtype, sz = dt[kinds[k]]
            off = ptrs[k] - self.ws.data_ptr()
            out.append(self.ws[off:off + counts[k] * sz].view(dtype))
        return out

    def note_declined(self, res):
        """A one-pass selection / the fast bins declined (on every rank alike): switch that engine off for this raster."""
        add = _lib.ITER_RADIX_BINS if res.select_engine else _lib.ITER_RADIX_ONLY
        self.flags |= add
        self.sticky_flags |= add

    def result(self):
        torch = self.torch
        L = _lib.lib()
        res = _lib.IterResult()
        _lib.check(L.dcr_band_result(C.byref(self.args), C.c_void_p(self.ws.data_ptr()), self.ws.numel(), C.byref(res),
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        return res

    def apply_shift(self, dx, dy, dz, defer=True):
        """``apply_xy_shift`` + ``apply_z_shift`` on this rank's band (halo rows included: they stay consistent).  The z
        shift is deferred like in the native single-GPU loop: the next front kernels add the pending float32 shifts, in
        order, while they load the source tile; every third one is written into the band (``flush_z``)."""
        g = list(self.src.gt)
        g[0] += dx
        g[3] += dy
        self.src.gt = tuple(g)
        self.pending_dz.append(float(np.float32(dz)))
        if not defer or len(self.pending_dz) >= 3:
            self.flush_z()

    def flush_z(self):
        """Write the pending z shifts into the source band (same ordered float32 adds as one apply_z_shift per iteration)."""
        if not self.pending_dz:
            return
        L = _lib.lib()
        torch = self.torch
        b = self.src
        dz = (C.c_float * len(self.pending_dz))(*self.pending_dz)
        _lib.check(L.dcr_apply_z_shift_seq(C.c_void_p(b.data.data_ptr()), int(b.data.shape[0]), b.width, b.data.stride(0),
                                           b.ndv if b.ndv is not None else 0.0, dz, len(self.pending_dz),
                                           C.c_void_p(torch.cuda.current_stream().cuda_stream)))
        self.pending_dz = []


def iteration_distributed(al, group=None, **kw):
    """One row-band iteration under ``torch.distributed`` (NCCL): phases + all-reduces driven from the interpreter (28 phase
    boundaries per iteration).  Returns IterResult.  ``iteration_native`` is the production path."""
    import torch.distributed as dist
    nph = al.begin_iteration(**kw)
    for attempt in range(3):
        for ph in range(nph):
            for t in al.run_phase(ph):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        res = al.result()
        if res.status != _lib.STATUS_RERUN_RADIX:
            break
        al.note_declined(res)
    return res


class Comm(object):
    """The library's own NCCL communicator (``dcr_comm``): rank 0 makes the ncclUniqueId, ``torch.distributed`` carries it to
    the other ranks (any transport would do), every rank joins with ``dcr_comm_init``."""

    def __init__(self, rank=None, world=None, group=None):
        import torch.distributed as dist
        torch = _lib.require_cuda()
        L = _lib.lib()
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        ident = torch.zeros(128, dtype=torch.uint8)
        if self.rank == 0:
            buf = (C.c_ubyte * 128)()
            _lib.check(L.dcr_comm_unique_id(buf, 128))
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        if self.world > 1:
            dev = ident.cuda() if dist.get_backend(group) == "nccl" else ident
            dist.broadcast(dev, src=0, group=group)
            ident = dev.cpu()
        raw = (C.c_ubyte * 128)(*[int(v) for v in ident.tolist()])
        h = C.c_void_p()
        _lib.check(L.dcr_comm_init(raw, 128, self.rank, self.world, C.byref(h)))
        self.handle = h

    def close(self):
        if self.handle:
            _lib.lib().dcr_comm_destroy(self.handle)
            self.handle = None


def iteration_native(al, comm, **kw):
    """One row-band iteration as ONE native call (``dcr_band_iteration``): every phase and its ncclAllReduce enqueued back to
    back on the current stream, no interpreter work in between.  Collective over ``comm``.  Returns IterResult."""
    torch = al.torch
    L = _lib.lib()
    al.begin_iteration(**kw)
    res = _lib.IterResult()
    _lib.check(L.dcr_band_iteration(C.byref(al.args), al.flags, comm.handle, C.byref(res), C.c_void_p(al.ws.data_ptr()),
                                    al.ws.numel(), C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    # an engine that declined inside the call (and was rerun natively) stays off for the next iterations of this raster
    if not res.select_engine:
        al.sticky_flags |= _lib.ITER_RADIX_ONLY
    elif not res.bin_engine:
        al.sticky_flags |= _lib.ITER_RADIX_BINS
    return res


def exchange_halos_native(al, comm):
    """Halo exchange of the three band rasters through the library's communicator (grouped ncclSend / ncclRecv)."""
    torch = al.torch
    L = _lib.lib()
    al.flush_z()
    al._exchange = lambda x: exchange_halos_native(x, comm)
    a, b = al.owned
    for band in al.bands():
        nrows = int(band.data.shape[0])
        up = _halo_of_neighbour(band, al.rank - 1, al.world, below=True) if al.rank > 0 else 0
        dn = _halo_of_neighbour(band, al.rank + 1, al.world, below=False) if al.rank < al.world - 1 else 0
        row_bytes = band.data.stride(0) * band.data.element_size()
        _lib.check(L.dcr_band_exchange_halos(comm.handle, C.c_void_p(band.data.data_ptr()), row_bytes,
                                             a, up, 0, a, b - dn, dn, b, nrows - b,
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)))


class VirtualGroup(object):
    """G virtual ranks on ONE device: the same phases, with an in-process sum instead of the collective."""

    def __init__(self, ref, src, mask, gt, ndv, world, halo=16):
        torch = _lib.require_cuda()
        H = ref.shape[0]
        self.world = world
        self.ranks = []
        for r in range(world):
            r0, r1 = band_rows(H, r, world)
            self.ranks.append(BandAligner(torch.from_numpy(np.ascontiguousarray(ref[r0:r1])).cuda(),
                                          torch.from_numpy(np.ascontiguousarray(src[r0:r1])).cuda(),
                                          torch.from_numpy(np.ascontiguousarray(mask[r0:r1])).cuda() if mask is not None else None,
                                          H, gt, ndv, r, world, halo))
        self.exchange_halos()

    def exchange_halos(self, only=None):
        """In-process stand-in for the NCCL halo exchange: copy the neighbours' owned rows into the halos."""
        for al in self.ranks:
            al.flush_z()             # (under NCCL every rank flushes inside the collective exchange)
        for r, al in enumerate(self.ranks):
            if only is not None and al is not only:
                continue
            al._exchange = lambda x: self.exchange_halos(only=x)
            for name in ("ref", "src", "mask"):
                b = getattr(al, name)
                if b is None:
                    continue
                a, e = al.owned
                nrows = b.data.shape[0]
                if r > 0 and a > 0:
                    nb = getattr(self.ranks[r - 1], name)
                    na, ne = self.ranks[r - 1].owned
                    b.data[0:a] = nb.data[ne - a:ne]
                if r < self.world - 1 and nrows > e:
                    nb = getattr(self.ranks[r + 1], name)
                    na, ne = self.ranks[r + 1].owned
                    b.data[e:nrows] = nb.data[na:na + (nrows - e)]

    def iteration(self, radix_only=False, **kw):
        nph = 0
        for al in self.ranks:
            nph = al.begin_iteration(**kw)
            if radix_only:
                al.flags |= _lib.ITER_RADIX_ONLY
        for attempt in range(3):
            for ph in range(nph):
                outs = [al.run_phase(ph) for al in self.ranks]
                for k in range(len(outs[0])):
                    tot = outs[0][k].clone()
                    for o in outs[1:]:
                        tot += o[k]
                    for o in outs:
                        o[k].copy_(tot)
            res = [al.result() for al in self.ranks]
            if res[0].status != _lib.STATUS_RERUN_RADIX:
                break
            for al, r in zip(self.ranks, res):
                al.note_declined(r)
        return res

    def apply_shift(self, dx, dy, dz):
        for al in self.ranks:
            al.apply_shift(dx, dy, dz)
