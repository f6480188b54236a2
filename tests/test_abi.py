"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the header declares,
its structs match the ctypes mirror, and the host-only entry points (geometry, fit) agree with the oracle."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    txt = open(os.path.join(ROOT, "include", "demcoreg_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dcr_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol(built_lib):
    from demcoreg_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(built_lib, s), "missing export: %s" % s
        assert s in _lib.SYMBOLS, "ctypes binding missing for %s" % s
    assert built_lib.dcr_version() >= 100


def test_struct_layouts_match(built_lib):
    from demcoreg_b200 import _lib
    for i, t in enumerate([_lib.RasterDesc, _lib.WarpGeom, _lib.Grid, _lib.FrontArgs, _lib.NuthStats, _lib.IterResult,
                           _lib.AlignIter, _lib.AlignState, _lib.Poly2d]):
        assert C.sizeof(t) == built_lib.dcr_struct_size(i), t.__name__
    assert built_lib.dcr_struct_size(99) == -1


def _desc(gt, w, h, ndv=-9999.0):
    from demcoreg_b200 import _lib
    d = _lib.RasterDesc()
    for k in range(6):
        d.gt[k] = gt[k]
    d.width, d.height, d.ndv, d.has_ndv = w, h, ndv, 1
    return d


@pytest.mark.parametrize("shift", [(0.0, 0.0), (3.2, -1.7), (-3.2, 1.7), (40.3, 12.9), (45.0, -15.0), (-0.0004, 0.0003)])
def test_plan_intersection_matches_oracle_geometry(built_lib, shift):
    from demcoreg_b200 import _lib
    from oracle import pygeotools_restated as pg
    gt = (500000.0, 30.0, 0.0, 4100000.0, 0.0, -30.0)
    gts = (gt[0] + shift[0], 30.0, 0.0, gt[3] + shift[1], 0.0, -30.0)
    W, H = 2048, 1024
    a, b = _desc(gt, W, H), _desc(gts, W, H)
    grid, ga, gb = _lib.Grid(), _lib.WarpGeom(), _lib.WarpGeom()
    assert built_lib.dcr_plan_intersection(C.byref(a), C.byref(b), 0.0, C.byref(grid), C.byref(ga), C.byref(gb)) == 0
    z = np.zeros((H, W), dtype=np.float32)
    oa, ob = pg.MemDataset(z, gt), pg.MemDataset(z, gts)
    xmin, ymax, ow, oh, ext = pg.intersection_grid([oa, ob], 30.0)
    assert (grid.width, grid.height) == (ow, oh)
    assert grid.gt[0] == xmin and grid.gt[3] == ymax
    outs = pg.memwarp_multi([oa, ob])
    for g, ods, out in ((ga, oa, outs[0]), (gb, ob, outs[1])):
        if out is ods:
            assert g.identity == 1 and g.ix0 == 0 and g.fx == 0.0
            continue
        assert g.identity == 0
        og = pg.warp_geometry(ods, xmin, ymax, 30.0)
        assert (g.ix0, g.iy0, g.nx0, g.ny0) == (og['ix0'], og['iy0'], og['nx0'], og['ny0'])
        assert g.fx == og['dx'] and g.fy == og['dy']
        assert tuple(g.wx) == pg.cubic_weights(og['dx']) and tuple(g.wy) == pg.cubic_weights(og['dy'])


def test_plan_errors(built_lib):
    from demcoreg_b200 import _lib
    gt = (0.0, 10.0, 0.0, 1000.0, 0.0, -10.0)
    a = _desc(gt, 100, 100)
    far = _desc((5000.0, 10.0, 0.0, 1000.0, 0.0, -10.0), 100, 100)
    grid, ga, gb = _lib.Grid(), _lib.WarpGeom(), _lib.WarpGeom()
    assert built_lib.dcr_plan_intersection(C.byref(a), C.byref(far), 0.0, C.byref(grid), C.byref(ga), C.byref(gb)) == -2
    assert b"intersection" in built_lib.dcr_last_error()
    coarse = _desc((0.0, 20.0, 0.0, 1000.0, 0.0, -20.0), 50, 50)
    assert built_lib.dcr_plan_intersection(C.byref(a), C.byref(coarse), 0.0, C.byref(grid), C.byref(ga), C.byref(gb)) == -3
    assert built_lib.dcr_plan_intersection(None, None, 0.0, None, None, None) == -1


def test_fit_nuth_matches_scipy_curve_fit(built_lib):
    """dcr_fit_nuth (closed-form LSQ) lands on the minimum scipy's LM reaches from x0 = [0, 0, c]."""
    import scipy.optimize
    from demcoreg_b200 import engine
    from oracle import coreglib as ocl
    rng = np.random.default_rng(5)
    x = np.arange(360) + 0.5
    keep = rng.random(360) > 0.2
    for a, b, c in ((3.7, 61.0, -0.4), (0.11, 250.0, 2.5), (12.0, 359.0, 0.0)):
        y = ocl.nuth_func(x, a, b, c) + 0.05 * rng.standard_normal(360)
        want = scipy.optimize.curve_fit(ocl.nuth_func, x[keep], y[keep], np.array([0.0, 0.0, c]))[0]
        got = engine.fit_nuth(x[keep], y[keep])
        for f in (np.sin, np.cos):
            assert abs(got[0] * f(np.deg2rad(got[1])) - want[0] * f(np.deg2rad(want[1]))) < 1e-6
        assert abs(got[2] - want[2]) < 1e-6
    with pytest.raises(Exception):
        engine.fit_nuth([1.0, 2.0], [1.0, 2.0])


def test_product_path_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under demcoreg_b200/ may import it."""
    pkg = os.path.join(ROOT, "demcoreg_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith(".py"):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), fn


def test_compute_entry_points_fail_loudly_without_cuda(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    from demcoreg_b200 import _lib
    from demcoreg_b200.raster import Raster
    with pytest.raises(_lib.DcrError):
        Raster(np.zeros((4, 4), dtype=np.float32), (0, 1, 0, 4, 0, -1))


def test_rasterio_lite_roundtrip(tmp_path):
    from demcoreg_b200 import rasterio_lite as rio
    rng = np.random.default_rng(0)
    a = rng.standard_normal((37, 53)).astype(np.float32)
    gt = (500000.0, 30.0, 0.0, 4100000.0, 0.0, -30.0)
    for ext in (".tif", ".npz"):
        fn = str(tmp_path / ("r" + ext))
        rio.write_raster(fn, a, gt, -9999.0)
        r = rio.read_array(fn)
        assert np.array_equal(r["array"], a) and tuple(r["gt"]) == gt and r["ndv"] == -9999.0
    fn = str(tmp_path / "m.tif")
    m = (rng.random((20, 10)) < 0.5).astype(np.uint8)
    rio.write_raster(fn, m, gt, None)
    r = rio.read_array(fn)
    assert np.array_equal(r["array"], m) and r["ndv"] is None


def test_cli_parser_matches_reference_options():
    from demcoreg_b200 import dem_align
    p = dem_align.getparser()
    a = p.parse_args(["ref.tif", "src.tif"])
    assert (a.mode, a.mask_list, a.tiltcorr, a.polyorder, a.tol, a.max_offset, a.max_dz, a.res, tuple(a.slope_lim),
            a.max_iter, a.outdir) == ('nuth', [], False, 1, 0.02, 100, 100, 'mean', (0.1, 40), 30, None)
    a = p.parse_args(["r", "s", "-mode", "ncc", "-max_offset", "200", "-res", "min", "-slope_lim", "1", "30"])
    assert a.mode == 'ncc' and a.max_offset == 200 and a.res == 'min' and a.slope_lim == [1.0, 30.0]


def test_reference_import_name_resolves():
    """``from demcoreg import coreglib`` (reference ``demcoreg/__init__.py:3``, ``dem_align.py:19``) resolves to the shim."""
    import demcoreg
    from demcoreg import coreglib
    import demcoreg.dem_align as da
    from demcoreg_b200 import coreglib as c2
    assert coreglib is c2 and da.getparser().parse_args(['a.tif', 'b.tif']).mode == 'nuth'
    for name in ('compute_offset_nuth', 'compute_offset_ncc', 'compute_offset_sad', 'apply_xy_shift', 'apply_z_shift',
                 'nuth_func', 'find_subpixel_peak_position'):
        assert callable(getattr(coreglib, name))


def test_flag_constants_match_the_header():
    """The Python mirror of the C-ABI's #define constants (iteration flags, mask bits, statuses) equals include/*.h."""
    import re
    from demcoreg_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "demcoreg_b200.h")).read()
    defs = {m.group(1): int(m.group(2).rstrip("u"), 0) for m in re.finditer(r"#define\s+(DCR_\w+)\s+(0x[0-9a-fA-F]+u?|\d+u?)\b", hdr)}
    pairs = {"DCR_ITER_REMOVE_OUTLIERS": _lib.ITER_REMOVE_OUTLIERS, "DCR_ITER_PROFILE": _lib.ITER_PROFILE,
             "DCR_ITER_RADIX_ONLY": _lib.ITER_RADIX_ONLY, "DCR_ITER_APPLY_Z": _lib.ITER_APPLY_Z,
             "DCR_ITER_TAN_SLOPE": _lib.ITER_TAN_SLOPE, "DCR_ITER_RADIX_BINS": _lib.ITER_RADIX_BINS,
             "DCR_STATUS_RERUN_RADIX": _lib.STATUS_RERUN_RADIX, "DCR_M_BASE": _lib.M_BASE, "DCR_M_SLOPE": _lib.M_SLOPE}
    for name, val in pairs.items():
        assert defs[name] == val, (name, defs[name], val)
    flags = [defs[k] for k in defs if k.startswith("DCR_ITER_")]
    assert len(set(flags)) == len(flags) and all(f & (f - 1) == 0 for f in flags), "iteration flags must be distinct bits"


def test_host_helpers_of_the_cli():
    """Pure host logic next to the kernels: dates in file names (compute_diff -rate), memwarp_multi's target resolution,
    band row split, phase count of the row-band iteration."""
    import datetime
    from demcoreg_b200 import band, compute_diff, engine
    from demcoreg_b200._lib import lib
    f = compute_diff.fn_getdatetime
    assert f("/x/20150203_1234_1020_abc-DEM_8m.tif") == datetime.datetime(2015, 2, 3, 12, 34)
    assert f("WV01_20190712_102001008A_x.tif") == datetime.datetime(2019, 7, 12)
    assert f("SETSM_WV02_20121301_bad_20120131T101112.tif") == datetime.datetime(2012, 1, 31, 10, 11, 12)
    assert f("nodate_1234567.tif") is None

    class R(object):
        def __init__(self, res):
            self.gt = (0.0, res, 0.0, 0.0, 0.0, -res)
    rs = [R(2.0), R(8.0)]
    assert engine.target_res(rs, 'max') == 8.0 and engine.target_res(rs, 'min') == 2.0
    assert engine.target_res(rs, 'mean') == 5.0 and engine.target_res(rs, '4') == 4.0 and engine.target_res(rs, 3) == 3.0
    rows = [band.band_rows(32768, r, 8) for r in range(8)]
    assert rows[0][0] == 0 and rows[-1][1] == 32768 and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
    assert lib().dcr_band_nphases() == 34
