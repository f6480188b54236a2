"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle on the same seeded inputs.

Bar (BASELINE.json north_star): masks, valid-pixel counts and per-bin counts bit-exact;
warped / dh / slope / aspect rasters within 1e-4 relative (float32); offsets within 1e-3 m.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

RTOL = 1e-4   # rasters, float32
OFFTOL = 1e-3  # metres


def _pair(n=768, res=30.0, seed=20260010, **kw):
    from demcoreg_b200 import synth
    return synth.make_pair(n=n, res=res, seed=seed, **kw)


def _ods(p, key, gt=None):
    from oracle import pygeotools_restated as pg
    return pg.MemDataset(p[key], gt or p['gt'], p['ndv'])


def _gds(p, key, gt=None):
    from demcoreg_b200.raster import Raster
    return Raster(p[key], gt or p['gt'], p.get('ndv', -9999.0))


def _shifted_gt(gt, dx, dy):
    g = list(gt)
    g[0] += dx
    g[3] += dy
    return tuple(g)


def _check_raster(got, want, ndv, name):
    gv = got != np.float32(ndv)
    wv = want != np.float32(ndv)
    assert np.array_equal(gv, wv), "%s: nodata pattern differs (%d px)" % (name, int((gv != wv).sum()))
    np.testing.assert_allclose(got[gv], want[wv], rtol=RTOL, atol=0, err_msg=name)


@pytest.mark.parametrize("shift", [(0.0, 0.0), (3.4217, -1.8140), (-40.3, 12.9), (95.5, 61.0)])
def test_front_kernel_matches_oracle(cuda, shift):
    """warp(ref), warp(src), dh, masks, Horn slope/aspect of one iteration vs the oracle restatement."""
    from demcoreg_b200 import engine, _lib
    from oracle import pygeotools_restated as pg, dem_align as oda
    p = _pair(nodata_frac=0.01)
    gt_s = _shifted_gt(p['gt'], *shift)
    ref_o, src_o = _ods(p, 'ref'), _ods(p, 'src', gt_s)
    grid, bufs = engine.front_only(_gds(p, 'ref'), _gds(p, 'src', gt_s))
    cuda.cuda.synchronize()
    rc, sc = pg.memwarp_multi([ref_o, src_o], res='max', extent='intersection')
    assert (grid.height, grid.width) == rc.array.shape
    assert tuple(grid.gt) == tuple(sc.GetGeoTransform())
    _check_raster(bufs.ref_w.cpu().numpy(), rc.array, rc.ndv, "ref_w")
    _check_raster(bufs.src_w.cpu().numpy(), sc.array, sc.ndv, "src_w")
    # warped rasters: the FP64 evaluation order is replicated, so they are in fact bit-identical
    assert np.array_equal(bufs.src_w.cpu().numpy(), sc.array)
    assert np.array_equal(bufs.ref_w.cpu().numpy(), rc.array)
    # dh + base / max_dz masks
    diff = pg.ds_getma(sc) - pg.ds_getma(rc)
    diff = np.ma.masked_outside(diff, -100, 100)
    bits = bufs.maskbits.cpu().numpy()
    gm = (bits & (_lib.M_BASE | _lib.M_MAXDZ)) != 0
    assert np.array_equal(gm, np.ma.getmaskarray(diff))
    dh = bufs.dh.cpu().numpy()
    assert np.array_equal(dh[~gm], diff.compressed())
    # slope / aspect
    slope = oda.get_filtered_slope(sc)
    aspect = pg.gdaldem_mem_ds(sc, 'aspect')
    assert np.array_equal((bits & _lib.M_SLOPE) != 0, np.ma.getmaskarray(slope))
    assert np.array_equal((bits & _lib.M_ASPECT) != 0, np.ma.getmaskarray(aspect))
    sv = ~np.ma.getmaskarray(slope)
    np.testing.assert_allclose(bufs.slope.cpu().numpy()[sv], slope.data[sv], rtol=RTOL)
    av = ~np.ma.getmaskarray(aspect)
    np.testing.assert_allclose(bufs.aspect.cpu().numpy()[av], aspect.data[av], rtol=RTOL, atol=1e-4)
    # aspect bins are mask-deciding: floor() must agree everywhere
    assert np.array_equal(np.floor(bufs.aspect.cpu().numpy()[av]), np.floor(aspect.data[av]))


def test_standalone_warp_and_horn(cuda):
    """dcr_warp_cubic / dcr_horn_slope_aspect (the memwarp / gdaldem entry points) vs the oracle."""
    from demcoreg_b200 import engine
    from demcoreg_b200.raster import plan_intersection
    from oracle import pygeotools_restated as pg
    p = _pair(n=512, nodata_frac=0.01)
    gt_s = _shifted_gt(p['gt'], 17.3, -44.1)
    ref_g, src_g = _gds(p, 'ref'), _gds(p, 'src', gt_s)
    grid, gr, gs = plan_intersection(ref_g, src_g)
    rw = engine.warp_cubic(ref_g, grid, gr)
    sw = engine.warp_cubic(src_g, grid, gs)
    rc, sc = pg.memwarp_multi([_ods(p, 'ref'), _ods(p, 'src', gt_s)], res='max', extent='intersection')
    assert np.array_equal(rw.numpy(), rc.array)
    assert np.array_equal(sw.numpy(), sc.array)
    s, a = engine.horn(sw)
    so = pg.gdaldem_slope(sc.array, sc.ndv, sc.gt[1], sc.gt[5])
    ao = pg.gdaldem_aspect(sc.array, sc.ndv)
    _check_raster(s.cpu().numpy(), so, -9999.0, "slope")
    ga, wa = a.cpu().numpy(), ao
    assert np.array_equal(ga == -9999.0, wa == -9999.0)
    v = wa != -9999.0
    np.testing.assert_allclose(ga[v], wa[v], rtol=RTOL, atol=1e-4)


@pytest.mark.parametrize("n", [1, 2, 5, 1000, 100003, 3_000_000])
def test_select_quantiles_exact(cuda, n):
    """Exact order statistics + numpy lerp vs the oracle's exact-rank percentile."""
    from demcoreg_b200 import engine
    from oracle import pygeotools_restated as pg
    rng = np.random.default_rng(n)
    x = (rng.standard_normal(n) * 3 + 0.5).astype(np.float32)
    x[rng.integers(0, n, size=max(1, n // 50))] = 0.5  # duplicates
    m = (rng.random(n) < 0.3).astype(np.uint8)
    if m.all():
        m[0] = 0
    xt, mt = cuda.from_numpy(x).cuda(), cuda.from_numpy(m).cuda()
    qs = [50.0, 25.0, 75.0, 15.9, 84.1, 0.0, 100.0]
    got, cnt = engine.select_quantiles(xt, qs, mt, 1)
    xv = x[m == 0]
    assert cnt == xv.size
    for q, g in zip(qs, got):
        assert g == pg.exact_percentile(xv, q), (q, g)
    # |x - c| transform (malib.mad)
    c = np.float32(got[0])
    (g,), _ = engine.select_quantiles(xt, [50.0], mt, 1, transform=1, center=float(c))
    assert g == pg.exact_percentile(np.fabs(xv - c), 50)


def test_select_empty_and_negative(cuda):
    from demcoreg_b200 import engine
    x = cuda.tensor([-3.0, -1.0, -2.0, 4.0], dtype=cuda.float32, device="cuda")
    m = cuda.ones(4, dtype=cuda.uint8, device="cuda")
    got, cnt = engine.select_quantiles(x, [50.0], m, 1)
    assert cnt == 0 and np.isnan(got[0])
    got, cnt = engine.select_quantiles(x, [50.0, 0.0, 100.0])
    assert cnt == 4 and got[0] == np.float32(-1.5) and got[1] == -3.0 and got[2] == 4.0


@pytest.mark.parametrize("stride", [1, 2, 11])
def test_compress_strided(cuda, stride):
    """compressed()[::stride] in row-major order of the unmasked pixels (malib.get_stats subsample)."""
    from demcoreg_b200 import engine
    rng = np.random.default_rng(7)
    n = 1_000_003
    x = rng.standard_normal(n).astype(np.float32)
    m = (rng.random(n) < 0.4).astype(np.uint8) * 4
    got = engine.compress_strided(cuda.from_numpy(x).cuda(), cuda.from_numpy(m).cuda(), 4, stride).cpu().numpy()
    want = x[m == 0][::stride]
    assert np.array_equal(got, want)


def test_moments(cuda):
    from demcoreg_b200 import engine
    rng = np.random.default_rng(3)
    x = (rng.standard_normal(500_000) * 2 + 1).astype(np.float32)
    m = (rng.random(x.size) < 0.2).astype(np.uint8)
    r = engine.moments(cuda.from_numpy(x).cuda(), cuda.from_numpy(m).cuda(), 1)
    xv = x[m == 0].astype(np.float64)
    assert r['count'] == xv.size and r['min'] == xv.min() and r['max'] == xv.max()
    assert abs(r['mean'] - xv.mean()) < 1e-12 * max(1, abs(xv.mean())) + 1e-10
    assert abs(r['std'] - xv.std()) < 1e-9


def _compare_iteration(res, det, odx, ody, odz):
    assert res.count_base == det['count_base']
    assert res.count_final == det['count_final']
    nd = det['nuth']
    assert res.nuth.n_common == nd['n_common']
    assert res.nuth.n_final == nd['n_final']
    assert np.array_equal(np.array(res.nuth.bin_count[:]), nd['bin_count'].astype(np.int64))
    gm = np.array(res.nuth.bin_med[:], dtype=np.float64)
    ok = ~np.isnan(nd['bin_med'])
    assert np.array_equal(np.isnan(gm), ~ok)
    np.testing.assert_allclose(gm[ok], nd['bin_med'][ok], rtol=1e-4, atol=1e-4)
    assert res.med_dh == nd['med_dh']
    assert abs(res.med_slope - nd['med_slope']) <= 1e-5 * abs(nd['med_slope'])  # FP32 Horn fast path: few ulp
    assert res.dz_med == det['dz_med']
    assert abs(res.dx - odx) < OFFTOL and abs(res.dy - ody) < OFFTOL and abs(res.dz - odz) < OFFTOL


@pytest.mark.parametrize("cfg", [dict(), dict(nodata_frac=0.01, static_mask_frac=0.3),
                                 dict(shift_gt=(3.4, -1.8)), dict(n=1024, res=8.0, shift_gt=(-13.7, 5.2), nodata_frac=0.01)])
def test_nuth_iteration_matches_oracle(cuda, cfg):
    """One full iteration (dcr_nuth_iteration): masks/counts exact, bin medians 1e-4, offsets 1e-3 m."""
    from demcoreg_b200 import engine, _lib
    from demcoreg_b200.raster import MaskSource
    from oracle import dem_align as oda
    cfg = dict(cfg)
    sh = cfg.pop('shift_gt', (0.0, 0.0))
    p = _pair(n=cfg.pop('n', 768), res=cfg.pop('res', 30.0), **cfg)
    gt_s = _shifted_gt(p['gt'], *sh)
    ms_g = MaskSource(p['mask'], p['gt']) if p['mask'] is not None else None
    ms_o = oda.MaskSource(p['mask'], p['gt']) if p['mask'] is not None else None
    res, grid, bufs = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), ms_g)
    det = {}
    odx, ody, odz, omask, _ = oda.compute_offset(_ods(p, 'ref'), _ods(p, 'src', gt_s), mask_source=ms_o, details=det)
    gmask = ((bufs.maskbits & _lib.M_DIFF) != 0).cpu().numpy()
    assert np.array_equal(gmask, omask)
    _compare_iteration(res, det, odx, ody, odz)
    # DCR_ITER_TAN_SLOPE (what the dem_align loop runs): the slope raster holds tan(slope); same masks, counts and bins
    res_t, _, bufs_t = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), ms_g, tan_slope=True)
    assert np.array_equal(bufs_t.maskbits.cpu().numpy(), bufs.maskbits.cpu().numpy())
    assert np.array_equal(bufs_t.dh.cpu().numpy(), bufs.dh.cpu().numpy())
    _compare_iteration(res_t, det, odx, ody, odz)
    ok = ((bufs.maskbits & _lib.M_SLOPE) == 0).cpu().numpy()
    np.testing.assert_allclose(bufs_t.slope.cpu().numpy()[ok], np.tan(np.deg2rad(det['slope'].data[ok].astype(np.float64))),
                               rtol=2e-6)


def test_strided_stats_branch(cuda):
    """count >= 4e6 -> the dz median is taken on compressed()[::stride] (malib.get_stats)."""
    from demcoreg_b200 import engine
    from oracle import dem_align as oda
    p = _pair(n=3072, res=30.0, period=1024)
    res, grid, bufs = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src'))
    det = {}
    odx, ody, odz, omask, _ = oda.compute_offset(_ods(p, 'ref'), _ods(p, 'src'), details=det)
    assert res.stats_stride == 2
    _compare_iteration(res, det, odx, ody, odz)


def test_align_loop_matches_oracle(cuda):
    """The whole dem_align loop (config 1 geometry at reduced size): same iteration count, totals within 1e-3 m."""
    from demcoreg_b200 import dem_align
    from oracle import dem_align as oda
    p = _pair(n=1024, res=30.0)
    out = dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), verbose=False)
    oo = oda.align(_ods(p, 'ref'), _ods(p, 'src'), final_stats=False)
    assert out['n_iter'] == oo['n_iter']
    for k in ('dx', 'dy', 'dz'):
        assert abs(out[k] - oo[k]) < OFFTOL, (k, out[k], oo[k])
    # known answer: injected (+3.2, -1.7, +0.5) m; cubic-warp bias is ~0.005 px
    assert abs(out['dx'] - 3.2) < 0.2 and abs(out['dy'] + 1.7) < 0.2 and abs(out['dz'] + 0.5) < 0.01
    # shifted source raster: identical float32 values (same ordered f32 adds of dz)
    assert np.array_equal(out['src_align'].numpy(), oo['src_align'].array)


def test_apply_shifts(cuda):
    from demcoreg_b200 import coreglib
    from oracle import coreglib as ocl
    p = _pair(n=256, nodata_frac=0.02)
    g, o = _gds(p, 'src'), _ods(p, 'src')
    g2 = coreglib.apply_z_shift(coreglib.apply_xy_shift(g, 1.25, -2.5), np.float32(-0.4375))
    o2 = ocl.apply_z_shift(ocl.apply_xy_shift(o, 1.25, -2.5), np.float32(-0.4375))
    assert g2.GetGeoTransform() == o2.GetGeoTransform()
    assert np.array_equal(g2.numpy(), o2.array)
    assert g2 is not g and np.array_equal(g.numpy(), p['src'])


def test_compute_offset_nuth_api(cuda):
    """coreglib.compute_offset_nuth(dh, slope, aspect) with numpy masked arrays vs the oracle's (scipy curve_fit)."""
    from demcoreg_b200 import coreglib
    from oracle import dem_align as oda, coreglib as ocl
    p = _pair(n=512)
    det = {}
    oda.compute_offset(_ods(p, 'ref'), _ods(p, 'src'), details=det)
    fit, fig = coreglib.compute_offset_nuth(det['dh'], det['slope'], det['aspect'])
    ofit = det['fit']
    assert fig is None
    for f in (np.sin, np.cos):
        assert abs(fit[0] * f(np.deg2rad(fit[1])) - ofit[0] * f(np.deg2rad(ofit[1]))) < 1e-5
    assert abs(fit[2] - ofit[2]) < 1e-5


def test_gpu_against_golden_vectors(cuda):
    """The CUDA path against the committed regression vectors (tests/golden/make_golden.py)."""
    import os
    from demcoreg_b200 import engine, _lib
    from demcoreg_b200.raster import Raster, MaskSource
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "nuth_iter_128.npz"))
    ref = Raster(g['ref'], tuple(g['gt']), -9999.0)
    src = Raster(g['src'], tuple(g['gt_src']), -9999.0)
    res, grid, bufs = engine.nuth_iteration(ref, src, MaskSource(g['mask'], tuple(g['gt'])), keep_warped=True)
    assert np.array_equal(bufs.src_w.cpu().numpy(), g['src_w']) and np.array_equal(bufs.ref_w.cpu().numpy(), g['ref_w'])
    sm = ((bufs.maskbits & _lib.M_DIFF) != 0).cpu().numpy()
    assert np.array_equal(np.packbits(sm), g['static_mask_packed'])
    assert np.array_equal(np.array(res.nuth.bin_count[:]), g['bin_count'])
    assert [res.count_base, res.count_final, res.nuth.n_common, res.nuth.n_final] == list(g['counts'])
    np.testing.assert_allclose(np.array(res.nuth.bin_med[:], dtype=np.float64), g['bin_med'], rtol=1e-4, atol=1e-4,
                               equal_nan=True)
    assert res.dz_med == g['dz_med']
    np.testing.assert_allclose([res.dx, res.dy, res.dz], g['shift'], rtol=0, atol=OFFTOL)


@pytest.mark.parametrize("kind", ["normal", "ties", "constant", "two_values", "heavy_tail", "sorted"])
@pytest.mark.parametrize("n", [8192, 8193, 70001, 2_000_003])
def test_bracket_select_equals_radix_and_oracle(cuda, kind, n):
    """One-pass sample-bracket select (with its verified fallback) == 3-pass radix select == oracle."""
    from demcoreg_b200 import engine
    from oracle import pygeotools_restated as pg
    rng = np.random.default_rng(n + len(kind))
    if kind == "normal":
        x = rng.standard_normal(n)
    elif kind == "ties":
        x = np.round(rng.standard_normal(n) * 3)          # integer-valued: massive ties -> fallback path
    elif kind == "constant":
        x = np.full(n, 2.5)
    elif kind == "two_values":
        x = (rng.random(n) < 0.5).astype(np.float64) * 7 - 3
    elif kind == "heavy_tail":
        x = rng.standard_cauchy(n)
    else:
        x = np.linspace(-5, 5, n)
    x = x.astype(np.float32)
    m = (rng.random(n) < 0.25).astype(np.uint8) * 2
    xt, mt = cuda.from_numpy(x).cuda(), cuda.from_numpy(m).cuda()
    qs = [50.0, 25.0, 84.1, 0.5, 99.9]
    a, ca = engine.select_quantiles(xt, qs, mt, 2)
    b, cb = engine.select_quantiles(xt, qs, mt, 2, radix_only=True)
    xv = x[m == 0]
    assert ca == cb == xv.size
    for q, ga, gb in zip(qs, a, b):
        want = pg.exact_percentile(xv, q)
        assert ga == want and gb == want, (kind, n, q, ga, gb, want)


def test_iteration_uses_fast_selects_and_matches_radix(cuda):
    from demcoreg_b200 import engine
    p = _pair(n=1024, nodata_frac=0.01, static_mask_frac=0.3)
    from demcoreg_b200.raster import MaskSource
    ms = MaskSource(p['mask'], p['gt'])
    gt_s = _shifted_gt(p['gt'], 3.4, -1.8)
    r1, _, _ = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), ms)
    r2, _, _ = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), ms, radix_only=True)
    assert r1.select_engine == 1 and r2.select_engine == 0
    for f in ("med1", "nmad1", "mad_lo", "mad_hi", "dz_med", "med_dh", "med_slope", "count_mad", "count_final", "dx", "dy", "dz"):
        assert getattr(r1, f) == getattr(r2, f), f
    assert list(r1.nuth.bin_count) == list(r2.nuth.bin_count)
    # fast two-pass bin medians == 3-pass radix bin medians, bit for bit
    assert np.array_equal(np.array(r1.nuth.bin_med[:]), np.array(r2.nuth.bin_med[:]), equal_nan=True)
    assert r1.nuth.n_final == r2.nuth.n_final


def test_bin_stats_fast_path_fallback_on_ties(cuda):
    """Heavily tied y values (quantised dh, constant slope) force the fast bin path to decline; the radix bins
    take over and the result still equals the oracle's binned_statistic."""
    from demcoreg_b200 import engine, _lib
    from oracle import pygeotools_restated as pg
    rng = np.random.default_rng(4)
    n = 400_000
    dh = np.round(rng.standard_normal(n) * 2).astype(np.float32)       # ~10 distinct values
    slope = np.full(n, 10.0, dtype=np.float32)
    aspect = (rng.random(n) * 360).astype(np.float32)
    bits = np.zeros(n, dtype=np.uint8)
    st = engine.nuth_bin_stats(cuda.from_numpy(dh).cuda(), cuda.from_numpy(slope).cuda(), cuda.from_numpy(aspect).cuda(),
                               cuda.from_numpy(bits).cuda())
    y = (dh / np.tan(np.deg2rad(slope))).astype(np.float32)
    u, sg = y.mean(), y.std()
    keep = (y >= u - 3 * sg) & (y <= u + 3 * sg)
    cnt, _, _ = pg.bin_stats(aspect[keep], y[keep], stat='count', nbins=360, bin_range=(0., 360.))
    med, _, _ = pg.bin_stats(aspect[keep], y[keep], stat='median', nbins=360, bin_range=(0., 360.))
    assert np.array_equal(np.array(st.bin_count[:]), cnt.filled(0).astype(np.int64))
    np.testing.assert_allclose(np.array(st.bin_med[:]), med.filled(np.nan), rtol=1e-5, equal_nan=True)


def _ncc_pair(n=320, p=6, masked=False, seed=5):
    pr = _pair(n=n, noise=0.05, outlier_frac=0.0, shift=(0.0, 0.0, 0.0), seed=20260030 + seed)
    ref = pr['ref']
    # src = ref shifted by a non-integer amount: roll + blend -> sub-pixel peak
    a = np.roll(ref, (2, -3), axis=(0, 1)).astype(np.float64)
    b = np.roll(ref, (3, -3), axis=(0, 1)).astype(np.float64)
    src = (0.65 * a + 0.35 * b).astype(np.float32)
    m1 = np.zeros(ref.shape, dtype=bool)
    m2 = np.zeros(ref.shape, dtype=bool)
    if masked:
        rng = np.random.default_rng(seed)
        m1[rng.integers(0, n, 40), rng.integers(0, n, 40)] = True
        m2[40:60, 100:130] = True
    return np.ma.array(ref, mask=m1), np.ma.array(src, mask=m2)


@pytest.mark.parametrize("masked", [False, True])
@pytest.mark.parametrize("p", [4, 9, 16])
def test_ncc_matches_oracle(cuda, masked, p):
    """compute_offset_ncc vs the oracle (scipy.signal.correlate2d in float64), shared noise arrays."""
    from demcoreg_b200 import coreglib
    from oracle import coreglib as ocl
    ref, src = _ncc_pair(n=200 + 8 * p, p=p, masked=masked)
    rng = np.random.default_rng(9)
    noise = (rng.standard_normal(ref.shape), rng.standard_normal((ref.shape[0] - 2 * p, ref.shape[1] - 2 * p)))
    m, io, so, fig = coreglib.compute_offset_ncc(ref, src, pad=(p, p), noise=noise)
    mo, ioo, soo, _ = ocl.compute_offset_ncc(ref, src, pad=(p, p), noise=noise)
    assert m.shape == mo.shape == (2 * p + 1, 2 * p + 1) and fig is None
    np.testing.assert_allclose(m, mo, rtol=0, atol=2e-6 * np.abs(mo).max())
    assert tuple(io) == tuple(ioo)
    np.testing.assert_allclose(so, soo, atol=1e-4)


@pytest.mark.parametrize("masked", [False, True])
def test_sad_matches_oracle(cuda, masked):
    from demcoreg_b200 import coreglib
    from oracle import coreglib as ocl
    p = 3
    ref, src = _ncc_pair(n=160, p=p, masked=masked)
    m, io, so = coreglib.compute_offset_sad(ref, src, pad=(p, p))
    mo, ioo, soo = ocl.compute_offset_sad(ref, src, pad=(p, p))
    np.testing.assert_allclose(m, mo, rtol=2e-6)
    assert tuple(io) == tuple(ioo)
    np.testing.assert_allclose(so, soo, atol=1e-4)


@pytest.mark.parametrize("masked", [False, True])
def test_sad_batched_equals_per_offset(cuda, masked):
    """The batched SAD kernels (sample brackets, two passes over all offsets) vs the exact per-offset path (materialised
    diff + two radix selections per offset) on a raster larger than the sample, ragged tile edges."""
    from demcoreg_b200 import coreglib, engine
    p = 5
    ref, src = _ncc_pair(n=731, p=p, masked=masked)
    d1, m1 = coreglib._to_device_masked(ref)
    d2, m2 = coreglib._to_device_masked(src)
    info = {}
    mb = engine.sad_search(d1, m1, d2, m2, p, info=info)
    mo = engine.sad_search(d1, m1, d2, m2, p, per_offset=True)
    assert info['n_per_offset'] == 0
    np.testing.assert_allclose(mb, mo, rtol=1e-6)


def test_sad_heavy_ties_take_the_per_offset_path(cuda):
    """Quantised rasters: thousands of equal differences around the quartiles -> the batched path declines those offsets
    (candidate overflow / degenerate bracket) and the exact per-offset path answers; the map still equals the oracle's."""
    from demcoreg_b200 import coreglib, engine
    from oracle import coreglib as ocl
    p = 2
    ref, src = _ncc_pair(n=300, p=p, masked=True)
    ref = np.ma.array(np.round(ref.data / 4) * 4, mask=ref.mask).astype(np.float32)
    src = np.ma.array(np.round(src.data / 4) * 4, mask=src.mask).astype(np.float32)
    d1, m1 = coreglib._to_device_masked(ref)
    d2, m2 = coreglib._to_device_masked(src)
    info = {}
    mb = engine.sad_search(d1, m1, d2, m2, p, info=info)
    assert info['n_per_offset'] > 0
    mo, _, _ = ocl.compute_offset_sad(ref, src, pad=(p, p))
    np.testing.assert_allclose(-mb, mo, rtol=2e-6)


@pytest.mark.parametrize("p,n", [(30, 340), (51, 420)])
def test_ncc_large_search_window(cuda, p, n):
    """Search radii beyond one CTA's offset rows (dem_align -max_offset 100 at 2 m needs pad 51): blocks of offset rows."""
    from demcoreg_b200 import coreglib
    from oracle import coreglib as ocl
    ref, src = _ncc_pair(n=n, p=p, masked=True)
    rng = np.random.default_rng(3)
    noise = (rng.standard_normal(ref.shape), rng.standard_normal((ref.shape[0] - 2 * p, ref.shape[1] - 2 * p)))
    m, io, so, _ = coreglib.compute_offset_ncc(ref, src, pad=(p, p), noise=noise)
    mo, ioo, soo, _ = ocl.compute_offset_ncc(ref, src, pad=(p, p), noise=noise)
    np.testing.assert_allclose(m, mo, rtol=0, atol=2e-6 * np.abs(mo).max())
    assert tuple(io) == tuple(ioo)
    np.testing.assert_allclose(so, soo, atol=1e-4)


@pytest.mark.parametrize("mode", ["ncc", "sad"])
def test_dem_align_corr_modes(cuda, mode):
    """dem_align.compute_offset(mode='ncc'|'sad') vs the oracle's (noise injected for NCC)."""
    from demcoreg_b200 import dem_align, corr
    from oracle import dem_align as oda
    n, res = 192, 30.0
    p = _pair(n=n, res=res, shift=(33.0, -21.0, 0.5), noise=0.05)
    max_offset = 60.0 if mode == 'ncc' else 45.0   # pad = int(max_offset/res + 1)
    pad = int(max_offset / res + 1)
    rng = np.random.default_rng(1)
    noise = (rng.standard_normal((n, n)), rng.standard_normal((n - 2 * pad, n - 2 * pad)))
    det = {}
    gdx, gdy, gdz, gmask, _ = corr.compute_offset_corr(_gds(p, 'ref'), _gds(p, 'src'), mode, True, max_offset, 100, (0.1, 40),
                                                       None, False, ncc_noise=noise, _details=det)
    odx, ody, odz, omask, _ = oda.compute_offset(_ods(p, 'ref'), _ods(p, 'src'), mode=mode, max_offset=max_offset,
                                                 ncc_noise=noise)
    assert np.array_equal(gmask, omask)
    assert abs(gdx - odx) < OFFTOL and abs(gdy - ody) < OFFTOL and abs(gdz - odz) < OFFTOL, (gdx, odx, gdy, ody)


def test_robust_stats_and_stats_dict(cuda):
    """robust_stats.py outputs and malib.get_stats_dict(full=True) vs the oracle."""
    from demcoreg_b200 import robust_stats as grs
    from oracle import robust_stats as ors, pygeotools_restated as pg
    rng = np.random.default_rng(11)
    a = (rng.standard_normal((700, 900)) * 0.8 + 0.1).astype(np.float32)
    a = np.round(a * 512) / 512   # quantised like real dz rasters -> meaningful mode
    mask = rng.random(a.shape) < 0.2
    ma = np.ma.array(a, mask=mask)
    g = grs.robust_stats(ma)
    o = ors.robust_stats(ma.compressed())
    assert g['count'] == o['count']
    for k in ('rmse', 'mean', 'std', 'med', 'p16', 'p84', 'spread', 'absmed', 'nmad', 'p68', 'p95'):
        assert abs(g[k] - float(o[k])) <= 1e-5 * max(1.0, abs(float(o[k]))), k
    gd = grs.get_stats_dict(ma)
    od = pg.get_stats_dict(ma, full=True)
    assert set(gd) == set(od)
    assert gd['count'] == od['count'] and gd['min'] == od['min'] and gd['max'] == od['max'] and gd['mode'] == od['mode']
    for k in ('ptp', 'mean', 'std', 'nmad', 'med', 'median', 'p16', 'p84', 'spread'):
        assert abs(gd[k] - od[k]) <= 1e-6 * max(1.0, abs(od[k])), k


@pytest.mark.parametrize("world", [2, 3, 4])
def test_row_band_virtual_ranks_match_single_gpu(cuda, world):
    """Row-band mode with G virtual ranks on one device (in-process sum instead of the all-reduce) reproduces
    the single-GPU iteration: integer results bit-identical, FP64-sum-derived values to rounding."""
    from demcoreg_b200 import engine, band
    from demcoreg_b200.raster import MaskSource
    p = _pair(n=1024, res=8.0, nodata_frac=0.01, static_mask_frac=0.3)
    gt_s = _shifted_gt(p['gt'], 13.7, -21.3)   # 1.7 px / 2.7 px: halos and intersection rows in play
    ms = MaskSource(p['mask'], p['gt'])
    r1, grid, bufs = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), ms, radix_only=True)
    vg = band.VirtualGroup(p['ref'], p['src'], p['mask'], p['gt'], p['ndv'], world, halo=8)
    for al in vg.ranks:
        al.src.gt = gt_s
    res = vg.iteration()
    rows = 0
    for al, r in zip(vg.ranks, res):
        for f in ("count_base", "count_maxdz", "count_mad", "count_final", "count_slope", "stats_stride", "med1", "nmad1",
                  "mad_lo", "mad_hi", "dz_med", "med_dh", "med_slope"):
            assert getattr(r, f) == getattr(r1, f), (f, getattr(r, f), getattr(r1, f))
        assert list(r.nuth.bin_count) == list(r1.nuth.bin_count)
        assert r.nuth.n_common == r1.nuth.n_common and r.nuth.n_final == r1.nuth.n_final
        np.testing.assert_allclose(np.array(r.nuth.bin_med[:]), np.array(r1.nuth.bin_med[:]), rtol=0, atol=0, equal_nan=True)
        assert abs(r.dx - r1.dx) < 1e-9 and abs(r.dy - r1.dy) < 1e-9 and r.dz == r1.dz
        # band outputs are the matching rows of the single-GPU rasters
        rb, re = al.args.row_begin, al.args.row_end
        assert np.array_equal(al.bufs[0].cpu().numpy(), bufs.dh[rb:re].cpu().numpy())
        assert np.array_equal(al.bufs[3].cpu().numpy(), bufs.maskbits[rb:re].cpu().numpy())
        rows += re - rb
    assert rows == grid.height


def test_row_band_strided_stats_and_loop(cuda):
    """count >= 4e6 in band mode (global compressed ranks across bands) + a 2-iteration loop with shifts applied."""
    from demcoreg_b200 import engine, band, coreglib
    p = _pair(n=3072, res=30.0, period=1024)
    vg = band.VirtualGroup(p['ref'], p['src'], None, p['gt'], p['ndv'], 4, halo=8)
    ref, src = _gds(p, 'ref'), _gds(p, 'src')
    import contextlib, io
    for it in range(2):
        r1, _, _ = engine.nuth_iteration(ref, src, radix_only=True)
        rs = vg.iteration()
        assert r1.stats_stride == 2
        for r in rs:
            assert r.dz_med == r1.dz_med and r.count_final == r1.count_final
            assert list(r.nuth.bin_count) == list(r1.nuth.bin_count)
            assert abs(r.dx - r1.dx) < 1e-9 and abs(r.dy - r1.dy) < 1e-9
        with contextlib.redirect_stdout(io.StringIO()):
            coreglib.apply_xy_shift(src, r1.dx, r1.dy, createcopy=False)
            coreglib.apply_z_shift(src, np.float32(r1.dz), createcopy=False)
        vg.apply_shift(rs[0].dx, rs[0].dy, rs[0].dz)


def test_cli_end_to_end(cuda, tmp_path):
    """dem_align.main on GeoTIFF files: same cumulative shift as the oracle loop, output raster + json written."""
    import json
    from demcoreg_b200 import dem_align, rasterio_lite as rio
    from oracle import dem_align as oda, pygeotools_restated as pg
    p = _pair(n=512)
    rfn, sfn = str(tmp_path / "ref.tif"), str(tmp_path / "src.tif")
    rio.write_raster(rfn, p['ref'], p['gt'], p['ndv'])
    rio.write_raster(sfn, p['src'], p['gt'], p['ndv'])
    info = dem_align.main([rfn, sfn, "-outdir", str(tmp_path / "out")])
    oo = oda.align(_ods(p, 'ref'), _ods(p, 'src'), final_stats=False)
    for k in ('dx', 'dy', 'dz'):
        assert abs(info['shift'][k] - oo[k]) < OFFTOL
    out = rio.read_array(info['align_fn'])
    # dem_align.py:466-483: the written product is the ORIGINAL source with the TOTAL shifts applied once
    from oracle import coreglib as ocl
    want = ocl.apply_z_shift(ocl.apply_xy_shift(_ods(p, 'src'), oo['dx'], oo['dy']), np.float32(oo['dz']))
    assert np.array_equal(out['array'], want.array)
    assert abs(out['gt'][0] - oo['src_align'].gt[0]) < 1e-6 and abs(out['gt'][3] - oo['src_align'].gt[3]) < 1e-6
    assert "_nuth_x%+0.2f_y%+0.2f_z%+0.2f" % (oo['dx'], oo['dy'], oo['dz']) in info['align_fn']
    js = json.load(open(info['align_fn'].replace('_align.tif', '_align_stats.json')))
    assert js['n_iter'] == oo['n_iter']
    # final-difference statistics (dem_align.py:366-392) vs the oracle's get_stats_dict(full=True)
    oo = oda.align(_ods(p, 'ref'), _ods(p, 'src'), final_stats=True)
    for key, okey in (('after', 'diff_align_stats'), ('after_filt', 'diff_align_filt_stats')):
        assert set(js[key]) == set(oo[okey])
        assert js[key]['count'] == oo[okey]['count'], key
        for k in ('min', 'max', 'ptp', 'mean', 'std', 'nmad', 'med', 'median', 'p16', 'p84', 'spread', 'mode'):
            assert abs(js[key][k] - oo[okey][k]) <= 1e-5 * max(1.0, abs(oo[okey][k])), (key, k, js[key][k], oo[okey][k])
    assert set(js) >= {'src_fn', 'ref_fn', 'align_fn', 'res', 'center_coord', 'shift', 'before', 'before_filt', 'after', 'after_filt'}
    d = rio.read_array(info['align_fn'].replace('_align.tif', '_align_diff.tif'))
    dm = np.ma.masked_values(d['array'], -9999.0)
    assert dm.count() == oo['diff_align_stats']['count']
    # *_align_filt.tif (dem_align.py:488-499): the aligned source, masked by the cubic-warped filtered difference map
    f = rio.read_array(info['align_fn'].replace('_align.tif', '_align_filt.tif'))
    dfilt = pg.MemDataset(oo['diff_align_filt'].filled(-9999.0).astype(np.float32), tuple(d['gt']), -9999.0)
    res_px = want.gt[1]
    geom = pg.warp_geometry(dfilt, want.gt[0], want.gt[3], res_px)
    dfw = pg.gdal_cubic_warp(dfilt.array, -9999.0, geom, want.array.shape[0], want.array.shape[1], dst_ndv=0.0)
    keep = ~np.ma.getmaskarray(np.ma.masked_values(dfw, 0.0))
    assert np.array_equal(f['array'] != np.float32(-9999.0), keep & (want.array != np.float32(-9999.0)))
    assert np.array_equal(f['array'][keep], want.array[keep])


def test_compute_diff(cuda, tmp_path):
    from demcoreg_b200 import compute_diff
    from oracle import pygeotools_restated as pg
    p = _pair(n=300, nodata_frac=0.02)
    gt_s = _shifted_gt(p['gt'], 47.0, -13.0)
    diff, gt = compute_diff.compute_diff(_gds(p, 'ref'), _gds(p, 'src', gt_s))
    rc, sc = pg.memwarp_multi([_ods(p, 'ref'), _ods(p, 'src', gt_s)])
    want = pg.ds_getma(sc) - pg.ds_getma(rc)
    assert np.array_equal(np.ma.getmaskarray(diff), np.ma.getmaskarray(want))
    assert np.array_equal(diff.compressed(), want.compressed()) and gt == tuple(sc.gt)


def test_deferred_z_shift_equals_eager(cuda):
    """dcr_front_args.src_dz (z shifts applied while the front kernel loads the source) == dcr_apply_z_shift first."""
    from demcoreg_b200 import engine
    p = _pair(n=512, nodata_frac=0.02)
    gt_s = _shifted_gt(p['gt'], 3.4, -1.8)
    dzs = [np.float32(-0.4375), np.float32(0.0013), np.float32(0.25)]
    eager = _gds(p, 'src', gt_s)
    for dz in dzs:
        engine.apply_z_shift_inplace(eager, dz)
    lazy = _gds(p, 'src', gt_s)
    lazy.pending_dz = [float(d) for d in dzs]
    r1, _, b1 = engine.nuth_iteration(_gds(p, 'ref'), eager, keep_warped=True)
    r2, _, b2 = engine.nuth_iteration(_gds(p, 'ref'), lazy, keep_warped=True)
    assert len(lazy.pending_dz) == 3   # still deferred
    assert np.array_equal(b1.src_w.cpu().numpy(), b2.src_w.cpu().numpy())
    assert np.array_equal(b1.maskbits.cpu().numpy(), b2.maskbits.cpu().numpy())
    assert (r1.dx, r1.dy, r1.dz) == (r2.dx, r2.dy, r2.dz)
    assert np.array_equal(lazy.numpy(), eager.numpy()) and lazy.pending_dz == []


def test_front_kernel_tma_equals_manual_staging(cuda, monkeypatch):
    """The TMA-staged front kernel and the coalesced-row-load staging (DCR_NO_TMA=1, also the path for rasters whose
    row pitch is not a multiple of 16 bytes) produce bit-identical outputs."""
    from demcoreg_b200 import engine
    p = _pair(n=640, nodata_frac=0.02, static_mask_frac=0.2)
    from demcoreg_b200.raster import MaskSource
    outs = []
    for gt_shift in ((0.0, 0.0), (37.3, -95.2), (-61.0, 8.4)):
        gt_s = _shifted_gt(p['gt'], *gt_shift)
        res = []
        for no_tma in (False, True):
            if no_tma:
                monkeypatch.setenv("DCR_NO_TMA", "1")
            else:
                monkeypatch.delenv("DCR_NO_TMA", raising=False)
            grid, bufs = engine.front_only(_gds(p, 'ref'), _gds(p, 'src', gt_s), MaskSource(p['mask'], p['gt']))
            res.append([t.cpu().numpy() for t in (bufs.dh, bufs.slope, bufs.aspect, bufs.maskbits, bufs.ref_w, bufs.src_w)])
        for a, b in zip(*res):
            assert np.array_equal(a, b, equal_nan=True)
    monkeypatch.delenv("DCR_NO_TMA", raising=False)
    # odd width: row pitch not a multiple of 16 bytes -> the library must pick the manual path by itself
    q = _pair(n=256)
    ref = _gds(dict(ref=q['ref'][:, :253].copy()), 'ref', q['gt'])
    src = _gds(dict(src=q['src'][:, :253].copy()), 'src', q['gt'])
    from oracle import pygeotools_restated as pg, dem_align as oda
    grid, bufs = engine.front_only(ref, src)
    det = {}
    oda.compute_offset(pg.MemDataset(q['ref'][:, :253], q['gt'], q['ndv']), pg.MemDataset(q['src'][:, :253], q['gt'], q['ndv']),
                       details=det)
    assert np.array_equal(bufs.src_w.cpu().numpy(), det['src_w'])


@pytest.mark.parametrize("n,nodata", [(768, 0.0), (640, 0.02)])
def test_native_align_loop_equals_python_loop(cuda, n, nodata):
    """``dcr_align_nuth`` (whole loop in one native call, deferred z shifts applied by the front kernel through the
    TMA-staged tile) against the interpreter-driven loop (one ``dcr_nuth_iteration`` + device z shift per iteration):
    same iterations, same totals, bit-identical aligned band."""
    from demcoreg_b200 import dem_align
    from demcoreg_b200.raster import MaskSource
    p = _pair(n=n, res=30.0, nodata_frac=nodata, static_mask_frac=0.2)
    ms = MaskSource(p['mask'], p['gt'])
    a = dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), mask_source=ms, verbose=False)                 # native
    b = dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), mask_source=ms, verbose=False, records=[])     # python loop
    assert a['n_iter'] == b['n_iter'] and a['n_iter'] >= 2
    for ia, ib in zip(a['iters'], b['iters']):
        for k in ('dx', 'dy', 'dz'):
            assert ia[k] == ib[k], (ia, ib)
    for k in ('dx', 'dy', 'dz'):
        assert a[k] == b[k]
    assert a['src_align'].gt == b['src_align'].gt
    assert np.array_equal(a['src_align'].numpy(), b['src_align'].numpy())


def test_native_align_stepper_many_steps(cuda):
    """Stepping: the native state is the only owner of the deferred z shifts and step() / run() fold them into the band
    before they return, so `src` is consistent (band + geotransform) whenever the interpreter looks at it; a run(4) in the
    middle passes DCR_ALIGN_MAX_PENDING inside one native call.  The band equals the python loop's after the same number of
    iterations (tol=0)."""
    from demcoreg_b200 import dem_align, engine
    p = _pair(n=512, res=30.0)
    ref, src = _gds(p, 'ref'), _gds(p, 'src')
    al = engine.NuthAligner(ref, src, tol=0.0, max_iter=1 << 30)
    r = al.step()
    assert r.status == 0 and al.st.src_ndz == 0 and src.pending_dz == []
    one = dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), tol=0.0, max_iter=1, verbose=False, records=[])
    assert src.gt == one['src_align'].gt and np.array_equal(src.numpy(), one['src_align'].numpy())   # no double shift
    assert len(al.run(max_steps=4)) == 4
    al.step()
    assert al.st.n_done == 6 and al.st.src_ndz == 0 and len(al.iters) == 6
    al.finish()
    b = dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), tol=0.0, max_iter=6, verbose=False, records=[])
    assert b['n_iter'] == 6
    assert src.gt == b['src_align'].gt
    assert np.array_equal(src.numpy(), b['src_align'].numpy())
    dxt, dyt, dzt = al.totals
    assert dxt == b['dx'] and dyt == b['dy'] and dzt == b['dz']


def test_bracket_select_staging_overflow(cuda):
    """10 % of the elements tie exactly at the median: the bracket then holds ~16 % of the data, more than a CTA's
    staging buffer (4096 keys) at this size, so the collector's overflow path (direct global reservations per warp)
    runs -- the result must stay exact (candidate buffer n/6 is still large enough: no radix fallback needed)."""
    from demcoreg_b200 import engine
    n = 40_000_000
    rng = np.random.default_rng(7)
    x = rng.standard_normal(n).astype(np.float32)
    x[rng.random(n) < 0.10] = np.float32(0.0)
    m = (rng.random(n) < 0.25).astype(np.uint8)
    xt, mt = cuda.from_numpy(x).cuda(), cuda.from_numpy(m).cuda()
    a, ca = engine.select_quantiles(xt, [50.0], mt, 1)
    b, cb = engine.select_quantiles(xt, [50.0], mt, 1, radix_only=True)
    xv = x[m == 0]
    want = np.float32(np.percentile(xv, 50.0))   # ranks sit inside the tie block: value 0 exactly
    assert ca == cb == xv.size
    assert a[0] == b[0] == want == np.float32(0.0)
    # a quantile outside the tie block on the same data
    a, _ = engine.select_quantiles(xt, [90.0], mt, 1)
    b, _ = engine.select_quantiles(xt, [90.0], mt, 1, radix_only=True)
    assert a[0] == b[0]


def _crop(p, key, r0, r1, c0, c1):
    """A sub-raster with its own geotransform (north-up): rows [r0, r1), cols [c0, c1)."""
    gt = list(p['gt'])
    gt[0] += c0 * gt[1]
    gt[3] += r0 * gt[5]
    return np.ascontiguousarray(p[key][r0:r1, c0:c1]), tuple(gt)


@pytest.mark.parametrize("case", ["partial_overlap", "src_inside_ref", "non_square_shifted"])
def test_iteration_ragged_inputs(cuda, case):
    """Rasters of different sizes / origins (partial overlap, containment, non-square grids with a sub-pixel shift):
    the intersection grid, masks and counts must equal the oracle's memwarp_multi(extent='intersection') path."""
    from demcoreg_b200 import engine, _lib
    from demcoreg_b200.raster import Raster, MaskSource
    from oracle import dem_align as oda, pygeotools_restated as pg
    p = _pair(n=1024, res=30.0, nodata_frac=0.01, static_mask_frac=0.2)
    if case == "partial_overlap":
        (ra, rgt), (sa, sgt) = _crop(p, 'ref', 40, 900, 25, 1000), _crop(p, 'src', 0, 801, 131, 1024)
        sh = (0.0, 0.0)
    elif case == "src_inside_ref":
        (ra, rgt), (sa, sgt) = _crop(p, 'ref', 0, 1024, 0, 1024), _crop(p, 'src', 200, 777, 150, 905)
        sh = (7.3, -11.9)
    else:
        (ra, rgt), (sa, sgt) = _crop(p, 'ref', 100, 600, 0, 1024), _crop(p, 'src', 64, 700, 301, 1001)
        sh = (-44.1, 17.6)
    sgt = _shifted_gt(sgt, *sh)
    res, grid, bufs = engine.nuth_iteration(Raster(ra, rgt, p['ndv']), Raster(sa, sgt, p['ndv']), MaskSource(p['mask'], p['gt']))
    det = {}
    odx, ody, odz, omask, _ = oda.compute_offset(pg.MemDataset(ra, rgt, p['ndv']), pg.MemDataset(sa, sgt, p['ndv']),
                                                 mask_source=oda.MaskSource(p['mask'], p['gt']), details=det)
    assert (grid.height, grid.width) == omask.shape
    assert grid.height != grid.width
    gmask = ((bufs.maskbits & _lib.M_DIFF) != 0).cpu().numpy()
    assert np.array_equal(gmask, omask)
    _compare_iteration(res, det, odx, ody, odz)


def test_exits_and_degenerate_inputs(cuda):
    """The reference's sys.exit paths (dem_align.py:88-89, coreglib.py:206-209) and library errors."""
    from demcoreg_b200 import engine, dem_align, _lib
    from demcoreg_b200.raster import Raster, MaskSource
    p = _pair(n=256, res=30.0)
    ref, src = _gds(p, 'ref'), _gds(p, 'src')
    # 1. disjoint extents: no intersection (library error -2, as memwarp_multi would fail)
    far = Raster(p['src'], _shifted_gt(p['gt'], 1e6, 0.0), p['ndv'])
    with pytest.raises(_lib.DcrError, match="no intersection"):
        engine.nuth_iteration(ref, far)
    # 2. everything masked by the static mask -> "No overlapping, unmasked pixels shared between input DEMs"
    allm = MaskSource(np.ones((256, 256), np.uint8), p['gt'])
    res, _, _ = engine.nuth_iteration(ref, src, allm)
    assert res.status == 1 and res.count_base == 0
    with pytest.raises(SystemExit, match="No overlapping"):
        dem_align.compute_offset(ref, src, None, mask_list=[], mask_source=allm)
    with pytest.raises(SystemExit, match="No overlapping"):
        dem_align.align(ref, src, mask_source=allm, verbose=False)          # native loop: same exit
    # 3. fewer than 100 samples left -> "Not enough dh samples"
    few = np.ones((256, 256), np.uint8); few[100:108, 100:110] = 0          # 80 unmasked pixels
    res, _, _ = engine.nuth_iteration(ref, src, MaskSource(few, p['gt']))
    assert res.status == 2 and 0 < res.count_final < 100
    with pytest.raises(SystemExit, match="Not enough dh samples"):
        dem_align.align(ref, src, mask_source=MaskSource(few, p['gt']), verbose=False)
    # 4. a raster smaller than one kernel tile (40 x 50): runs, too few bins for a fit -> dx = dy = 0 (status 4)
    from oracle import dem_align as oda, pygeotools_restated as pg
    (ra, rgt), (sa, sgt) = _crop(p, 'ref', 10, 50, 20, 70), _crop(p, 'src', 10, 50, 20, 70)
    res, grid, bufs = engine.nuth_iteration(Raster(ra, rgt, p['ndv']), Raster(sa, sgt, p['ndv']))
    det = {}
    odx, ody, odz, omask, _ = oda.compute_offset(pg.MemDataset(ra, rgt, p['ndv']), pg.MemDataset(sa, sgt, p['ndv']), details=det)
    assert (grid.height, grid.width) == (40, 50)
    assert np.array_equal(((bufs.maskbits & _lib.M_DIFF) != 0).cpu().numpy(), omask)
    assert res.count_final == det['count_final'] and res.dz_med == det['dz_med']
    assert (res.status == 4) == (det.get('fit') is None)
    assert abs(res.dx - odx) < OFFTOL and abs(res.dy - ody) < OFFTOL


def _tilted_pair(n=512, res=30.0):
    p = _pair(n=n, res=res, nodata_frac=0.01)
    jj, ii = np.meshgrid(np.arange(n), np.arange(n))
    tilt = (0.8 + 2.0 * (jj / n - 0.5) - 1.2 * (ii / n - 0.5)).astype(np.float32)
    src = p['src'].copy()
    v = src != p['ndv']
    src[v] += tilt[v]
    p['src'] = src
    return p


@pytest.mark.parametrize("order", [1, 2])
def test_polyfit2d_surface(cuda, order):
    """dcr_polyfit2d / dcr_polyval2d against least squares in float64: the oracle's geolib.ma_fitpoly restatement for
    order 1 (raw map coordinates), a centred-coordinate lstsq for order 2 (where raw coordinates are rank-deficient in
    float64 and np.linalg.lstsq returns its regularised answer).  See the note in oracle ma_fitpoly."""
    from demcoreg_b200 import tiltcorr as tc
    from oracle import pygeotools_restated as pg
    rng = np.random.default_rng(5 + order)
    h, w = 300, 420
    gt = (512345.0, 30.0, 0.0, 4101234.0, 0.0, -30.0)
    pX, pY = np.meshgrid(np.arange(w), np.arange(h))
    x, y = pg.pixel_to_map(pX, pY, gt)
    u, v = (x - x.mean()) / 6000.0, (y - y.mean()) / 6000.0
    z = 0.3 + 1.5 * u - 0.7 * v + 0.4 * u * v + (0.25 * u * u - 0.1 * v * v if order == 2 else 0.0)
    z = (z + 0.05 * rng.standard_normal(z.shape)).astype(np.float32)
    m = (rng.random(z.shape) < 0.3).astype(np.uint8) * 4
    zt, mt = cuda.from_numpy(z).cuda(), cuda.from_numpy(m).cuda()
    model, cnt = tc.polyfit2d(zt, mt, 4, gt, order)
    assert cnt == int((m == 0).sum())
    vals = tc.polyval2d(model, gt, h, w).cpu().numpy().astype(np.float64)
    bma = np.ma.array(z, mask=m != 0)
    ovals, _, omodel = pg.ma_fitpoly(bma, order=order, gt=gt, perc=(0, 100), origmask=False)
    np.testing.assert_allclose([model.x0, model.y0, model.xs, model.ys], [omodel.x0, omodel.y0, omodel.xs, omodel.ys], rtol=1e-15)
    np.testing.assert_allclose([model.coeff_norm[k] for k in range((order + 1) ** 2)], omodel.coeff_norm, rtol=0, atol=1e-8)
    if order == 1:
        # the raw-basis expansion (json 'coeff') describes the same surface
        mine = pg.polyval2d(x, y, np.array(tc.coeff_list(model)))
        assert np.abs(mine - vals).max() < 1e-3
    assert np.abs(vals - ovals).max() < 1e-4, np.abs(vals - ovals).max()
    # removal: diff -= vals in float64, rounded once
    d = zt.clone()
    tc.subtract_polyval(d, gt, model)
    assert np.abs(d.cpu().numpy().astype(np.float64) - (z.astype(np.float64) - ovals)).max() < 2e-4


def test_tiltcorr_align_matches_oracle(cuda):
    """dem_align -tiltcorr (dem_align.py:396-447): loop, ONE plane fit of the filtered residuals, removal, loop again."""
    from demcoreg_b200 import dem_align
    from oracle import dem_align as oda
    p = _tilted_pair()
    out = dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), verbose=False, tiltcorr=True, polyorder=1)
    oo = oda.align(_ods(p, 'ref'), _ods(p, 'src'), final_stats=False, tiltcorr=True, polyorder=1)
    assert out['n_iter'] == oo['n_iter']
    for k in ('dx', 'dy', 'dz'):
        assert abs(out[k] - oo[k]) < OFFTOL, (k, out[k], oo[k])
    gv = out['tiltcorr']['vals'].cpu().numpy().astype(np.float64)
    ov = np.asarray(oo['tiltcorr']['vals'], dtype=np.float64)
    assert gv.shape == ov.shape
    assert np.abs(gv - ov).max() < 1e-3
    assert np.ptp(ov) > 2.5                                   # the injected +-1.6 m tilt was found ...
    a, b = out['src_align'].numpy(), oo['src_align'].array
    valid = b != np.float32(p['ndv'])
    assert np.array_equal(a != np.float32(p['ndv']), valid)
    assert np.abs(a[valid].astype(np.float64) - b[valid]).max() < 2e-3   # ... and removed identically (f32 elevations ~1e3 m)
    for k in ('count', 'min', 'max', 'mean', 'std', 'med'):
        assert abs(out['tiltcorr']['val_stats'][k] - oo['tiltcorr']['val_stats'][k]) <= 1e-3 * max(1.0, abs(oo['tiltcorr']['val_stats'][k]))
    # without the correction the residual tilt stays in the difference map
    plain = dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), verbose=False)
    d0 = dem_align.diff_stats(_gds(p, 'ref'), plain['src_align'])['stats_filt']['std']
    d1 = dem_align.diff_stats(_gds(p, 'ref'), out['src_align'])['stats_filt']['std']
    assert d1 < 0.7 * d0


def test_cli_tiltcorr(cuda, tmp_path):
    import json
    from demcoreg_b200 import dem_align, rasterio_lite as rio
    p = _tilted_pair(n=384)
    rfn, sfn = str(tmp_path / "ref.tif"), str(tmp_path / "src.tif")
    rio.write_raster(rfn, p['ref'], p['gt'], p['ndv'])
    rio.write_raster(sfn, p['src'], p['gt'], p['ndv'])
    info = dem_align.main([rfn, sfn, "-tiltcorr", "-outdir", str(tmp_path / "out")])
    assert "out_tiltcorr" in info['align_fn']
    js = json.load(open(info['align_fn'].replace('_align.tif', '_align_stats.json')))
    assert len(js['tiltcorr']['coeff']) == 4 and js['tiltcorr']['val_stats']['count'] > 0
    assert js['after_filt']['std'] < js['before_filt']['std']


@pytest.mark.parametrize("extent", ["union", "first", "last"])
def test_compute_diff_extents(cuda, extent, tmp_path):
    """compute_diff -te union / first / last (memwarp_multi extents beyond 'intersection'): output pixels outside a raster
    are nodata, the rest is the same cubic warp -- bit-identical to the oracle."""
    from demcoreg_b200 import compute_diff, rasterio_lite as rio
    from demcoreg_b200.raster import Raster
    from oracle import pygeotools_restated as pg
    p = _pair(n=512, nodata_frac=0.02)
    (ra, rgt), (sa, sgt) = _crop(p, 'ref', 40, 400, 25, 480), _crop(p, 'src', 100, 512, 131, 500)
    sgt = _shifted_gt(sgt, 47.0, -13.0)
    diff, gt = compute_diff.compute_diff(Raster(ra, rgt, p['ndv']), Raster(sa, sgt, p['ndv']), extent=extent)
    rc, sc = pg.memwarp_multi([pg.MemDataset(ra, rgt, p['ndv']), pg.MemDataset(sa, sgt, p['ndv'])], extent=extent)
    want = pg.ds_getma(sc) - pg.ds_getma(rc)
    assert diff.shape == want.shape and diff.count() > 0
    assert np.array_equal(np.ma.getmaskarray(diff), np.ma.getmaskarray(want))
    assert np.array_equal(diff.compressed(), want.compressed())
    np.testing.assert_allclose(gt, sc.gt if hasattr(sc, 'gt') else sc.GetGeoTransform(), rtol=0, atol=1e-9)
    if extent == "union":   # the CLI path
        f1, f2 = str(tmp_path / "a.tif"), str(tmp_path / "b.tif")
        rio.write_raster(f1, ra, rgt, p['ndv'])
        rio.write_raster(f2, sa, sgt, p['ndv'])
        dst = compute_diff.main([f1, f2, "-te", "union", "-outdir", str(tmp_path)])
        d = rio.read_array(dst)
        assert d['array'].shape == want.shape
        assert np.array_equal(d['array'] == -9999.0, np.ma.getmaskarray(want))


def test_iteration_without_outlier_removal(cuda):
    """compute_offset(remove_outliers=False) (dem_align.py:91): no max_dz / MAD filter on dh and no 3-sigma trim of y."""
    from demcoreg_b200 import engine, _lib
    from oracle import dem_align as oda
    p = _pair(n=768, res=30.0, nodata_frac=0.01, outlier_frac=0.01)
    gt_s = _shifted_gt(p['gt'], 3.4, -1.8)
    res, grid, bufs = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), None, remove_outliers=False)
    det = {}
    odx, ody, odz, omask, _ = oda.compute_offset(_ods(p, 'ref'), _ods(p, 'src', gt_s), remove_outliers=False, details=det)
    gmask = ((bufs.maskbits & _lib.M_DIFF) != 0).cpu().numpy()
    assert np.array_equal(gmask, omask)
    assert res.count_mad == res.count_maxdz == res.count_base      # nothing filtered
    _compare_iteration(res, det, odx, ody, odz)
    # and it differs from the filtered run (the 1 % injected outliers are kept)
    res2, _, _ = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), None, remove_outliers=True)
    assert res2.count_final < res.count_final


@pytest.mark.parametrize("case", ["max_dz_5", "slope_lim_1_20", "mask_other_grid", "different_ndv"])
def test_iteration_parameter_corners(cuda, case):
    from demcoreg_b200 import engine, _lib
    from demcoreg_b200.raster import Raster, MaskSource
    from oracle import dem_align as oda, pygeotools_restated as pg
    p = _pair(n=768, res=30.0, nodata_frac=0.01, static_mask_frac=0.25, outlier_frac=0.01)
    gt_s = _shifted_gt(p['gt'], -7.7, 4.4)
    kw_g, kw_o = {}, {}
    ref_g, src_g = _gds(p, 'ref'), _gds(p, 'src', gt_s)
    ref_o, src_o = _ods(p, 'ref'), _ods(p, 'src', gt_s)
    ms_g, ms_o = MaskSource(p['mask'], p['gt']), oda.MaskSource(p['mask'], p['gt'])
    if case == "max_dz_5":
        kw_g = dict(max_dz=5); kw_o = dict(max_dz=5)
    elif case == "slope_lim_1_20":
        kw_g = dict(slope_lim=(1.0, 20.0)); kw_o = dict(slope_lim=(1.0, 20.0))
    elif case == "mask_other_grid":
        # the world-fixed mask raster covers another extent, offset by a non-integer number of pixels
        mk, mgt = _crop(p, 'mask', 30, 700, 55, 768)
        mgt = _shifted_gt(mgt, 11.0, -4.0)
        ms_g, ms_o = MaskSource(mk, mgt), oda.MaskSource(mk, mgt)
    else:
        ref2 = p['ref'].copy(); ref2[ref2 == p['ndv']] = np.float32(-32768.0)
        ref_g, ref_o = Raster(ref2, p['gt'], -32768.0), pg.MemDataset(ref2, p['gt'], -32768.0)
    res, grid, bufs = engine.nuth_iteration(ref_g, src_g, ms_g, **kw_g)
    det = {}
    odx, ody, odz, omask, _ = oda.compute_offset(ref_o, src_o, mask_source=ms_o, details=det, **kw_o)
    gmask = ((bufs.maskbits & _lib.M_DIFF) != 0).cpu().numpy()
    assert np.array_equal(gmask, omask)
    _compare_iteration(res, det, odx, ody, odz)


def test_max_offset_exit(cuda):
    from demcoreg_b200 import dem_align
    p = _pair(n=512, res=30.0)
    with pytest.raises(SystemExit, match="Total offset exceeded specified max_offset"):
        dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), max_offset=1.0, verbose=False)
    with pytest.raises(SystemExit, match="Total offset exceeded specified max_offset"):
        dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), max_offset=1.0, verbose=False, records=[])
    out = dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), max_iter=2, tol=0.0, verbose=False)
    assert out['n_iter'] == 2


def test_stream_pairs_prefetch(cuda):
    """batch.stream_pairs: double-buffered uploads give the same iteration results as direct uploads, pair by pair."""
    from demcoreg_b200 import batch, engine
    from demcoreg_b200.raster import MaskSource
    pairs, want = [], []
    for k in range(3):
        p = _pair(n=384, res=30.0, seed=20260010 + 7 * k, static_mask_frac=0.2)
        gt_s = _shifted_gt(p['gt'], 2.0 * k - 1.5, 0.7 * k)
        pairs.append(dict(ref=p['ref'], src=p['src'], mask=p['mask'], ref_gt=p['gt'], src_gt=gt_s, mask_gt=p['gt'], ndv=p['ndv']))
        r, _, _ = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), MaskSource(p['mask'], p['gt']))
        want.append((r.dx, r.dy, r.dz, r.count_final))
    got = []
    for ref, src, ms in batch.stream_pairs(pairs):
        r, _, _ = engine.nuth_iteration(ref, src, ms)
        got.append((r.dx, r.dy, r.dz, r.count_final))
    assert got == want


def test_gpu_against_golden_horn_and_tiltcorr_loop(cuda):
    """The CUDA path against the second committed fixture: standalone Horn kernel, loop with -tiltcorr."""
    import os
    from demcoreg_b200 import engine, dem_align
    from demcoreg_b200.raster import Raster
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "horn_loop_96.npz"))
    gt = tuple(g['gt'])
    s, a = engine.horn(Raster(g['src'], gt, -9999.0), True, True)
    _check_raster(s.cpu().numpy(), g['slope'], -9999.0, "slope")
    _check_raster(a.cpu().numpy(), g['aspect'], -9999.0, "aspect")
    assert np.array_equal(np.floor(a.cpu().numpy()), np.floor(g['aspect']))
    out = dem_align.align(Raster(g['ref'], gt, -9999.0), Raster(g['src_tilted'], gt, -9999.0), verbose=False, tiltcorr=True)
    assert out['n_iter'] == int(g['loop_n_iter'])
    np.testing.assert_allclose([out['dx'], out['dy'], out['dz']], g['loop_shift'], rtol=0, atol=OFFTOL)
    m = out['tiltcorr']['model']
    np.testing.assert_allclose([m.coeff_norm[k] for k in range(4)], g['tilt_coeff_norm'], rtol=0, atol=1e-4)
    got, want = out['src_align'].numpy(), g['src_align']
    valid = want != np.float32(-9999.0)
    assert np.array_equal(got != np.float32(-9999.0), valid)
    assert np.abs(got[valid].astype(np.float64) - want[valid]).max() < 2e-3


@pytest.mark.parametrize("case", ["same_lattice", "fractional_lattice"])
def test_align_loop_ragged_inputs(cuda, case):
    """The whole loop on rasters of different extents: like the reference (dem_align.py:294-299) both are first clipped --
    or, when their pixel lattices differ by a fraction of a pixel, cubic-resampled -- onto their intersection."""
    from demcoreg_b200 import dem_align
    from demcoreg_b200.raster import Raster
    from oracle import dem_align as oda, pygeotools_restated as pg
    p = _pair(n=768, res=30.0, nodata_frac=0.01)
    (ra, rgt), (sa, sgt) = _crop(p, 'ref', 30, 700, 20, 768), _crop(p, 'src', 0, 640, 100, 740)
    if case == "fractional_lattice":
        sgt = _shifted_gt(sgt, 11.3, -7.9)       # the source grid sits between the reference's pixels
    out = dem_align.align(Raster(ra, rgt, p['ndv']), Raster(sa, sgt, p['ndv']), verbose=False)
    oo = oda.align(pg.MemDataset(ra, rgt, p['ndv']), pg.MemDataset(sa, sgt, p['ndv']), final_stats=False)
    assert out['n_iter'] == oo['n_iter']
    for k in ('dx', 'dy', 'dz'):
        assert abs(out[k] - oo[k]) < OFFTOL, (k, out[k], oo[k])
    assert out['src_align'].numpy().shape == oo['src_align'].array.shape
    assert np.array_equal(out['src_align'].numpy(), oo['src_align'].array)
    np.testing.assert_allclose(out['src_align'].gt, oo['src_align'].gt, rtol=0, atol=OFFTOL)   # origin + cumulative shift
    assert np.array_equal(out['ref'].numpy(), oo['ref'].array)


@pytest.mark.parametrize("shape", [(333, 517), (768, 768), (57, 1031)])
def test_guard_zones_stay_intact(cuda, shape, monkeypatch):
    """Out-of-bounds writes (compute-sanitizer is not available on the GPU pool, so the check is our own): every output
    raster and the workspace of an iteration sit between 64 KB guard zones filled with a pattern; after a one-pass and a
    radix iteration -- ragged sizes, partial tiles, sub-pixel shift, nodata, static mask -- the guards are untouched and the
    results equal those of the ordinary allocation."""
    from demcoreg_b200 import engine
    from demcoreg_b200.raster import MaskSource
    h, w = shape
    n = max(h, w)
    p = _pair(n=((n + 127) // 128) * 128, res=30.0, nodata_frac=0.01, static_mask_frac=0.2)
    q = {k: (np.ascontiguousarray(v[:h, :w]) if isinstance(v, np.ndarray) else v) for k, v in p.items()}
    gt_s = _shifted_gt(q['gt'], 13.7, -21.3)
    ms = MaskSource(q['mask'], q['gt'])
    want, _, wb = engine.nuth_iteration(_gds(q, 'ref'), _gds(q, 'src', gt_s), ms)
    G = 65536
    PAT = 0xA5

    class Arena(object):
        def __init__(self):
            self.guards = []

        def alloc(self, nbytes):
            nb = (int(nbytes) + 255) // 256 * 256
            t = cuda.full((G + nb + G,), PAT, dtype=cuda.uint8, device="cuda")
            self.guards.append((t, nb))
            return t[G:G + nb]

        def check(self):
            for t, nb in self.guards:
                assert bool((t[:G] == PAT).all()) and bool((t[G + nb:] == PAT).all()), "guard zone overwritten"
    ar = Arena()
    gh, gw = wb.h, wb.w
    bufs = engine.IterBuffers.__new__(engine.IterBuffers)
    bufs.h, bufs.w = gh, gw
    bufs.dh = ar.alloc(gh * gw * 4).view(cuda.float32)[:gh * gw].view(gh, gw)
    bufs.slope = ar.alloc(gh * gw * 4).view(cuda.float32)[:gh * gw].view(gh, gw)
    bufs.aspect = ar.alloc(gh * gw * 4).view(cuda.float32)[:gh * gw].view(gh, gw)
    bufs.maskbits = ar.alloc(gh * gw)[:gh * gw].view(gh, gw)
    bufs.ref_w = bufs.src_w = None
    monkeypatch.setattr(engine, "workspace", lambda nbytes, device=None: ar.alloc(nbytes))
    for radix in (False, True):
        res, _, _ = engine.nuth_iteration(_gds(q, 'ref'), _gds(q, 'src', gt_s), ms, bufs=bufs, radix_only=radix)
        cuda.cuda.synchronize()
        ar.check()
        assert (res.dx, res.dy, res.dz, res.count_final) == (want.dx, want.dy, want.dz, want.count_final) or radix
        assert res.count_final == want.count_final and list(res.nuth.bin_count) == list(want.nuth.bin_count)
        assert np.array_equal(bufs.maskbits.cpu().numpy(), wb.maskbits.cpu().numpy())
        assert np.array_equal(bufs.dh.cpu().numpy(), wb.dh.cpu().numpy())


@pytest.mark.parametrize("res", ["max", "min", "mean"])
def test_memwarp_multi_unequal_resolution_matches_oracle(cuda, res):
    """f3: the general GDAL cubic (dcr_resample_cubic) through engine.memwarp_multi on a 2:1 pair with nodata holes and a
    source origin off the reference lattice: same grid as the oracle, same untouched pass-through, identical rasters
    (down-sampled footprint for res='max', 4-sample up-sampling for 'min', both for 'mean')."""
    from demcoreg_b200 import engine, synth
    from demcoreg_b200.raster import Raster
    from oracle import pygeotools_restated as pg
    p = synth.make_pair(n=256, res=8.0, seed=12, nodata_frac=0.01)
    q = synth.make_pair(n=512, res=4.0, seed=12, nodata_frac=0.01)
    gt_s = _shifted_gt(q['gt'], 13.7, -21.3)
    ga = engine.memwarp_multi([Raster(p['ref'], p['gt'], p['ndv']), Raster(q['src'], gt_s, q['ndv'])], res=res)
    oa = pg.memwarp_multi([pg.MemDataset(p['ref'], p['gt'], p['ndv']), pg.MemDataset(q['src'], gt_s, q['ndv'])], res=res)
    for g, o in zip(ga, oa):
        assert g.RasterYSize == o.RasterYSize and g.RasterXSize == o.RasterXSize
        np.testing.assert_allclose(g.gt, o.GetGeoTransform(), rtol=0, atol=1e-9)
        assert (g.ndv if g.ndv is not None else None) == o.ndv
        assert np.array_equal(g.numpy(), o.array)


def test_dem_align_unequal_resolution(cuda):
    """dem_align on a pair of different pixel size (ref 8 m, src 4 m; -res max): the one-time memwarp_multi brings both to
    the 8 m intersection grid, then the loop runs as usual -- same iterations and offsets as the oracle."""
    from demcoreg_b200 import dem_align, synth
    from demcoreg_b200.raster import Raster
    from oracle import dem_align as oda, pygeotools_restated as pg
    p = synth.make_pair(n=384, res=8.0, seed=13)
    q = synth.make_pair(n=768, res=4.0, seed=13)          # (another realisation of the terrain: only the plumbing matters)
    src4 = np.repeat(np.repeat(p['src'], 2, axis=0), 2, axis=1)
    gt4 = (p['gt'][0], 4.0, 0.0, p['gt'][3], 0.0, -4.0)
    del q
    out = dem_align.align(Raster(p['ref'], p['gt'], p['ndv']), Raster(src4, gt4, p['ndv']), verbose=False, res='max')
    ra, sa = pg.memwarp_multi([pg.MemDataset(p['ref'], p['gt'], p['ndv']), pg.MemDataset(src4, gt4, p['ndv'])], res='max')
    oo = oda.align(ra, sa, final_stats=False)
    assert out['n_iter'] == oo['n_iter']
    for k in ('dx', 'dy', 'dz'):
        assert abs(out[k] - oo[k]) < OFFTOL, (k, out[k], oo[k])


def test_compute_diff_tr_and_rate(cuda, tmp_path):
    """compute_diff -tr on rasters of different pixel size and -rate from the dates in the file names
    (compute_diff.py:81-93, 132-136)."""
    from demcoreg_b200 import compute_diff, rasterio_lite as rio, synth
    from oracle import pygeotools_restated as pg
    p = synth.make_pair(n=256, res=8.0, seed=14)
    src4 = np.repeat(np.repeat(p['src'], 2, axis=0), 2, axis=1)
    gt4 = (p['gt'][0], 4.0, 0.0, p['gt'][3], 0.0, -4.0)
    f1 = str(tmp_path / "20150101_ref-DEM.npz")
    f2 = str(tmp_path / "20170101_src-DEM.npz")
    rio.write_raster(f1, p['ref'], p['gt'], p['ndv'])
    rio.write_raster(f2, src4, gt4, p['ndv'])
    dst = compute_diff.main([f1, f2, "-tr", "max", "-rate", "-outdir", str(tmp_path)])
    d = rio.read_array(dst)
    ra, sa = pg.memwarp_multi([pg.MemDataset(p['ref'], p['gt'], p['ndv']), pg.MemDataset(src4, gt4, p['ndv'])], res='max')
    want = pg.ds_getma(sa) - pg.ds_getma(ra)
    got = np.ma.masked_values(d['array'], -9999.0)
    assert np.array_equal(np.ma.getmaskarray(got), np.ma.getmaskarray(want))
    assert np.array_equal(got.compressed(), want.compressed())
    r = rio.read_array(dst.replace("_diff.", "_diff_rate."))
    years = (731 * 86400.0) / (365.25 * 86400.0)          # 2015-01-01 -> 2017-01-01
    rate = np.ma.masked_values(r['array'], -9999.0)
    np.testing.assert_allclose(rate.compressed(), (want / np.float32(years)).compressed(), rtol=1e-6)


@pytest.mark.parametrize("shape", [(257, 1031), (8, 41000), (300, 300)])
def test_axis_median_matches_numpy(cuda, shape):
    """dcr_axis_median = np.ma.median / np.ma.count along an axis: odd and even counts, duplicates, fully masked lines,
    lines longer than the shared-memory staging (global spill)."""
    from demcoreg_b200 import engine
    rng = np.random.default_rng(5)
    h, w = shape
    a = np.round(rng.normal(size=shape) * 4).astype(np.float32) / 4          # many duplicates
    m = rng.random(shape) < 0.3
    m[3, :] = True
    m[:, 7] = True
    a[5, ::2] = 1.25
    ma = np.ma.array(a, mask=m)
    for axis in (1, 0):
        med, cnt = engine.axis_median(cuda.from_numpy(a).cuda(), cuda.from_numpy(m.astype(np.uint8)).cuda(), axis)
        want = np.ma.median(ma, axis=axis)
        got = med.cpu().numpy()
        assert np.array_equal(cnt.cpu().numpy(), np.ma.count(ma, axis=axis))
        assert np.array_equal(np.isnan(got), np.ma.getmaskarray(want))
        ok = ~np.isnan(got)
        assert np.array_equal(got[ok], np.asarray(want[ok], dtype=np.float32))


@pytest.mark.parametrize("kw", [dict(), dict(first_axis=0), dict(first_axis_only=True), dict(sav_filter=True, sg_window=31)])
def test_successive_med_matches_oracle(cuda, kw):
    """coreglib.successive_med (coreglib.py:490-584) against the numpy/scipy restatement."""
    from demcoreg_b200 import coreglib
    from oracle import coreglib as ocl
    rng = np.random.default_rng(6)
    h, w = 420, 530
    yy, xx = np.mgrid[0:h, 0:w]
    a = (rng.normal(size=(h, w)) + 0.5 * np.sin(yy / 40.0) + 0.3 * np.cos(xx / 25.0)).astype(np.float32)
    m = rng.random((h, w)) < 0.2
    m[:, :30] |= rng.random((h, 30)) < 0.9          # columns with too few samples -> correction 0
    ma = np.ma.array(a, mask=m)
    got = coreglib.successive_med(ma, **kw)
    want = ocl.successive_med(ma, **kw)
    assert len(got) == len(want)
    exact = not kw.get('sav_filter')
    for g, o in zip(got, want):
        assert np.shape(g) == np.shape(o)
        assert np.array_equal(np.ma.getmaskarray(g), np.ma.getmaskarray(o))
        if exact:
            assert np.array_equal(np.ma.filled(g, 0), np.ma.filled(o, 0))
        else:
            np.testing.assert_allclose(np.ma.filled(g, 0), np.ma.filled(o, 0), rtol=0, atol=2e-6)
    assert got[0].dtype == want[0].dtype


def test_dem_mask_raster_pieces(cuda):
    """dem_mask.py:111-161, 403-409, 448-565 (raster parts): class / threshold masks and the dilated final mask."""
    from scipy import ndimage
    from demcoreg_b200 import dem_mask
    rng = np.random.default_rng(8)
    l = rng.choice(np.array([11, 12, 31, 41, 42, 43, 52, 71], dtype=np.uint8), size=(300, 411))
    for f in ('rock', 'rock+ice', 'rock+ice+water', 'not_forest', 'not_forest+not_water'):
        want = {'rock': l == 31, 'rock+ice': (l == 31) | (l == 12), 'rock+ice+water': (l == 31) | (l == 12) | (l == 11),
                'not_forest': ~((l == 41) | (l == 42) | (l == 43)),
                'not_forest+not_water': ~((l == 41) | (l == 42) | (l == 43) | (l == 11))}[f]
        assert np.array_equal(dem_mask.get_nlcd_mask(l, f), want)
    assert dem_mask.get_nlcd_mask(l, 'bogus') is None
    bg = (rng.random((300, 411)) * 100).astype(np.float32)
    assert np.array_equal(dem_mask.get_bareground_mask(bg, 60), bg > 60)
    toa = np.ma.array(rng.random((300, 411)).astype(np.float32), mask=rng.random((300, 411)) < 0.1)
    assert np.array_equal(dem_mask.get_toa_mask(toa, 0.4), ~np.ma.getmaskarray(np.ma.masked_greater(toa, 0.4)))
    valid = [dem_mask.get_nlcd_mask(l, 'not_forest'), bg > 60]
    for nit in (None, 1, 3):
        got = dem_mask.combine_masks(l.shape, valid, dilate=nit)
        newmask = ~(valid[0] & valid[1])
        if nit is not None:
            newmask = ~(ndimage.binary_dilation(~newmask, iterations=nit))
        assert np.array_equal(got, newmask)


def test_row_band_one_pass_selects_and_radix_rerun(cuda):
    """Row bands use the one-pass sample-bracket selects (sample keys, counts and histograms summed over the ranks); heavy
    ties make a bracket decline on every rank alike and the iteration reruns with the radix selects.  Both ways the integer
    results equal the single-GPU iteration's."""
    from demcoreg_b200 import engine, band
    from demcoreg_b200.raster import MaskSource
    p = _pair(n=1024, res=8.0, nodata_frac=0.01, static_mask_frac=0.3)
    for quant in (False, True):
        ref, src = p['ref'], p['src']
        if quant:   # quantised elevations: thousands of equal dh values around the median
            ref = np.where(ref == p['ndv'], p['ndv'], np.round(ref)).astype(np.float32)
            src = np.where(src == p['ndv'], p['ndv'], np.round(src)).astype(np.float32)
        q = dict(p, ref=ref, src=src)
        ms = MaskSource(p['mask'], p['gt'])
        r1, _, _ = engine.nuth_iteration(_gds(q, 'ref'), _gds(q, 'src'), ms, radix_only=True)
        vg = band.VirtualGroup(ref, src, p['mask'], p['gt'], p['ndv'], 3, halo=8)
        res = vg.iteration()
        for r in res:
            assert r.status == r1.status
            for f in ("count_base", "count_maxdz", "count_mad", "count_final", "count_slope", "med1", "nmad1", "dz_med", "med_dh",
                      "med_slope"):
                assert getattr(r, f) == getattr(r1, f), (quant, f, getattr(r, f), getattr(r1, f))
            assert list(r.nuth.bin_count) == list(r1.nuth.bin_count)
        if not quant:
            assert res[0].select_engine == 1, "the one-pass selects should have run on the bands"
        print("band selects: quantised=%s engine=%d" % (quant, res[0].select_engine))


def test_row_band_halo_grows_with_the_shift(cuda):
    """A cumulative y shift larger than the halo: the bands are re-laid with a deeper halo and the new rows fetched from the
    neighbours (round 1 returned -5 "re-exchange halos" and nothing re-exchanged); results equal the single-GPU iteration."""
    from demcoreg_b200 import engine, band
    from demcoreg_b200.raster import MaskSource
    p = _pair(n=1024, res=8.0, nodata_frac=0.01, static_mask_frac=0.3)
    gt_s = _shifted_gt(p['gt'], 13.7, -101.3)      # 12.7 rows: beyond the 8-row halo
    ms = MaskSource(p['mask'], p['gt'])
    r1, grid, bufs = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), ms, radix_only=True)
    vg = band.VirtualGroup(p['ref'], p['src'], p['mask'], p['gt'], p['ndv'], 3, halo=8)
    for al in vg.ranks:
        al.src.gt = gt_s
    res = vg.iteration()
    assert all(al.halo_regrown == 1 and al.halo >= 16 for al in vg.ranks)
    for r in res:
        for f in ("count_base", "count_final", "count_slope", "med1", "nmad1", "dz_med", "med_dh", "med_slope"):
            assert getattr(r, f) == getattr(r1, f), (f, getattr(r, f), getattr(r1, f))
        assert list(r.nuth.bin_count) == list(r1.nuth.bin_count)
        assert abs(r.dx - r1.dx) < 1e-9 and abs(r.dy - r1.dy) < 1e-9
    rows = 0
    for al in vg.ranks:
        rb, re = al.args.row_begin, al.args.row_end
        assert np.array_equal(al.bufs[0].cpu().numpy(), bufs.dh[rb:re].cpu().numpy())
        rows += re - rb
    assert rows == grid.height


def test_golden_resample_and_successive_med(cuda):
    """The CUDA path against the third regression fixture: unequal-resolution memwarp_multi (res = max and min) and
    coreglib.successive_med, bit for bit."""
    import os
    from demcoreg_b200 import coreglib, engine
    from demcoreg_b200.raster import Raster
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "resample_successive_64.npz"))

    def pair():
        return [Raster(z['ref'], tuple(z['gt_ref']), -9999.0), Raster(z['src'], tuple(z['gt_src']), -9999.0)]
    mx = engine.memwarp_multi(pair(), res='max')
    mn = engine.memwarp_multi(pair(), res='min')
    assert np.array_equal(mx[0].numpy(), z['max_ref']) and np.array_equal(mx[1].numpy(), z['max_src'])
    assert np.array_equal(mn[0].numpy(), z['min_ref']) and np.array_equal(mn[1].numpy(), z['min_src'])
    np.testing.assert_allclose(mx[1].gt, z['max_gt'], rtol=0, atol=1e-9)
    sm = coreglib.successive_med(np.ma.array(z['sm_in'], mask=z['sm_mask']), min_axes_count=40)
    assert np.array_equal(sm[0].filled(0).astype(np.float32), z['sm_out'])
    assert np.array_equal(np.ma.filled(sm[1], 0).astype(np.float32), z['sm_first'])
    assert np.array_equal(np.ma.filled(sm[2], 0).astype(np.float32), z['sm_second'])
