"""GPU parity at the BASELINE.json configuration sizes (the sizes bench.py runs at), CUDA path vs the CPU oracle.

config 1: 2048 x 2048 at 30 m, the whole dem_align loop;
config 2: 8192 x 8192 at 8 m with a stable-surface mask and nodata, one full iteration (period = size: no tiling);
config 3: 4096 x 4096, +-16 px search window: the NCC map against scipy.signal.correlate2d, the SAD map against the
          oracle on a subset of the offsets (the oracle needs ~1.5 s per offset) and against the exact per-offset CUDA path
          on all of them.
The oracle needs about a minute per test at these sizes.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

OFFTOL = 1e-3  # metres


def _ods(p, key, gt=None):
    from oracle import pygeotools_restated as pg
    return pg.MemDataset(p[key], gt or p['gt'], p['ndv'])


def _gds(p, key, gt=None):
    from demcoreg_b200.raster import Raster
    return Raster(p[key], gt or p['gt'], p.get('ndv', -9999.0))


def test_config1_2048_loop(cuda):
    """BASELINE configs[0]: dem_align -mode nuth on a 2048^2 pair at 30 m, injected shift (+3.2, -1.7, +0.5) m."""
    from demcoreg_b200 import dem_align, synth
    from oracle import dem_align as oda
    p = synth.make_pair(n=2048, res=30.0, seed=20260010)
    out = dem_align.align(_gds(p, 'ref'), _gds(p, 'src'), verbose=False)
    oo = oda.align(_ods(p, 'ref'), _ods(p, 'src'), final_stats=False)
    assert out['n_iter'] == oo['n_iter']
    for k in ('dx', 'dy', 'dz'):
        assert abs(out[k] - oo[k]) < OFFTOL, (k, out[k], oo[k])
    assert abs(out['dx'] - 3.2) < 0.2 and abs(out['dy'] + 1.7) < 0.2 and abs(out['dz'] + 0.5) < 0.01
    assert np.array_equal(out['src_align'].numpy(), oo['src_align'].array)


def test_config2_8192_iteration(cuda):
    """BASELINE configs[1] (the bench workload): one iteration of an 8192^2 pair at 8 m, 30 % stable-surface mask, 1 % nodata,
    after a sub-pixel shift of the source grid (both rasters resampled).  Masks, every count and the 360 bin counts exact."""
    from demcoreg_b200 import engine, synth, _lib
    from demcoreg_b200.raster import MaskSource
    from oracle import dem_align as oda
    p = synth.make_pair(n=8192, res=8.0, seed=20260020, static_mask_frac=0.3, nodata_frac=0.01, mask_tile=8192)
    g = list(p['gt'])
    g[0] += 3.4
    g[3] += -1.8
    gt_s = tuple(g)
    det = {}
    odx, ody, odz, omask, _ = oda.compute_offset(_ods(p, 'ref'), _ods(p, 'src', gt_s),
                                                 mask_source=oda.MaskSource(p['mask'], p['gt']), details=det)
    nd = det['nuth']
    oc = nd['bin_count'].astype(np.int64)
    # both flavours of the slope raster: degrees (the standalone iteration) and tan(slope) (what the dem_align loop runs)
    for tan_slope in (False, True):
        res, grid, bufs = engine.nuth_iteration(_gds(p, 'ref'), _gds(p, 'src', gt_s), MaskSource(p['mask'], p['gt']),
                                                tan_slope=tan_slope)
        assert res.select_engine == 1, "the one-pass selects fell back to the radix select"
        gmask = ((bufs.maskbits & _lib.M_DIFF) != 0).cpu().numpy()
        assert np.array_equal(gmask, omask)
        assert res.count_base == det['count_base'] and res.count_final == det['count_final']
        assert res.count_slope == int(np.ma.count(det['slope']))
        assert res.stats_stride == int(np.around(det['count_final'] / 4e6))
        assert res.dz_med == det['dz_med'] and res.med_dh == nd['med_dh']
        assert abs(res.med_slope - nd['med_slope']) <= 1e-5 * nd['med_slope']
        assert res.nuth.n_common == nd['n_common']
        # 3-sigma trim of y: float32 comparisons against float32 thresholds; numpy's float32 tan is not correctly rounded, so
        # a pixel within ~1 ulp of a threshold may legitimately differ (~1 pixel in 4e7) -- report it
        gc = np.array(res.nuth.bin_count[:])
        ndiff = int(np.abs(gc - oc).sum())
        assert abs(res.nuth.n_final - nd['n_final']) <= 2 and ndiff <= 2, (res.nuth.n_final, nd['n_final'], ndiff)
        if ndiff == 0:
            assert res.nuth.n_final == nd['n_final']
        gm = np.array(res.nuth.bin_med[:], dtype=np.float64)
        np.testing.assert_allclose(gm, nd['bin_med'], rtol=1e-4, atol=1e-4)
        assert abs(res.dx - odx) < OFFTOL and abs(res.dy - ody) < OFFTOL and abs(res.dz - odz) < OFFTOL
        print("config2 parity (tan_slope=%s): bin-count differences %d, n_final %d vs %d" %
              (tan_slope, ndiff, res.nuth.n_final, nd['n_final']))
        del bufs


def _config3_pair():
    """4096^2 at 8 m, source = reference terrain displaced by (5.3, -3.6) px (Fourier shift) + noise, with the masks
    dem_align hands to the correlators (a few per cent of the pixels, dem_align.py:128-139)."""
    from demcoreg_b200 import synth
    res = 8.0
    p = synth.make_pair(n=4096, res=res, seed=20260030, shift=(5.3 * res, -3.6 * res, 0.5), noise=0.3, outlier_frac=0.0,
                        static_mask_frac=0.05, mask_tile=4096)
    m = p['mask'].astype(bool)
    return np.ma.array(p['ref'], mask=m), np.ma.array(p['src'], mask=m)


def test_config3_4096_ncc(cuda):
    """BASELINE configs[2]: compute_offset_ncc with a +-16 px window on a 4096^2 pair vs scipy.signal.correlate2d (float64)."""
    from demcoreg_b200 import coreglib
    from oracle import coreglib as ocl
    ref, src = _config3_pair()
    p = 16
    rng = np.random.default_rng(9)
    noise = (rng.standard_normal(ref.shape, dtype=np.float32),
             rng.standard_normal((ref.shape[0] - 2 * p, ref.shape[1] - 2 * p), dtype=np.float32))
    m, io, so, _ = coreglib.compute_offset_ncc(ref, src, pad=(p, p), noise=noise)
    mo, ioo, soo, _ = ocl.compute_offset_ncc(ref, src, pad=(p, p), noise=noise)
    np.testing.assert_allclose(m, mo, rtol=0, atol=2e-6 * np.abs(mo).max())
    assert tuple(io) == tuple(ioo)
    np.testing.assert_allclose(so, soo, atol=1e-4)
    # (the correlation peak of z-scored fractal terrain is broad: NCC lands within a pixel of the injected (3.6, 5.3) px
    # displacement -- exactly where scipy's correlate2d puts it; the sharper SAD minimum recovers it to 0.05 px, below)
    assert abs(so[1] - 5.3) < 1.5 and abs(so[0] - 3.6) < 1.5, so


def test_config3_4096_sad(cuda):
    """BASELINE configs[2]: compute_offset_sad, +-16 px on 4096^2: the batched kernels vs (a) the exact per-offset CUDA path on
    all 1089 offsets, (b) the oracle (numpy percentiles) on the 3x3 neighbourhood of the peak and 6 far offsets."""
    from demcoreg_b200 import coreglib, engine
    from oracle import pygeotools_restated as pg
    ref, src = _config3_pair()
    p = 16
    d1, m1 = coreglib._to_device_masked(ref)
    d2, m2 = coreglib._to_device_masked(src)
    info = {}
    mb = engine.sad_search(d1, m1, d2, m2, p, info=info)
    assert info['n_per_offset'] == 0, "batched SAD fell back to the per-offset path for %d offsets" % info['n_per_offset']
    mp = engine.sad_search(d1, m1, d2, m2, p, per_offset=True)
    np.testing.assert_allclose(mb, mp, rtol=2e-6)
    m, io, so = coreglib.compute_offset_sad((d1, m1), (d2, m2), pad=(p, p))
    pi, pj = np.unravel_index(np.argmax(m), m.shape)
    assert 0 < pi < 2 * p and 0 < pj < 2 * p
    kernel = src[p:-p, p:-p]
    offs = [(pi + a, pj + b) for a in (-1, 0, 1) for b in (-1, 0, 1)] + [(0, 0), (32, 32), (0, 32), (7, 25), (16, 16), (30, 3)]
    mo = {}
    for (i, j) in offs:
        diff = ref[i:i + kernel.shape[0], j:j + kernel.shape[1]] - kernel
        iqr = pg.calcperc(diff, (25, 75))
        diff = np.ma.masked_outside(diff, *iqr)
        mo[(i, j)] = -float(np.ma.abs(diff).sum() / diff.count())   # the reference's own expression (float32 sum)
    for (i, j), v in mo.items():
        assert abs(m[i, j] - v) <= 1e-5 * abs(v), ((i, j), m[i, j], v)
    # sub-pixel peak from the oracle's 3x3 values (parabolic, coreglib.py:455-457)
    c, cl, cr = mo[(pi, pj)], mo[(pi - 1, pj)], mo[(pi + 1, pj)]
    cd, cu = mo[(pi, pj - 1)], mo[(pi, pj + 1)]
    sp_o = (pi + (cl - cr) / (2 * cl - 4 * c + 2 * cr) - p, pj + (cd - cu) / (2 * cd - 4 * c + 2 * cu) - p)
    assert tuple(io) == (pi - p, pj - p)
    np.testing.assert_allclose(so, sp_o, atol=1e-4)
    assert abs(so[1] - 5.3) < 0.2 and abs(so[0] - 3.6) < 0.2, so
    print("config3 SAD stages (ms):", info['stage_ms'])
