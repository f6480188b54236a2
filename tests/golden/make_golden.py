"""Generate the regression vectors under tests/golden/.

The reference (dshean/demcoreg) ships no golden vectors and cannot be imported in this image
(pygeotools / GDAL absent), so these vectors are produced by the CPU oracle restatement
(oracle/) and pin ITS behaviour: they are regression vectors, not reference outputs
("parity unpinned").  Run from the repo root:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from demcoreg_b200 import synth  # noqa: E402
from oracle import pygeotools_restated as pg, dem_align as oda  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def nuth_iter(n=128, name="nuth_iter_128.npz"):
    p = synth.make_pair(n=n, res=30.0, seed=20260777, static_mask_frac=0.2, nodata_frac=0.02, median_slope=14.0)
    gt = p['gt']
    gt_src = (gt[0] + 3.4217, gt[1], 0.0, gt[3] - 1.8140, 0.0, gt[5])
    det = {}
    ms = oda.MaskSource(p['mask'], gt)
    dx, dy, dz, sm, _ = oda.compute_offset(pg.MemDataset(p['ref'], gt, p['ndv']), pg.MemDataset(p['src'], gt_src, p['ndv']),
                                           mask_source=ms, details=det)
    np.savez_compressed(os.path.join(HERE, name), ref=p['ref'], src=p['src'], mask=p['mask'], gt=np.array(gt),
                        gt_src=np.array(gt_src), ref_w=det['ref_w'], src_w=det['src_w'],
                        static_mask_packed=np.packbits(sm), bin_count=det['nuth']['bin_count'].astype(np.int64),
                        bin_med=det['nuth']['bin_med'], shift=np.array([dx, dy, dz], dtype=np.float64),
                        fit=np.array(det['fit'], dtype=np.float64), dz_med=np.float32(det['dz_med']),
                        counts=np.array([det['count_base'], det['count_final'], det['nuth']['n_common'],
                                         det['nuth']['n_final']], dtype=np.int64))
    print(name, "shift", dx, dy, dz, "bins>=9:", int((det['nuth']['bin_count'] >= 9).sum()))


def horn_and_loop(n=96, name="horn_loop_96.npz"):
    """Horn slope / aspect of a raster with voids, and the whole loop with -tiltcorr on a tilted pair."""
    p = synth.make_pair(n=n, res=30.0, seed=20260888, nodata_frac=0.02, median_slope=14.0)
    gt = p['gt']
    ds = pg.MemDataset(p['src'], gt, p['ndv'])
    slope = pg.gdaldem_mem_ds(ds, 'slope', returnma=True)
    aspect = pg.gdaldem_mem_ds(ds, 'aspect', returnma=True)
    jj, ii = np.meshgrid(np.arange(n), np.arange(n))
    tilt = (0.4 + 1.0 * (jj / n - 0.5) - 0.6 * (ii / n - 0.5)).astype(np.float32)
    src_t = p['src'].copy()
    v = src_t != p['ndv']
    src_t[v] += tilt[v]
    out = oda.align(pg.MemDataset(p['ref'], gt, p['ndv']), pg.MemDataset(src_t, gt, p['ndv']), final_stats=False,
                    tiltcorr=True, polyorder=1)
    np.savez_compressed(os.path.join(HERE, name), ref=p['ref'], src=p['src'], src_tilted=src_t, gt=np.array(gt),
                        slope=slope.filled(-9999.0).astype(np.float32), aspect=aspect.filled(-9999.0).astype(np.float32),
                        loop_shift=np.array([out['dx'], out['dy'], out['dz']], dtype=np.float64),
                        loop_n_iter=np.int64(out['n_iter']),
                        tilt_coeff_norm=np.asarray(out['tiltcorr']['coeff'].coeff_norm, dtype=np.float64),
                        tilt_vals=np.asarray(out['tiltcorr']['vals'], dtype=np.float64),
                        src_align=out['src_align'].array, src_align_gt=np.array(out['src_align'].gt))
    print(name, "loop", out['dx'], out['dy'], out['dz'], out['n_iter'], "tilt ptp", float(np.ptp(out['tiltcorr']['vals'])))


def resample_and_successive(name="resample_successive_64.npz"):
    """Unequal-resolution memwarp_multi (2:1 pair, nodata, source origin off the reference lattice; res = max and min) and
    successive_med of a masked difference map with row / column biases."""
    from oracle import coreglib as ocl
    p = synth.make_pair(n=64, res=8.0, seed=20260999, nodata_frac=0.02)
    q = synth.make_pair(n=128, res=4.0, seed=20260999, nodata_frac=0.02)
    gt8 = p['gt']
    gt4 = (q['gt'][0] + 5.3, 4.0, 0.0, q['gt'][3] - 9.1, 0.0, -4.0)
    a, b = pg.MemDataset(p['ref'], gt8, p['ndv']), pg.MemDataset(q['src'], gt4, q['ndv'])
    mx = pg.memwarp_multi([a, b], res='max')
    mn = pg.memwarp_multi([a, b], res='min')
    rng = np.random.default_rng(5)
    yy, xx = np.mgrid[0:90, 0:110]
    d = (rng.normal(size=(90, 110)) + 0.5 * np.sin(yy / 9.0) + 0.3 * np.cos(xx / 7.0)).astype(np.float32)
    m = rng.random((90, 110)) < 0.2
    sm = ocl.successive_med(np.ma.array(d, mask=m), min_axes_count=40)
    np.savez_compressed(os.path.join(HERE, name), ref=p['ref'], src=q['src'], gt_ref=np.array(gt8), gt_src=np.array(gt4),
                        max_ref=mx[0].array, max_src=mx[1].array, max_gt=np.array(mx[1].GetGeoTransform()),
                        min_ref=mn[0].array, min_src=mn[1].array, min_gt=np.array(mn[0].GetGeoTransform()),
                        sm_in=d, sm_mask=m, sm_out=sm[0].filled(0).astype(np.float32), sm_first=np.ma.filled(sm[1], 0).astype(np.float32),
                        sm_second=np.ma.filled(sm[2], 0).astype(np.float32))
    print(name, mx[1].array.shape, mn[0].array.shape)


if __name__ == "__main__":
    resample_and_successive()
    nuth_iter()
    horn_and_loop()
