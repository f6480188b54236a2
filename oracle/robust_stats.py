"""CPU restatement of reference ``demcoreg/robust_stats.py:47-86`` (the printed statistics, as a dict).

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).  numpy calls are the reference's own; for float32
inputs beyond 2**24 elements the percentile rank is the exact one.
"""
import numpy as np

from .pygeotools_restated import exact_percentile


def robust_stats(dz_m):
    dz_m = np.asarray(dz_m)
    dz_m_abs = np.abs(dz_m)
    out = {}
    out['count'] = dz_m.shape[0] - 1                    # robust_stats.py:55 prints n - 1
    out['rmse'] = np.sqrt(np.sum(dz_m ** 2) / dz_m.size)
    out['mean'] = np.mean(dz_m)
    out['std'] = np.std(dz_m)
    out['med'] = np.median(dz_m)
    out['p16'], out['p84'] = exact_percentile(dz_m, 15.9), exact_percentile(dz_m, 84.2)
    out['spread'] = out['p84'] - out['p16']
    out['absmed'] = np.median(dz_m_abs)
    mad = np.median(np.abs(dz_m - out['med']))
    out['nmad'] = 1.4826 * mad
    out['p68'], out['p95'] = exact_percentile(dz_m_abs, 68.3), exact_percentile(dz_m_abs, 95.0)
    return out
