"""One NCC + one SAD search (+-16 px, 4096^2) for profiling: python profiles/run_config3.py [reps]"""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from demcoreg_b200 import engine, synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
n, res, p = 4096, 8.0, 16
pr = synth.make_pair_torch(n=n, res=res, seed=20260030, shift=(5.3 * res, -3.6 * res, 0.5), noise=0.3, outlier_frac=0.0,
                           static_mask_frac=0.05, nodata_frac=0.0)
m = pr['mask']
for _ in range(reps):
    engine.ncc_corr(pr['ref'], m, pr['src'], m, p, None, None)
    info = {}
    engine.sad_search(pr['ref'], m, pr['src'], m, p, info=info)
torch.cuda.synchronize()
print(info)
