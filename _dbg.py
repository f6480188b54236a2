import sys; sys.path.insert(0,'.')
import numpy as np, torch
from bench import make_workload
from demcoreg_b200 import engine
from demcoreg_b200.raster import Raster, MaskSource
for size in (2048, 4096, 8192):
    p = make_workload(size, 8.0)
    ref = Raster(p['ref'], p['gt'], p['ndv']); src = Raster(p['src'], p['gt'], p['ndv']); ms = MaskSource(p['mask'], p['gt'])
    r,_,_ = engine.nuth_iteration(ref, src, ms)
    print(size, 'engine', r.select_engine, r.count_final, r.dx, r.dy, r.dz, flush=True)
