"""Multi-GPU tests (need >= 2 CUDA devices; skipped otherwise): row-band mode under real NCCL."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("native", [False, True])
def test_row_band_nccl_matches_single_gpu(cuda, native):
    """Row bands under real NCCL against the single-GPU iteration: interpreter-driven phases over torch.distributed, and the
    native path (the library's own communicator, dcr_band_exchange_halos + dcr_band_iteration)."""
    n = cuda.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29612" if native else "29611", "-m",
           "demcoreg_b200.tools.band_check", "--size", "2048", "--iters", "3"] + (["--native"] if native else [])
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["ok"] and all(i["same_as_single_gpu"] for i in d["iterations"])


def test_two_devices_in_one_process(cuda):
    """One process, two GPUs, one host thread each: every cached piece of library state (shared-memory opt-ins, SM
    count, side streams / events, pinned result block) is keyed by the calling thread's current device."""
    import threading
    import numpy as np
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from demcoreg_b200 import engine, synth
    from demcoreg_b200.raster import Raster
    p = synth.make_pair(n=640, res=30.0, seed=20260077)
    out = {}

    def work(dev):
        try:
            cuda.cuda.set_device(dev)
            ref = Raster(p['ref'], p['gt'], p['ndv'], device="cuda:%d" % dev)
            src = Raster(p['src'], p['gt'], p['ndv'], device="cuda:%d" % dev)
            for _ in range(2):
                res, grid, bufs = engine.nuth_iteration(ref, src)
            out[dev] = (res.dx, res.dy, res.dz, int(res.count_final), np.array(res.nuth.bin_count[:]))
        except Exception as e:   # surfaced by the assertions below
            out[dev] = e
    ts = [threading.Thread(target=work, args=(d,)) for d in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for d in (0, 1):
        assert not isinstance(out[d], Exception), out[d]
    assert out[0][:4] == out[1][:4] and np.array_equal(out[0][4], out[1][4])
    # and sequentially from ONE thread switching devices
    work(1)
    work(0)
    assert not isinstance(out[0], Exception) and not isinstance(out[1], Exception)
