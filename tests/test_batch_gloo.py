"""Pair-sharded batch logic under a real world_size-2 process group (gloo, CPU)."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_align(pair):
    # deterministic stand-in for dem_align.align: "recovers" the shift stored in the pair
    return dict(dx=pair["shift"][0], dy=pair["shift"][1], dz=-pair["shift"][2], n_iter=3 + pair["id"] % 2)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from demcoreg_b200 import batch
    pairs = [dict(id=k, shift=(0.5 * k, -0.25 * k, 0.1 * k)) for k in range(7)]
    mine = batch.shard(len(pairs), rank, world)
    out = batch.align_batch(pairs, _fake_align)
    q.put((rank, mine, out))
    dist.barrier()
    dist.destroy_process_group()


def test_pair_sharding_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert res[0][1] == [0, 2, 4, 6] and res[1][1] == [1, 3, 5]
    for rank, mine, out in res:
        assert len(out) == 7
        for k, r in enumerate(out):
            assert r == dict(dx=0.5 * k, dy=-0.25 * k, dz=-0.1 * k, n_iter=3 + k % 2)


def test_shard_single_process():
    from demcoreg_b200 import batch
    assert batch.shard(5, 0, 1) == [0, 1, 2, 3, 4]
    assert sorted(sum((batch.shard(10, r, 4) for r in range(4)), [])) == list(range(10))
    out = batch.align_batch([dict(id=0, shift=(1.0, 2.0, 3.0))], _fake_align)
    assert out == [dict(dx=1.0, dy=2.0, dz=-3.0, n_iter=3)]


def _halo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from demcoreg_b200 import band
    H, W, halo = 37, 5, 3
    full = torch.arange(H * W, dtype=torch.float32).reshape(H, W)
    r0, r1 = band.band_rows(H, rank, world)
    b, owned = band.make_band(full[r0:r1], r0, r1, H, halo, (0, 1, 0, H, 0, -1), -9999.0, torch, fill=-1.0, device="cpu")
    band.exchange_halos(b, owned, rank, world)
    lo = max(0, r0 - halo)
    hi = min(H, r1 + halo)
    q.put((rank, bool(torch.equal(b.data, full[lo:hi])), b.row0 == lo))
    dist.barrier()
    dist.destroy_process_group()


def test_halo_exchange_world3():
    """Row-band halo exchange (send/recv with both neighbours) under a real 3-rank gloo group on CPU tensors."""
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_halo_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, True), (1, True, True), (2, True, True)]


def test_band_rows_partition():
    from demcoreg_b200 import band
    for H, G in ((32768, 8), (1000, 3), (7, 4)):
        rows = [band.band_rows(H, r, G) for r in range(G)]
        assert rows[0][0] == 0 and rows[-1][1] == H
        assert all(rows[k][1] == rows[k + 1][0] for k in range(G - 1))


def test_band_halo_bookkeeping_is_consistent_between_neighbours():
    """Row bands, host logic only: what rank r sends to a neighbour is exactly what that neighbour's halo holds, for every
    split (the native exchange, dcr_band_exchange_halos, pairs one ncclSend with one ncclRecv of the same size)."""
    from demcoreg_b200 import band

    class FakeBand(object):
        def __init__(self, height, halo):
            self.height, self.halo = height, halo
    for height, world, halo in ((1000, 2, 16), (1001, 3, 8), (32768, 8, 64), (257, 4, 100), (64, 8, 3)):
        rows = [band.band_rows(height, r, world) for r in range(world)]
        assert rows[0][0] == 0 and rows[-1][1] == height and all(a[1] == b[0] for a, b in zip(rows, rows[1:]))
        b = FakeBand(height, halo)
        for r in range(world):
            r0, r1 = rows[r]
            top = r0 - max(0, r0 - halo)                    # rows this rank stores above / below its own
            bot = min(height, r1 + halo) - r1
            if r > 0:        # the neighbour above sends its bottom rows: as many as our top halo holds
                assert band._halo_of_neighbour(b, r, world, below=False) == top
                # and we send it our top rows: as many as ITS bottom halo holds
                pr0, pr1 = rows[r - 1]
                assert band._halo_of_neighbour(b, r - 1, world, below=True) == min(height, pr1 + halo) - pr1
            if r < world - 1:
                assert band._halo_of_neighbour(b, r, world, below=True) == bot
