"""Multi-GPU control plane on CPU: world_size-2 gloo.  Images shard by contiguous index ranges with no
data-path collective; the only collectives are the assignment broadcast and the stats gather."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_items, ret):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from jpeg_decoder_b200 import workload
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    table = workload.broadcast_assignment(n_items, world, dist)
    lo, hi = int(table[rank, 0]), int(table[rank, 1])
    # stand-in for the per-image work: checksum of the image index range this rank owns
    checksum = float(sum(i * i for i in range(lo, hi)))
    stats = workload.gather_stats([hi - lo, checksum, float(rank)], dist)
    ret[rank] = (table.tolist(), stats.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_assignment_broadcast_and_gather_world2():
    world, n_items = 2, 1025
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29731, n_items, ret), nprocs=world, join=True)
    t0, s0 = ret[0]
    t1, s1 = ret[1]
    assert t0 == t1 and s0 == s1           # every rank sees the same table / stats
    assert t0[0][0] == 0 and t0[-1][1] == n_items and t0[0][1] == t0[1][0]   # contiguous cover
    stats = np.array(s0)
    assert stats[:, 0].sum() == n_items
    assert stats[:, 1].sum() == float(sum(i * i for i in range(n_items)))
    assert stats[:, 2].tolist() == [0.0, 1.0]


def test_shard_range_properties():
    sys.path.insert(0, ROOT)
    from jpeg_decoder_b200.workload import shard_range
    for n in (0, 1, 7, 1024, 8192, 8193):
        for world in (1, 2, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1
