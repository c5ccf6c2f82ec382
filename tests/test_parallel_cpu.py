"""world_size-2 gloo test of the N > 1 path's host logic: clip sharding + the all-gather of scores (uneven shards,
global ordering).  The arithmetic per rank is a stand-in; on the GPU box the same code runs over NCCL."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, n_items, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kvq_b200 import parallel
    lo, hi = parallel.shard_bounds(n_items, world, rank)
    local = torch.arange(lo, hi, dtype=torch.float32) * 0.5 + 1.0          # "score" of clip i = 0.5 i + 1
    full = parallel.all_gather_scores(local, n_items)
    out[rank] = full.tolist()
    dist.destroy_process_group()


def _run(n_items, world=2):
    import sys
    from conftest import PKG
    if PKG not in sys.path:
        sys.path.insert(0, PKG)
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), n_items, out), nprocs=world, join=True)
    return out


def test_all_gather_even_and_uneven_shards():
    for n in (8, 7, 1):
        out = _run(n)
        expect = [0.5 * i + 1.0 for i in range(n)]
        assert out[0] == expect and out[1] == expect, (n, dict(out))
