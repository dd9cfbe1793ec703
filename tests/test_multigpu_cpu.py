"""world_size-2 gloo test (CPU) of the multi-rank plumbing used by bench.py: disjoint state slices per rank and the
max-over-ranks reduction of the step time.  The data path has no collective (states are independent)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mecano_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, b = sharding.slice_for_rank(n, rank, world)
    # each rank "processes" its slice: here, marks it
    mark = torch.zeros(n, dtype=torch.int32)
    mark[a:b] = rank + 1
    gathered = [torch.zeros_like(mark) for _ in range(world)]
    dist.all_gather(gathered, mark)
    total = torch.stack(gathered).sum(0)
    assert (total > 0).all() and int((torch.stack(gathered) > 0).sum()) == n  # disjoint cover
    t = sharding.max_over_ranks(1.0 + rank)
    s = sharding.sum_over_ranks(float(b - a))
    np.save(os.path.join(out_dir, "r%d.npy" % rank), np.array([t, s, a, b]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slicing_and_reductions(tmp_path):
    world, n = 2, 1001
    mp.spawn(_worker, args=(world, _free_port(), n, str(tmp_path)), nprocs=world, join=True)
    r0, r1 = (np.load(os.path.join(str(tmp_path), "r%d.npy" % r)) for r in range(2))
    assert r0[0] == 2.0 and r1[0] == 2.0          # max over ranks
    assert r0[1] == n and r1[1] == n              # the slices add up to the batch
    assert r0[3] == r1[2]                         # contiguous
