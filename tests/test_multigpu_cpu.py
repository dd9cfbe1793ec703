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


def _engine_worker(rank, world, port, seed, n, out_dir):
    """One rank of the data path as bench.py / the multi-device engine run it: its slice of the batch through the kernels' own
    per-state code (compiled for the host by tests/emu), results gathered on rank 0."""
    import emu_lib as el
    import treedesc as td

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(seed)  # every rank builds the same tree and the same batch, then keeps its slice
    t = td.humanoid(rng, 2)
    q, qd, qdd, tau = td.random_states(rng, t, n)
    a, b = sharding.slice_for_rank(n, rank, world)
    e = el.Emu(t)
    sl = lambda x: np.ascontiguousarray(x[:, a:b])  # noqa: E731
    parts = {"rnea": e.rnea(sl(q), sl(qd), sl(qdd)), "aba": e.aba(sl(q), sl(qd), sl(tau)), "crba": e.crba(sl(q)).reshape(t.nv * t.nv, b - a)}
    for name, mine in parts.items():
        full = torch.zeros((mine.shape[0], n), dtype=torch.float64)
        full[:, a:b] = torch.from_numpy(np.ascontiguousarray(mine))
        dist.reduce(full, dst=0, op=dist.ReduceOp.SUM)  # disjoint slices: the sum is the concatenation, bit for bit
        if rank == 0:
            np.save(os.path.join(out_dir, name + ".npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_slices_concatenate_to_the_single_process_result(tmp_path):
    """the slices of N ranks concatenate to the bit-identical single-device result -- checked here on the CPU with
    the kernels' per-state code under gloo (on hardware: tests/test_gpu_host_path.py::test_multi_device_slices_are_bit_identical)."""
    import emu_lib as el
    import treedesc as td

    world, n, seed = 2, 37, 4242
    mp.spawn(_engine_worker, args=(world, _free_port(), seed, n, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(seed)
    t = td.humanoid(rng, 2)
    q, qd, qdd, tau = td.random_states(rng, t, n)
    e = el.Emu(t)
    whole = {"rnea": e.rnea(q, qd, qdd), "aba": e.aba(q, qd, tau), "crba": e.crba(q).reshape(t.nv * t.nv, n)}
    for name, want in whole.items():
        got = np.load(os.path.join(str(tmp_path), name + ".npy"))
        assert np.array_equal(got, want), name
