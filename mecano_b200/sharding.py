"""Multi-GPU partition of a batch of states.

States are independent: the batch is cut into disjoint contiguous slices, one per rank
(one process per GPU), and there is no data-path collective.  torch.distributed is used only by the benchmark
for the barrier and the max-over-ranks of the elapsed time.
"""


def slice_for_rank(n_states, rank, world_size):
    """Contiguous slice [start, stop) of rank `rank`: sizes differ by at most one, slices are disjoint and cover
    [0, n_states)."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("invalid rank / world_size")
    if n_states < 0:
        raise ValueError("n_states must be >= 0")
    base, extra = divmod(n_states, world_size)
    start = rank * base + min(rank, extra)
    stop = start + base + (1 if rank < extra else 0)
    return start, stop


def max_over_ranks(value, device=None):
    """Max of a python float over all ranks (identity when torch.distributed is not initialised)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())
