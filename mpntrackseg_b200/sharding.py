"""Host-side helpers for running independent frame windows on several GPUs (one process per GPU).

Windows never exchange data: each rank takes a contiguous, balanced slice of
the window list and evaluates it alone; ``torch.distributed`` is only used to agree on the timing
(max over ranks) and to add up the work counters.  The reference itself is single-process
(scripts/train.py:65, ``gpus=1``); its sliding-window loop (tracker/mpn_tracker.py:167) is what gets
partitioned here.
"""
import torch


def shard_range(num_items, rank, world_size):
    """[start, end) of the contiguous slice of ``num_items`` that ``rank`` owns; slice sizes differ by
    at most one and cover every item exactly once."""
    if not (0 <= rank < world_size):
        raise ValueError(f'rank {rank} outside world of {world_size}')
    base, extra = divmod(int(num_items), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def reduce_step_stats(elapsed_ms, counters, device=None, group=None):
    """(max over ranks of ``elapsed_ms``, sum over ranks of every entry of ``counters``).
    Works with any initialised backend (NCCL on GPUs, gloo on CPU); without an initialised process
    group it returns its inputs."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(elapsed_ms), [float(c) for c in counters]
    t = torch.tensor([float(elapsed_ms)], dtype=torch.float64, device=device)
    c = torch.tensor([float(v) for v in counters], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    dist.all_reduce(c, op=dist.ReduceOp.SUM, group=group)
    return float(t[0]), c.tolist()


def all_reduce_sum_(bucket, group=None):
    """Sums the flat gradient ``bucket`` over the ranks IN PLACE (one collective: NCCL on GPUs, gloo on CPU) and returns
    the world size, so the caller can take the mean inside its optimizer kernel (``grad_scale = 1 / world``).  Without an
    initialised process group, or alone in it, nothing happens and 1 is returned.  This is the only exchange step of the
    path: data-parallel training, pl_module.py:137-141 / scripts/train.py:76 spread over ranks."""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(bucket, op=dist.ReduceOp.SUM, group=group)
        return dist.get_world_size(group)
    return 1
