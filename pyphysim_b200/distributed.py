"""Realization sharding over GPUs: one process per GPU (torchrun), rank g simulates a contiguous
slice of the realization indices of the current SNR point, and ONE all-reduce(sum) of the 4 int64
error counters per SNR point makes every rank build identical Results (so `_keep_going` decisions
agree).  Because the Philox stream is keyed by the global realization index, the summed counters are
identical for 1, 2, 4 or 8 GPUs.  Backend: NCCL on GPUs, gloo in the CPU tests."""
import os


def is_initialized():
    import torch.distributed as dist
    return dist.is_available() and dist.is_initialized()


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op when WORLD_SIZE <= 1)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get('WORLD_SIZE', '1'))
    if world <= 1 or is_initialized():
        return world_size()
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    if backend == 'nccl':
        local = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(local)
        dist.init_process_group(backend, device_id=torch.device('cuda', local))
    else:
        dist.init_process_group(backend)
    return world_size()


def rank():
    import torch.distributed as dist
    return dist.get_rank() if is_initialized() else 0


def world_size():
    import torch.distributed as dist
    return dist.get_world_size() if is_initialized() else 1


def shard(n_units, first_unit=0, rank_=None, world=None):
    """(first_unit, count) of this rank's contiguous slice of [first_unit, first_unit + n_units):
    rank g gets [g*R/G, (g+1)*R/G) (integer division, so every unit is covered exactly once)."""
    g = rank() if rank_ is None else rank_
    G = world_size() if world is None else world
    lo = (n_units * g) // G
    hi = (n_units * (g + 1)) // G
    return first_unit + lo, hi - lo


def allreduce_counters(counters):
    """In-place sum of a counters tensor (int64[4]) over all ranks; returns it."""
    import torch.distributed as dist
    if is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters
