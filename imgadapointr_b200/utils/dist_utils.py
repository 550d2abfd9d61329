"""Drop-in for the collective helpers of the reference's ``utils/dist_utils.py`` (:9-54) plus the batch-sharding
helpers the multi-GPU path needs.  Every cloud is independent, so the only collective is an all-reduce of tiny
sum/count vectors (NCCL on GPUs over NVLink; gloo in the CPU tests)."""
from __future__ import annotations

import os

import torch
from torch import distributed as dist


def init_dist(launcher="pytorch", backend="nccl", **kwargs):
    """utils/dist_utils.py:9-24 (uses LOCAL_RANK when present instead of rank % num_gpus)."""
    if launcher != "pytorch":
        raise ValueError(f"Invalid launcher type: {launcher}")
    rank = int(os.environ["RANK"])
    if backend == "nccl":
        local = int(os.environ.get("LOCAL_RANK", rank % max(torch.cuda.device_count(), 1)))
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        dist.init_process_group(backend=backend, **kwargs)
    return dist.get_rank(), dist.get_world_size()


def get_dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def reduce_tensor(tensor, args=None):
    """all_reduce(SUM) / world_size (utils/dist_utils.py:41-48); ``args.world_size`` is honoured when given."""
    rt = tensor.clone()
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(rt, op=dist.ReduceOp.SUM)
        world = getattr(args, "world_size", None) or dist.get_world_size()
        rt /= world
    return rt


class _Done:
    """Stand-in for a finished collective (single process)."""

    def wait(self):
        return True


def reduce_tensor_async(tensor, args=None):
    """``reduce_tensor`` off the critical path: returns ``(rt, work)``; the all-reduce runs on the collective's own
    stream behind the kernels already queued, and the launch stream only waits for it when the consumer calls
    ``work.wait()`` (a stream-side wait, not a host block) - i.e. at read-back.  ``rt`` holds the SUM until then and the
    mean (SUM / world_size, utils/dist_utils.py:41-48) after ``finish_reduce(rt, work)``."""
    rt = tensor.clone()
    if dist.is_available() and dist.is_initialized():
        return rt, dist.all_reduce(rt, op=dist.ReduceOp.SUM, async_op=True)
    return rt, _Done()


def finish_reduce(rt, work, args=None):
    """Consume a ``reduce_tensor_async`` result: wait on the launch stream, then divide by the world size."""
    work.wait()
    if dist.is_available() and dist.is_initialized():
        rt /= getattr(args, "world_size", None) or dist.get_world_size()
    return rt


def gather_tensor(tensor, args=None):
    """utils/dist_utils.py:50-54."""
    if not (dist.is_available() and dist.is_initialized()):
        return tensor.clone()
    world = getattr(args, "world_size", None) or dist.get_world_size()
    outs = [torch.empty_like(tensor) for _ in range(world)]
    dist.all_gather(outs, tensor)
    return torch.cat(outs, dim=0)


def shard_range(total: int, rank: int, world: int):
    """Contiguous, balanced [begin, end) slice of ``total`` independent clouds for ``rank`` (first total % world
    ranks get one extra)."""
    base, extra = divmod(total, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def allreduce_sums(vec: torch.Tensor) -> torch.Tensor:
    """Sum a small [sum..., count...] vector over all ranks (the path's only collective); identity when not
    distributed.  Means are formed AFTER the reduction so the global mean equals the single-GPU mean even when the
    shards are uneven (the reference averages per-rank means, identical only for equal shards)."""
    out = vec.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM)
    return out
