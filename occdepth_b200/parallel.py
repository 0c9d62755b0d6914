"""Multi-GPU plumbing for the forward path: one process per GPU, frames sharded across ranks.

Frames are independent: the data path needs NO collective -- every rank runs the whole
forward on its own frames.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used only for the
barrier around a timed region, the max-over-ranks reduction of device times and for gathering predictions.
"""
import os

import torch
import torch.distributed as dist


def env_world():
    return int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))


def init(backend=None, device=None):
    """Initialise the default process group from the torchrun environment (no-op for world size 1)."""
    world, rank, local = env_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl" and device is not None:
            kw["device_id"] = device
        dist.init_process_group(backend, **kw)
    return world, rank, local


def frame_shard(n_frames, rank, world):
    """Contiguous, balanced shard [lo, hi) of n_frames for `rank` (first n_frames % world ranks get one extra)."""
    base, extra = divmod(n_frames, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def barrier():
    if dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (device times must be reduced as the MAX over ranks)."""
    if not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64,
                     device=device if device is not None else torch.device("cpu"))
    if t.is_cuda:
        t = t.float()
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(count, device=None):
    """sum of an integer over ranks (e.g. frames processed)"""
    if not dist.is_initialized():
        return int(count)
    t = torch.tensor([int(count)], dtype=torch.int64, device=device if device is not None else torch.device("cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
