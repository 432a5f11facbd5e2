"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous / barriers / gathers.

The path shards along the ensemble's natural axes:
  * training   -- classifiers: rank r builds classifiers r, r+W, r+2W, ... with per-classifier seeds
                  (no data-path collective; finished classifiers are gathered to rank 0);
  * prediction -- samples: rank r scores the contiguous slice shard_range(n, r, W) (no collective),
                  or classifiers: each rank accumulates its classifiers' weighted posteriors and one
                  all-reduce (NCCL over NVLink) sums them before finalisation.
"""
import os

import numpy as np


def env_rank():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), \
        int(os.environ.get("WORLD_SIZE", "1"))


def init(backend=None):
    """Initialise torch.distributed from the torchrun environment (no-op for WORLD_SIZE=1)."""
    import torch.distributed as dist
    rank, local_rank, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29512")
        if backend is None:
            import torch
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, local_rank, world


def shard_range(n, rank, world):
    """contiguous slice [begin, end) of n items for `rank` (sizes differ by at most one)"""
    base, rem = divmod(n, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def classifier_indices(n_total, rank, world):
    """global classifier indices built by `rank`: rank, rank+world, ..."""
    return list(range(rank, n_total, world))


def barrier():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def max_over_ranks(value, device=None):
    """max of a python float over all ranks (the timing rule: slowest rank defines the step)"""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_classifiers(local, world_indices=None):
    """Gather per-rank lists of (global_index, classifier dict) on every rank, ordered by global
    index -- the model a single process would have built."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return [c for _, c in sorted(local, key=lambda x: x[0])]
    bucket = [None] * dist.get_world_size()
    dist.all_gather_object(bucket, local)
    merged = [x for part in bucket for x in part]
    return [c for _, c in sorted(merged, key=lambda x: x[0])]


def allreduce_partial(acc):
    """sum the [n_samp, n_cells+3] partial posterior buffers of all ranks in place
    (torch tensor on the rank's device; NCCL over NVLink when the backend is nccl)"""
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)
    return acc


def shutdown():
    """Tear the process group down (no-op on a single rank)."""
    import torch.distributed as td
    if td.is_available() and td.is_initialized():
        td.destroy_process_group()
