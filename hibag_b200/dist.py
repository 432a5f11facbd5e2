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


def sub_model(model, rank, world):
    """the classifiers rank, rank+world, ... of `model` as a model of their own (classifier shard)"""
    from . import api
    sub = api.HLAModel(model.n_snp, model.n_hla, model.hla_allele, model.snp_id)
    for k in classifier_indices(model.num_classifiers(), rank, world):
        c = model.classifier(k)
        sub.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"])
    return sub


def predict_sample_sharded(model, geno_dev, n_total, rank, world, want_prob=True):
    """Prediction with the SAMPLES sharded over ranks (reference cluster split R/HIBAG.R:766-807):
    every rank holds the whole model and scores the slice shard_range(n_total, rank, world) of the
    device-resident int8 genotype matrix geno_dev [n_total, n_snp]. No collective; the slices are
    bit-identical to a single-GPU run. Returns (begin, end, dict of device tensors for the slice)."""
    import torch
    b, e = shard_range(n_total, rank, world)
    n = e - b
    dev = geno_dev.device
    out = dict(h1=torch.empty(n, dtype=torch.int32, device=dev), h2=torch.empty(n, dtype=torch.int32, device=dev),
               prob=torch.empty(n, dtype=torch.float64, device=dev), matching=torch.empty(n, dtype=torch.float64, device=dev))
    if want_prob:
        out["postprob"] = torch.empty((n, model.n_cells), dtype=torch.float64, device=dev)
    if n > 0:
        model.predict_device(geno_dev[b:e].data_ptr(), n, out["h1"].data_ptr(), out["h2"].data_ptr(),
                             out["prob"].data_ptr(), out["matching"].data_ptr(), 0,
                             out["postprob"].data_ptr() if want_prob else 0,
                             stream=torch.cuda.current_stream().cuda_stream, sync=True)
    return b, e, out


def predict_classifier_sharded(sub, snp_weight_dev, n_hla, geno_dev, tile=262144, want_prob=True, timing=None):
    """Prediction with the CLASSIFIERS sharded over ranks: `sub` holds this
    rank's classifiers (sub_model), snp_weight_dev the WHOLE model's per-SNP weights (int32 device
    tensor, reference _GetSNPWeights src/LibHLA.cpp:2484-2496). Per tile of samples every rank
    accumulates its classifiers' weighted posteriors, ONE all-reduce (NCCL over NVLink) sums the
    [tile, n_cells + 3] fp64 partial buffers, and every rank finalises the same reduced buffer
    (normalise, best guess, matching -- :2479-2480, :2370-2382). The classifier sum is reassociated,
    so posteriors agree with the sequential order to ~1e-15 relative (tested at 1e-10), calls differ
    only on exact ties. timing: optional dict, gets allreduce_ms / allreduce_bytes added."""
    import ctypes as C
    import torch
    from . import api
    n = geno_dev.shape[0]
    dev = geno_dev.device
    n_cells = n_hla * (n_hla + 1) // 2
    out = dict(h1=torch.empty(n, dtype=torch.int32, device=dev), h2=torch.empty(n, dtype=torch.int32, device=dev),
               prob=torch.empty(n, dtype=torch.float64, device=dev), matching=torch.empty(n, dtype=torch.float64, device=dev))
    if want_prob:
        out["postprob"] = torch.empty((n, n_cells), dtype=torch.float64, device=dev)
    acc = torch.empty((min(tile, max(n, 1)), n_cells + 3), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ar_ms, ar_bytes = 0.0, 0
    for b in range(0, n, tile):
        e = min(n, b + tile)
        a = acc[:e - b]
        a.zero_()
        sub.predict_partial_device(geno_dev[b:e].data_ptr(), e - b, snp_weight_dev.data_ptr(), a.data_ptr(),
                                   stream=st, sync=False)
        e0.record()
        allreduce_partial(a)
        e1.record()
        po = api.PredictOut(out["h1"][b:e].data_ptr(), out["h2"][b:e].data_ptr(), out["prob"][b:e].data_ptr(),
                            out["matching"][b:e].data_ptr(), None,
                            out["postprob"][b:e].data_ptr() if want_prob else None)
        rc = api.lib().hibag_b200_predict_finalize_device(n_hla, e - b, C.c_void_p(a.data_ptr()), C.byref(po),
                                                          C.c_void_p(st) if st else None, 1)
        if rc != 0:
            raise RuntimeError("hibag_b200: " + api.lib().hibag_b200_last_error().decode())
        ar_ms += e0.elapsed_time(e1)
        ar_bytes += a.numel() * 8
    if timing is not None:
        timing["allreduce_ms"] = timing.get("allreduce_ms", 0.0) + ar_ms
        timing["allreduce_bytes"] = timing.get("allreduce_bytes", 0) + ar_bytes
    return out


def shutdown():
    """Tear the process group down (no-op on a single rank)."""
    import torch.distributed as td
    if td.is_available() and td.is_initialized():
        td.destroy_process_group()
