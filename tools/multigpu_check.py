"""Multi-GPU check (run under torchrun, one rank per GPU, NCCL):
  1. classifier-sharded training: the gathered model equals a single-rank model;
  2. sample-sharded prediction: concatenated slices equal the single-rank result bit for bit;
  3. classifier-sharded prediction with ONE NCCL all-reduce of the partial posterior sums:
     calls equal, posteriors within 1e-10 relative of the single-rank result.
Prints one JSON line on rank 0."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hibag_b200 import api, dist as hd, synth  # noqa: E402


def cls_equal(a, b):
    return all(np.array_equal(a[k], b[k]) for k in ("snpidx", "freq", "hla", "packed"))


def main():
    rank, local_rank, world = hd.init()
    torch.cuda.set_device(local_rank)
    api.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    coh = synth.make_cohort(1200, 160, 16, seed=11)
    mtry = api.default_mtry(coh.n_snp)
    n_cls = 4 * world
    # 1. training shard
    m = api.HLAModel(coh.n_snp, coh.n_hla)
    m.set_training(coh.geno, coh.h1, coh.h2)
    mine = hd.classifier_indices(n_cls, rank, world)
    m.train(len(mine), mtry, seed=77, per_classifier_seed=True, first_index=rank, index_stride=world,
            n_concurrent=2)
    local = [(k, m.classifier(j)) for j, k in enumerate(mine)]
    merged = hd.gather_classifiers(local)
    full = api.HLAModel(coh.n_snp, coh.n_hla)
    ok_train = True
    if rank == 0:
        single = api.HLAModel(coh.n_snp, coh.n_hla)
        single.set_training(coh.geno, coh.h1, coh.h2)
        single.train(n_cls, mtry, seed=77, per_classifier_seed=True)
        ok_train = all(cls_equal(merged[k], single.classifier(k)) for k in range(n_cls))
    for c in merged:
        full.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"])
    # 2. sample-sharded prediction
    new = synth.draw_more(coh, 20000, seed=5)
    g_all = np.ascontiguousarray(new.geno, dtype=np.int8)
    b, e = hd.shard_range(len(g_all), rank, world)
    part = full.predict(g_all[b:e])
    gathered = [None] * world
    dist.all_gather_object(gathered, {k: part[k] for k in ("h1", "h2", "prob", "matching", "postprob")})
    ref = full.predict(g_all) if rank == 0 else None
    ok_samp = True
    if rank == 0:
        for key in ("h1", "h2", "prob", "matching", "postprob"):
            cat = np.concatenate([g[key] for g in gathered])
            ok_samp &= bool(np.array_equal(cat, ref[key], equal_nan=True))
    # 3. classifier-sharded prediction, one NCCL all-reduce
    sub = api.HLAModel(coh.n_snp, coh.n_hla)
    for k in hd.classifier_indices(n_cls, rank, world):
        c = merged[k]
        sub.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"])
    n = len(g_all)
    n_cells = full.n_cells
    g_dev = torch.from_numpy(g_all).to(dev)
    wts = torch.from_numpy(full.snp_weights()).to(dev)
    acc = torch.zeros((n, n_cells + 3), dtype=torch.float64, device=dev)
    torch.cuda.synchronize(); hd.barrier()
    t0 = time.time()
    sub.predict_partial_device(g_dev.data_ptr(), n, wts.data_ptr(), acc.data_ptr())
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    hd.allreduce_partial(acc)
    e1.record()
    torch.cuda.synchronize()
    ar_ms = e0.elapsed_time(e1)
    h1 = torch.zeros(n, dtype=torch.int32, device=dev); h2 = torch.zeros_like(h1)
    pp = torch.zeros((n, n_cells), dtype=torch.float64, device=dev)
    mt = torch.zeros(n, dtype=torch.float64, device=dev)
    out = api.PredictOut(h1.data_ptr(), h2.data_ptr(), None, mt.data_ptr(), None, pp.data_ptr())
    rc = api.lib().hibag_b200_predict_finalize_device(full.n_hla, n, C.c_void_p(acc.data_ptr()),
                                                      C.byref(out), None, 1)
    assert rc == 0
    dt = time.time() - t0
    ok_cls, max_rel = True, 0.0
    if rank == 0:
        ok_cls = bool(np.array_equal(h1.cpu().numpy(), ref["h1"]) and np.array_equal(h2.cpu().numpy(), ref["h2"]))
        a, r_ = pp.cpu().numpy(), ref["postprob"]
        den = np.maximum(np.abs(r_), 1e-300)
        max_rel = float(np.max(np.abs(a - r_) / den))
        ok_cls &= max_rel <= 1e-10
        print(json.dumps({"world": world, "nproc_host": os.cpu_count(), "train_shard_equals_single": ok_train,
                          "sample_sharded_bit_exact": ok_samp, "classifier_sharded_calls_equal": ok_cls,
                          "classifier_sharded_max_rel_err": max_rel, "allreduce_ms": ar_ms,
                          "allreduce_bytes": int(acc.numel() * 8), "partial+reduce+finalize_s": dt}))
    hd.barrier()
    assert ok_train and ok_samp and ok_cls


if __name__ == "__main__":
    main()
