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
    # 3. classifier-sharded prediction, one NCCL all-reduce per tile of samples (package helpers)
    sub = hd.sub_model(full, rank, world)
    n = len(g_all)
    n_cells = full.n_cells
    g_dev = torch.from_numpy(g_all).to(dev)
    wts = torch.from_numpy(full.snp_weights()).to(dev)
    torch.cuda.synchronize(); hd.barrier()
    t0 = time.time()
    tm = {}
    res = hd.predict_classifier_sharded(sub, wts, full.n_hla, g_dev, tile=8192, timing=tm)
    torch.cuda.synchronize()
    dt = time.time() - t0
    ar_ms, ar_bytes = tm["allreduce_ms"], tm["allreduce_bytes"]
    h1, h2, pp, mt = res["h1"], res["h2"], res["postprob"], res["matching"]
    # 2b. the sample-sharded helper on device tensors: slices equal the single-rank result
    b2, e2, part2 = hd.predict_sample_sharded(full, g_dev, n, rank, world)
    ok_samp_dev = True
    ok_cls, max_rel = True, 0.0
    mine2 = {k: part2[k].cpu().numpy() for k in ("h1", "h2", "prob", "matching", "postprob")}
    g2 = [None] * world
    dist.all_gather_object(g2, (b2, e2, mine2))
    if rank == 0:
        for (bb, ee, pr) in g2:
            for key in ("h1", "h2", "prob", "matching", "postprob"):
                ok_samp_dev &= bool(np.array_equal(pr[key], ref[key][bb:ee], equal_nan=True))
        ok_samp &= ok_samp_dev
        ok_cls = bool(np.array_equal(h1.cpu().numpy(), ref["h1"]) and np.array_equal(h2.cpu().numpy(), ref["h2"]))
        a, r_ = pp.cpu().numpy(), ref["postprob"]
        den = np.maximum(np.abs(r_), 1e-300)
        max_rel = float(np.max(np.abs(a - r_) / den))
        ok_cls &= max_rel <= 1e-10
        print(json.dumps({"world": world, "nproc_host": os.cpu_count(), "train_shard_equals_single": ok_train,
                          "sample_sharded_bit_exact": ok_samp, "classifier_sharded_calls_equal": ok_cls,
                          "classifier_sharded_max_rel_err": max_rel, "allreduce_ms": ar_ms,
                          "allreduce_bytes": int(ar_bytes), "partial+reduce+finalize_s": dt}))
    hd.barrier()
    assert ok_train and ok_samp and ok_cls


if __name__ == "__main__":
    main()
