"""Run on the GPU box: describe the config-2 workload for bench.py's reference arm.

Trains classifiers 0..N-1 of the benchmark cohort (bench.py seeds) with the B200 path and writes
  gpurun_out/c2_workload.json : per classifier the pair evaluations done when each SNP was
                                accepted + the total (the reference arm extrapolates bounded
                                prefixes of the same classifiers with these counts)
  gpurun_out/c2_model.npz     : the first 6 trained classifiers (prediction benchmark model)
  gpurun_out/pipe_peaks.json  : measured POPC / FP64 / LDS issue rates (roofline denominators)
Copy them to profiles/ and tests/golden/ afterwards (gpurun only brings gpurun_out/ back).
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from hibag_b200 import api  # noqa: E402

n_cls = int(sys.argv[1]) if len(sys.argv) > 1 else 16
out_dir = os.path.join(ROOT, "gpurun_out")
os.makedirs(out_dir, exist_ok=True)
api.set_device(0)
info = api.device_info()
peaks = {}
for w, key in enumerate(["popc32_per_s", "lop3_per_s", "fp64_ops_per_s", "dfma_per_s", "lds64_per_s", "iadd3_per_s"]):
    best = 0
    for _ in range(3):
        ops, ms = api.pipe_peak(w)
        best = max(best, ops)
    peaks[key] = best
peaks["device"] = info
peaks["how"] = "hibag_b200_pipe_peak microbenchmarks (hibag_b200/csrc/kernels.cu), best of 3, lane-ops/s"
json.dump(peaks, open(os.path.join(out_dir, "pipe_peaks.json"), "w"), indent=1)
print(peaks, flush=True)

coh = bench.make_cohort()
m = api.HLAModel(bench.N_SNP, coh.n_hla)
m.set_training(np.ascontiguousarray(coh.geno, dtype=np.int8), coh.h1, coh.h2)
wl = {"n_samp": bench.N_SAMP, "n_snp": bench.N_SNP, "n_hla": coh.n_hla, "mtry": bench.MTRY,
      "train_seed": bench.TRAIN_SEED, "cohort_seed": bench.COHORT_SEED, "classifiers": []}
for k in range(n_cls):
    t0 = time.time()
    m.train(1, bench.MTRY, prune=True, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=k)
    dt = time.time() - t0
    tr = m.train_trace()
    rows = tr[tr[:, 0] == k]
    acc = [[int(r[1]), int(r[2])] for r in rows if r[1] >= 0]
    fin = rows[rows[:, 1] < 0][0]
    c = m.classifier(k)
    wl["classifiers"].append(dict(index=k, accepted_pairs=acc, total_pairs=int(fin[2]), em_runs=int(fin[3]),
                                  n_snp=len(c["snpidx"]), n_haplo=len(c["freq"]), oob_acc=c["oob_acc"],
                                  b200_seconds=round(dt, 3)))
    print(k, round(dt, 2), "s", len(c["snpidx"]), "SNPs", len(c["freq"]), "haplotypes", int(fin[2]), "pair evals",
          flush=True)
    json.dump(wl, open(os.path.join(out_dir, "c2_workload.json"), "w"))
st = m.train_stats()
print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items()})
n_keep = min(6, n_cls)
cls = [m.classifier(k) for k in range(n_keep)]
np.savez_compressed(
    os.path.join(out_dir, "c2_model.npz"), n_hla=np.int64(coh.n_hla), n_snp=np.int64(bench.N_SNP),
    snp_off=np.cumsum([0] + [len(c["snpidx"]) for c in cls]), snpidx=np.concatenate([c["snpidx"] for c in cls]),
    hap_off=np.cumsum([0] + [len(c["freq"]) for c in cls]), freq=np.concatenate([c["freq"] for c in cls]),
    hla=np.concatenate([c["hla"] for c in cls]), packed=np.concatenate([c["packed"] for c in cls]),
    oob_acc=np.array([c["oob_acc"] for c in cls]))
