"""GPU probe: one classifier in flight (config 2) -- duration of the out-of-bag and in-bag gather launches alone."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hibag_b200 import api
api.set_device(0)
coh = bench.make_cohort()
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
m = api.HLAModel(bench.N_SNP, coh.n_hla); m.set_training(g, coh.h1, coh.h2)
m.train(1, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=0, n_concurrent=1)
s0 = m.train_stats()
m.train(2, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=1, n_concurrent=1)
s1 = m.train_stats()
d = {k: s1[k] - s0[k] for k in s1}
ib_ms, ib_n = d["gather_ib_kernel_ms"], d["gather_ib_launches"]
oob_ms, oob_n = d["gather_kernel_ms"] - ib_ms, d["gather_kernel_launches"] - ib_n
print("in-bag: %d launches, avg %.3f ms, %.3e popc/s | out-of-bag: %d launches, avg %.3f ms, %.3e popc/s | whole passes %.1f ms/classifier" % (
    ib_n, ib_ms / max(ib_n, 1), d["gather_ib_popc32"] / max(ib_ms * 1e-3, 1e-12), oob_n, oob_ms / max(oob_n, 1),
    (d["popc32_issued"] - d["gather_ib_popc32"]) / max(oob_ms * 1e-3, 1e-12), d["cell_kernel_ms"] / 2))
