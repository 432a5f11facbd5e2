"""Short training run for ncu: one classifier of the benchmark workload (config 2)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hibag_b200 import api
api.set_device(0)
coh = bench.make_cohort()
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
m = api.HLAModel(bench.N_SNP, coh.n_hla); m.set_training(g, coh.h1, coh.h2)
m.train(1, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=0,
        em_on_device=(len(sys.argv) < 2 or sys.argv[1] != "host"))
print(m.train_stats())
