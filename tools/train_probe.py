"""GPU probe: config-2 training, phase breakdown for several host-thread counts."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hibag_b200 import api
api.set_device(0)
coh = bench.make_cohort()
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
for nt in [int(x) for x in (sys.argv[1:] or ["16", "23"])]:
    m = api.HLAModel(bench.N_SNP, coh.n_hla); m.set_training(g, coh.h1, coh.h2)
    m.train(1, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=0, n_threads=nt)
    s0 = m.train_stats(); t0 = time.time()
    m.train(4, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=1, n_threads=nt)
    dt = time.time() - t0; s1 = m.train_stats()
    d = {k: s1[k] - s0[k] for k in s1}
    print("threads %d: %.3f s/classifier | prepare %.3f oob-phase %.3f ib-phase %.3f other %.3f | em_sum %.2f wait_sum %.2f | cell_ms_sum %.0f launches %d" % (
        nt, dt / 4, d["seconds_prepare"] / 4, d["seconds_phase_oob"] / 4, d["seconds_phase_ib"] / 4,
        (dt - d["seconds_prepare"] - d["seconds_phase_oob"] - d["seconds_phase_ib"]) / 4,
        d["seconds_em"] / 4, d["seconds_gpu_wait"] / 4, d["cell_kernel_ms"] / 4, d["cell_kernel_launches"] / 4), flush=True)
