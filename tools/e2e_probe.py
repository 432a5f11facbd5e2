"""GPU probe: where the end-to-end call (host arrays -> trained model) spends its time."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hibag_b200 import api
api.set_device(0)
coh = bench.make_cohort()
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
lanes = int(sys.argv[1]) if len(sys.argv) > 1 else 24
for rep in range(3):
    t0 = time.time()
    m = api.HLAModel(bench.N_SNP, coh.n_hla)
    t1 = time.time()
    m.set_training(g, coh.h1, coh.h2)
    t2 = time.time()
    m.train(3 * lanes, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=rep * 3 * lanes,
            n_threads=2 * lanes, n_concurrent=lanes)
    t3 = time.time()
    s = m.train_stats()
    del m
    t4 = time.time()
    print("rep %d: new %.3f set_training %.3f train %.3f (inside train_model %.3f) free %.3f -> %.1f /min; prepare %.2f em-phase %.2f score-phase %.2f (sums over lanes)" % (
        rep, t1 - t0, t2 - t1, t3 - t2, s["seconds_total"], t4 - t3, 60 * 3 * lanes / (t4 - t0),
        s["seconds_prepare"], s["seconds_phase_oob"], s["seconds_phase_ib"]), flush=True)
