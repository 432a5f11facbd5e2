"""GPU probe: config-2 training, phase breakdown for several (host threads, classifiers in flight)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hibag_b200 import api
api.set_device(0)
coh = bench.make_cohort()
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
for spec in (sys.argv[1:] or ["0:1", "0:2", "0:3", "0:4"]):
    nt, nc, dev, scr = (int(x) for x in (spec + ":1:1").split(":")[:4])
    n = max(4, 2 * nc)
    m = api.HLAModel(bench.N_SNP, coh.n_hla); m.set_training(g, coh.h1, coh.h2)
    m.train(nc, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=0, n_threads=nt, n_concurrent=nc, em_on_device=bool(dev), screening=bool(scr))
    s0 = m.train_stats(); t0 = time.time()
    m.train(n, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=nc, n_threads=nt, n_concurrent=nc, em_on_device=bool(dev), screening=bool(scr))
    dt = time.time() - t0; s1 = m.train_stats()
    d = {k: s1[k] - s0[k] for k in s1}
    print("screen %d nominal %.3e executed %.3e (%.1f%%) fallback %d | " % (scr, d["pair_evals_nominal"], d["pair_evals"], 100.0 * d["pair_evals"] / max(1, d["pair_evals_nominal"]), d["n_screen_fallback"]), end="")
    print("threads %d lanes %d devEM %d (em kernel %.0f ms, host fallbacks %d): %.3f s/classifier (%.1f /min) | per classifier: prepare %.3f em-phase %.3f score-phase %.3f | em_sum %.2f wait_sum %.2f | cell_ms %.0f launches %d -> %.3e pair-evals/s in kernel" % (
        nt, nc, dev, d["em_kernel_ms"] / n, d["n_em_host_fallback"], dt / n, 60 * n / dt, d["seconds_prepare"] / n, d["seconds_phase_oob"] / n, d["seconds_phase_ib"] / n,
        d["seconds_em"] / n, d["seconds_gpu_wait"] / n, d["cell_kernel_ms"] / n, d["cell_kernel_launches"] / n,
        d["pair_evals"] / (d["cell_kernel_ms"] * 1e-3)), flush=True)
