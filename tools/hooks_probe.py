"""GPU probe: the ten-hook plugin path at config 2 (our host driver calling the hooks exactly as the
reference host does): classifiers/min and where a classifier's time goes."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hibag_b200 import api
api.set_device(0)
coh = bench.make_cohort()
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
m = api.HLAModel(bench.N_SNP, coh.n_hla); m.set_training(g, coh.h1, coh.h2)
m.train(1, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=0, use_legacy_hooks=True)
s0 = m.train_stats(); t0 = time.time()
m.train(n, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=1, use_legacy_hooks=True)
dt = time.time() - t0; s1 = m.train_stats()
d = {k: s1[k] - s0[k] for k in s1}
print("hooks: %.3f s/classifier (%.1f /min) | per classifier: prepare %.3f candidates(EM, host pool) %.3f hooks(sequential) %.3f | "
      "hook wait %.3f, gpu kernel span %.1f ms, cell kernel %.1f ms, %d oob + %d ib evals, h2d %.1f MB d2h %.2f MB" % (
          dt / n, 60 * n / dt, d["seconds_prepare"] / n, d["seconds_phase_oob"] / n, d["seconds_phase_ib"] / n,
          d["seconds_gpu_wait"] / n, d["gpu_kernel_ms"] / n, d["cell_kernel_ms"] / n, d["n_oob_evals"] / n, d["n_ib_evals"] / n,
          d["h2d_bytes"] / n / 1e6, d["d2h_bytes"] / n / 1e6), flush=True)
import hashlib
print("digest", hashlib.sha1(b"".join(m.classifier(k)["freq"].tobytes() + m.classifier(k)["snpidx"].tobytes() for k in range(n + 1))).hexdigest()[:12])
del m
