"""Short predict-only run for ncu: the benchmark's 100-classifier-style model (tests/golden/
c2_model.npz, classifiers cycled) on N samples of the benchmark cohort's distribution."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hibag_b200 import api, synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n_cls = int(sys.argv[2]) if len(sys.argv) > 2 else 12
api.set_device(0)
wl = np.load(os.path.join(bench.ROOT, "tests", "golden", "c2_model.npz"))
coh = bench.make_cohort()
m = api.HLAModel(bench.N_SNP, int(wl["n_hla"]))
n_src = len(wl["snp_off"]) - 1
for c in range(n_cls):
    k = c % n_src
    a, b = wl["snp_off"][k:k + 2]; q, r = wl["hap_off"][k:k + 2]
    m.add_classifier(wl["snpidx"][a:b], wl["freq"][q:r], wl["hla"][q:r], wl["packed"][q:r])
g = np.ascontiguousarray(synth.draw_more(coh, n, seed=99).geno, dtype=np.int8)
for rep in range(2):
    t0 = time.time(); s0 = m.predict_stats()
    r = m.predict(g)
    s1 = m.predict_stats(); dt = time.time() - t0
    d = {k: s1[k] - s0[k] for k in s1}
    print("rep %d: %.3f s, %.0f samples/s, cell kernel %.1f ms of %.1f ms, %.3e pair-evals/s" % (
        rep, dt, n / dt, d["cell_kernel_ms"], d["gpu_kernel_ms"], d["pair_evals"] / max(d["cell_kernel_ms"] * 1e-3, 1e-12)))
