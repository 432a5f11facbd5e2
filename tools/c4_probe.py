"""Config-4-shaped robustness/throughput probe: synthetic HLA-DRB1-like training, 10,000 samples x
800 SNPs x 100 alleles (large haplotype lists, 5,050 allele-pair cells), then prediction."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hibag_b200 import api, synth
api.set_device(0)
n_samp, n_snp, n_hla = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (10000, 800, 100)))
n_cls = int(sys.argv[4]) if len(sys.argv) > 4 else 2
coh = synth.make_cohort(n_samp, n_snp, n_hla, seed=2)
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
mtry = api.default_mtry(n_snp)
lanes = int(sys.argv[5]) if len(sys.argv) > 5 else min(n_cls, 3)
screening = (sys.argv[6] != "0") if len(sys.argv) > 6 else True
m = api.HLAModel(n_snp, coh.n_hla); m.set_training(g, coh.h1, coh.h2)
t0 = time.time()
m.train(n_cls, mtry, seed=7, per_classifier_seed=True, n_concurrent=lanes, screening=screening)
dt = time.time() - t0
st = m.train_stats()
print("screening %s lanes %d: executed %.3e of %.3e pair evaluations (%.1f %%), uncertified %d" % (
    screening, lanes, st["pair_evals"], st["pair_evals_nominal"], 100.0 * st["pair_evals"] / max(1, st["pair_evals_nominal"]),
    st["n_screen_fallback"]))
import hashlib
print("model digest", hashlib.sha1(b"".join(m.classifier(k)["freq"].tobytes() + m.classifier(k)["packed"].tobytes() for k in range(n_cls))).hexdigest())
print("alleles %d, mtry %d: %d classifiers in %.1f s (%.2f /min); %s" % (coh.n_hla, mtry, n_cls, dt, 60 * n_cls / dt,
      [(len(m.classifier(k)["snpidx"]), len(m.classifier(k)["freq"]), round(m.classifier(k)["oob_acc"], 4)) for k in range(n_cls)]))
print("pair evals %.3e, cell kernel %.0f ms -> %.3e /s; em kernel %.0f ms, host fallbacks %d, em runs %d" % (
    st["pair_evals"], st["cell_kernel_ms"], st["pair_evals"] / (st["cell_kernel_ms"] * 1e-3), st["em_kernel_ms"],
    st["n_em_host_fallback"], st["n_em"]))
new = synth.draw_more(coh, 20000, seed=3)
t0 = time.time()
r = m.predict(np.ascontiguousarray(new.geno, dtype=np.int8))
dt = time.time() - t0
acc = np.mean((np.minimum(r["h1"], r["h2"]) == np.minimum(new.h1, new.h2)) & (np.maximum(r["h1"], r["h2"]) == np.maximum(new.h1, new.h2)))
print("predict 20000 samples: %.2f s, call accuracy %.3f, posterior row sums in [%.12f, %.12f]" % (
    dt, acc, r["postprob"].sum(axis=1).min(), r["postprob"].sum(axis=1).max()))
