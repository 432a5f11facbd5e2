"""GPU probe: prediction through the host-array entry point (HLAModel.predict: H2D of the raw
genotypes, D2H of calls, dosages and the posterior matrix inside) against the device-resident call,
100 classifiers (the golden HLA-A model's, 34... alleles of the bench cohort) x N samples."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from hibag_b200 import api, synth
api.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
coh = bench.make_cohort()
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
m = api.HLAModel(bench.N_SNP, coh.n_hla); m.set_training(g, coh.h1, coh.h2)
m.train(8, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, n_threads=16, n_concurrent=8)
big = api.HLAModel(bench.N_SNP, coh.n_hla)
for c in range(100):
    k = m.classifier(c % 8)
    big.add_classifier(k["snpidx"], k["freq"], k["hla"], k["packed"])
host = np.ascontiguousarray(synth.draw_more(coh, n, seed=99).geno, dtype=np.int8)
prev = None
for rep in range(4):
    t0 = time.time()
    res = big.predict(host, want_prob=True, want_dosage=True)
    dt = time.time() - t0
    chk = (int(res["h1"].sum()), float(res["postprob"][::997].sum()))
    print("rep %d: %.3f s -> %.0f samples/s  check %s" % (rep, dt, n / dt, chk), flush=True)
    assert prev is None or prev == chk
    prev = chk
    del res
