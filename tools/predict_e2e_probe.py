"""GPU probe: prediction of N samples with 100 classifiers of the bench cohort -- device-resident call
and host-array call (HLAModel.predict: H2D of the raw genotypes, D2H of calls, dosages and the
posterior matrix inside) for a list of environment variants; every variant must give the same bits.

  python tools/predict_e2e_probe.py 200000 "" "HIBAG_B200_PREDICT_DEDUP=0" "HIBAG_B200_PREDICT_TILE=65536"
"""
import hashlib, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from hibag_b200 import api, synth
api.set_device(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
variants = sys.argv[2:] or [""]
coh = bench.make_cohort()
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
m = api.HLAModel(bench.N_SNP, coh.n_hla); m.set_training(g, coh.h1, coh.h2)
m.train(100, bench.MTRY, seed=bench.TRAIN_SEED, per_classifier_seed=True, n_threads=48, n_concurrent=40)
big = api.HLAModel(bench.N_SNP, coh.n_hla)
for c in range(100):
    k = m.classifier(c)
    big.add_classifier(k["snpidx"], k["freq"], k["hla"], k["packed"])
host = np.ascontiguousarray(synth.draw_more(coh, n, seed=99).geno, dtype=np.int8)
dev = torch.device("cuda:0")
gd = torch.from_numpy(host).to(dev)
nc = big.n_cells
h1 = torch.empty(n, dtype=torch.int32, device=dev); h2 = torch.empty_like(h1)
mp_ = torch.empty(n, dtype=torch.float64, device=dev); mt = torch.empty_like(mp_)
ds = torch.empty((n, coh.n_hla), dtype=torch.float64, device=dev)
pp = torch.empty((n, nc), dtype=torch.float64, device=dev)
prev = None
for spec in variants:
    for k in ("HIBAG_B200_PREDICT_DEDUP", "HIBAG_B200_PREDICT_TILE"):
        os.environ.pop(k, None)
    for kv in spec.split():
        k, v = kv.split("=", 1); os.environ[k] = v
    def run():
        big.predict_device(gd.data_ptr(), n, h1.data_ptr(), h2.data_ptr(), mp_.data_ptr(), mt.data_ptr(),
                           ds.data_ptr(), pp.data_ptr(), stream=torch.cuda.current_stream().cuda_stream, sync=True)
    run()
    s0 = big.predict_stats(); torch.cuda.synchronize(); t0 = time.time()
    run()
    torch.cuda.synchronize(); dt = time.time() - t0; s1 = big.predict_stats()
    d = {k: s1[k] - s0[k] for k in s1}
    res = big.predict(host); del res
    t0 = time.time(); res = big.predict(host); dt2 = time.time() - t0
    dig = hashlib.sha1(res["postprob"].tobytes() + res["h1"].tobytes() + res["dosage"].tobytes()).hexdigest()[:12]
    dig_dev = hashlib.sha1(pp.cpu().numpy().tobytes() + h1.cpu().numpy().tobytes() + ds.cpu().numpy().tobytes()).hexdigest()[:12]
    print("%-34s resident %.3f s = %7.0f samples/s | host arrays %.3f s = %7.0f | distinct %.3f, cell kernel %.0f ms of %.0f, "
          "%.3e pair evals/s in it | digest %s %s" % (spec or "(default)", dt, n / dt, dt2, n / dt2,
          d["positions_scored"] / max(d["positions_total"], 1), d["cell_kernel_ms"], d["gpu_kernel_ms"],
          d["pair_evals"] / max(d["cell_kernel_ms"] * 1e-3, 1e-12), dig, dig_dev), flush=True)
    assert dig == dig_dev and (prev is None or prev == dig)
    prev = dig
    del res
