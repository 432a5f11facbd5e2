"""GPU probe: config-2 training throughput and the in-region roofline fraction of the gather kernel
for a list of environment variants, each in its own process (the library reads its switches once).

  python tools/sweep_probe.py "LANES=24" "LANES=24 HIBAG_B200_GATHER_EXCL=1" ...
LANES / STEPS / THREADS are consumed here; everything else is exported to the child.
"""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import os, sys, time, json
import numpy as np
sys.path.insert(0, %r)
import bench
from hibag_b200 import api
api.set_device(0)
lanes = int(os.environ.get("LANES", "24")); steps = int(os.environ.get("STEPS", "3"))
nt = int(os.environ.get("THREADS", "0")) or bench.default_threads(1, lanes)
coh = bench.make_cohort()
g = np.ascontiguousarray(coh.geno, dtype=np.int8)
m = api.HLAModel(bench.N_SNP, coh.n_hla); m.set_training(g, coh.h1, coh.h2)
kw = dict(seed=bench.TRAIN_SEED, per_classifier_seed=True, n_threads=nt, n_concurrent=lanes)
m.train(lanes, bench.MTRY, first_index=0, **kw)
api.sm_time(reset=True)
import resource
ru0 = resource.getrusage(resource.RUSAGE_SELF)
s0 = m.train_stats(); t0 = time.time()
for s in range(steps):
    m.train(lanes, bench.MTRY, first_index=(1 + s) * lanes, **kw)
dt = time.time() - t0; s1 = m.train_stats()
ru1 = resource.getrusage(resource.RUSAGE_SELF)
host_cores = ((ru1.ru_utime - ru0.ru_utime) + (ru1.ru_stime - ru0.ru_stime)) / dt
acct = api.sm_time()
info = api.device_info()
sm_total = info["sm_count"] * dt * info["clock_khz"] * 1e3
d = {k: s1[k] - s0[k] for k in s1}
peak = json.load(open(bench.PEAKS_JSON))["popc32_per_s"]
n = steps * lanes
ib = d["gather_ib_popc32"] / max(d["gather_ib_kernel_ms"] * 1e-3, 1e-12) / peak
oob_ms = d["gather_kernel_ms"] - d["gather_ib_kernel_ms"]
oob = (d["popc32_issued"] - d["gather_ib_popc32"]) / max(oob_ms * 1e-3, 1e-12) / peak
import hashlib
dig = hashlib.sha1(b"".join(m.classifier(k)["freq"].tobytes() + m.classifier(k)["snpidx"].tobytes() for k in range(lanes))).hexdigest()[:12]
print(json.dumps(dict(per_min=round(60 * n / dt, 1), frac=round(d["popc32_issued"] / (d["gather_kernel_ms"] * 1e-3) / peak, 4),
      ib_frac=round(ib, 4), oob_frac=round(oob, 4), gather_ms_per_cls=round(d["gather_kernel_ms"] / n, 2),
      ib_ms_per_cls=round(d["gather_ib_kernel_ms"] / n, 2), pass_ms_per_cls=round(d["cell_kernel_ms"] / n, 2),
      em_ms_per_cls=round(d["em_kernel_ms"] / n, 1), wall_ms_per_cls=round(1e3 * dt / n, 2), lanes=lanes, digest=dig,
      em_kcyc_per_iter=round(acct["em_cta_cycles"] / max(d["em_iterations"], 1) / 1e3, 1), em_iters_per_cand=round(d["em_iterations"] / max(d["n_em"], 1), 1),
      share={k: round(v / sm_total, 3) for k, v in acct.items() if k != "em_cta_cycles" and v > 0},
      em_fallbacks=d["n_em_host_fallback"], host_cores=round(host_cores, 2))), flush=True)
''' % ROOT
for spec in sys.argv[1:]:
    env = dict(os.environ)
    for kv in spec.split():
        k, v = kv.split("=", 1)
        env[k] = v
    r = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True, timeout=900)
    last = (r.stdout.strip().splitlines() or ["<no output> " + r.stderr[-400:]])[-1]
    print("%-60s %s" % (spec, last), flush=True)
