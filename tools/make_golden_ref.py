"""Golden fixtures on the BASELINE-shaped cohorts, trained by the compiled, UNMODIFIED reference
(oracle/_ref/libhibag_ref.so).  Run in the build container (needs /root/reference to have built
oracle/_ref); the GPU tests only read the committed outputs.

  python tools/make_golden_ref.py worker c2 K TARGET OUT.npz [BUDGET_S]
      one classifier (global index K, set.seed(2024 + K)) of config 2 = bench.py's cohort
      (5,000 x 500, cohort seed 1, mtry 23, prune) under kernel target TARGET with
      verbose.detail on; writes the classifier (if it finished inside BUDGET_S), the
      accepted-SNP trace the reference prints (LibHLA.cpp:2104-2111) and the wall seconds.
  python tools/make_golden_ref.py worker c4 K TARGET OUT.npz BUDGET_S
      the same on config 4 = tools/c4_probe.py's cohort (10,000 x 800, 100 alleles drawn,
      cohort seed 2, mtry 29, train seed 7 + K); a full classifier is ~10 CPU-hours, so the
      fixture is the accepted-SNP PREFIX reached inside BUDGET_S.
  python tools/make_golden_ref.py fromlog c4 K TARGET TRACE.log OUT.npz SECONDS
      the same prefix from the captured stderr of a worker that is still running: after its last
      accepted SNP the reference scores every remaining SNP (LibHLA.cpp:2113-2119), hours at this
      scale, and reaches CheckInterrupt only at an accepted one (:2112)
  python tools/make_golden_ref.py merge c2 OUT1.npz OUT2.npz ...  -> tests/golden/c2_ref.npz
  python tools/make_golden_ref.py merge c4 OUT.npz               -> tests/golden/c4_ref_prefix.npz
  python tools/make_golden_ref.py extend c2 OUT7.npz ...         -> adds classifiers to the committed c2_ref.npz

Targets: `base` is the parity oracle; `avx2` is bit-identical to it (checked again
by merge: classifiers present under both targets must be equal); `max` only for timing.
"""
import os
import re
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (n_samp, n_snp, n_hla drawn, cohort seed, mtry, train seed base)
    "c2": (5000, 500, 40, 1, 23, 2024),
    "c4": (10000, 800, 100, 2, 29, 7),
}
LINE = re.compile(r"^\s*(\d+), SNP: (\d+), loss: (\S+), oob acc: (\S+)%, # of haplo: (\d+)")


def parse_trace(text):
    """[(position, snp index 0-based, loss string, oob acc %, n_haplo)] from verbose.detail lines"""
    out = []
    for ln in text.splitlines():
        m = LINE.match(ln)
        if m:
            out.append((int(m.group(1)), int(m.group(2)) - 1, m.group(3), m.group(4), int(m.group(5))))
    return out


def worker(cfg, k, target, out, budget):
    from oracle import refpy
    from hibag_b200 import synth
    n_samp, n_snp, n_hla, cseed, mtry, tseed = CONFIGS[cfg]
    ref = refpy.RefLib()
    info = ref.set_target(target)
    ref.set_gpu_procs(None)
    coh = synth.make_cohort(n_samp, n_snp, n_hla, seed=cseed)
    m = ref.new_model()
    m.init_training(coh.geno, coh.h1, coh.h2, coh.n_hla)
    # the reference prints through Rprintf -> stderr (oracle/ref_driver.cpp): capture fd 2
    fd, log = tempfile.mkstemp(suffix=".log")
    saved = os.dup(2)
    os.dup2(fd, 2)
    if budget > 0:
        ref.set_interrupt(seconds=budget)
    # wall-clock time at which each accepted-SNP line appears (the ctypes call releases the GIL):
    # the time profile bench.py's reference arm uses to extrapolate a bounded prefix
    import threading
    stamps, stop = [], threading.Event()

    def poll():
        seen = 0
        while not stop.is_set():
            n = len(parse_trace(open(log).read()))
            while seen < n:
                stamps.append(time.time()); seen += 1
            stop.wait(0.05)
    th = threading.Thread(target=poll, daemon=True)
    t0 = time.time()
    th.start()
    rc = m.build(1, mtry, prune=True, verbose=2, reseed_base=tseed, first_index=k, allow_interrupt=True)
    dt = time.time() - t0
    stop.set(); th.join()
    os.dup2(saved, 2)
    os.close(fd)
    log_text = open(log).read()
    trace = parse_trace(log_text)
    os.unlink(log)
    stamps = [x - t0 for x in stamps][:len(trace)]
    stamps += [dt] * (len(trace) - len(stamps))
    res = dict(config=np.array(cfg), k=np.int64(k), target=np.array(target), cpu=np.array(info),
               seconds=np.float64(dt), finished=np.bool_(rc == 0),
               trace_snp=np.array([t[1] for t in trace], dtype=np.int32),
               trace_loss=np.array([t[2] for t in trace]), trace_acc=np.array([t[3] for t in trace]),
               trace_n_haplo=np.array([t[4] for t in trace], dtype=np.int32),
               trace_seconds=np.array(stamps, dtype=np.float64))
    # BuildClassifiers checks for an interrupt once more AFTER a classifier is complete (LibHLA.cpp:2300):
    # a run that ends there reports "interrupted" with the finished classifier in the model -- keep it
    # (the summary line "[k] date, oob acc: ..., # of SNPs: n, # of haplo: m" is printed just before that check)
    done = rc == 0 or re.search(r"^\[\d+\] .*# of SNPs: \d+, # of haplo: \d+", log_text, re.M) is not None
    res["finished"] = np.bool_(done)
    if done:
        c = m.classifier(0)
        for key in ("snpidx", "samp_num", "freq", "hla", "packed"):
            res[key] = np.asarray(c[key])
        res["oob_acc"] = np.float64(c["oob_acc"])
    np.savez_compressed(out, **res)
    print("%s k=%d target=%s: %s in %.1f s, %d accepted SNPs %s" % (
        cfg, k, target, "finished" if done else "interrupted", dt, len(trace), [t[1] for t in trace]), flush=True)


def fromlog(cfg, k, target, log, out, seconds):
    trace = parse_trace(open(log).read())
    res = dict(config=np.array(cfg), k=np.int64(k), target=np.array(target), cpu=np.array("build container"),
               seconds=np.float64(seconds), finished=np.bool_(False),
               trace_snp=np.array([t[1] for t in trace], dtype=np.int32),
               trace_loss=np.array([t[2] for t in trace]), trace_acc=np.array([t[3] for t in trace]),
               trace_n_haplo=np.array([t[4] for t in trace], dtype=np.int32),
               trace_seconds=np.zeros(len(trace)))
    np.savez_compressed(out, **res)
    print("%s k=%d target=%s: %d accepted SNPs from %s" % (cfg, k, target, len(trace), log))


def merge(cfg, paths, extend=False):
    parts = [np.load(p) for p in paths]
    n_samp, n_snp, n_hla, cseed, mtry, tseed = CONFIGS[cfg]
    out = dict(n_samp=np.int64(n_samp), n_snp=np.int64(n_snp), n_hla_drawn=np.int64(n_hla), cohort_seed=np.int64(cseed),
               mtry=np.int64(mtry), train_seed=np.int64(tseed))
    old_ks = []
    if extend:
        # keep the classifiers of the committed fixture, add the new parts
        name = "c2_ref.npz" if cfg == "c2" else "c4_ref_prefix.npz"
        old = np.load(os.path.join(ROOT, "tests", "golden", name))
        old_ks = [int(k) for k in old["ks"]]
        for key in old.files:
            if key.startswith("c") and key.split("_")[0][1:].isdigit():
                out[key] = old[key]
    by_k = {}
    for p in parts:
        if str(p["target"]) == "max":
            continue
        by_k.setdefault(int(p["k"]), []).append(p)
    for k, ps in sorted(by_k.items()):
        a = ps[0]
        for b in ps[1:]:                      # same classifier under two bit-identical targets
            n = min(len(a["trace_snp"]), len(b["trace_snp"]))
            assert np.array_equal(a["trace_snp"][:n], b["trace_snp"][:n]), (k, "targets disagree on the SNP set")
            assert np.array_equal(a["trace_loss"][:n], b["trace_loss"][:n])
            if bool(a["finished"]) and bool(b["finished"]):
                for key in ("snpidx", "samp_num", "freq", "hla", "packed"):
                    assert np.array_equal(a[key], b[key]), (k, key)
                print("classifier", k, "identical under", str(a["target"]), "and", str(b["target"]))
            if len(b["trace_snp"]) > len(a["trace_snp"]) or (bool(b["finished"]) and not bool(a["finished"])):
                a = b
        pre = "c%d_" % k
        for key in a.files:
            if key not in ("config", "k", "cpu"):
                out[pre + key] = a[key]
    out["ks"] = np.array(sorted(set(by_k) | set(old_ks)), dtype=np.int64)
    name = "c2_ref.npz" if cfg == "c2" else "c4_ref_prefix.npz"
    path = os.path.join(ROOT, "tests", "golden", name)
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; classifiers", sorted(by_k))


if __name__ == "__main__":
    if sys.argv[1] == "worker":
        worker(sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5], float(sys.argv[6]) if len(sys.argv) > 6 else 0.0)
    elif sys.argv[1] == "fromlog":
        fromlog(sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5], sys.argv[6], float(sys.argv[7]))
    elif sys.argv[1] == "extend":
        merge(sys.argv[2], sys.argv[3:], extend=True)
    else:
        merge(sys.argv[2], sys.argv[3:])
