#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 scoring path (driver contract in the task statement).

Workload (BASELINE.json configs[1]): synthetic HLA-A training, 5,000 samples x 500 flanking SNPs,
seed-fixed, mtry = ceil(sqrt(500)) = 23, prune on.  A STEP = growing LANES classifiers per GPU
(bootstrap, greedy SNP search: pair matching + EM + pair scoring on the GPU, decisions on the
host), LANES of them in flight at a time.  Weak scaling: every rank builds its own classifiers
(classifier-sharded ensemble, no data-path collective).

  value  classifiers/min with the cohort resident in HBM (persistent training session: genotype
         matrix and bit planes stay on the device between steps)
  e2e    classifiers/min through the public call hlaAttrBagging(hla, snp, ...) on HOST arrays:
         model creation, H2D of the cohort, every per-round list upload and result download inside
         the timed region
  e2e_legacy_hooks  the same through the ten TypeGPUExtProc hooks exactly as the reference host
         calls them (full TGenotype[5000] + haplotype list from host memory per candidate SNP)
  predict.* (secondary, configs[2]): samples/s of a 100-classifier model on 200,000 samples

`--impl reference` times the reference's own CPU implementation (compiled unmodified sources,
oracle/_ref, kernel target "max") on the box's host cores, on bounded prefixes of the same
classifiers, extrapolated with the workload's pair-evaluation counts (profiles/c2_workload.json).
"""
import argparse
import json
import math
import multiprocessing as mp
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
# before any CUDA context exists (torch creates it): see hibag_b200/api.py lib()
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

N_SAMP, N_SNP, N_HLA_REQ, COHORT_SEED = 5000, 500, 40, 1
TRAIN_SEED = 2024
MTRY = 23
GATHER_NCU_CSV = ("r02_gather_cellform_ncu_raw.csv", "r01_gather_final_ncu_raw.csv")   # newest capture first
LANES = 40           # classifiers in flight per GPU (one step = LANES classifiers per GPU)
N_PREDICT = 200000
N_PREDICT_CLS = 100
WORKLOAD_JSON = os.path.join(ROOT, "profiles", "c2_workload.json")
PEAKS_JSON = os.path.join(ROOT, "profiles", "pipe_peaks.json")


def gather_traffic_from_ncu():
    """dram__bytes_read.sum + dram__bytes_write.sum of the largest cell_gather_kernel launch in the
    committed `ncu --set full` capture (profiles/*.csv, --page raw --csv): bytes per launch"""
    import csv
    for name in GATHER_NCU_CSV:
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        try:
            kn, rd, wr = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        except ValueError:
            continue
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        best = None
        for r in rows[2:]:
            if len(r) > wr and "cell_gather_kernel" in r[kn]:
                b = float(r[rd]) * scale.get(units[rd], 1.0) + float(r[wr]) * scale.get(units[wr], 1.0)
                best = b if best is None or b > best else best
        if best is not None:
            return int(best), "profiles/" + name
    return None, None


def make_cohort():
    from hibag_b200 import synth
    return synth.make_cohort(N_SAMP, N_SNP, N_HLA_REQ, seed=COHORT_SEED)


# ---------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi, recipe in B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "200",
                 "-i", str(self.device)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()          # exact PID we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            rows = [r.split(",") for r in open(self.path).read().strip().splitlines() if r.strip()]
            sm = [float(r[1]) for r in rows]
            out["samples"] = len(sm)
            if sm:
                out["sm_mhz"] = float(np.median(sm))
                out["sm_max_mhz"] = float(rows[0][2])
                out["power_w_max"] = max(float(r[3]) for r in rows)
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for k, nm in enumerate(names):
                if any(r[5 + k].strip().lower().startswith("active") for r in rows):
                    out["reasons"].append(nm)
        except Exception as e:          # never let telemetry break the benchmark
            out["error"] = str(e)
        finally:
            try:
                os.unlink(self.path)
            except OSError:
                pass
        return out


# ---------------------------------------------------------------------------------------------
# reference CPU arm (oracle/_ref): bounded prefixes of the same classifiers
# ---------------------------------------------------------------------------------------------
def load_workload():
    if os.path.exists(WORKLOAD_JSON):
        return json.load(open(WORKLOAD_JSON))
    return None


def _ref_train_worker(args):
    """One host process: the reference's BuildClassifiers on classifier `k` of the workload until
    `budget` seconds have passed (checked at accepted SNPs), verbose.detail on so that the accepted
    SNPs (src/LibHLA.cpp:2104-2111) can be compared with the B200 model's; returns (k, number of
    accepted SNPs, seconds, finished, cpu info, accepted SNP indices)."""
    k, budget, target = args
    from oracle import refpy
    ref = refpy.RefLib()
    info = ref.set_target(target)
    ref.set_gpu_procs(None)
    coh = make_cohort()
    m = ref.new_model()
    m.init_training(coh.geno, coh.h1, coh.h2, coh.n_hla)
    fd, log = tempfile.mkstemp(suffix=".log")      # the reference prints through Rprintf -> fd 2
    saved = os.dup(2)
    os.dup2(fd, 2)
    ref.set_interrupt(seconds=budget)
    t0 = time.time()
    try:
        rc = m.build(1, MTRY, prune=True, verbose=2, reseed_base=TRAIN_SEED, first_index=k, allow_interrupt=True)
    finally:
        dt = time.time() - t0
        os.dup2(saved, 2)
        os.close(fd); os.close(saved)
    snps = []
    for ln in open(log).read().splitlines():
        mt = re.match(r"^\s*(\d+), SNP: (\d+), loss:", ln)
        if mt:
            snps.append(int(mt.group(2)) - 1)
    os.unlink(log)
    ref.set_interrupt(0.0, -1)
    return k, len(snps), dt, rc == 0, info, snps


def _ref_predict_worker(args):
    """reference PredictHLA (target 'max') on successive 4-sample chunks for `budget` seconds"""
    p, budget = args
    from oracle import refpy
    from hibag_b200 import synth
    ref = refpy.RefLib()
    ref.set_target("max")
    ref.set_gpu_procs(None)
    wl = np.load(os.path.join(ROOT, "tests", "golden", "c2_model.npz"))
    coh = make_cohort()
    new = synth.draw_more(coh, 64 * (p + 1), seed=4242).geno[64 * p:].astype(np.int32)
    m = ref.new_model()
    m.init_predict(N_SNP, 1, int(wl["n_hla"]))
    n_src = len(wl["snp_off"]) - 1
    for c in range(N_PREDICT_CLS):
        k = c % n_src
        a, b = wl["snp_off"][k:k + 2]; q, r = wl["hap_off"][k:k + 2]
        m.add_classifier(wl["snpidx"][a:b], wl["freq"][q:r], wl["hla"][q:r], wl["packed"][q:r])
    done, t0 = 0, time.time()
    while time.time() - t0 < budget and done + 4 <= len(new):
        m.predict(new[done:done + 4])
        done += 4
    return done, time.time() - t0


TIME_PROFILE_JSON = os.path.join(ROOT, "profiles", "ref_time_profile.json")


def load_time_profile():
    if os.path.exists(TIME_PROFILE_JSON):
        return json.load(open(TIME_PROFILE_JSON))
    return None


def reference_train_rate(step, procs, budget, wl, target="max", parity_worker=False):
    """classifiers/min of the reference CPU path with `procs` worker processes, each growing one
    classifier of the workload for `budget` seconds (until the next accepted SNP). The prefix is
    extrapolated to the whole classifier with the reference's own TIME PROFILE of that classifier
    (profiles/ref_time_profile.json: seconds at every accepted SNP and to completion, measured once
    with the same binary and target): est = t_box(n) * T_full / T(n). Round 1 extrapolated by pair
    evaluations, which overestimates the classifier 2.4x (early rounds run short lists at a far
    lower pair rate; profiles/ref_full_classifier.json) -- kept only as the fallback for classifiers
    without a profile. With parity_worker one more process runs classifier 0 under target avx2
    (== base, the parity oracle) for the prefix check only."""
    prof = (load_time_profile() or {}).get("classifiers", {}) if target == "max" else {}
    cal = sorted(int(k) for k in prof)
    n_tr = len(wl["classifiers"]) if wl else 0
    if cal:
        ks = [cal[(step * procs + p) % len(cal)] for p in range(procs)]
    else:
        ks = [(step * procs + p) % max(n_tr, 1) for p in range(procs)]
    jobs = [(k, budget, target) for k in ks] + ([(0, budget, "avx2")] if parity_worker else [])
    with mp.get_context("spawn").Pool(len(jobs)) as pool:
        res = pool.map(_ref_train_worker, jobs)
    rate, detail = 0.0, []
    for k, accepted, dt, finished, info, snps in res[:procs]:
        how = "finished"
        if finished:
            est = dt
        elif str(k) in prof and 0 < accepted <= len(prof[str(k)]["seconds_at_accepted_snp"]):
            pk = prof[str(k)]
            est = dt * pk["full_seconds"] / pk["seconds_at_accepted_snp"][accepted - 1]
            how = "time profile"
        elif wl and k < n_tr:
            tr = wl["classifiers"][k]
            cum = {a: p for a, p in tr["accepted_pairs"]}
            part = cum.get(accepted) or cum.get(max([x for x in cum if x <= accepted] or [0]), None)
            est = dt * tr["total_pairs"] / part if part else float("nan")
            how = "pair evaluations (overestimates)"
        else:
            est = float("nan")
        if est == est and est > 0:
            rate += 60.0 / est
        detail.append(dict(classifier=k, accepted_snps=accepted, seconds=round(dt, 2),
                           est_seconds_per_classifier=round(est, 1), extrapolated_by=how, snps=snps))
    extra = [dict(classifier=r[0], snps=r[5], target="avx2") for r in res[procs:]]
    return rate, detail, res[0][4], extra


def load_calibration():
    """one line on how the bounded prefix is extrapolated and how well that works"""
    tp = load_time_profile()
    if tp:
        cl = tp["classifiers"]
        return "time profile of classifiers %s run to completion (%s s); cross-profile error of a 15 s prefix %s" % (
            ",".join(sorted(cl)), "/".join("%.0f" % cl[k]["full_seconds"] for k in sorted(cl)),
            "/".join("%+.0f%%" % (100.0 * (cl[k]["cross_profile_estimate_from_15s_prefix"] / cl[k]["full_seconds"] - 1))
                     for k in sorted(cl)))
    return None


def cpu_baseline(procs, budget, wl, with_predict=True, snp_lookup=None):
    """the reference's CPU path on this box's host cores (bounded sample). snp_lookup(k) -> the B200
    model's accepted SNPs of classifier k: the prefix the reference reached must equal them."""
    rate, detail, info, extra = reference_train_rate(0, procs, budget, wl, parity_worker=snp_lookup is not None)
    out = {"value": rate, "unit": "classifiers/min", "cores": procs, "kind": "reference",
           "target": info,
           "sample": "%d procs x 1 classifier x %.0f s of reference BuildClassifiers (target max), extrapolated with the "
                     "classifier's measured time profile (profiles/ref_time_profile.json)" % (procs, budget),
           "calibration": load_calibration(),
           "per_process": [{k: v for k, v in dd.items() if k != "snps"} for dd in detail[:4]]}
    if snp_lookup is not None:
        ok, n_cmp, ok_max = True, 0, True
        for dd in extra:                    # target avx2 == base: the parity oracle
            mine = list(snp_lookup(dd["classifier"]))
            ok = ok and len(dd["snps"]) > 0 and mine[:len(dd["snps"])] == dd["snps"]
            n_cmp += len(dd["snps"])
        for dd in detail:                   # target max (AVX-512 sums are reassociated: informative only)
            mine = list(snp_lookup(dd["classifier"]))
            ok_max = ok_max and mine[:len(dd["snps"])] == dd["snps"]
        out["parity_prefix_ok"] = bool(ok)
        out["parity_prefix_snps"] = n_cmp
        out["parity_prefix_ok_target_max"] = bool(ok_max)
        out["parity_prefix_snps_target_max"] = sum(len(dd["snps"]) for dd in detail)
    if with_predict and os.path.exists(os.path.join(ROOT, "tests", "golden", "c2_model.npz")):
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_ref_predict_worker, [(p, budget / 2) for p in range(procs)])
        out["predict_value"] = sum(dn / t for dn, t in res)
        out["predict_unit"] = "samples/s"
        out["predict_sample"] = "%d samples over %d processes, 100-classifier model" % (
            sum(dn for dn, _ in res), procs)
    return out


def run_reference_arm(args):
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    if rank != 0:
        return
    wl = load_workload()
    procs = args.cpu_procs or os.cpu_count() or 1
    rates, detail, info = [], None, ""
    t0 = time.time()
    for s in range(args.warmup):
        reference_train_rate(s, procs, min(args.cpu_seconds, 3.0), wl)
    tw = time.time()
    for s in range(args.steps):
        r, detail, info, _ = reference_train_rate(args.warmup + s, procs, args.cpu_seconds, wl)
        rates.append(r)
    timed = time.time() - tw
    value = float(np.mean(rates))
    full = {
        "impl": "reference", "metric": "classifiers/min trained (HLA-A 5k x 500 SNP)", "value": value,
        "unit": "classifiers/min", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        # the arm's real wall clock per step (a step = one bounded sample on every host core)
        "ms_per_step": 1e3 * timed / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.gpus, args.lanes or default_lanes(args.gpus)),
        "cpu_baseline": {"value": value, "unit": "classifiers/min", "cores": procs, "kind": "reference",
                         "target": info,
                         "sample": "per step: %d procs x 1 classifier x %.0f s of reference BuildClassifiers (target "
                                   "max), extrapolated with its time profile" % (procs, args.cpu_seconds),
                         "calibration": load_calibration(),
                         "extrapolated_seconds_per_classifier": 60.0 * procs / value if value > 0 else None,
                         "per_process": [{k: v for k, v in dd.items() if k != "snps"} for dd in (detail or [])[:4]]},
        "e2e": {"value": value, "unit": "classifiers/min", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": round(time.time() - t0, 1), "timed_wall_s": round(timed, 1),
    }
    side = write_detail(full, args.gpus)
    line = compact_line(full)
    line["detail_file"] = side
    emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# the stdout line: bounded (< 4 KB), flat scalars under the contract's keys; everything else goes
# to gpurun_out/bench_detail_n<N>.json and stderr
# ---------------------------------------------------------------------------------------------
LINE_LIMIT = 4096


def _short(v, n=120):
    if isinstance(v, str):
        return v if len(v) <= n else v[:n - 3] + "..."
    if isinstance(v, float):
        return float("%.6g" % v)
    return v


def _flat(dst, src, keys, prefix=""):
    for k in keys:
        if src and src.get(k) is not None and not isinstance(src.get(k), (dict, list)):
            dst[prefix + k] = _short(src[k])


def compact_line(detail):
    """The contract's keys with flat scalar members only; never longer than LINE_LIMIT bytes
    (tests/test_bench_contract.py)."""
    line = {k: _short(detail.get(k)) for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                                               "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
                                               "gpu_launches")}
    if detail.get("impl"):
        line["impl"] = detail["impl"]
    cfg = detail.get("config") or {}
    line["config"] = {k: _short(v) for k, v in cfg.items() if not isinstance(v, (dict, list))}
    if detail.get("clocks"):
        c = detail["clocks"]
        line["clocks"] = {k: c.get(k) for k in ("sm_mhz", "sm_max_mhz", "reasons", "samples", "power_w_max") if k in c}
    e = {}
    _flat(e, detail.get("e2e"), ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "api"))
    hooks, pred = detail.get("e2e_legacy_hooks"), detail.get("predict")
    if hooks:
        e["legacy_hooks_value"] = _short(hooks.get("value"))
    if pred:
        e["predict_value"] = _short(pred.get("value"))
        e["predict_e2e_value"] = _short((pred.get("e2e") or {}).get("value"))
        e["predict_unit"] = pred.get("unit")
        for k in ("distinct_genotype_fraction", "sharded_by_classifier_value", "allreduce_ms", "allreduce_bytes"):
            if pred.get(k) is not None:
                e["predict_" + k] = _short(pred[k])
    line["e2e"] = e or None
    r = {}
    rf = detail.get("roofline")
    _flat(r, rf, ("bound", "kernel", "achieved", "peak", "unit", "frac", "traffic", "avg_launch_ms", "launches",
                  "pair_evals_per_s", "fp64_frac", "peak_source", "timing"))
    if rf:
        _flat(r, rf.get("in_bag_launches"), ("frac", "avg_launch_ms"), "in_bag_")
        _flat(r, rf.get("out_of_bag_launches"), ("frac", "avg_launch_ms"), "out_of_bag_")
        _flat(r, rf.get("alone"), ("frac", "in_bag_frac"), "alone_")
        _flat(r, rf.get("screening"), ("executed_fraction",), "screen_")
        _flat(r, rf.get("em"), ("frac", "cycles_per_iteration", "longest_chain_mean"), "em_")
        _flat(r, rf.get("sm_time"), ("scoring_share", "em_share", "other_share", "busy"), "sm_time_")
        _flat(r, rf, ("frac_launch_events", "frac_of_held_sm_time_in_bag"))
    if detail.get("roofline_unscreened"):
        r["unscreened_frac"] = _short(detail["roofline_unscreened"].get("frac"))
    if pred and pred.get("roofline"):
        r["predict_frac"] = _short(pred["roofline"].get("frac"))
    if detail.get("large_list"):
        r["global_operand_frac"] = _short(detail["large_list"].get("frac"))
    line["roofline"] = r or None
    c = {}
    cb = detail.get("cpu_baseline")
    _flat(c, cb, ("value", "unit", "cores", "kind", "target", "sample", "predict_value", "predict_unit",
                  "parity_prefix_ok", "parity_prefix_snps", "calibration"))
    line["cpu_baseline"] = c or None
    # hard bound: drop the longest strings first, then optional members
    def size():
        return len(json.dumps(line))
    for key, sub in (("cpu_baseline", "sample"), ("roofline", "peak_source"), ("config", "l2"), ("e2e", "api"),
                     ("roofline", "timing"), ("config", "workload")):
        if size() <= LINE_LIMIT:
            break
        if line.get(key) and sub in line[key]:
            line[key][sub] = _short(line[key][sub], 48)
    assert size() <= LINE_LIMIT, size()
    return line


def write_detail(detail, world):
    """full record -> gpurun_out/bench_detail_n<N>.json (+ stderr); returns the relative path"""
    rel = os.path.join("gpurun_out", "bench_detail_%sn%d.json" % ("ref_" if detail.get("impl") else "", world))
    try:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        with open(os.path.join(ROOT, rel), "w") as f:
            json.dump(detail, f, indent=1)
    except OSError:
        rel = None
    sys.stderr.write("bench detail: " + json.dumps(detail) + "\n")
    sys.stderr.flush()
    return rel


_REAL_STDOUT = None


def emit(text):
    out = _REAL_STDOUT or sys.stdout
    out.write(text + "\n")
    out.flush()


def default_lanes(world):
    """classifiers in flight per GPU: enough to cover the per-round latency chain of a lane (EM
    launch, two scoring passes, host decisions). The SAME at every N (weak scaling of one per-GPU
    configuration): the lanes block on CUDA events, so they do not need a host core each."""
    return LANES


def default_threads(world, lanes):
    """host threads per rank handed to the trainer (split over the lanes' pools): two per lane when
    the rank can count on that many cores, one per lane otherwise"""
    cores = max(1, (os.cpu_count() or 1) // max(world, 1))
    return 2 * lanes if cores >= 12 else lanes


def workload_config(n_gpus, lanes):
    return {"workload": "configs[1]: synthetic HLA-A training 5000 samples x 500 SNPs, 34 alleles, mtry 23, prune; "
                        "step = %d classifiers/GPU" % lanes,
            "n_samp": N_SAMP, "n_snp": N_SNP, "n_hla": 34, "mtry": MTRY, "cohort_seed": COHORT_SEED,
            "train_seed": TRAIN_SEED, "parallelism": "classifier-sharded x%d" % n_gpus,
            "lanes": lanes, "classifiers_per_step": lanes * n_gpus,
            "l2": "inputs > L2: each lane rebuilds ~0.5 GB of cell matrix, need lists, bounds per selection round"}


# ---------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    from hibag_b200 import api, dist as hd, synth
    rank, local_rank, world = hd.init()
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node == --gpus"
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    api.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    info = api.device_info()
    lanes = args.lanes
    if not lanes:
        lanes = default_lanes(world)
    n_threads = args.threads or default_threads(world, lanes)

    coh = make_cohort()
    geno = np.ascontiguousarray(coh.geno, dtype=np.int8)

    def sync_all():
        torch.cuda.synchronize()
        hd.barrier()
        torch.cuda.synchronize()

    # ---- resident path --------------------------------------------------------------------------
    model = api.HLAModel(N_SNP, coh.n_hla)
    model.set_training(geno, coh.h1, coh.h2)
    dev_em = not args.host_em

    def step_resident(step):
        # classifiers (step*world + rank)*lanes ... + lanes-1 of the global sequence
        model.train(lanes, MTRY, prune=True, seed=TRAIN_SEED, per_classifier_seed=True,
                    first_index=(step * world + rank) * lanes, n_threads=n_threads,
                    n_concurrent=lanes, em_on_device=dev_em)

    for s in range(args.warmup):
        step_resident(s)
    st0 = model.train_stats()
    sampler = ClockSampler(local_rank)
    sync_all()
    api.sm_time(reset=True)          # held SM-time per kernel class over the timed region
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    import resource
    ru0 = resource.getrusage(resource.RUSAGE_SELF)
    e0.record()
    t0 = time.time()
    for s in range(args.steps):
        step_resident(args.warmup + s)
    e1.record()
    sync_all()
    wall = time.time() - t0
    ru1 = resource.getrusage(resource.RUSAGE_SELF)
    host_cpu_s = (ru1.ru_utime - ru0.ru_utime) + (ru1.ru_stime - ru0.ru_stime)
    ms_local = e0.elapsed_time(e1)
    ms = hd.max_over_ranks(ms_local, dev)
    clocks = sampler.stop() if rank == 0 else None
    acct = api.sm_time()
    st1 = model.train_stats()
    d = {k: st1[k] - st0[k] for k in st1}
    value = world * args.steps * lanes / (ms / 60000.0)

    # ---- end to end through the public API on host arrays -------------------------------------------
    e2e = None
    e2e_hooks = None
    if not args.no_e2e:
        def run_public(first_step, n_steps, **kw):
            # ONE call of the public API per rank: the cohort goes in as host numpy arrays
            m = api.hlaAttrBagging((coh.h1, coh.h2), geno, nclassifier=n_steps * lanes, mtry=MTRY, prune=True,
                                   mono_rm=False, seed=TRAIN_SEED, nthread=n_threads, per_classifier_seed=True,
                                   first_index=first_step * world * lanes + rank, index_stride=world,
                                   n_concurrent=lanes, em_on_device=dev_em, **kw)
            return m.train_stats()
        run_public(0, 1)
        sync_all()
        e0.record()
        stt = run_public(args.warmup, args.steps)
        e1.record()
        sync_all()
        ms2 = hd.max_over_ranks(e0.elapsed_time(e1), dev)
        e2e = {"value": world * args.steps * lanes / (ms2 / 60000.0), "unit": "classifiers/min",
               "h2d_bytes_per_step": int((stt["h2d_bytes"] + geno.nbytes + 8 * N_SAMP) / args.steps),
               "d2h_bytes_per_step": int(stt["d2h_bytes"] / args.steps),
               "api": "one hlaAttrBagging(hla, snp, nclassifier=steps*lanes) call on host numpy arrays "
                      "(new model, cohort H2D, per-round list uploads, accuracy/ratio/frequency downloads)"}
        # the reference-facing plugin struct, driven with the reference's own call sequence
        n_hook = max(1, min(3, args.steps))
        hook_threads = max(2, (os.cpu_count() or 2) // max(world, 1))    # the host EM pool: one thread per core

        def run_hooks(first, count):
            m = api.hlaAttrBagging((coh.h1, coh.h2), geno, nclassifier=count, mtry=MTRY, prune=True,
                                   mono_rm=False, seed=TRAIN_SEED, nthread=hook_threads,
                                   per_classifier_seed=True, use_legacy_hooks=True,
                                   first_index=rank + world * first, index_stride=world)
            return m.train_stats()
        run_hooks(0, 1)
        sync_all()
        e0.record()
        sth = run_hooks(1, n_hook)
        e1.record()
        sync_all()
        ms3 = hd.max_over_ranks(e0.elapsed_time(e1), dev)
        e2e_hooks = {"value": world * n_hook / (ms3 / 60000.0), "unit": "classifiers/min",
                     "classifiers": n_hook * world,
                     "h2d_bytes_per_classifier": int(sth["h2d_bytes"] / n_hook),
                     "d2h_bytes_per_classifier": int(sth["d2h_bytes"] / n_hook),
                     "api": "hlaAttrBagging(..., use_legacy_hooks=True): ten TypeGPUExtProc hooks, "
                            "TGenotype[5000] + haplotype list from host memory per candidate SNP, scalar "
                            "results back per hook call, strictly sequential as the reference host calls them"}

    # ---- roofline of the pair-scoring kernel, training region -----------------------------------------
    peaks = json.load(open(PEAKS_JSON)) if os.path.exists(PEAKS_JSON) else None
    popc_peak = (peaks or {}).get("popc32_per_s", 148 * 16 * 1.965e9)
    peak_src = "measured (profiles/pipe_peaks.json, POPC.32 microbenchmark on this pool's B200)" if peaks \
        else "nominal 148 SM x 16 POPC/clk x 1.965 GHz (no measured file)"
    screened = d["gather_kernel_launches"] > 0
    k_ms = d["gather_kernel_ms"] if screened else d["cell_kernel_ms"]
    n_l = max(d["gather_kernel_launches"] if screened else d["cell_kernel_launches"], 1)
    avg_ms = k_ms / n_l
    pair_rate = d["pair_evals"] / max(k_ms * 1e-3, 1e-12)
    issued_rate = d["popc32_issued"] / max(k_ms * 1e-3, 1e-12)
    eff_rate = d["pair_evals_nominal"] / max(d["cell_kernel_ms"] * 1e-3, 1e-12)
    roofline = {
        "bound": "popc", "kernel": "cell_gather_kernel" if screened else "cell_pass_kernel",
        "achieved": issued_rate / 1e9, "peak": popc_peak / 1e9, "unit": "Gpopc32/s",
        "frac": issued_rate / popc_peak,
        "achieved_reference_formulation": pair_rate * 4 / 1e9,
        "frac_reference_formulation": pair_rate * 4 / popc_peak,
        "pair_evals_per_s": pair_rate, "avg_launch_ms": avg_ms, "launches": int(n_l),
        "pair_evals_per_launch": d["pair_evals"] / n_l,
        "screening": {
            "pair_evals_executed": int(d["pair_evals"]), "pair_evals_reference": int(d["pair_evals_nominal"]),
            "executed_fraction": d["pair_evals"] / max(d["pair_evals_nominal"], 1),
            "uncertified_sums_rescored": int(d["n_screen_fallback"]),
            "scoring_pass_ms": d["cell_kernel_ms"],
            "effective_pair_evals_per_s": eff_rate,
            "effective_frac_reference_formulation": eff_rate * 4 / popc_peak,
            "note": "exact screening (DESIGN.md 4.5): cells proven irrelevant for the reference's outputs are "
                    "not scored; 'effective' = the pair evaluations the reference performs for the same "
                    "passes / the summed durations of the whole screened passes (bounds, need lists, gather "
                    "launch, reduction)"},
        "note": "achieved = POPC.32 the pair-scoring kernel issues (1 per executed pair evaluation per 32 "
                "SNPs: the one-popcount distance) / the time its launches hold the GPU (see timing). POPC (XU pipe) "
                "and the lane-private LDS.64 table lookup co-bind at the same 16 /clk/SM. *_reference_formulation "
                "counts the 4 POPC.32 per pair evaluation of the reference's hamm_d (SURVEY.md 8d). 'alone' holds "
                "the same launches with one classifier in flight (every launch alone on the GPU, CUDA events), "
                "'in_bag_launches' the launches that carry the work, 'launch_events' the overlapped CUDA-event "
                "figure. The plain kernel's fraction (every cell, full warps) is under roofline_unscreened and "
                "predict.roofline.",
        "peak_source": peak_src,
        # dram__bytes_read.sum + dram__bytes_write.sum of one in-bag launch (ncu --set full,
        # profiles/r01_gather_r2_ncu_raw.csv): the surviving entries of the cell matrix written once and
        # the need lists read once; the path is not HBM-bound
        "traffic": gather_traffic_from_ncu()[0],
        "traffic_note": "dram bytes read + written by the largest in-bag gather launch of the ncu --set full capture "
                        "%s (a late selection round of one classifier); the launches of a step differ in "
                        "size" % gather_traffic_from_ncu()[1],
        "fp64_frac": pair_rate * 3 / (peaks or {}).get("fp64_ops_per_s", 148 * 64 * 1.965e9),
        "em_kernel_ms": d["em_kernel_ms"],
    }
    # ---- where the GPU's time went: held SM-time per kernel class (counters inside the kernels) ------
    sm_clock = info["clock_khz"] * 1e3
    sm_total = info["sm_count"] * ms_local * 1e-3 * sm_clock             # SM-cycles of the timed region
    share = {k: v / sm_total for k, v in acct.items() if k != "em_cta_cycles"}
    popc_per_sm_cycle = popc_peak / (info["sm_count"] * sm_clock)
    roofline["sm_time"] = {
        "scoring_share": share["gather_ib"] + share["gather_oob"] + share["cell_pass"],
        "gather_ib_share": share["gather_ib"], "gather_oob_share": share["gather_oob"],
        "em_share": share["em"],
        "other_share": share["screen_bound"] + share["screen_need"] + share["screen_tasks"] + share["reduce_oob"] + share["reduce_ib"] + share.get("screen_dedup", 0.0),
        "busy": sum(share.values()),
        "by_class": {k: round(v, 4) for k, v in share.items()},
        "note": "share of the GPU's SM-cycles in the timed region HELD by each kernel class: sum over its CTAs of resident "
                "cycles x the fraction of an SM a CTA of that launch occupies (1 / CTAs that fit an SM by registers, threads "
                "and shared memory). Kernels of the lanes overlap, so CUDA-event durations (em_kernel_ms, gather ms) are "
                "NOT exclusive time and do not add up to the step; held SM-time does. Pair preparation (cub sort, "
                "pair matching) is not instrumented."}
    if screened and acct["gather_ib"] > 0:
        # ---- the headline fraction: POPC issued / the SM-time the kernel's CTAs HELD in the timed region.
        # 40 lanes overlap their launches on six streams and share every SM with other lanes' EM CTAs, so a
        # launch's CUDA-event duration is not its kernel time ("CUDA-event sums are not
        # exclusive time ... measure SM-time"): the kernels count their own resident cycles (devutil.cuh:
        # SmAcct), weighted by the share of an SM a CTA occupies (1/8: 128 threads x 64 registers).
        held_cycles = acct["gather_ib"] + acct["gather_oob"]                 # SM-cycles
        held_s = held_cycles / (info["sm_count"] * sm_clock)                 # seconds of the WHOLE GPU
        held = d["gather_ib_popc32"] / (acct["gather_ib"] * popc_per_sm_cycle)
        roofline["frac_of_held_sm_time_in_bag"] = held
        roofline["frac_of_held_sm_time"] = d["popc32_issued"] / (held_cycles * popc_per_sm_cycle)
        roofline["launch_events"] = {
            "achieved": roofline["achieved"], "frac": roofline["frac"], "avg_launch_ms": roofline["avg_launch_ms"],
            "pair_evals_per_s": roofline["pair_evals_per_s"],
            "note": "the same POPC count / the summed CUDA-event durations of the launches (events on the launching "
                    "stream): every overlapped launch is charged the whole time it shares the GPU, so this is a lower "
                    "bound of the kernel's efficiency, not a roofline fraction"}
        roofline["achieved"] = d["popc32_issued"] / held_s / 1e9
        roofline["frac"] = roofline["frac_of_held_sm_time"]
        roofline["avg_launch_ms"] = held_s * 1e3 / n_l
        roofline["pair_evals_per_s"] = d["pair_evals"] / held_s
        roofline["frac_reference_formulation"] = d["pair_evals"] * 4 / held_s / popc_peak
        roofline["fp64_frac"] = d["pair_evals"] * 3 / held_s / (peaks or {}).get("fp64_ops_per_s", 148 * 64 * 1.965e9)
        roofline["frac_launch_events"] = roofline["launch_events"]["frac"]
        roofline["timing"] = ("sm_time: avg_launch_ms = SM-cycles the kernel's CTAs held in the timed region (in-kernel clock64 "
                              "counters, x 1/8 SM per CTA) / (SMs x clock) / launches, i.e. the duration a launch would have "
                              "with the GPU to itself at the same efficiency; CUDA-event durations overlap (launch_events)")
    # ---- the EM kernel (time-dominant when serialised) against ITS bound: the latency of the longest
    # chain of dependent fp64 adds of every M step (16.9 cycles per dependent DADD, profiles/pipe_peaks.json)
    if d.get("em_chain_adds", 0) > 0 and acct["em_cta_cycles"] > 0:
        dadd = (peaks or {}).get("dadd_dependent_cycles", 16.9)
        floor_cycles = d["em_chain_adds"] * dadd
        roofline["em"] = {
            "kernel": "em_chain_kernel", "bound": "latency of the dependent fp64 add chain (M step), %.1f cycles per add" % dadd,
            "achieved": floor_cycles / 1e9, "peak": acct["em_cta_cycles"] / 1e9, "unit": "Gcycles (chain floor / CTA-resident)",
            "frac": floor_cycles / acct["em_cta_cycles"], "sm_time_share": share["em"],
            "iterations": int(d["em_iterations"]), "candidates": int(d["n_em"]),
            "cycles_per_iteration": acct["em_cta_cycles"] / max(d["em_iterations"], 1),
            "longest_chain_mean": d["em_chain_adds"] / max(d["em_iterations"], 1),
            "pair_updates_per_iteration": d["em_pair_updates"] / max(d["em_iterations"], 1),
            "note": "one CTA per candidate SNP; an iteration cannot be shorter than its longest chain of sequential "
                    "fp64 adds (bit-exact M step). frac = that floor / the CTA's resident cycles; the rest is the three "
                    "E passes, barriers and L2 latency (DESIGN.md 4.4)"}
    if screened and d["gather_ib_launches"] > 0:
        ib_rate = d["gather_ib_popc32"] / max(d["gather_ib_kernel_ms"] * 1e-3, 1e-12)
        roofline["in_bag_launches"] = {
            "launches": int(d["gather_ib_launches"]), "avg_launch_ms": d["gather_ib_kernel_ms"] / d["gather_ib_launches"],
            "achieved": ib_rate / 1e9, "frac": ib_rate / popc_peak,
            "share_of_gather_ms": d["gather_ib_kernel_ms"] / max(d["gather_kernel_ms"], 1e-12),
            "share_of_pair_evals": d["gather_ib_popc32"] / max(d["popc32_issued"], 1),
            "note": "the in-bag launches alone (86 of 595 cells per sample survive: full warps); the out-of-bag "
                    "launches (2.5 cells per sample) are latency-bound"}
    # the same kernel timed ALONE: one lane, its launches serialised, nothing else on the GPU
    if screened and not args.no_unscreened:
        m3 = api.HLAModel(N_SNP, coh.n_hla)
        m3.set_training(geno, coh.h1, coh.h2)
        n_alone = 2
        m3.train(n_alone, MTRY, prune=True, seed=TRAIN_SEED, per_classifier_seed=True, first_index=rank * n_alone,
                 n_threads=n_threads, n_concurrent=1, em_on_device=dev_em)
        torch.cuda.synchronize()
        sa = m3.train_stats()
        a_all = sa["popc32_issued"] / max(sa["gather_kernel_ms"] * 1e-3, 1e-12)
        a_ib = sa["gather_ib_popc32"] / max(sa["gather_ib_kernel_ms"] * 1e-3, 1e-12)
        roofline["alone"] = {
            "classifiers": n_alone, "lanes": 1,
            "launches": int(sa["gather_kernel_launches"]), "avg_launch_ms": sa["gather_kernel_ms"] / max(sa["gather_kernel_launches"], 1),
            "achieved": a_all / 1e9, "frac": a_all / popc_peak,
            "in_bag_launches": int(sa["gather_ib_launches"]),
            "in_bag_avg_launch_ms": sa["gather_ib_kernel_ms"] / max(sa["gather_ib_launches"], 1),
            "in_bag_achieved": a_ib / 1e9, "in_bag_frac": a_ib / popc_peak,
            "note": "same kernel, same workload, ONE classifier in flight: every gather launch has the GPU to "
                    "itself, so launch duration = kernel time (the burst figure). In the timed region the lanes "
                    "overlap their launches on six streams and share the SMs with the EM CTAs of other lanes."}
        del m3
    # one step with screening off: the plain pair-scoring kernel alone on its stream
    roofline_plain = None
    if screened and not args.no_unscreened:
        m2 = api.HLAModel(N_SNP, coh.n_hla)
        m2.set_training(geno, coh.h1, coh.h2)
        n_plain = min(lanes, 6)
        m2.train(n_plain, MTRY, prune=True, seed=TRAIN_SEED, per_classifier_seed=True, first_index=rank * n_plain,
                 n_threads=n_threads, n_concurrent=n_plain, em_on_device=dev_em, screening=False)
        torch.cuda.synchronize()
        sp = m2.train_stats()
        r_issued = sp["popc32_issued"] / max(sp["cell_kernel_ms"] * 1e-3, 1e-12)
        roofline_plain = {"kernel": "cell_pass_kernel", "classifiers": n_plain, "achieved": r_issued / 1e9,
                          "peak": popc_peak / 1e9, "unit": "Gpopc32/s", "frac": r_issued / popc_peak,
                          "pair_evals_per_s": sp["pair_evals"] / max(sp["cell_kernel_ms"] * 1e-3, 1e-12),
                          "launches": int(sp["cell_kernel_launches"]),
                          "avg_launch_ms": sp["cell_kernel_ms"] / max(sp["cell_kernel_launches"], 1),
                          "classifiers_per_min": 60.0 * n_plain / max(sp["seconds_total"], 1e-9),
                          "note": "screening off: every cell of every candidate scored by cell_pass_kernel, "
                                  "launches serialised on one stream"}
        del m2

    # ---- secondary metric: prediction (configs[2]) ---------------------------------------------------
    predict = None
    if not args.no_predict:
        predict = bench_predict(api, hd, synth, torch, model, coh, rank, world, dev, args)
        if predict and peaks:
            predict["roofline"]["peak"] = popc_peak / 1e9
            predict["roofline"]["frac"] = predict["roofline"]["achieved"] * 1e9 / popc_peak
            predict["roofline"]["frac_reference_formulation"] = \
                predict["roofline"]["achieved_reference_formulation"] * 1e9 / popc_peak

    # ---- PLINK BED import: an HBM-bound byte kernel ----------------------------
    bed = None
    large = None
    if rank == 0 and not args.no_predict:
        bed = bench_bed_decode(api, torch, dev)
        large = bench_large_list(api, torch, dev, popc_peak)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            # classifiers 0..lanes-1 were built by warm-up step 0 (rank 0): the reference's accepted-SNP
            # prefixes of the same classifiers must equal them
            def snp_lookup(k):
                return [int(x) for x in model.classifier(k)["snpidx"]] if k < model.num_classifiers() and args.warmup > 0 else []
            cpu = cpu_baseline(args.cpu_procs or os.cpu_count() or 1, args.cpu_seconds, load_workload(),
                               snp_lookup=snp_lookup if args.warmup > 0 else None)
        except Exception as ex:        # the checker library may be absent on some boxes
            cpu = {"value": None, "unit": "classifiers/min", "cores": 0, "kind": "reference",
                   "sample": "unavailable: %s" % ex}

    if rank == 0:
        n_snps = [len(model.classifier(k)["snpidx"]) for k in range(model.num_classifiers())]
        n_haps = [len(model.classifier(k)["freq"]) for k in range(model.num_classifiers())]
        detail = {
            "metric": "classifiers/min trained (HLA-A 5k x 500 SNP)", "value": value,
            "unit": "classifiers/min", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(world, lanes),
            "e2e": e2e, "e2e_legacy_hooks": e2e_hooks, "gpu_launches": int(d["kernel_launches"]),
            "clocks": clocks, "roofline": roofline, "roofline_unscreened": roofline_plain,
            "cpu_baseline": cpu, "predict": predict, "bed_decode": bed, "large_list": large,
            "train_detail": {
                "host_threads": n_threads, "lanes": lanes, "em_on_device": dev_em,
                "host_cores": os.cpu_count(), "host_cpu_seconds": host_cpu_s,
                "host_cores_busy": host_cpu_s / max(wall, 1e-9),
                "em_kernel_ms": d["em_kernel_ms"], "em_host_fallbacks": int(d["n_em_host_fallback"]),
                "seconds_em_sum": d["seconds_em"],
                "seconds_prepare": d["seconds_prepare"], "seconds_candidates": d["seconds_phase_oob"],
                "seconds_gpu_wait_sum": d["seconds_gpu_wait"], "gpu_kernel_span_ms": d["gpu_kernel_ms"],
                "pair_evals": int(d["pair_evals"]), "pair_evals_reference": int(d["pair_evals_nominal"]),
                "oob_evals": int(d["n_oob_evals"]),
                "ib_evals": int(d["n_ib_evals"]), "em_runs": int(d["n_em"]),
                "h2d_bytes_per_step": int(d["h2d_bytes"] / args.steps),
                "d2h_bytes_per_step": int(d["d2h_bytes"] / args.steps), "wall_s": wall,
                "classifiers": {"count": len(n_snps), "n_snp_mean": float(np.mean(n_snps)), "n_snp_min": int(min(n_snps)),
                                "n_snp_max": int(max(n_snps)), "n_haplo_mean": float(np.mean(n_haps)),
                                "n_haplo_min": int(min(n_haps)), "n_haplo_max": int(max(n_haps))}},
            "device": info,
        }
        # the full record goes to a side file (and stderr); stdout carries ONE bounded line
        side = write_detail(detail, world)
        line = compact_line(detail)
        line["detail_file"] = side
        emit(json.dumps(line))
    hd.shutdown()


def bench_predict(api, hd, synth, torch, model, coh, rank, world, dev, args):
    """samples/s of a 100-classifier model on 200,000 fresh samples, sample-sharded over ranks"""
    n_src = model.num_classifiers()
    big = api.HLAModel(N_SNP, coh.n_hla)
    for c in range(N_PREDICT_CLS):
        k = model.classifier(c % n_src)
        big.add_classifier(k["snpidx"], k["freq"], k["hla"], k["packed"])
    n_total = args.predict_samples
    b, e = hd.shard_range(n_total, rank, world)
    n = e - b
    all_host = np.ascontiguousarray(synth.draw_more(coh, n_total, seed=99).geno, dtype=np.int8)
    host = np.ascontiguousarray(all_host[b:e])
    pinned = torch.from_numpy(host).pin_memory()
    g = pinned.to(dev, non_blocking=False)
    nc = big.n_cells
    h1 = torch.empty(n, dtype=torch.int32, device=dev); h2 = torch.empty_like(h1)
    mp_ = torch.empty(n, dtype=torch.float64, device=dev); mt = torch.empty_like(mp_)
    ds = torch.empty((n, coh.n_hla), dtype=torch.float64, device=dev)
    pp = torch.empty((n, nc), dtype=torch.float64, device=dev)

    def run():
        big.predict_device(g.data_ptr(), n, h1.data_ptr(), h2.data_ptr(), mp_.data_ptr(), mt.data_ptr(),
                           ds.data_ptr(), pp.data_ptr(), stream=torch.cuda.current_stream().cuda_stream,
                           sync=True)
    warm = n        # full-size warm-up: every tile buffer is allocated before the timed pass
    big.predict_device(g.data_ptr(), warm, h1.data_ptr(), h2.data_ptr(), mp_.data_ptr(), mt.data_ptr(),
                       ds.data_ptr(), pp.data_ptr(), stream=torch.cuda.current_stream().cuda_stream, sync=True)
    s0 = big.predict_stats()
    torch.cuda.synchronize(); hd.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    run()
    e1.record()
    torch.cuda.synchronize(); hd.barrier()
    ms = hd.max_over_ranks(e0.elapsed_time(e1), dev)
    s1 = big.predict_stats()
    d = {k: s1[k] - s0[k] for k in s1}
    # end to end with host buffers (H2D of raw genotypes, D2H of every output inside); one untimed
    # call first, as for every timed region here: it page-locks the result blocks the library recycles
    res = big.predict(host, want_prob=True, want_dosage=True)
    del res
    torch.cuda.synchronize(); hd.barrier()
    t0 = time.time()
    res = big.predict(host, want_prob=True, want_dosage=True)
    torch.cuda.synchronize(); hd.barrier()
    ms2 = hd.max_over_ranks((time.time() - t0) * 1e3, dev)
    same = bool(np.array_equal(res["h1"], h1.cpu().numpy()))
    rate = d["pair_evals"] / max(d["cell_kernel_ms"] * 1e-3, 1e-12)
    out = {
        "metric": "samples/s predicted (100 classifiers, type=response+prob)", "value": n_total / (ms * 1e-3),
        "unit": "samples/s", "samples": n_total, "ms": ms,
        "e2e": {"value": n_total / (ms2 * 1e-3), "unit": "samples/s", "h2d_bytes": int(host.nbytes),
                "d2h_bytes": int(n * (8 + 16 + 8 * coh.n_hla + 8 * nc)), "api": "HLAModel.predict (host numpy in, page-locked numpy out; 1 warm-up call)"},
        "model": "100 classifiers = the %d classifiers trained above, cycled" % n_src,
        # each distinct packed genotype of a tile is scored once per classifier (exact: equal inputs,
        # equal bits); the roofline below counts EXECUTED pair evaluations only
        "distinct_genotype_fraction": d["positions_scored"] / max(d["positions_total"], 1),
        "pair_evals_executed": int(d["pair_evals"]), "pair_evals_reference": int(d["pair_evals_nominal"]),
        "calls_equal_between_paths": same,
        "roofline": {"bound": "popc", "kernel": "cell_pass_kernel",
                     "achieved": d["popc32_issued"] / max(d["cell_kernel_ms"] * 1e-3, 1e-12) / 1e9,
                     "achieved_reference_formulation": rate * 4 / 1e9,
                     "peak": 148 * 16 * 1.965, "unit": "Gpopc32/s",
                     "frac": d["popc32_issued"] / max(d["cell_kernel_ms"] * 1e-3, 1e-12) / (148 * 16 * 1.965e9),
                     "pair_evals_per_s": rate, "launches": int(d["cell_kernel_launches"]),
                     "avg_launch_ms": d["cell_kernel_ms"] / max(d["cell_kernel_launches"], 1),
                     "cell_kernel_share_of_step": d["cell_kernel_ms"] / max(d["gpu_kernel_ms"], 1e-9)},
        "gpu_launches": int(d["kernel_launches"]),
    }
    if world > 1:
        # the other multi-GPU mode (SURVEY.md 8e, config 5's shape of work): classifiers sharded, every
        # rank scores ALL samples with its share of the classifiers, one NCCL all-reduce of the
        # [tile, n_cells + 3] fp64 partial posterior sums per tile of up to 262,144 samples, then finalise
        sub = hd.sub_model(big, rank, world)
        wts = torch.from_numpy(big.snp_weights()).to(dev)
        g_all = torch.from_numpy(all_host).to(dev)
        hd.predict_classifier_sharded(sub, wts, coh.n_hla, g_all)                 # full-size warm-up (tile buffers, NCCL)
        torch.cuda.synchronize(); hd.barrier()
        tm = {}
        e0.record()
        res_cs = hd.predict_classifier_sharded(sub, wts, coh.n_hla, g_all, timing=tm)
        e1.record()
        torch.cuda.synchronize(); hd.barrier()
        ms3 = hd.max_over_ranks(e0.elapsed_time(e1), dev)
        same_cs = bool(torch.equal(res_cs["h1"][b:e], h1) and torch.equal(res_cs["h2"][b:e], h2))
        rel, worst = 0.0, None
        if n > 0:
            # relative to the larger of the two values (entries below 1e-290 are in or next to the
            # denormal range, where fp64 has no relative precision left: compared absolutely)
            x, y = res_cs["postprob"][b:e], pp
            den = torch.maximum(x.abs(), y.abs()).clamp_min(1e-290)
            r = (x - y).abs() / den
            rel = float(r.max().item())
            k = int(r.argmax().item())
            worst = [float(x.flatten()[k].item()), float(y.flatten()[k].item()), k // nc, k % nc]
        out.update({"sharded_by_classifier_value": n_total / (ms3 * 1e-3), "sharded_by_classifier_ms": ms3,
                    "allreduce_ms": tm["allreduce_ms"], "allreduce_bytes": int(tm["allreduce_bytes"]),
                    "allreduce_gbs": tm["allreduce_bytes"] / max(tm["allreduce_ms"] * 1e-3, 1e-12) / 1e9,
                    "sharded_by_classifier_calls_equal": same_cs, "sharded_by_classifier_max_rel_err": rel,
                    "sharded_by_classifier_worst_entry": worst,
                    "sharded_by_classifier_note": "%d classifiers per rank x all %d samples; one NCCL all-reduce (fp64 sum) of "
                                                  "[min(n, 262144), %d] per tile; calls equal / posteriors vs the sample-sharded "
                                                  "(sequential classifier order) result on this rank's slice" % (
                                                      sub.num_classifiers(), n_total, nc + 3)})
    return out


def bench_large_list(api, torch, dev, popc_peak, n_hap=15000, n_hla=40, n_snp=30, n_samp=16384):
    """The global-memory operand path of the pair-scoring kernel (cell_pass_kernel<.., SMEM=false>): a
    haplotype list too large for an SM's shared memory (15,000 x 16 B = 240 KB > 227 KB; the DRB1-scale
    lists of configs[3] still fit), scored for every cell of 16,384 samples through the prediction
    entry point. The records are then read with __ldg through L1/L2 instead of LDS.128 broadcasts."""
    rng = np.random.default_rng(5)
    lens = rng.multinomial(n_hap - n_hla, rng.dirichlet(np.ones(n_hla))) + 1
    hla = np.repeat(np.arange(n_hla), lens).astype(np.int32)
    packed = rng.integers(0, 2 ** 63, size=(n_hap, 2), dtype=np.int64).astype(np.uint64) & np.uint64((1 << n_snp) - 1)
    freq = rng.random(n_hap) ** 3 + 1e-7
    freq /= freq.sum()
    m = api.HLAModel(n_snp, n_hla)
    m.add_classifier(np.arange(n_snp, dtype=np.int32), freq, hla, packed)
    geno = rng.integers(0, 3, size=(n_samp, n_snp)).astype(np.int8)
    g = torch.from_numpy(geno).to(dev)
    h1 = torch.empty(n_samp, dtype=torch.int32, device=dev); h2 = torch.empty_like(h1)
    mp_ = torch.empty(n_samp, dtype=torch.float64, device=dev); mt = torch.empty_like(mp_)

    def run(n):
        m.predict_device(g.data_ptr(), n, h1.data_ptr(), h2.data_ptr(), mp_.data_ptr(), mt.data_ptr(), 0, 0,
                         stream=torch.cuda.current_stream().cuda_stream, sync=True)
    run(2048)
    s0 = m.predict_stats()
    run(n_samp)
    s1 = m.predict_stats()
    d = {k: s1[k] - s0[k] for k in s1}
    rate = d["popc32_issued"] / max(d["cell_kernel_ms"] * 1e-3, 1e-12)
    return {"kernel": "cell_pass_kernel<SMEM=false>", "n_hap": n_hap, "n_hla": n_hla, "n_snp": n_snp, "samples": n_samp,
            "list_bytes": n_hap * 16, "pair_evals": int(d["pair_evals"]), "cell_kernel_ms": d["cell_kernel_ms"],
            "achieved": rate / 1e9, "peak": popc_peak / 1e9, "unit": "Gpopc32/s", "frac": rate / popc_peak,
            "note": "haplotype records from global memory (__ldg, L1/L2) because the list exceeds the 227 KB of shared "
                    "memory; same 6-instruction pair evaluation otherwise"}


def bench_bed_decode(api, torch, dev, n_samp=N_PREDICT, n_snp=4096, reps=5):
    """GB/s of the SNP-major BED decode kernel on a synthetic file of the prediction cohort's size
    (200,000 samples x 4,096 SNPs: 205 MB packed in, 819 MB int8 out -- larger than L2), device
    resident, and end to end from a host byte string"""
    bps = (n_samp + 3) // 4
    gen = torch.Generator(device=dev); gen.manual_seed(5)
    payload = torch.randint(0, 256, (n_snp * bps,), dtype=torch.uint8, device=dev, generator=gen)
    out = torch.empty((n_samp, n_snp), dtype=torch.int8, device=dev)
    st = torch.cuda.current_stream().cuda_stream

    def run():
        api.bed_decode_device(payload.data_ptr(), 1, n_samp, n_snp, 0, n_snp, out.data_ptr(), st)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    alg_bytes = n_snp * bps + n_samp * n_snp            # 0.25 B read + 1 B written per genotype
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    else:
        peak, src = 6650.0, "of fallback (B200_PROFILING.md: 6.65 TB/s)"
    # spot check against an independent decoding of a corner of the file
    cvt = torch.tensor([2, -1, 1, 0], dtype=torch.int8, device=dev)
    want = cvt[((payload.view(n_snp, bps)[:64, :16].long().unsqueeze(-1) >> torch.tensor([0, 2, 4, 6], device=dev)) & 3)]
    ok = bool(torch.equal(out[:64, :64], want.reshape(64, 64).t().contiguous()))
    host = np.concatenate([np.array([0x6C, 0x1B, 1], dtype=np.uint8), payload.cpu().numpy()])
    t0 = time.time()
    g = api.bed_decode(host, n_samp, n_snp)
    e2e_ms = (time.time() - t0) * 1e3
    achieved = alg_bytes / (ms * 1e-3) / 1e9
    return {"metric": "PLINK BED decode (SNP-major, 200,000 samples x 4,096 SNPs)", "genotypes_per_s": n_samp * n_snp / (ms * 1e-3),
            "ms": ms, "correct": ok and bool(np.array_equal(g[:64, :64], out[:64, :64].cpu().numpy())),
            "roofline": {"bound": "hbm", "kernel": "bed_snp_major_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": None, "peak_source": src,
                         "algorithmic_bytes": alg_bytes},
            "e2e": {"ms": e2e_ms, "genotypes_per_s": n_samp * n_snp / (e2e_ms * 1e-3),
                    "h2d_bytes": int(host.nbytes), "d2h_bytes": int(n_samp * n_snp),
                    "api": "hibag_b200_bed_decode on a host byte string (pageable), int8 matrix back"}}


def main():
    # exactly ONE line on stdout (the JSON): libraries that print to fd 1 (NCCL's version banner)
    # go to stderr instead
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--threads", type=int, default=0, help="host threads per rank (default cores/ranks)")
    ap.add_argument("--lanes", type=int, default=0, help="classifiers in flight per GPU (0: %d, fewer on "
                    "boxes with few host cores per GPU)" % LANES)
    ap.add_argument("--host-em", action="store_true", help="candidate EM on the host thread pool")
    ap.add_argument("--predict-samples", type=int, default=N_PREDICT)
    ap.add_argument("--no-predict", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-unscreened", action="store_true", help="skip the extra step with screening off")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--cpu-procs", type=int, default=0)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
