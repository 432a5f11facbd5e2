"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against the oracle,
the compiled reference and the reference's golden model. Integers are bit-exact; fp64 outputs
are required to agree within 1e-10 relative (north_star) and in practice are bit-identical, which
is asserted too where the operation order is the reference's."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import refpy
from tests import helpers

pytestmark = pytest.mark.gpu

SNP_COUNTS = [1, 7, 31, 32, 33, 64, 65, 96, 128]
RTOL = 1e-10      # tolerance stated by BASELINE.json north_star for fp64 outputs


def rel_close(a, b):
    a, b = np.asarray(a), np.asarray(b)
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return False
    m = ~np.isnan(a)
    return np.all(np.abs(a[m] - b[m]) <= RTOL * np.maximum(np.abs(a[m]), np.abs(b[m])))


# ---- kernel level -------------------------------------------------------------------------------

@pytest.mark.parametrize("n_snp", SNP_COUNTS)
def test_scoring_matches_oracle(gpu, orc, n_snp):
    rng = np.random.default_rng(2000 + n_snp)
    h, n_hla, _ = helpers.random_haplo_list(rng, 13, n_snp)
    g = helpers.random_genotypes(rng, 777, n_snp, n_hla, haplo=h)
    a1, a2 = gpu.best_guess(h, n_hla, n_snp, g)
    o1, o2 = orc.best_guess(h, n_hla, n_snp, g)
    assert np.array_equal(a1, o1) and np.array_equal(a2, o2)
    assert np.array_equal(gpu.post_prob(h, n_hla, n_snp, g), orc.post_prob(h, n_hla, n_snp, g),
                          equal_nan=True)
    p, s = gpu.post_prob2(h, n_hla, n_snp, g)
    q, t = orc.post_prob2(h, n_hla, n_snp, g)
    assert np.array_equal(s, t) and np.array_equal(p, q, equal_nan=True)


def test_scoring_edge_cases(gpu, orc):
    rng = np.random.default_rng(9)
    # a single genotype, a single allele, a single haplotype
    h = refpy.make_haplo(np.array([[5, 0]], dtype=np.uint64), [1.0], [0])
    g = helpers.random_genotypes(rng, 1, 3, 1)
    assert np.array_equal(gpu.post_prob2(h, 1, 3, g)[0], orc.post_prob2(h, 1, 3, g)[0])
    # all SNPs missing
    h, n_hla, n_snp = helpers.random_haplo_list(rng, 6, 20)
    g = refpy.pack_geno(np.full((40, 20), -1), boot=np.zeros(40), a1=np.zeros(40, int), a2=np.ones(40, int))
    p, s = gpu.post_prob2(h, n_hla, n_snp, g)
    q, t = orc.post_prob2(h, n_hla, n_snp, g)
    assert np.array_equal(p, q) and np.array_equal(s, t)
    # everything scores exactly zero -> (NA, NA) and NaN ratio
    h0 = refpy.make_haplo(np.zeros((2, 2), dtype=np.uint64), [0.5, 0.5], [0, 1])
    g0 = refpy.pack_geno(np.full((5, 128), 2), boot=np.zeros(5), a1=np.zeros(5, int), a2=np.ones(5, int))
    a1, a2 = gpu.best_guess(h0, 2, 128, g0)
    assert np.all(a1 == refpy.NA_INTEGER) and np.all(a2 == refpy.NA_INTEGER)
    assert np.all(np.isnan(gpu.post_prob(h0, 2, 128, g0)))
    # denormal products are honoured (distances 62..64)
    g1 = refpy.pack_geno(np.concatenate([np.full((1, 32), 2), np.full((1, 96), -1)], axis=1),
                         boot=[0], a1=[0], a2=[1])
    assert np.array_equal(gpu.post_prob2(h0, 2, 128, g1)[1], orc.post_prob2(h0, 2, 128, g1)[1])


def test_scoring_large_list_many_alleles(gpu, orc):
    """DRB1-like shape: 60 alleles, 2 words, long list (global-memory operand path when the list
    exceeds shared memory is exercised by the 9000-haplotype case)"""
    rng = np.random.default_rng(31)
    h, n_hla, n_snp = helpers.random_haplo_list(rng, 60, 100, max_per_allele=12, empty_frac=0.1)
    g = helpers.random_genotypes(rng, 300, n_snp, n_hla, haplo=h)
    a1, a2 = gpu.best_guess(h, n_hla, n_snp, g)
    o1, o2 = orc.best_guess(h, n_hla, n_snp, g)
    assert np.array_equal(a1, o1) and np.array_equal(a2, o2)
    h, n_hla, n_snp = helpers.random_haplo_list(rng, 3, 70, max_per_allele=4000, empty_frac=0.0)
    g = helpers.random_genotypes(rng, 40, n_snp, n_hla, haplo=h)
    assert np.array_equal(gpu.post_prob(h, n_hla, n_snp, g), orc.post_prob(h, n_hla, n_snp, g),
                          equal_nan=True)


# ---- the drop-in plugin driven by the reference's own host code -------------------------------------

def test_reference_host_with_gpu_hooks_reproduces_golden(gpu, ref):
    """the reference's BuildClassifiers (compiled, unmodified) calling OUR ten hooks:
    set.seed(100) -> the first 30 classifiers of inst/extdata/ModelList.RData bit for bit"""
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    m = ref.new_model()
    m.init_training(geno, h1, h2, len(al))
    ref.set_seed(int(ml["seed"]))
    ref.set_gpu_procs(gpu.get_procs())
    try:
        m.build(30, int(ml["mtry"]))
    finally:
        ref.set_gpu_procs(None)
    for k in range(30):
        helpers.assert_classifier_equals_golden(m.classifier(k), ml, k)


def test_reference_predict_with_gpu_hook_matches_cpu(gpu, ref):
    """the reference's PredictHLA calling predict_init / predict_avg_prob / predict_done"""
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    m = ref.new_model()
    m.init_predict(geno.shape[1], geno.shape[0], len(al))
    for k in range(100):
        c = helpers.golden_classifier(ml, k)
        m.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"], acc=c["oob_acc"])
    test = geno.astype(np.int32).copy()
    rng = np.random.default_rng(8)
    test[rng.random(test.shape) < 0.04] = -1
    test[5, :] = -1
    cpu = m.predict(test)
    ref.set_gpu_procs(gpu.get_procs())
    try:
        dev = m.predict(test)
    finally:
        ref.set_gpu_procs(None)
    assert np.array_equal(cpu["h1"], dev["h1"]) and np.array_equal(cpu["h2"], dev["h2"])
    for key in ("prob", "matching", "dosage", "postprob"):
        assert rel_close(cpu[key], dev[key]), key
        assert np.array_equal(cpu[key], dev[key], equal_nan=True), key + " (bit-exact)"


def test_haplomatch_hook_matches_oracle(gpu, orc):
    """build_haplomatch body on the GPU == the CPU restatement, record for record"""
    rng = np.random.default_rng(123)
    for n_snp, n_hla, per in ((0, 5, 1), (5, 8, 4), (40, 12, 9), (64, 6, 30), (65, 9, 12), (128, 7, 20)):
        haplo, n_hla, _ = helpers.random_haplo_list(rng, n_hla=n_hla, n_snp=max(n_snp, 1), max_per_allele=per)
        geno = helpers.random_genotypes(rng, 700, max(n_snp, 1), n_hla, haplo=haplo)
        if n_snp == 0:
            geno["s1"][:] = 0; geno["s2"][:] = np.uint64(0xFFFFFFFFFFFFFFFF)
        a1 = np.minimum(geno["a1"], geno["a2"]); a2 = np.maximum(geno["a1"], geno["a2"])
        geno["a1"], geno["a2"] = a1, a2
        lens = np.bincount(haplo["hla"], minlength=n_hla)
        got = gpu.haplomatch(haplo, lens, n_snp, geno)
        want = orc.haplomatch_records(haplo, lens, n_snp, geno)
        assert got.shape == want.shape and np.array_equal(got, want), n_snp


def test_reference_host_with_haplomatch_hook(gpu, ref):
    """the unmodified reference host driven through ALL hooks including build_haplomatch: pair
    lists arrive in record order, so EM sums differ from the CPU path in the last bits only --
    same SNPs, same haplotypes, frequencies within 1e-9 relative (see include/hibag_b200.h)"""
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    try:
        ref.set_gpu_procs(gpu.get_procs(with_haplomatch=True))
        r = ref.new_model()
        r.init_training(geno, h1, h2, len(al))
        ref.set_seed(int(ml["seed"]))
        r.build(6, int(ml["mtry"]), prune=True)
        got = [r.classifier(k) for k in range(6)]
    finally:
        ref.set_gpu_procs(None)
    for k in range(6):
        g = helpers.golden_classifier(ml, k)
        assert np.array_equal(got[k]["snpidx"], g["snpidx"]), k
        assert np.array_equal(got[k]["packed"], g["packed"]) and np.array_equal(got[k]["hla"], g["hla"]), k
        assert np.allclose(got[k]["freq"], g["freq"], rtol=1e-9, atol=0), k


# ---- own host driver ------------------------------------------------------------------------------

def test_trainer_reproduces_reference_golden_model(gpu):
    """own C++ driver + CUDA scoring, set.seed(100), 100 classifiers: every bootstrap sample, SNP
    set, haplotype list, frequency and OOB accuracy of ModelList.RData bit for bit"""
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    m = gpu.HLAModel(geno.shape[1], len(al), al)
    m.set_training(geno, h1, h2)
    m.train(100, int(ml["mtry"]), prune=True, seed=int(ml["seed"]), n_threads=8, em_on_device=True)
    assert m.num_classifiers() == 100
    for k in range(100):
        helpers.assert_classifier_equals_golden(m.classifier(k), ml, k)
    st = m.train_stats()
    assert st["pair_evals"] > 0 and st["kernel_launches"] > 0


def test_trainer_legacy_hook_mode_and_thread_count_invariance(gpu):
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    for kwargs in (dict(use_legacy_hooks=True, n_threads=3), dict(n_threads=1),
                   dict(n_threads=4, em_on_device=False)):
        m = gpu.HLAModel(geno.shape[1], len(al), al)
        m.set_training(geno, h1, h2)
        m.train(8, int(ml["mtry"]), prune=True, seed=int(ml["seed"]), **kwargs)
        for k in range(8):
            helpers.assert_classifier_equals_golden(m.classifier(k), ml, k)


@pytest.mark.parametrize("chain", ["1", "0"])
def test_both_device_em_kernels_reproduce_the_reference(gpu, monkeypatch, chain):
    """em_chain_kernel (128 threads, frequencies and scale factors in shared memory, both steps as walks
    over chains of 4-byte records) and em_kernel (the streaming form, also for rounds too large for an
    SM) are bit-identical to the reference's CAlg_EM: the golden model, and the seeded many-allele
    cohort whose classifiers the compiled reference trained"""
    from hibag_b200 import synth
    monkeypatch.setenv("HIBAG_B200_EM_CHAIN", chain)
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    m = gpu.HLAModel(geno.shape[1], len(al), al)
    m.set_training(geno, h1, h2)
    m.train(12, int(ml["mtry"]), prune=True, seed=int(ml["seed"]), n_threads=4, em_on_device=True)
    for k in range(12):
        helpers.assert_classifier_equals_golden(m.classifier(k), ml, k)
    assert m.train_stats()["em_iterations"] > 0
    gd = helpers.load_golden("synth_many_alleles_ref.npz")
    coh = synth.make_cohort(int(gd["n_samp"]), int(gd["n_snp"]), int(gd["n_hla"]), seed=int(gd["cohort_seed"]))
    for lanes in (1, 2):
        s = gpu.HLAModel(coh.n_snp, coh.n_hla)
        s.set_training(coh.geno, coh.h1, coh.h2)
        s.train(int(gd["n_cls"]), gpu.default_mtry(coh.n_snp), prune=True, seed=int(gd["train_seed"]),
                per_classifier_seed=True, n_concurrent=lanes)
        for k in range(int(gd["n_cls"])):
            want = dict(snpidx=gd["c%d_snpidx" % k], samp_num=gd["c%d_samp_num" % k], freq=gd["c%d_freq" % k],
                        hla=gd["c%d_hla" % k], packed=gd["c%d_packed" % k], oob_acc=float(gd["c%d_oob_acc" % k]))
            assert helpers.classifier_diff(s.classifier(k), want) == "", (chain, lanes, k)


def test_device_em_host_fallback_path(gpu, monkeypatch):
    """candidates whose stopping test the device cannot decide are re-estimated on the host:
    force that path for every third candidate and require the golden model all the same"""
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    monkeypatch.setenv("HIBAG_B200_EM_FORCE_FALLBACK", "3")
    m = gpu.HLAModel(geno.shape[1], len(al), al)
    m.set_training(geno, h1, h2)
    m.train(6, int(ml["mtry"]), prune=True, seed=int(ml["seed"]), n_threads=4, em_on_device=True)
    for k in range(6):
        helpers.assert_classifier_equals_golden(m.classifier(k), ml, k)
    assert m.train_stats()["n_em_host_fallback"] > 0


def _synthetic():
    from hibag_b200 import synth
    return synth.make_cohort(400, 120, 12, seed=3)


def test_trainer_matches_reference_on_synthetic(gpu, ref):
    """per-classifier seeding (the multi-GPU convention), prune on and off, vs reference 'base'"""
    coh = _synthetic()
    mtry = gpu.default_mtry(coh.n_snp)
    for prune in (True, False):
        r = ref.new_model()
        r.init_training(coh.geno, coh.h1, coh.h2, coh.n_hla)
        r.build(3, mtry, prune=prune, reseed_base=500, first_index=0)
        m = gpu.HLAModel(coh.n_snp, coh.n_hla)
        m.set_training(coh.geno, coh.h1, coh.h2)
        m.train(3, mtry, prune=prune, seed=500, per_classifier_seed=True, n_threads=6)
        for k in range(3):
            d = helpers.classifier_diff(m.classifier(k), r.classifier(k))
            assert d == "", (prune, k, d)
        h = gpu.HLAModel(coh.n_snp, coh.n_hla)
        h.set_training(coh.geno, coh.h1, coh.h2)
        h.train(3, mtry, prune=prune, seed=500, per_classifier_seed=True, n_threads=6, em_on_device=False)
        for k in range(3):
            assert helpers.classifier_diff(h.classifier(k), r.classifier(k)) == "", (prune, k, "host EM")
    # sharding: classifiers {1, 3} built as a strided shard equal classifiers 1 and 3 of a full run
    full = gpu.HLAModel(coh.n_snp, coh.n_hla); full.set_training(coh.geno, coh.h1, coh.h2)
    full.train(4, mtry, seed=500, per_classifier_seed=True)
    shard = gpu.HLAModel(coh.n_snp, coh.n_hla); shard.set_training(coh.geno, coh.h1, coh.h2)
    shard.train(2, mtry, seed=500, per_classifier_seed=True, first_index=1, index_stride=2)
    for j, k in enumerate((1, 3)):
        assert helpers.classifier_diff(shard.classifier(j), full.classifier(k)) == ""
    # several classifiers in flight on one GPU: same model, in global classifier order
    conc = gpu.HLAModel(coh.n_snp, coh.n_hla); conc.set_training(coh.geno, coh.h1, coh.h2)
    conc.train(4, mtry, seed=500, per_classifier_seed=True, n_concurrent=3, n_threads=6)
    assert conc.num_classifiers() == 4
    for k in range(4):
        assert helpers.classifier_diff(conc.classifier(k), full.classifier(k)) == ""
    st = conc.train_stats()
    assert st["n_oob_evals"] == full.train_stats()["n_oob_evals"]
    assert st["pair_evals"] == full.train_stats()["pair_evals"]


def test_exact_screening_changes_nothing_but_the_work(gpu, monkeypatch):
    """exact screening of the training passes: with it and without it the same
    classifiers and the same search trace; it skips pair evaluations; samples whose in-bag sum the
    screen cannot certify take the rescoring path (forced here) with the same result"""
    from hibag_b200 import synth
    coh = synth.make_cohort(700, 140, 25, seed=5)
    mtry = gpu.default_mtry(coh.n_snp)

    def run(**kw):
        m = gpu.HLAModel(coh.n_snp, coh.n_hla)
        m.set_training(coh.geno, coh.h1, coh.h2)
        m.train(3, mtry, prune=True, seed=77, per_classifier_seed=True, n_threads=6, **kw)
        return m

    plain, scr = run(screening=False), run(screening=True)
    for k in range(3):
        assert helpers.classifier_diff(scr.classifier(k), plain.classifier(k)) == "", k
    assert np.array_equal(scr.train_trace(), plain.train_trace())
    sp, ss = plain.train_stats(), scr.train_stats()
    assert sp["pair_evals"] == sp["pair_evals_nominal"] == ss["pair_evals_nominal"]
    assert 0 < ss["pair_evals"] < sp["pair_evals"]
    assert sp["n_screen_fallback"] == 0
    # the second level of the in-bag screen (class bound over two heterozygous SNPs; opt-in) only removes
    # work, also with the rescue path of the reduction forced
    monkeypatch.setenv("HIBAG_B200_SCREEN_REFINE", "1")
    lvl2 = run(screening=True)
    monkeypatch.setenv("HIBAG_B200_SCREEN_FORCE_RESCUE", "5")
    lvl2r = run(screening=True, n_concurrent=2)
    monkeypatch.delenv("HIBAG_B200_SCREEN_FORCE_RESCUE")
    monkeypatch.delenv("HIBAG_B200_SCREEN_REFINE")
    for k in range(3):
        assert helpers.classifier_diff(lvl2.classifier(k), plain.classifier(k)) == "", k
        assert helpers.classifier_diff(lvl2r.classifier(k), plain.classifier(k)) == "", k
    assert np.array_equal(lvl2.train_trace(), plain.train_trace())
    assert lvl2.train_stats()["pair_evals"] < ss["pair_evals"] < sp["pair_evals"]
    # uncertified sums: rescued inside the reduction (every 5th position forced), or -- the host's
    # safety net for a ratio the device reports as uncertified -- rescored with the plain kernel
    for var in ("HIBAG_B200_SCREEN_FORCE_RESCUE", "HIBAG_B200_SCREEN_FORCE_FALLBACK"):
        monkeypatch.setenv(var, "5")
        forced = run(screening=True, n_concurrent=2)
        monkeypatch.delenv(var)
        for k in range(3):
            assert helpers.classifier_diff(forced.classifier(k), plain.classifier(k)) == "", (var, k)
        assert forced.train_stats()["n_screen_fallback"] > ss["n_screen_fallback"], var


def test_position_classes_change_nothing_but_the_work(gpu, monkeypatch):
    """Positions of a candidate list whose samples carry the same packed genotype (candidate SNP
    included) and the same true type are screened and scored ONCE (kernels.h ScreenArgs::rep): same
    classifiers and search trace as with every position scored (HIBAG_B200_TRAIN_DEDUP=0) and as
    without screening, fewer pair evaluations -- also with the second screen level, the rescue path of
    the reduction and the host's rescoring fallback forced, and for two-word genotypes."""
    from hibag_b200 import synth
    coh = synth.make_cohort(900, 140, 12, seed=6, noise=0.002, missing=0.002)    # few founders: many repeats
    mtry = gpu.default_mtry(coh.n_snp)

    def run(**kw):
        m = gpu.HLAModel(coh.n_snp, coh.n_hla)
        m.set_training(coh.geno, coh.h1, coh.h2)
        m.train(3, mtry, prune=True, seed=78, per_classifier_seed=True, n_threads=6, **kw)
        return m

    plain = run(screening=False)
    monkeypatch.setenv("HIBAG_B200_TRAIN_DEDUP", "0")
    every = run(screening=True)
    monkeypatch.delenv("HIBAG_B200_TRAIN_DEDUP")
    cls = run(screening=True, n_concurrent=3)
    for k in range(3):
        assert helpers.classifier_diff(every.classifier(k), plain.classifier(k)) == "", k
        assert helpers.classifier_diff(cls.classifier(k), plain.classifier(k)) == "", k
    assert np.array_equal(cls.train_trace(), plain.train_trace())
    assert 0 < cls.train_stats()["pair_evals"] < 0.8 * every.train_stats()["pair_evals"]
    for env in ({"HIBAG_B200_SCREEN_REFINE": "1"}, {"HIBAG_B200_SCREEN_FORCE_RESCUE": "3"},
                {"HIBAG_B200_SCREEN_FORCE_FALLBACK": "1"}, {"HIBAG_B200_SCREEN_DEVICE_RESCUE": "1"},
                {"HIBAG_B200_GATHER_FLAT": "7"}):
        for k_, v_ in env.items():
            monkeypatch.setenv(k_, v_)
        m = run(screening=True, n_concurrent=2)
        for k_ in env:
            monkeypatch.delenv(k_)
        for k in range(3):
            assert helpers.classifier_diff(m.classifier(k), plain.classifier(k)) == "", (env, k)
    # two-word genotypes (37 SNPs): the thermometer cohort of the reference fixture, every position vs classes
    gd = helpers.load_golden("synth_thermo_ref.npz")
    geno, h1, h2 = synth.make_thermo_cohort(int(gd["s0_n_samp"]), int(gd["s0_n_hla"]),
                                            seed=int(gd["s0_cohort_seed"]), noise=float(gd["s0_noise"]))
    want = dict(snpidx=gd["s0_snpidx"], samp_num=gd["s0_samp_num"], freq=gd["s0_freq"],
                hla=gd["s0_hla"], packed=gd["s0_packed"], oob_acc=float(gd["s0_oob_acc"]))
    m = gpu.HLAModel(geno.shape[1], int(gd["s0_n_hla"]))
    m.set_training(geno, h1, h2)
    m.train(1, geno.shape[1], prune=True, seed=int(gd["s0_train_seed"]), per_classifier_seed=True)
    assert helpers.classifier_diff(m.classifier(0), want) == ""
    assert m.train_stats()["pair_evals"] < m.train_stats()["pair_evals_nominal"]


def test_trainer_matches_reference_many_alleles(gpu):
    """DRB1-like shape (many alleles, long haplotype lists, EM clusters of several CTAs):
    bit-identical classifiers vs the reference's base target (fixture generated from the compiled
    reference by tools/make_golden_synth.py)"""
    from hibag_b200 import synth
    gd = helpers.load_golden("synth_many_alleles_ref.npz")
    coh = synth.make_cohort(int(gd["n_samp"]), int(gd["n_snp"]), int(gd["n_hla"]), seed=int(gd["cohort_seed"]))
    n_cls = int(gd["n_cls"])
    for kwargs in (dict(n_concurrent=2), dict(em_on_device=False, n_threads=8)):
        m = gpu.HLAModel(coh.n_snp, coh.n_hla)
        m.set_training(coh.geno, coh.h1, coh.h2)
        m.train(n_cls, gpu.default_mtry(coh.n_snp), prune=True, seed=int(gd["train_seed"]),
                per_classifier_seed=True, **kwargs)
        for k in range(n_cls):
            want = dict(snpidx=gd["c%d_snpidx" % k], samp_num=gd["c%d_samp_num" % k], freq=gd["c%d_freq" % k],
                        hla=gd["c%d_hla" % k], packed=gd["c%d_packed" % k], oob_acc=float(gd["c%d_oob_acc" % k]))
            d = helpers.classifier_diff(m.classifier(k), want)
            assert d == "", (kwargs, k, d)


def test_trainer_matches_reference_many_snps(gpu):
    """Classifiers of 37 and 67 SNPs (two- and four-word packed genotypes; the BASELINE-shaped cohorts
    stop at 20-27 SNPs): bit-identical to the reference's base target with the screened passes and the
    device EM, and with every cell scored and the host EM (fixture generated from the compiled reference
    by tools/make_golden_thermo.py)"""
    from hibag_b200 import synth
    gd = helpers.load_golden("synth_thermo_ref.npz")
    for i in range(int(gd["n_specs"])):
        geno, h1, h2 = synth.make_thermo_cohort(int(gd["s%d_n_samp" % i]), int(gd["s%d_n_hla" % i]),
                                                seed=int(gd["s%d_cohort_seed" % i]), noise=float(gd["s%d_noise" % i]))
        want = dict(snpidx=gd["s%d_snpidx" % i], samp_num=gd["s%d_samp_num" % i], freq=gd["s%d_freq" % i],
                    hla=gd["s%d_hla" % i], packed=gd["s%d_packed" % i], oob_acc=float(gd["s%d_oob_acc" % i]))
        assert len(want["snpidx"]) > (32, 64)[i]
        for kwargs in (dict(n_concurrent=1), dict(em_on_device=False, n_threads=8, screening=False)):
            m = gpu.HLAModel(geno.shape[1], int(gd["s%d_n_hla" % i]))
            m.set_training(geno, h1, h2)
            m.train(1, geno.shape[1], prune=True, seed=int(gd["s%d_train_seed" % i]), per_classifier_seed=True, **kwargs)
            d = helpers.classifier_diff(m.classifier(0), want)
            assert d == "", (i, kwargs, d)
            if "screening" not in kwargs:
                st = m.train_stats()
                assert st["pair_evals"] < st["pair_evals_nominal"]          # the screen was on


@pytest.mark.parametrize("flat", ["0", "1", "2", "7"])
def test_gather_forms_reproduce_the_reference(gpu, monkeypatch, flat):
    """the forms of the screened pair-scoring kernel -- (cell, block of positions) tasks with
    warp-uniform cell bounds, CTAs walking all lists of the launch (0) or serving one list each (2, the
    default for the out-of-bag passes), and entry-flat tasks where every lane walks the chain of its own
    (cell, position) entry (1; 7 = for both passes) -- all produce the reference's classifiers.
    One-, two- and four-word genotypes (golden HapMap model, 37- and 67-SNP cohorts), many alleles."""
    from hibag_b200 import synth
    monkeypatch.setenv("HIBAG_B200_GATHER_FLAT", flat)
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    m = gpu.HLAModel(geno.shape[1], len(al), al)
    m.set_training(geno, h1, h2)
    m.train(8, int(ml["mtry"]), prune=True, seed=int(ml["seed"]), n_threads=4)
    for k in range(8):
        helpers.assert_classifier_equals_golden(m.classifier(k), ml, k)
    st = m.train_stats()
    assert st["pair_evals"] < st["pair_evals_nominal"]
    gd = helpers.load_golden("synth_thermo_ref.npz")
    for i in range(int(gd["n_specs"])):
        g2, a1, a2 = synth.make_thermo_cohort(int(gd["s%d_n_samp" % i]), int(gd["s%d_n_hla" % i]),
                                              seed=int(gd["s%d_cohort_seed" % i]), noise=float(gd["s%d_noise" % i]))
        want = dict(snpidx=gd["s%d_snpidx" % i], samp_num=gd["s%d_samp_num" % i], freq=gd["s%d_freq" % i],
                    hla=gd["s%d_hla" % i], packed=gd["s%d_packed" % i], oob_acc=float(gd["s%d_oob_acc" % i]))
        s = gpu.HLAModel(g2.shape[1], int(gd["s%d_n_hla" % i]))
        s.set_training(g2, a1, a2)
        s.train(1, g2.shape[1], prune=True, seed=int(gd["s%d_train_seed" % i]), per_classifier_seed=True)
        assert helpers.classifier_diff(s.classifier(0), want) == "", (flat, i)
    gm = helpers.load_golden("synth_many_alleles_ref.npz")
    coh = synth.make_cohort(int(gm["n_samp"]), int(gm["n_snp"]), int(gm["n_hla"]), seed=int(gm["cohort_seed"]))
    t = gpu.HLAModel(coh.n_snp, coh.n_hla)
    t.set_training(coh.geno, coh.h1, coh.h2)
    t.train(int(gm["n_cls"]), gpu.default_mtry(coh.n_snp), prune=True, seed=int(gm["train_seed"]),
            per_classifier_seed=True, n_concurrent=2)
    for k in range(int(gm["n_cls"])):
        want = dict(snpidx=gm["c%d_snpidx" % k], samp_num=gm["c%d_samp_num" % k], freq=gm["c%d_freq" % k],
                    hla=gm["c%d_hla" % k], packed=gm["c%d_packed" % k], oob_acc=float(gm["c%d_oob_acc" % k]))
        assert helpers.classifier_diff(t.classifier(k), want) == "", (flat, k)


def _golden_model(gpu, ref, n_cls=100):
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    m = gpu.HLAModel(geno.shape[1], len(al), al)
    r = ref.new_model()
    r.init_predict(geno.shape[1], geno.shape[0], len(al))
    for k in range(n_cls):
        c = helpers.golden_classifier(ml, k)
        m.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"], oob_acc=c["oob_acc"])
        r.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"], acc=c["oob_acc"])
    return m, r, geno


def test_batched_predict_matches_reference(gpu, ref):
    m, r, geno = _golden_model(gpu, ref)
    rng = np.random.default_rng(21)
    test = np.tile(geno, (5, 1)).astype(np.int8)
    test[rng.random(test.shape) < 0.05] = -1
    test[7, :] = -1                                    # no usable SNP: (NA, NA), prob 0, NaN matching
    test[11, ::2] = 3                                  # out-of-range codes are missing too
    want = r.predict(test.astype(np.int32))
    got = gpu.hlaPredict(m, test, type="response+prob")
    assert np.array_equal(want["h1"], got["h1"]) and np.array_equal(want["h2"], got["h2"])
    assert got["h1"][7] == refpy.NA_INTEGER and got["prob"][7] == 0 and np.isnan(got["matching"][7])
    for key in ("prob", "matching", "dosage", "postprob"):
        assert rel_close(want[key], got[key]), key
        assert np.array_equal(want[key], got[key], equal_nan=True), key + " (bit-exact)"


def test_predict_on_synthetic_model_matches_reference(gpu, ref):
    from hibag_b200 import synth
    coh = _synthetic()
    mtry = gpu.default_mtry(coh.n_snp)
    m = gpu.HLAModel(coh.n_snp, coh.n_hla); m.set_training(coh.geno, coh.h1, coh.h2)
    m.train(6, mtry, seed=900, per_classifier_seed=True)
    r = ref.new_model(); r.init_predict(coh.n_snp, coh.n_samp, coh.n_hla)
    for k in range(6):
        c = m.classifier(k)
        r.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"], acc=c["oob_acc"])
    new = synth.draw_more(coh, 3000, seed=77)
    want = r.predict(new.geno.astype(np.int32))
    got = m.predict(new.geno)
    assert np.array_equal(want["h1"], got["h1"]) and np.array_equal(want["h2"], got["h2"])
    for key in ("prob", "matching", "dosage", "postprob"):
        assert np.array_equal(want[key], got[key], equal_nan=True), key
    acc = np.mean((np.minimum(got["h1"], got["h2"]) == np.minimum(new.h1, new.h2)) &
                  (np.maximum(got["h1"], got["h2"]) == np.maximum(new.h1, new.h2)))
    assert acc > 0.5          # sanity: the ensemble actually predicts


@pytest.mark.parametrize("n_snp", [20, 50, 100])
def test_predict_scores_each_distinct_genotype_once_and_exactly(gpu, ref, monkeypatch, n_snp):
    """Prediction scores every DISTINCT packed genotype of a tile once per classifier and shares its
    column of the cell matrix (kernels.h: launch_dedup_genotypes). Same bits as scoring every sample
    (HIBAG_B200_PREDICT_DEDUP=0) and as the reference's per-sample loop (src/LibHLA.cpp:2451-2464),
    for 1-, 2- and 4-word genotypes, duplicates inside and across tiles, missing and all-missing rows."""
    rng = np.random.default_rng(100 + n_snp)
    n_hla, n_total_snp = 7, n_snp + 9
    m = gpu.HLAModel(n_total_snp, n_hla)
    r = ref.new_model(); r.init_predict(n_total_snp, 1, n_hla)
    for k in range(3):
        haplo, _, _ = helpers.random_haplo_list(rng, n_hla=n_hla, n_snp=n_snp, max_per_allele=5)
        idx = np.sort(rng.choice(n_total_snp, size=n_snp, replace=False)).astype(np.int32)
        m.add_classifier(idx, haplo["freq"], haplo["hla"], haplo["packed"])
        r.add_classifier(idx, haplo["freq"], haplo["hla"], haplo["packed"])
    base = rng.integers(0, 3, size=(60, n_total_snp)).astype(np.int8)
    test = base[rng.integers(0, 60, size=3000)]                      # ~50 copies of each row
    test[rng.random(test.shape) < 0.002] = -1                        # a few near-duplicates
    test[17, :] = -1
    test[2500:2520] = rng.integers(0, 3, size=(20, n_total_snp))      # rows seen nowhere else
    want = r.predict(test.astype(np.int32))
    outs, stats = {}, {}
    for name, env in (("dedup", {}), ("every sample", {"HIBAG_B200_PREDICT_DEDUP": "0"}),
                      ("dedup, tiles of 1024", {"HIBAG_B200_PREDICT_TILE": "1024"})):
        for k in ("HIBAG_B200_PREDICT_DEDUP", "HIBAG_B200_PREDICT_TILE"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        s0 = m.predict_stats()
        outs[name] = m.predict(test)
        s1 = m.predict_stats()
        stats[name] = {k: s1[k] - s0[k] for k in s1}
    for name, got in outs.items():
        assert np.array_equal(want["h1"], got["h1"]) and np.array_equal(want["h2"], got["h2"]), name
        for key in ("prob", "matching", "dosage", "postprob"):
            assert np.array_equal(want[key], got[key], equal_nan=True), (name, key)
    d, e, t = stats["dedup"], stats["every sample"], stats["dedup, tiles of 1024"]
    assert e["positions_scored"] == e["positions_total"] == 3 * 3000
    assert e["pair_evals"] == e["pair_evals_nominal"] == d["pair_evals_nominal"]
    assert d["positions_total"] == 3 * 3000 and d["positions_scored"] < 0.2 * d["positions_total"]
    assert d["pair_evals"] < 0.2 * d["pair_evals_nominal"]
    assert d["positions_scored"] <= t["positions_scored"] < 0.5 * t["positions_total"]


def test_classifier_sharded_predict_agrees_within_tolerance(gpu, ref):
    """classifiers split over two 'ranks', partial sums added (what the NCCL all-reduce does), then
    finalised: calls equal, posteriors within 1e-10 of the single-rank result"""
    import torch
    m, r, geno = _golden_model(gpu, ref, n_cls=40)
    geno_t = torch.from_numpy(np.tile(geno, (3, 1)).astype(np.int8)).cuda()
    n = geno_t.shape[0]
    n_cells = m.n_cells
    full = gpu.hlaPredict(m, geno_t.cpu().numpy(), type="response+prob")
    wts = torch.from_numpy(m.snp_weights()).cuda()
    parts = []
    for rank in range(2):
        sub = gpu.HLAModel(m.n_snp, m.n_hla)
        for k in range(rank, 40, 2):
            c = m.classifier(k)
            sub.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"])
        acc = torch.zeros((n, n_cells + 3), dtype=torch.float64, device="cuda")
        sub.predict_partial_device(geno_t.data_ptr(), n, wts.data_ptr(), acc.data_ptr())
        parts.append(acc)
    total = parts[0] + parts[1]
    h1 = torch.zeros(n, dtype=torch.int32, device="cuda"); h2 = torch.zeros_like(h1)
    pp = torch.zeros((n, n_cells), dtype=torch.float64, device="cuda")
    mt = torch.zeros(n, dtype=torch.float64, device="cuda")
    out = gpu.PredictOut(h1.data_ptr(), h2.data_ptr(), None, mt.data_ptr(), None, pp.data_ptr())
    rc = gpu.lib().hibag_b200_predict_finalize_device(m.n_hla, n, C.c_void_p(total.data_ptr()),
                                                      C.byref(out), None, 1)
    assert rc == 0
    assert np.array_equal(h1.cpu().numpy(), full["h1"]) and np.array_equal(h2.cpu().numpy(), full["h2"])
    assert rel_close(pp.cpu().numpy(), full["postprob"]) and rel_close(mt.cpu().numpy(), full["matching"])


def test_round_trip_properties_at_scale(gpu):
    """size-independent properties on a larger batch: posterior rows sum to 1, dosages to 2,
    best guess = argmax of the posterior row, a sample equal to two model haplotypes is called
    with those alleles' pair among the top, and prediction is invariant to batch splitting"""
    from hibag_b200 import synth
    coh = synth.make_cohort(1500, 200, 20, seed=5)
    m = gpu.HLAModel(coh.n_snp, coh.n_hla); m.set_training(coh.geno, coh.h1, coh.h2)
    m.train(4, gpu.default_mtry(coh.n_snp), seed=1, per_classifier_seed=True)
    new = synth.draw_more(coh, 70000, seed=9)
    got = m.predict(new.geno)
    ok = got["h1"] != refpy.NA_INTEGER
    assert ok.mean() > 0.99
    assert np.allclose(got["postprob"][ok].sum(axis=1), 1.0, rtol=0, atol=1e-12)
    assert np.allclose(got["dosage"][ok].sum(axis=1), 2.0, rtol=0, atol=1e-12)
    am = got["postprob"].argmax(axis=1)
    n_hla = m.n_hla
    idx = got["h2"] + got["h1"] * (2 * n_hla - got["h1"] - 1) // 2
    assert np.array_equal(am[ok], idx[ok])
    assert np.array_equal(got["prob"][ok], got["postprob"][ok, am[ok]])
    part = m.predict(new.geno[12345:23456])
    for key in ("h1", "h2", "prob", "matching", "dosage", "postprob"):
        assert np.array_equal(part[key], got[key][12345:23456], equal_nan=True), key


# ---- PLINK BED import --------------------------------------------------------

def test_bed_decode_matches_oracle(gpu, orc, tmp_path):
    """bed_snp_major_kernel / bed_ind_major_kernel through the C ABI vs the restated HIBAG_ConvBED:
    both file modes, ragged sizes, SNP selection, the reference's example file set, the file-level
    mirror of hlaBED2Geno, and the reference's error for a bad prefix"""
    rng = np.random.default_rng(12)
    na = np.iinfo(np.int32).min
    for mode in (0, 1):
        for n_samp, n_snp in ((1, 1), (5, 7), (127, 129), (128, 128), (1000, 517), (3001, 260)):
            rows, per = (n_samp, n_snp) if mode == 0 else (n_snp, n_samp)
            pay = rng.integers(0, 256, size=rows * ((per + 3) // 4), dtype=np.uint8)
            raw = np.concatenate([np.array([0x6C, 0x1B, mode], dtype=np.uint8), pay])
            for flag in (None, (rng.random(n_snp) < 0.5).astype(np.int32)):
                if flag is not None:
                    flag[n_snp // 2] = 1
                want = orc.bed_decode(raw, n_samp, n_snp, flag)
                got = gpu.bed_decode(raw, n_samp, n_snp, flag)
                assert got.dtype == np.int8 and got.shape == want.shape
                assert np.array_equal(got, np.where(want == na, -1, want)), (mode, n_samp, n_snp)
    pk = helpers.load_golden("hapmap_ceu_plink.npz")
    hm = helpers.load_golden("hapmap_ceu.npz")
    # write the file set and import it the way a user would
    (tmp_path / "x.bed").write_bytes(pk["bed"].tobytes())
    (tmp_path / "x.fam").write_text("".join("%s\t%s\t0\t0\t0\t-9\n" % (f, i) for f, i in zip(pk["fam_family"], pk["fam_id"])))
    (tmp_path / "x.bim").write_text("".join("%s\t%s\t0\t%d\t%s\t%s\n" % r for r in zip(
        pk["bim_chr"], pk["bim_snp"], pk["bim_pos"], pk["bim_a1"], pk["bim_a2"])))
    lo, hi = int(hm["snp_position"].min()), int(hm["snp_position"].max())
    geno = gpu.hlaBED2Geno(str(tmp_path / "x.bed"), str(tmp_path / "x.fam"), str(tmp_path / "x.bim"),
                           region=("6", lo, hi))
    si = [geno["sample_id"].index(s) for s in hm["sample_id"]]
    sj = [geno["snp_id"].index(s) for s in hm["snp_id"]]
    want = hm["genotype"].astype(np.int16)
    assert np.array_equal(geno["genotype"][np.ix_(si, sj)], np.where((want < 0) | (want > 2), -1, want))
    with pytest.raises(RuntimeError, match="Invalid prefix in the PLINK BED file"):
        gpu.bed_decode(np.array([1, 2, 3, 4, 5], dtype=np.uint8), 2, 2)


# ---- the BASELINE-shaped configurations against the compiled reference ------

def _capture_fd2(fn):
    """run fn() with file descriptor 2 redirected to a file; returns (result, text)"""
    import os, tempfile
    fd, path = tempfile.mkstemp(suffix=".log")
    saved = os.dup(2)
    os.dup2(fd, 2)
    try:
        out = fn()
    finally:
        os.dup2(saved, 2)
        os.close(fd); os.close(saved)
    text = open(path).read()
    os.unlink(path)
    return out, text


def _trace_lines(text):
    """(snp index 0-based, loss string, oob acc string, n_haplo) per accepted SNP from verbose.detail
    lines -- the reference prints them at src/LibHLA.cpp:2104-2111, the own driver in the same format"""
    import re
    pat = re.compile(r"^\s*(\d+), SNP: (\d+), loss: (\S+), oob acc: (\S+)%, # of haplo: (\d+)")
    out = []
    for ln in text.splitlines():
        m = pat.match(ln)
        if m:
            out.append((int(m.group(2)) - 1, m.group(3), m.group(4), int(m.group(5))))
    return out


def test_headline_config_classifiers_equal_the_reference(gpu):
    """configs[1] itself: HLA-A-shaped 5,000 samples x 500 SNPs, bench.py's cohort and seeds. Fixture
    tests/golden/c2_ref.npz = classifiers trained to completion by the compiled, unmodified reference
    (tools/make_golden_ref.py; targets base / avx2, 20-40 CPU-minutes each). The B200 path trains the
    same global indices the way bench.py does -- exact screening, device EM, 24 classifiers in flight --
    and must reproduce them bit for bit: bootstrap counts, SNP order, haplotypes, fp64 frequencies,
    out-of-bag accuracy. At this scale the paths the small cohorts never reach are live: uncertified
    in-bag sums rescored, device-EM host fallbacks, the dense EM shape."""
    import bench
    gd = helpers.load_golden("c2_ref.npz")
    assert (int(gd["n_samp"]), int(gd["n_snp"]), int(gd["mtry"]), int(gd["train_seed"])) == \
        (bench.N_SAMP, bench.N_SNP, bench.MTRY, bench.TRAIN_SEED)
    coh = bench.make_cohort()
    ks = [int(k) for k in gd["ks"] if bool(gd["c%d_finished" % k])]
    assert len(ks) >= 2
    m = gpu.HLAModel(bench.N_SNP, coh.n_hla)
    m.set_training(np.ascontiguousarray(coh.geno, dtype=np.int8), coh.h1, coh.h2)
    n = max(24, max(ks) + 1)
    m.train(n, bench.MTRY, prune=True, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=0,
            n_threads=48, n_concurrent=24)
    st = m.train_stats()
    assert st["pair_evals"] < 0.3 * st["pair_evals_nominal"]           # the exact screen was on
    for k in ks:
        want = {key: gd["c%d_%s" % (k, key)] for key in ("snpidx", "samp_num", "freq", "hla", "packed")}
        want["oob_acc"] = float(gd["c%d_oob_acc" % k])
        d = helpers.classifier_diff(m.classifier(k), want)
        assert d == "", "classifier %d differs from the reference in '%s'" % (k, d)
    # the same with the opt-in second level of the in-bag screen
    os.environ["HIBAG_B200_SCREEN_REFINE"] = "1"
    try:
        m2 = gpu.HLAModel(bench.N_SNP, coh.n_hla)
        m2.set_training(np.ascontiguousarray(coh.geno, dtype=np.int8), coh.h1, coh.h2)
        m2.train(n, bench.MTRY, prune=True, seed=bench.TRAIN_SEED, per_classifier_seed=True, first_index=0,
                 n_threads=48, n_concurrent=24)
    finally:
        del os.environ["HIBAG_B200_SCREEN_REFINE"]
    assert m2.train_stats()["pair_evals"] < st["pair_evals"]
    for k in ks:
        want = {key: gd["c%d_%s" % (k, key)] for key in ("snpidx", "samp_num", "freq", "hla", "packed")}
        want["oob_acc"] = float(gd["c%d_oob_acc" % k])
        assert helpers.classifier_diff(m2.classifier(k), want) == "", k
    # one of them again alone, every cell scored, EM on the host pool, with the accepted-SNP trace
    # (SNP, loss to 6 digits, out-of-bag accuracy, haplotypes) equal line by line
    k = ks[0]
    h = gpu.HLAModel(bench.N_SNP, coh.n_hla)
    h.set_training(np.ascontiguousarray(coh.geno, dtype=np.int8), coh.h1, coh.h2)
    _, text = _capture_fd2(lambda: h.train(1, bench.MTRY, prune=True, seed=bench.TRAIN_SEED, per_classifier_seed=True,
                                           first_index=k, screening=False, em_on_device=False, verbose=2))
    want = {key: gd["c%d_%s" % (k, key)] for key in ("snpidx", "samp_num", "freq", "hla", "packed")}
    want["oob_acc"] = float(gd["c%d_oob_acc" % k])
    assert helpers.classifier_diff(h.classifier(0), want) == ""
    got = _trace_lines(text)
    ref_trace = list(zip(gd["c%d_trace_snp" % k].tolist(), [str(x) for x in gd["c%d_trace_loss" % k]],
                         [str(x) for x in gd["c%d_trace_acc" % k]], gd["c%d_trace_n_haplo" % k].tolist()))
    assert got == ref_trace


def test_reference_host_with_gpu_hooks_at_headline_config(gpu, ref):
    """The drop-in itself at configs[1]: the reference's own BuildClassifiers (compiled, unmodified:
    CVariableSelection::Search src/LibHLA.cpp:1981-2122 with its sequential CAlg_EM on the host)
    calling OUR ten hooks per candidate SNP on the 5,000 x 500 cohort must build the classifier the
    reference's CPU path built (tests/golden/c2_ref.npz) bit for bit. Prints the seconds per
    classifier of that pairing (the CPU path alone: 590-610 s, profiles/ref_full_classifier.json)."""
    import time
    import bench
    gd = helpers.load_golden("c2_ref.npz")
    coh = bench.make_cohort()
    k = [int(k) for k in gd["ks"] if bool(gd["c%d_finished" % k])][0]
    m = ref.new_model()
    m.init_training(coh.geno, coh.h1, coh.h2, coh.n_hla)
    ref.set_gpu_procs(gpu.get_procs())
    try:
        t0 = time.time()
        m.build(1, bench.MTRY, prune=True, reseed_base=bench.TRAIN_SEED, first_index=k)
        dt = time.time() - t0
    finally:
        ref.set_gpu_procs(None)
    want = {key: gd["c%d_%s" % (k, key)] for key in ("snpidx", "samp_num", "freq", "hla", "packed")}
    want["oob_acc"] = float(gd["c%d_oob_acc" % k])
    assert helpers.classifier_diff(m.classifier(0), want) == ""
    print("\nreference host + B200 hooks, configs[1], classifier %d: %.2f s (reference CPU path: %.0f s)"
          % (k, dt, float(gd["c%d_seconds" % k])))


def test_drb1_scale_prefix_equals_the_reference(gpu):
    """configs[3]: HLA-DRB1-shaped 10,000 samples x 800 SNPs, 86 alleles (3,741 cells), haplotype lists
    up to 1,800. After its last accepted SNP the reference scores every remaining SNP at full list
    size (src/LibHLA.cpp:2113-2119): ~80 CPU-minutes to the last accepted SNP, many hours to the end.
    The fixture (tests/golden/c4_ref_prefix.npz, tools/make_golden_ref.py) is therefore the trace of
    accepted SNPs the compiled reference printed -- SNP, loss (6 digits), out-of-bag accuracy and
    haplotype count at each -- and every line must equal the B200 path's trace; a classifier the
    reference did finish is compared whole, bit for bit."""
    import os
    if not os.path.exists(os.path.join(helpers.GOLDEN, "c4_ref_prefix.npz")):
        pytest.skip("tests/golden/c4_ref_prefix.npz not generated")
    from hibag_b200 import synth
    gd = helpers.load_golden("c4_ref_prefix.npz")
    coh = synth.make_cohort(int(gd["n_samp"]), int(gd["n_snp"]), int(gd["n_hla_drawn"]), seed=int(gd["cohort_seed"]))
    for k in [int(k) for k in gd["ks"]]:
        m = gpu.HLAModel(coh.n_snp, coh.n_hla)
        m.set_training(np.ascontiguousarray(coh.geno, dtype=np.int8), coh.h1, coh.h2)
        _, text = _capture_fd2(lambda: m.train(1, int(gd["mtry"]), prune=True, seed=int(gd["train_seed"]),
                                               per_classifier_seed=True, first_index=k, verbose=2))
        got = _trace_lines(text)
        ref_trace = list(zip(gd["c%d_trace_snp" % k].tolist(), [str(x) for x in gd["c%d_trace_loss" % k]],
                             [str(x) for x in gd["c%d_trace_acc" % k]], gd["c%d_trace_n_haplo" % k].tolist()))
        assert len(ref_trace) >= 4
        assert got[:len(ref_trace)] == ref_trace
        assert m.classifier(0)["snpidx"][:len(ref_trace)].tolist() == [t[0] for t in ref_trace]
        if bool(gd["c%d_finished" % k]):
            want = {key: gd["c%d_%s" % (k, key)] for key in ("snpidx", "samp_num", "freq", "hla", "packed")}
            want["oob_acc"] = float(gd["c%d_oob_acc" % k])
            assert helpers.classifier_diff(m.classifier(0), want) == "", k


def test_single_stream_seed_continues_across_train_calls(gpu):
    """with per_classifier_seed = 0 a second train() call used to re-seed and append
    duplicates. The reference continues R's stream (set.seed once, BuildClassifiers repeatedly):
    2 classifiers in one call == 1 + 1 in two calls; a new seed or clear() re-seeds."""
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    a = gpu.HLAModel(geno.shape[1], len(al)); a.set_training(geno, h1, h2)
    a.train(3, 17, seed=100)
    b = gpu.HLAModel(geno.shape[1], len(al)); b.set_training(geno, h1, h2)
    b.train(1, 17, seed=100); b.train(2, 17, seed=100)
    assert b.num_classifiers() == 3
    for k in range(3):
        assert helpers.classifier_diff(b.classifier(k), a.classifier(k)) == ""
        helpers.assert_classifier_equals_golden(b.classifier(k), ml, k)
    b.clear(); b.train(1, 17, seed=100)
    helpers.assert_classifier_equals_golden(b.classifier(0), ml, 0)
