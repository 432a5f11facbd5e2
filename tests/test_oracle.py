"""CPU tests: the oracle (oracle/hibag_oracle.c) against the compiled reference and the
reference's own golden model; the compiled reference against the golden fixture."""
import numpy as np
import pytest

from oracle import refpy
from tests import helpers

SNP_COUNTS = [1, 2, 7, 31, 32, 33, 63, 64, 65, 100, 128]


def test_table_matches_reference(ref, orc):
    t = orc.table()
    assert np.array_equal(t, ref.table())
    assert t[0] == 1.0 and t[64] > 0 and t[65] == 0 and t[256] == 0      # denormal tail, then 0
    assert t[62] < 2.3e-308                                              # entries 62..64 are denormal


@pytest.mark.parametrize("n_snp", SNP_COUNTS)
def test_scoring_functions_match_reference(ref, orc, n_snp):
    rng = np.random.default_rng(1000 + n_snp)
    h, n_hla, _ = helpers.random_haplo_list(rng, 12, n_snp)
    g = helpers.random_genotypes(rng, 150, n_snp, n_hla, haplo=h)
    for i in range(0, 150, 37):
        for j in (0, len(h) - 1):
            assert orc.hamming(g[i:i + 1], h[j:j + 1], h[0:1], n_snp) == \
                ref.hamming(g[i:i + 1], h[j:j + 1], h[0:1], n_snp)
    a, b = ref.best_guess(h, n_hla, n_snp, g), orc.best_guess(h, n_hla, n_snp, g)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert np.array_equal(ref.post_prob(h, n_hla, n_snp, g), orc.post_prob(h, n_hla, n_snp, g),
                          equal_nan=True)
    p, s = ref.post_prob2(h, n_hla, n_snp, g)
    q, t = orc.post_prob2(h, n_hla, n_snp, g)
    assert np.array_equal(p, q, equal_nan=True) and np.array_equal(s, t)
    assert ref.acc_oob(h, n_hla, n_snp, g) == orc.acc_oob(h, n_hla, n_snp, g)
    x, y = ref.acc_ib(h, n_hla, n_snp, g), orc.acc_ib(h, n_hla, n_snp, g)
    assert x == y or (np.isnan(x) and np.isnan(y))


def test_all_missing_genotype_gives_prior(orc, ref):
    """every SNP missing: distance 0 everywhere, posterior = frequency products; BestGuess picks
    the first maximum"""
    rng = np.random.default_rng(5)
    h, n_hla, n_snp = helpers.random_haplo_list(rng, 6, 20)
    g = refpy.pack_geno(np.full((3, 20), -1), boot=[0, 1, 2], a1=[0, 1, 2], a2=[1, 2, 3])
    p, s = orc.post_prob2(h, n_hla, n_snp, g)
    assert np.allclose(s, 1.0, rtol=1e-12)       # sum over all pairs of (sum f)^2 = 1
    pr, sr = ref.post_prob2(h, n_hla, n_snp, g)
    assert np.array_equal(p, pr) and np.array_equal(s, sr)


def test_zero_posterior_is_na(orc, ref):
    """a genotype at distance >= 65 from every pair scores exactly 0 everywhere -> (NA, NA)"""
    n_snp = 128
    packed = np.zeros((2, 2), dtype=np.uint64)                  # two all-zero haplotypes
    h = refpy.make_haplo(packed, [0.5, 0.5], [0, 1])
    g = refpy.pack_geno(np.full((1, n_snp), 2), boot=[0], a1=[0], a2=[1])   # distance 256
    a1, a2 = orc.best_guess(h, 2, n_snp, g)
    assert a1[0] == refpy.NA_INTEGER and a2[0] == refpy.NA_INTEGER
    r1, r2 = ref.best_guess(h, 2, n_snp, g)
    assert r1[0] == a1[0] and r2[0] == a2[0]
    assert np.isnan(orc.post_prob(h, 2, n_snp, g)[0])


def test_int_to_snp_matches_reference(orc, ref):
    rng = np.random.default_rng(11)
    row = rng.integers(-1, 4, size=400).astype(np.int32)        # 3 and -1 are both "missing"
    for length in (0, 1, 7, 8, 9, 63, 64, 65, 127, 128):
        idx = rng.choice(400, size=length, replace=False).astype(np.int32)
        a, b = orc.int_to_snp(row, idx), ref.int_to_snp(row, idx)
        for f in ("s1", "s2"):
            assert np.array_equal(a[f], b[f]), (length, f)


def test_reference_reproduces_golden_model(ref):
    """the compiled reference, target base, set.seed(100): first 12 classifiers of
    inst/extdata/ModelList.RData bit for bit (all 100 are checked on the GPU box against the CUDA
    path; 12 keep the CPU suite short)"""
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    m = ref.new_model()
    m.init_training(geno, h1, h2, len(al))
    ref.set_seed(int(ml["seed"]))
    m.build(12, int(ml["mtry"]))
    for k in range(12):
        helpers.assert_classifier_equals_golden(m.classifier(k), ml, k)


def test_oracle_predict_matches_reference_on_golden_model(ref, orc):
    geno, h1, h2, al, ml = helpers.hapmap_a_training()
    n_hla = len(al)
    cls = [helpers.golden_classifier(ml, k) for k in range(25)]
    m = ref.new_model()
    m.init_predict(geno.shape[1], geno.shape[0], n_hla)
    for c in cls:
        m.add_classifier(c["snpidx"], c["freq"], c["hla"], c["packed"], acc=c["oob_acc"])
    test = geno.astype(np.int32).copy()
    rng = np.random.default_rng(2)
    test[rng.random(test.shape) < 0.05] = -1
    test[3, :] = -1                                   # a sample with every SNP missing
    r = m.predict(test)
    o = orc.predict(cls, n_hla, geno.shape[1], test)
    assert np.array_equal(r["h1"], o["h1"]) and np.array_equal(r["h2"], o["h2"])
    for key in ("prob", "matching", "dosage", "postprob"):
        assert np.array_equal(r[key], o[key], equal_nan=True), key
    assert r["h1"][3] == refpy.NA_INTEGER and np.isnan(r["matching"][3])


def test_haplomatch_records_follow_reference_pair_matcher(ref, orc):
    """the oracle's build_haplomatch records == the reference's own _PrepHaploMatch_def
    (src/LibHLA.cpp:1569-1637) run per in-bag sample on the same (un-doubled) list"""
    rng = np.random.default_rng(77)
    for n_snp in (3, 17, 64, 100):
        haplo, n_hla, _ = helpers.random_haplo_list(rng, n_hla=7, n_snp=n_snp, max_per_allele=5)
        geno = helpers.random_genotypes(rng, 60, n_snp, n_hla, haplo=haplo)
        a1 = np.minimum(geno["a1"], geno["a2"]); a2 = np.maximum(geno["a1"], geno["a2"])
        geno["a1"], geno["a2"] = a1, a2
        lens = np.bincount(haplo["hla"], minlength=n_hla)
        start = np.concatenate([[0], np.cumsum(lens)])
        rec = orc.haplomatch_records(haplo, lens, n_snp, geno)
        want, k = [], 0
        for s in range(len(geno)):
            if geno["boot"][s] <= 0:
                continue
            x, y = int(geno["a1"][s]), int(geno["a2"][s])
            for i1, i2 in ref.prep_haplo_match(geno[s:s + 1], haplo, int(start[x]), int(lens[x]),
                                               int(start[y]), int(lens[y]), n_snp):
                want.append((k, (i2 << 16) | i1))
            k += 1
        assert [tuple(r) for r in rec.tolist()] == want


# ---- PLINK BED import -----------------------------------------------------

def _np_bed_decode(raw, n_samp, n_snp):
    """independent numpy decoding of a .bed byte string -> int8 [n_samp][n_snp], missing -1"""
    cvt = np.array([2, -1, 1, 0], dtype=np.int8)
    mode = raw[2]
    rows, per = (n_samp, n_snp) if mode == 0 else (n_snp, n_samp)
    pay = raw[3:3 + rows * ((per + 3) // 4)].reshape(rows, -1)
    g = np.zeros((rows, pay.shape[1] * 4), dtype=np.int8)
    for k in range(4):
        g[:, k::4] = cvt[(pay >> (2 * k)) & 3]
    g = g[:, :per]
    return g if mode == 0 else np.ascontiguousarray(g.T)


def test_bed_decode_reproduces_reference_dataset(orc):
    """the reference's example PLINK files decode to the reference's own HapMap_CEU_Geno dataset
    (60 shared samples x 1,564 shared SNPs): the pin of the BED decoder"""
    pk = helpers.load_golden("hapmap_ceu_plink.npz")
    hm = helpers.load_golden("hapmap_ceu.npz")
    n_samp, n_snp = len(pk["fam_id"]), len(pk["bim_snp"])
    full = orc.bed_decode(pk["bed"], n_samp, n_snp)
    assert full.shape == (n_samp, n_snp) and pk["bed"][2] == 0          # an individual-major file
    g = np.where(full == np.iinfo(np.int32).min, -1, full).astype(np.int8)
    assert np.array_equal(g, _np_bed_decode(pk["bed"], n_samp, n_snp))
    si = [list(pk["fam_id"]).index(s) for s in hm["sample_id"]]
    snp_ix = {s: i for i, s in enumerate(pk["bim_snp"])}
    sj = [snp_ix[s] for s in hm["snp_id"]]
    want = hm["genotype"].astype(np.int16)
    want = np.where((want < 0) | (want > 2), -1, want)
    assert np.array_equal(g[np.ix_(si, sj)], want)
    assert np.array_equal(pk["bim_pos"][sj], hm["snp_position"])
    # SNP selection keeps the flagged columns in file order
    flag = np.zeros(n_snp, dtype=np.int32); flag[sj] = 1
    sel = orc.bed_decode(pk["bed"], n_samp, n_snp, flag)
    assert np.array_equal(sel, full[:, np.nonzero(flag)[0]])


def test_bed_decode_both_modes_and_ragged_sizes(orc):
    rng = np.random.default_rng(11)
    for mode in (0, 1):
        for n_samp, n_snp in ((1, 1), (5, 7), (64, 33), (131, 258)):
            rows, per = (n_samp, n_snp) if mode == 0 else (n_snp, n_samp)
            pay = rng.integers(0, 256, size=rows * ((per + 3) // 4), dtype=np.uint8)
            raw = np.concatenate([np.array([0x6C, 0x1B, mode], dtype=np.uint8), pay])
            flag = (rng.random(n_snp) < 0.6).astype(np.int32); flag[0] = 1
            got = orc.bed_decode(raw, n_samp, n_snp, flag)
            want = _np_bed_decode(raw, n_samp, n_snp)[:, np.nonzero(flag)[0]]
            assert np.array_equal(np.where(got == np.iinfo(np.int32).min, -1, got), want), (mode, n_samp, n_snp)
    with pytest.raises(RuntimeError, match="Invalid prefix"):
        orc.bed_decode(np.array([1, 2, 3, 4], dtype=np.uint8), 2, 2)


@pytest.mark.parametrize("n_snp", [9, 40, 100])
def test_reference_scores_depend_on_the_packed_genotype_and_true_type_only(ref, n_snp):
    """The premise of DESIGN.md 4.7 (each distinct genotype scored once), checked on the compiled
    reference itself: its per-sample outputs -- BestGuess, PostProb of the true type, the PostProb2
    posterior row -- are the same bits for two samples with the same SNP words (and, for PostProb, the
    same true type), wherever they sit in the array and whatever their bootstrap count; and the
    out-of-bag / in-bag totals are the sums over samples (in-bag: count x log ratio in array order,
    src/LibHLA.cpp:1966-1977), so a copied per-sample result changes nothing."""
    rng = np.random.default_rng(4700 + n_snp)
    h, n_hla, _ = helpers.random_haplo_list(rng, 9, n_snp)
    rows = rng.integers(0, 3, size=(12, n_snp)).astype(np.int32)
    rows[rng.random(rows.shape) < 0.05] = -1
    t1 = rng.integers(0, n_hla, 12); t2 = rng.integers(0, n_hla, 12)
    pick = rng.integers(0, 12, size=90)                     # every row ~7 times, scattered
    boot = rng.integers(0, 4, size=90)
    g = refpy.pack_geno(rows[pick], boot=boot, a1=t1[pick], a2=t2[pick])
    a1, a2 = ref.best_guess(h, n_hla, n_snp, g)
    pp = ref.post_prob(h, n_hla, n_snp, g)
    p2, s2 = ref.post_prob2(h, n_hla, n_snp, g)
    first = {}
    for i, k in enumerate(pick):
        j = first.setdefault(int(k), i)
        assert (a1[i], a2[i]) == (a1[j], a2[j])
        assert pp[i].tobytes() == pp[j].tobytes()
        assert p2[i].tobytes() == p2[j].tobytes() and s2[i].tobytes() == s2[j].tobytes()
    # the totals from one representative per distinct (genotype, type)
    oob = g[boot == 0]
    cnt = 0
    for i in np.flatnonzero(boot == 0):
        j = first[int(pick[i])]
        lo, hi = sorted((int(t1[pick[i]]), int(t2[pick[i]])))
        x, y = int(a1[j]), int(a2[j])
        c = 0
        if x == lo: c, lo = 1, -1
        elif x == hi: c, hi = 1, -1
        if y in (lo, hi): c += 1
        cnt += c
    assert ref.acc_oob(h, n_hla, n_snp, oob) == cnt
    ib = np.flatnonzero(boot > 0)
    if n_snp == 40:
        assert np.all(pp[ib] > 0)                           # (this case does take the in-bag branch)
    if np.all(pp[ib] > 0):
        ll = 0.0
        for i in ib:
            ll += int(boot[i]) * np.log(pp[first[int(pick[i])]])
        assert ref.acc_ib(h, n_hla, n_snp, g[ib]) == ll * -2
