"""Shared helpers for the test-suite (and __graft_entry__.smoke)."""
import os

import numpy as np

from oracle import refpy

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def hapmap_a_training():
    """Training inputs of the reference's golden HLA-A model (vignettes/HIBAG.Rmd:100-118) in the
    model's own sample / SNP / allele order. Returns geno[int8 60x266], h1, h2, alleles, golden."""
    hm = load_golden("hapmap_ceu.npz")
    ml = load_golden("modellist_a.npz")
    sid = list(hm["sample_id"]); snp = list(hm["snp_id"])
    si = [sid.index(s) for s in ml["sample_id"]]
    sj = [snp.index(s) for s in ml["snp_id"]]
    geno = hm["genotype"][np.ix_(si, sj)].astype(np.int8)
    hs = list(hm["hla_sample_id"]); al = list(ml["hla_allele"])
    h1 = np.array([al.index(hm["hla_A_1"][hs.index(s)]) for s in ml["sample_id"]], dtype=np.int32)
    h2 = np.array([al.index(hm["hla_A_2"][hs.index(s)]) for s in ml["sample_id"]], dtype=np.int32)
    return geno, h1, h2, al, ml


def golden_classifier(ml, k):
    a, b = ml["snp_off"][k:k + 2]
    p, q = ml["hap_off"][k:k + 2]
    return dict(snpidx=ml["snpidx"][a:b], samp_num=ml["samp_num"][k], freq=ml["freq"][p:q],
                hla=ml["hla_idx"][p:q], packed=ml["packed"][p:q], oob_acc=float(ml["oob_acc"][k]))


def classifier_diff(c, g):
    """'' when classifier dict c equals g bit for bit, else the first differing field."""
    for key in ("samp_num", "snpidx", "hla", "packed", "freq"):
        if key == "samp_num" and (c.get(key) is None or len(c[key]) == 0):
            continue
        if not np.array_equal(np.asarray(c[key]), np.asarray(g[key])):
            return key
    if c["oob_acc"] != g["oob_acc"]:
        return "oob_acc"
    return ""


def assert_classifier_equals_golden(c, ml, k):
    d = classifier_diff(c, golden_classifier(ml, k))
    assert d == "", "classifier %d differs from the reference golden in '%s'" % (k, d)


def random_haplo_list(rng, n_hla, n_snp, max_per_allele=8, empty_frac=0.15):
    """A seeded haplotype list grouped by allele; some alleles have no haplotype. Bits >= n_snp
    are filled with garbage on purpose (the reference never clears them)."""
    lens = rng.integers(1, max_per_allele + 1, size=n_hla)
    lens[rng.random(n_hla) < empty_frac] = 0
    if lens.sum() == 0:
        lens[0] = 1
    n = int(lens.sum())
    packed = rng.integers(0, 2**63, size=(n, 2), dtype=np.int64).astype(np.uint64)
    packed ^= rng.integers(0, 2, size=(n, 2), dtype=np.int64).astype(np.uint64) << np.uint64(63)
    freq = rng.random(n) ** 3 + 1e-7
    freq /= freq.sum()
    hla = np.repeat(np.arange(n_hla), lens).astype(np.int32)
    return refpy.make_haplo(packed, freq, hla), n_hla, n_snp


def random_genotypes(rng, n, n_snp, n_hla, missing=0.05, haplo=None):
    """Seeded packed genotypes. With `haplo` given, most samples are built from two haplotypes of
    the list plus noise so that posteriors are not all vanishing."""
    g = rng.integers(0, 3, size=(n, n_snp)).astype(np.int32)
    if haplo is not None and len(haplo) > 0:
        bits = np.zeros((len(haplo), n_snp), dtype=np.int32)
        for j in range(n_snp):
            bits[:, j] = (haplo["packed"][:, j >> 6] >> np.uint64(j & 63)) & np.uint64(1)
        i1 = rng.integers(0, len(haplo), size=n); i2 = rng.integers(0, len(haplo), size=n)
        gg = bits[i1] + bits[i2]
        flip = rng.random((n, n_snp)) < 0.03
        gg = np.where(flip, rng.integers(0, 3, size=(n, n_snp)), gg)
        keep = rng.random(n) < 0.85
        g = np.where(keep[:, None], gg, g).astype(np.int32)
    g[rng.random((n, n_snp)) < missing] = -1
    a1 = rng.integers(0, n_hla, size=n); a2 = rng.integers(0, n_hla, size=n)
    boot = rng.integers(0, 4, size=n)
    return refpy.pack_geno(g, boot=boot, a1=a1, a2=a2)
