"""Deterministic synthetic HLA/SNP cohorts of the shapes BASELINE.json names.

Allele frequencies ~ Dirichlet(0.35); 1-4 founder haplotypes per allele over n_snp SNPs
(per-SNP MAF ~ U(0.05, 0.5); founders of one allele differ from the allele's base haplotype by
8 % flips); each sample = two founders drawn by frequency, 1 % allele noise, 1 % missing
genotypes; alleles relabelled densely in order of first appearance in sorted order.
"""
import numpy as np


class Cohort:
    def __init__(self, geno, h1, h2, n_hla, founders, founder_allele, allele_freq):
        self.geno = geno                  # int8 [n_samp, n_snp]: 0/1/2, -1 missing
        self.h1, self.h2 = h1, h2         # int32 [n_samp], dense 0-based allele indices
        self.n_hla = n_hla
        self._founders = founders
        self._founder_allele = founder_allele
        self._allele_freq = allele_freq

    @property
    def n_samp(self):
        return self.geno.shape[0]

    @property
    def n_snp(self):
        return self.geno.shape[1]


def _draw(rng, founders, founder_allele, allele_freq, n_samp, noise, missing):
    n_found, n_snp = founders.shape
    w = allele_freq[founder_allele]
    cnt = np.bincount(founder_allele, minlength=len(allele_freq))
    w = w / cnt[founder_allele]
    w = w / w.sum()
    f1 = rng.choice(n_found, size=n_samp, p=w)
    f2 = rng.choice(n_found, size=n_samp, p=w)
    a = founders[f1] ^ (rng.random((n_samp, n_snp)) < noise)
    b = founders[f2] ^ (rng.random((n_samp, n_snp)) < noise)
    geno = (a.astype(np.int8) + b.astype(np.int8))
    geno[rng.random((n_samp, n_snp)) < missing] = -1
    return geno, founder_allele[f1].astype(np.int32), founder_allele[f2].astype(np.int32)


def make_cohort(n_samp, n_snp, n_hla, seed=1, noise=0.01, missing=0.01):
    """Training cohort; alleles that no sample carries are dropped and the rest relabelled."""
    rng = np.random.default_rng(seed)
    allele_freq = rng.dirichlet(np.full(n_hla, 0.35))
    maf = rng.uniform(0.05, 0.5, size=n_snp)
    n_per = rng.integers(1, 5, size=n_hla)
    founders, founder_allele = [], []
    for a in range(n_hla):
        base = rng.random(n_snp) < maf
        for _ in range(n_per[a]):
            founders.append(base ^ (rng.random(n_snp) < 0.08))
            founder_allele.append(a)
    founders = np.array(founders, dtype=bool)
    founder_allele = np.array(founder_allele, dtype=np.int64)
    geno, h1, h2 = _draw(rng, founders, founder_allele, allele_freq, n_samp, noise, missing)
    used = np.unique(np.concatenate([h1, h2]))
    remap = np.full(n_hla, -1, dtype=np.int64)
    remap[used] = np.arange(len(used))
    keep = remap[founder_allele] >= 0
    coh = Cohort(geno, remap[h1].astype(np.int32), remap[h2].astype(np.int32), len(used),
                 founders[keep], remap[founder_allele[keep]], allele_freq[used] / allele_freq[used].sum())
    return coh


def draw_more(cohort, n_samp, seed=2, noise=0.01, missing=0.01):
    """Fresh samples from the same founders (the prediction cohort of config 3)."""
    rng = np.random.default_rng(seed)
    geno, h1, h2 = _draw(rng, cohort._founders, cohort._founder_allele, cohort._allele_freq,
                         n_samp, noise, missing)
    return Cohort(geno, h1, h2, cohort.n_hla, cohort._founders, cohort._founder_allele,
                  cohort._allele_freq)


def make_thermo_cohort(n_samp, n_hla, seed=1, noise=0.0, missing=0.005):
    """Cohort whose classifiers need MANY SNPs (test shape, not a BASELINE config): allele a carries
    1 at the SNPs below a ("thermometer code", SNP order shuffled), so every SNP separates exactly
    one more pair of neighbouring alleles and the search keeps accepting SNPs up to n_hla - 1 of
    them; `noise` flips alleles per sample and grows the haplotype lists. Returns (geno int8
    [n_samp, n_hla - 1], h1, h2)."""
    rng = np.random.default_rng(seed)
    n_snp = n_hla - 1
    code = np.zeros((n_hla, n_snp), dtype=bool)
    for a in range(n_hla):
        code[a, :a] = True
    code = code[:, rng.permutation(n_snp)]
    a1 = rng.integers(0, n_hla, n_samp)
    a2 = rng.integers(0, n_hla, n_samp)
    x = code[a1] ^ (rng.random((n_samp, n_snp)) < noise)
    y = code[a2] ^ (rng.random((n_samp, n_snp)) < noise)
    geno = x.astype(np.int8) + y.astype(np.int8)
    geno[rng.random((n_samp, n_snp)) < missing] = -1
    return geno, a1.astype(np.int32), a2.astype(np.int32)
