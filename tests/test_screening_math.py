"""CPU tests of the arithmetic the exact screening of the training passes rests on:
the bound factor taken from the host's own rare-frequency table, the cell bound, the single-pair
lower bound x_ref, and the lo/hi certificate of the in-bag sum. The kernels themselves are
parity-tested on the GPU (tests/test_gpu_parity.py); here the claims are checked in numpy against
the C oracle's chain values on small seeded cases."""
import numpy as np
import pytest

from tests import helpers


@pytest.fixture(scope="module")
def consts(built):
    from hibag_b200 import api
    return api.host_screen_constants()


def test_floor_table_and_bound_factor(consts, orc):
    T, Tf, K = consts
    assert np.array_equal(T, orc.table())                    # the host table is the reference's
    assert np.array_equal(Tf, np.maximum(T, 1e-100))
    assert np.all(np.diff(T) <= 0)                            # what makes d >= c_i + c_j usable
    # T[d] <= (K/2) * T'[p] * T'[q] for every p + q <= d
    suffix = np.maximum.accumulate(T[::-1])[::-1]
    p, q = np.meshgrid(np.arange(257), np.arange(257), indexing="ij")
    ok = p + q <= 256
    lhs = suffix[np.minimum(p + q, 256)]
    assert np.all(lhs[ok] <= (K / 2) * Tf[p[ok]] * Tf[q[ok]])
    assert 2.0 <= K < 2.0 * (1 + 1e-6)
    assert np.all(Tf[p] * Tf[q] >= 1e-201)                   # far from underflow


def _decode(haplo, n_snp):
    bits = np.zeros((len(haplo), n_snp), dtype=np.int8)
    for w in range(2):
        lo, hi = 64 * w, min(n_snp, 64 * (w + 1))
        if hi > lo:
            bits[:, lo:hi] = (haplo["packed"][:, w][:, None] >> np.arange(hi - lo, dtype=np.uint64)) & np.uint64(1)
    return bits


def _bounds_and_xref(haplo, n_hla, n_snp, g, t1, t2, T, Tf, K):
    """numpy restatement of screen_bound_kernel for one genotype row g (0/1/2, -1 missing)"""
    H = _decode(haplo, n_snp)
    f = haplo["freq"]
    hom0, hom2, het = g == 0, g == 2, g == 1
    c = (H[:, hom0] == 1).sum(1) + (H[:, hom2] == 0).sum(1)
    u = f * Tf[c]
    U = np.array([u[haplo["hla"] == a].sum() for a in range(n_hla)])

    def term(i, j):
        d = c[i] + c[j] + int((H[i, het] == H[j, het]).sum())
        return (f[i] * f[i] if i == j else (2.0 * f[i]) * f[j]) * T[d]

    A = np.nonzero(haplo["hla"] == t1)[0]
    B = np.nonzero(haplo["hla"] == t2)[0]
    xref = 0.0
    if len(A) and len(B):
        i1 = A[np.argmax(u[A])]
        j2 = B[np.argmax(u[B])]
        cand = [term(i1, j) for j in B if t1 != t2 or j >= i1] + [term(i, j2) for i in A if t1 != t2 or i <= j2]
        xref = max(cand) if cand else 0.0
    iu = np.triu_indices(n_hla)
    return (U[iu[0]] * U[iu[1]]) * K, xref


def _genotype_rows(geno, n_snp):
    """TGenotype records -> int rows (the helpers pack them; decode the two bit planes)"""
    s1 = np.zeros((len(geno), n_snp), dtype=np.int8); s2 = np.zeros_like(s1)
    for w in range(2):
        lo, hi = 64 * w, min(n_snp, 64 * (w + 1))
        if hi > lo:
            sh = np.arange(hi - lo, dtype=np.uint64)
            s1[:, lo:hi] = (geno["s1"][:, w].astype(np.uint64)[:, None] >> sh) & np.uint64(1)
            s2[:, lo:hi] = (geno["s2"][:, w].astype(np.uint64)[:, None] >> sh) & np.uint64(1)
    g = np.where((s1 == 0) & (s2 == 0), 0, np.where((s1 == 1) & (s2 == 0), 1, np.where((s1 == 1) & (s2 == 1), 2, -1)))
    return g.astype(np.int8)


@pytest.mark.parametrize("n_snp", [5, 23, 40, 70])
def test_bound_xref_and_certificate_against_the_oracle_chain(consts, orc, n_snp):
    """every cell value the oracle's chain produces is below its bound; x_ref is below the true
    cell's value; the screened argmax and the certified sum equal the full ones"""
    T, Tf, K = consts
    rng = np.random.default_rng(100 + n_snp)
    n_hla = 9
    haplo, n_hla, n_snp = helpers.random_haplo_list(rng, n_hla=n_hla, n_snp=n_snp, max_per_allele=7)
    geno = helpers.random_genotypes(rng, 120, n_snp, n_hla, haplo=haplo)
    G = _genotype_rows(geno, n_snp)
    p2, s2 = orc.post_prob2(haplo, n_hla, n_snp, geno)         # normalised cells and their raw sums
    pp = p2 * s2[:, None]                                       # within 1 ulp of the raw chain values
    a1, a2 = orc.best_guess(haplo, n_hla, n_snp, geno)
    iu = np.triu_indices(n_hla)
    n_certified = 0
    for s in range(len(geno)):
        t1, t2 = int(geno["a1"][s]), int(geno["a2"][s])
        t1, t2 = min(t1, t2), max(t1, t2)
        bound, xref = _bounds_and_xref(haplo, n_hla, n_snp, G[s], t1, t2, T, Tf, K)
        x = pp[s]
        assert np.all(x <= bound * (1 + 1e-12)), "a cell value exceeds its bound"
        true_idx = t2 + t1 * (2 * n_hla - t1 - 1) // 2
        assert xref <= x[true_idx] * (1 + 1e-12)
        # out-of-bag: argmax over the surviving cells == full argmax (strict '<', first wins)
        need = (bound >= xref) & (bound > 0)
        need[true_idx] = True
        best, bi = 0.0, -1
        for c in np.nonzero(need)[0]:
            if best < x[c]:
                best, bi = x[c], c
        fb, fi = 0.0, -1
        for c in range(len(x)):
            if fb < x[c]:
                fb, fi = x[c], c
        assert bi == fi
        if fi >= 0:
            assert (iu[0][fi], iu[1][fi]) == (a1[s], a2[s])
        # in-bag: lo == hi certifies the sequential sum
        need = (bound >= xref * 2.0 ** -70) & (bound > 0)
        need[true_idx] = True
        lo = hi = full = 0.0
        for c in range(len(x)):
            full += x[c]
            if need[c]:
                lo += x[c]; hi += x[c]
            else:
                hi += bound[c]
        assert lo <= full <= hi
        if lo == hi:
            n_certified += 1
            assert lo == full
    assert n_certified > 0.9 * len(geno)


def test_certificate_is_monotone_sandwich():
    """the argument itself: for non-negative terms and 0 <= x_c <= b_c on the skipped cells, the
    sequential fp64 sums satisfy chain(0) <= chain(x) <= chain(b) (round-to-nearest is monotone)"""
    rng = np.random.default_rng(7)
    for _ in range(300):
        n = int(rng.integers(3, 60))
        x = rng.random(n) * 10.0 ** rng.integers(-30, 1, size=n)
        skip = rng.random(n) < 0.6
        b = x * (1 + rng.random(n) * 10.0 ** rng.integers(-3, 3, size=n))
        lo = hi = full = 0.0
        for c in range(n):
            full += x[c]
            lo += 0.0 if skip[c] else x[c]
            hi += b[c] if skip[c] else x[c]
        assert lo <= full <= hi


def _class_bounds(haplo, n_hla, n_snp, g, T, Tf, K3, k):
    """The class-split bound of DESIGN.md 9 (not built on the device yet): per-allele sums split by
    the haplotypes' alleles at the sample's first k heterozygous SNPs; a pair of classes (s, t) agrees
    on k_eff - popc(s ^ t) of them, and every agreement on a heterozygous SNP is one more mismatch."""
    H = _decode(haplo, n_snp)
    f = haplo["freq"]
    hom0, hom2 = g == 0, g == 2
    het = np.nonzero(g == 1)[0][:k]
    k_eff = len(het)
    c = (H[:, hom0] == 1).sum(1) + (H[:, hom2] == 0).sum(1)
    u = f * Tf[c]
    cls = (H[:, het].astype(np.int64) << np.arange(k_eff)).sum(1) if k_eff else np.zeros(len(f), dtype=np.int64)
    Uc = np.zeros((n_hla, 1 << k))
    np.add.at(Uc, (haplo["hla"], cls), u)
    M = np.array([[Tf[max(k_eff - bin(a ^ b).count("1"), 0)] for b in range(1 << k)] for a in range(1 << k)])
    B = (Uc @ M @ Uc.T) * K3
    iu = np.triu_indices(n_hla)
    return B[iu]


@pytest.mark.parametrize("n_snp", [5, 23, 40, 70])
def test_class_split_bound_is_a_bound_and_tighter(consts, orc, n_snp):
    """groundwork for the next screening level: the class-split bound holds
    against every cell value of the oracle's chain, never exceeds the product bound by more than its
    rounding slack, and keeps fewer cells at the in-bag threshold"""
    T, Tf, K = consts
    # T[d] <= kappa3 * T'[p] * T'[q] * T'[m] for p + q + m <= d, m <= 3 -- taken over the host table itself
    suffix = np.maximum.accumulate(T[::-1])[::-1]
    kappa3 = 1.0
    for m in range(4):
        p, q = np.meshgrid(np.arange(257), np.arange(257), indexing="ij")
        ok = p + q + m <= 256
        r = suffix[np.minimum(p + q + m, 256)][ok] / (Tf[p[ok]] * Tf[q[ok]] * Tf[m])
        kappa3 = max(kappa3, float(r.max()))
    assert kappa3 < 1 + 1e-6
    K3 = 2.0 * kappa3 * (1 + 1e-8)
    rng = np.random.default_rng(300 + n_snp)
    haplo, n_hla, n_snp = helpers.random_haplo_list(rng, n_hla=9, n_snp=n_snp, max_per_allele=7)
    geno = helpers.random_genotypes(rng, 100, n_snp, n_hla, haplo=haplo)
    G = _genotype_rows(geno, n_snp)
    p2, s2 = orc.post_prob2(haplo, n_hla, n_snp, geno)
    pp = p2 * s2[:, None]
    kept = {0: 0, 1: 0, 2: 0, 3: 0}
    for s in range(len(geno)):
        t1, t2 = sorted((int(geno["a1"][s]), int(geno["a2"][s])))
        b0, xref = _bounds_and_xref(haplo, n_hla, n_snp, G[s], t1, t2, T, Tf, K)
        kept[0] += int(((b0 >= xref * 2.0 ** -70) & (b0 > 0)).sum())
        for k in (1, 2, 3):
            bk = _class_bounds(haplo, n_hla, n_snp, G[s], T, Tf, K3, k)
            assert np.all(pp[s] <= bk * (1 + 1e-12)), "a cell value exceeds its class-split bound"
            assert np.all(bk <= b0 * (1 + 1e-6))
            kept[k] += int(((bk >= xref * 2.0 ** -70) & (bk > 0)).sum())
    assert kept[3] <= kept[2] <= kept[1] <= kept[0]
