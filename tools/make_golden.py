"""Generate tests/golden/*.npz from the reference's shipped fixtures (run in the dev container,
where /root/reference exists; the GPU box only sees the committed .npz files).

  hapmap_ceu.npz     : data/HapMap_CEU_Geno.rdata + data/HLA_Type_Table.rdata
                       (60 samples x 1564 SNPs, HLA types for A,B,C,DQA1,DQB1,DRB1)
  modellist_a.npz    : inst/extdata/ModelList.RData $A -- the 100-classifier HLA-A model that
                       vignettes/HIBAG.Rmd:113-118 builds with set.seed(100); this is the
                       reference's own golden vector for the whole training path.

Usage:  python tools/make_golden.py [/root/reference]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import rdx2  # noqa: E402

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

# hg19 gene coordinates used by hlaFlankingSNP (reference inst/doc/GeneInfo_hg19.txt)
GENE_HG19 = {"A": (29910247, 29913661), "B": (31321649, 31324989), "C": (31236526, 31239913),
             "DRB1": (32546546, 32557613), "DQA1": (32605169, 32612152),
             "DQB1": (32627241, 32634466), "DPB1": (33043703, 33057473)}


def main():
    os.makedirs(OUT, exist_ok=True)
    g = rdx2.load(os.path.join(REF, "data", "HapMap_CEU_Geno.rdata"))["HapMap_CEU_Geno"]
    n_snp, n_samp = (int(x) for x in g["genotype"].attr["dim"].value)
    geno = g["genotype"].value.reshape(n_samp, n_snp).copy()      # column-major [snp x samp]
    geno[geno == rdx2.NA_INTEGER] = -1
    assert geno.min() >= -1 and geno.max() <= 2
    t = rdx2.load(os.path.join(REF, "data", "HLA_Type_Table.rdata"))["HLA_Type_Table"]
    loci = ["A", "B", "C", "DQA1", "DQB1", "DRB1"]
    hla = {}
    for loc in loci:
        for k in (1, 2):
            col = t["%s.%d" % (loc, k)]
            v = col.value
            if isinstance(v, np.ndarray):        # factor
                lev = col.attr["levels"].value
                v = [None if x == rdx2.NA_INTEGER else lev[x - 1] for x in v]
            hla["hla_%s_%d" % (loc, k)] = np.array(["" if x is None else x for x in v])
    np.savez_compressed(
        os.path.join(OUT, "hapmap_ceu.npz"),
        genotype=geno.astype(np.int8), sample_id=np.array(g["sample.id"].value),
        snp_id=np.array(g["snp.id"].value),
        snp_position=np.asarray(g["snp.position"].value, dtype=np.int64),
        hla_sample_id=np.array(t["sample.id"].value),
        gene_names=np.array(sorted(GENE_HG19)),
        gene_start=np.array([GENE_HG19[k][0] for k in sorted(GENE_HG19)], dtype=np.int64),
        gene_end=np.array([GENE_HG19[k][1] for k in sorted(GENE_HG19)], dtype=np.int64),
        **hla)

    A = rdx2.load(os.path.join(REF, "inst", "extdata", "ModelList.RData"))["modellist"]["A"]
    alleles = list(A["hla.allele"].value)
    cls = A["classifiers"].value
    n_cls = len(cls)
    samp_num = np.zeros((n_cls, int(A["n.samp"].value[0])), dtype=np.int32)
    snp_off = [0]; snpidx = []
    hap_off = [0]; freq = []; hidx = []; packed = []
    acc = np.zeros(n_cls)
    for k, c in enumerate(cls):
        samp_num[k] = c["samp.num"].value
        s = np.asarray(c["snpidx"].value, dtype=np.int32) - 1       # fixture is 1-based
        snpidx.append(s); snp_off.append(snp_off[-1] + len(s))
        h = c["haplos"]
        freq.append(np.asarray(h["freq"].value, dtype=np.float64))
        hidx.append(np.array([alleles.index(x) for x in h["hla"].value], dtype=np.int32))
        for st in h["haplo"].value:
            assert len(st) == len(s)
            w = [0, 0]
            for j, ch in enumerate(st):
                if ch == "1":
                    w[j >> 6] |= 1 << (j & 63)
            packed.append(w)
        hap_off.append(hap_off[-1] + len(h["freq"].value))
        acc[k] = c["outofbag.acc"].value[0]
    np.savez_compressed(
        os.path.join(OUT, "modellist_a.npz"),
        sample_id=np.array(A["sample.id"].value), snp_id=np.array(A["snp.id"].value),
        hla_allele=np.array(alleles), samp_num=samp_num,
        snp_off=np.array(snp_off, dtype=np.int64), snpidx=np.concatenate(snpidx),
        hap_off=np.array(hap_off, dtype=np.int64), freq=np.concatenate(freq),
        hla_idx=np.concatenate(hidx), packed=np.array(packed, dtype=np.uint64),
        oob_acc=acc, seed=np.int64(100), mtry=np.int64(17), nclassifier=np.int64(n_cls))
    for f in ("hapmap_ceu.npz", "modellist_a.npz"):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()
