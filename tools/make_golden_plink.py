"""Golden fixture for the PLINK BED import: the reference's own example file set
inst/extdata/HapMap_CEU.{bed,bim,fam} (individual-major BED, 90 samples x 5,316 SNPs) stored as
arrays in tests/golden/hapmap_ceu_plink.npz. Its decoded genotypes must equal the reference's
HapMap_CEU_Geno dataset (tests/golden/hapmap_ceu.npz, made by tools/make_golden.py from
data/HapMap_CEU_Geno.rda) on the 60 samples x 1,564 SNPs the two share -- that equality is the pin of
the decoder (tests/test_oracle.py). Run in the build container (needs /root/reference)."""
import os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = "/root/reference/inst/extdata/HapMap_CEU"


def main():
    bed = np.fromfile(SRC + ".bed", dtype=np.uint8)
    fam = [l.split() for l in open(SRC + ".fam") if l.strip()]
    bim = [l.split() for l in open(SRC + ".bim") if l.strip()]
    out = dict(bed=bed, fam_family=np.array([f[0] for f in fam]), fam_id=np.array([f[1] for f in fam]),
               bim_chr=np.array([b[0] for b in bim]), bim_snp=np.array([b[1] for b in bim]),
               bim_pos=np.array([int(b[3]) for b in bim], dtype=np.int64),
               bim_a1=np.array([b[4] for b in bim]), bim_a2=np.array([b[5] for b in bim]))
    path = os.path.join(ROOT, "tests", "golden", "hapmap_ceu_plink.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
