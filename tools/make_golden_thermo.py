"""Golden fixture for the many-SNP training parity test: the compiled, unmodified reference
(oracle/_ref, target 'base') trains one classifier on each of two seeded "thermometer" cohorts
(hibag_b200/synth.py: make_thermo_cohort) whose classifiers grow to 33-64 and to more than 64 SNPs --
the two- and four-word forms of the packed genotypes, which the BASELINE-shaped cohorts never reach
(their classifiers stop at 20-27 SNPs). Stored in tests/golden/synth_thermo_ref.npz.
Run in the build container (needs /root/reference to have built oracle/_ref)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refpy
from hibag_b200 import synth

SPECS = [dict(n_samp=500, n_hla=40, cohort_seed=1, noise=0.01, train_seed=5),
         dict(n_samp=800, n_hla=70, cohort_seed=2, noise=0.002, train_seed=5)]


def main():
    ref = refpy.RefLib(); ref.set_target("base"); ref.set_gpu_procs(None)
    out = {"n_specs": np.array(len(SPECS))}
    for i, sp in enumerate(SPECS):
        geno, h1, h2 = synth.make_thermo_cohort(sp["n_samp"], sp["n_hla"], seed=sp["cohort_seed"], noise=sp["noise"])
        r = ref.new_model()
        r.init_training(geno, h1, h2, sp["n_hla"])
        r.build(1, geno.shape[1], prune=True, reseed_base=sp["train_seed"], first_index=0)
        c = r.classifier(0)
        for k, v in sp.items():
            out["s%d_%s" % (i, k)] = np.array(v)
        for key in ("snpidx", "samp_num", "freq", "hla", "packed"):
            out["s%d_%s" % (i, key)] = np.asarray(c[key])
        out["s%d_oob_acc" % i] = np.array(c["oob_acc"])
        print("spec", i, "SNPs", len(c["snpidx"]), "haplotypes", len(c["freq"]), "oob acc", c["oob_acc"])
    path = os.path.join(ROOT, "tests", "golden", "synth_thermo_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
