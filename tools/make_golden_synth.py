"""Golden fixture for the many-allele training parity test: the compiled, unmodified reference
(oracle/_ref, target 'base') trains 2 classifiers on a seeded synthetic cohort (900 samples x 160
SNPs, 45 alleles drawn); the classifiers are stored in tests/golden/synth_many_alleles_ref.npz.
Run in the build container (needs /root/reference to have built oracle/_ref)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import refpy
from hibag_b200 import synth, api

SPEC = dict(n_samp=900, n_snp=160, n_hla=45, cohort_seed=8, train_seed=900, n_cls=2)


def main():
    ref = refpy.RefLib(); ref.set_target("base"); ref.set_gpu_procs(None)
    coh = synth.make_cohort(SPEC["n_samp"], SPEC["n_snp"], SPEC["n_hla"], seed=SPEC["cohort_seed"])
    r = ref.new_model()
    r.init_training(coh.geno, coh.h1, coh.h2, coh.n_hla)
    r.build(SPEC["n_cls"], api.default_mtry(coh.n_snp), prune=True, reseed_base=SPEC["train_seed"], first_index=0)
    out = {k: np.array(v) for k, v in SPEC.items()}
    for k in range(SPEC["n_cls"]):
        c = r.classifier(k)
        for key in ("snpidx", "samp_num", "freq", "hla", "packed"):
            out["c%d_%s" % (k, key)] = np.asarray(c[key])
        out["c%d_oob_acc" % k] = np.array(c["oob_acc"])
    path = os.path.join(ROOT, "tests", "golden", "synth_many_alleles_ref.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
