"""Small training run for compute-sanitizer: four classifiers of the HapMap golden model (screened passes,
em_chain_kernel), compared with the reference's golden classifiers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hibag_b200 import api
from tests import helpers
api.set_device(0)
geno, h1, h2, al, ml = helpers.hapmap_a_training()
m = api.HLAModel(geno.shape[1], len(al), al)
m.set_training(geno, h1, h2)
m.train(4, int(ml["mtry"]), prune=True, seed=int(ml["seed"]), n_threads=2)
for k in range(4):
    helpers.assert_classifier_equals_golden(m.classifier(k), ml, k)
print("ok", m.train_stats()["em_iterations"])
