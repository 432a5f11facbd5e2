"""CPU tests: the C-ABI library loads, exports every symbol include/hibag_b200.h declares, the
POD layouts match the reference's, host-only entry points work, and compute entry points fail
loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from tests import helpers

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "hibag_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hibag_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(built):
    from hibag_b200 import api
    L = api.lib()
    declared = header_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), "libhibag_b200.so does not export " + name
    assert sorted(api.EXPORTS) == declared


def test_pod_layouts_match_reference(built):
    from hibag_b200 import api
    from oracle import refpy
    assert api.HAPLO_DT.itemsize == 32 and api.GENO_DT.itemsize == 48
    assert api.HAPLO_DT == refpy.HAPLO_DT and api.GENO_DT == refpy.GENO_DT
    assert api.HAPLO_DT.fields["freq"][1] == 16 and api.HAPLO_DT.fields["hla"][1] == 28
    assert api.GENO_DT.fields["s2"][1] == 16 and api.GENO_DT.fields["boot"][1] == 32
    assert api.GENO_DT.fields["a1"][1] == 36 and api.GENO_DT.fields["a2"][1] == 40


def test_plugin_struct_has_ten_hooks_in_reference_order(built):
    from hibag_b200 import api
    ptr = api.get_procs()
    assert ptr
    hooks = (C.c_void_p * 10).from_address(ptr)
    # build_haplomatch (slot 3) is optional and left NULL; all other hooks are installed
    for i in range(10):
        assert (hooks[i] is None) == (i == 3), i


def test_host_rng_is_r_mersenne_twister(built, ref):
    from hibag_b200 import api
    for seed in (1, 100, 2**31 - 1):
        ref.set_seed(seed)
        want = np.array([ref.lib.ref_unif_rand() for _ in range(1500)])
        assert np.array_equal(api.host_unif_rand(seed, 1500), want)


def test_host_task_list_covers_every_cell_once(built):
    from hibag_b200 import api
    rng = np.random.default_rng(4)
    for n_hla, n_snp in ((1, 5), (7, 20), (40, 23), (100, 128)):
        h, _, _ = helpers.random_haplo_list(rng, n_hla, n_snp, max_per_allele=9)
        cells, chunks, pairs = api.host_build_tasks(h, n_hla, n_snp, target_chunks=64)
        n_cells = n_hla * (n_hla + 1) // 2
        assert sorted(cells[:, 4]) == list(range(n_cells))
        lens = np.bincount(h["hla"], minlength=n_hla)
        starts = np.concatenate([[0], np.cumsum(lens)])
        cost, want_pairs = [], 0
        for c in cells:
            a = int(np.searchsorted(starts, c[0], side="right") - 1) if c[1] > 0 else None
            assert c[5] in (0, 1)
            if c[5]:
                assert c[0] == c[2] and c[1] == c[3]
                want_pairs += c[1] * (c[1] + 1) // 2
                cost.append(c[1] * (c[1] + 1) // 2 + 2 * c[1])
            else:
                want_pairs += c[1] * c[3]
                cost.append(c[1] * c[3] + 2 * c[1])
        assert pairs == want_pairs
        assert all(cost[i] >= cost[i + 1] for i in range(len(cost) - 1))     # longest chains first
        assert chunks[0, 0] == 0 and chunks[-1, 1] == n_cells
        assert np.array_equal(chunks[1:, 0], chunks[:-1, 1]) and np.all(chunks[:, 1] > chunks[:, 0])


def test_unsorted_haplotype_list_is_rejected(built):
    from hibag_b200 import api
    rng = np.random.default_rng(4)
    h, n_hla, n_snp = helpers.random_haplo_list(rng, 5, 10)
    h["hla"] = h["hla"][::-1].copy()
    with pytest.raises(RuntimeError, match="grouped by HLA allele"):
        api.host_build_tasks(h, n_hla, n_snp)


def test_compute_fails_loudly_without_gpu(built):
    from hibag_b200 import api
    if api.device_count() > 0:
        pytest.skip("a CUDA device is present")
    rng = np.random.default_rng(4)
    h, n_hla, n_snp = helpers.random_haplo_list(rng, 5, 10)
    g = helpers.random_genotypes(rng, 4, n_snp, n_hla)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        api.best_guess(h, n_hla, n_snp, g)
    m = api.HLAModel(10, 5)
    m.set_training(np.zeros((6, 10), dtype=np.int8), np.zeros(6, dtype=np.int32), np.ones(6, dtype=np.int32))
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        m.train(1, 3)
    # page-locked result buffers come from the CUDA runtime: no device, no buffer (and no pageable stand-in)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        api.pinned_empty((4, 4))
    m.add_classifier(np.arange(3, dtype=np.int32), h["freq"], h["hla"], h["packed"])
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        m.predict(np.zeros((2, 10), dtype=np.int8))


def test_default_mtry_rule():
    from hibag_b200 import api
    assert api.default_mtry(266) == 17 and api.default_mtry(500) == 23      # ceil(sqrt(n))
    assert api.default_mtry(100, "all") == 100 and api.default_mtry(100, "one") == 1
    assert api.default_mtry(100, 0.25) == 25 and api.default_mtry(10, 50) == 10


def test_ctypes_mirrors_have_the_header_struct_sizes(built, tmp_path):
    """hibag_b200/api.py mirrors the structs of include/hibag_b200.h by hand: compile the header as C
    and compare sizeof (a missing field would let the library write past the Python object)."""
    import ctypes as C
    import os
    import subprocess
    from hibag_b200 import api
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sizes.c"
    src.write_text('#include <stdio.h>\n#include "hibag_b200.h"\nint main(void) { printf("%zu %zu %zu %zu %zu %zu %zu\\n", '
                   'sizeof(hibag_haplotype), sizeof(hibag_genotype), sizeof(hibag_gpu_ext_proc), '
                   'sizeof(hibag_b200_train_opts), sizeof(hibag_b200_train_stats), sizeof(hibag_b200_predict_out), '
                   'sizeof(hibag_b200_predict_stats)); return 0; }\n')
    exe = tmp_path / "sizes"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)], text=True).split()]
    assert sizes[0] == api.HAPLO_DT.itemsize and sizes[1] == api.GENO_DT.itemsize
    assert sizes[2] == 10 * C.sizeof(C.c_void_p)
    assert sizes[3] == C.sizeof(api.TrainOpts)
    assert sizes[4] == C.sizeof(api.TrainStats)
    assert sizes[5] == C.sizeof(api.PredictOut)
    assert sizes[6] == C.sizeof(api.PredictStats)
