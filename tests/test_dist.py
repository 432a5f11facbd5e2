"""CPU tests of the multi-process plumbing (gloo, world_size 2)."""
import os
import subprocess
import sys

import numpy as np

from hibag_b200 import dist as hd

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shards_partition_everything():
    for n in (0, 1, 7, 100, 200001):
        for w in (1, 2, 3, 8):
            seen = []
            for r in range(w):
                b, e = hd.shard_range(n, r, w)
                assert 0 <= b <= e <= n
                seen += list(range(b, e)) if n < 1000 else [(b, e)]
            if n < 1000:
                assert seen == list(range(n))
            else:
                assert seen[0][0] == 0 and seen[-1][1] == n
                assert all(seen[i][1] == seen[i + 1][0] for i in range(w - 1))
            got = sorted(k for r in range(w) for k in hd.classifier_indices(min(n, 50), r, w))
            assert got == list(range(min(n, 50)))


WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
from hibag_b200 import dist as hd
rank, local_rank, world = hd.init(backend="gloo")
assert world == 2
mine = [(k, dict(snpidx=np.arange(k + 1), tag=k * 10)) for k in hd.classifier_indices(5, rank, world)]
model = hd.gather_classifiers(mine)
assert [c["tag"] for c in model] == [0, 10, 20, 30, 40], model
assert hd.max_over_ranks(1.0 + rank) == 2.0 and hd.sum_over_ranks(1.0 + rank) == 3.0
acc = torch.full((4, 6), float(rank + 1), dtype=torch.float64)
hd.allreduce_partial(acc)
assert torch.all(acc == 3.0)
b, e = hd.shard_range(11, rank, world)
part = torch.zeros(11, dtype=torch.float64); part[b:e] = 1
hd.allreduce_partial(part)
assert torch.all(part == 1.0)
hd.barrier()
print("rank", rank, "ok")
"""


def test_two_process_gloo_gather_and_reduce(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29533", WORLD_SIZE="2")
    procs = []
    for r in range(2):
        e = dict(env, RANK=str(r), LOCAL_RANK=str(r))
        procs.append(subprocess.Popen([sys.executable, str(script)], env=e, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for r, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0, o
        assert "rank %d ok" % r in o
