"""Multi-GPU parity on REAL ranks: one process per GPU under torch.distributed.run,
NCCL. Needs >= 2 visible GPUs (skipped on a one-GPU box; `gpurun --gpus 2 -- python -m pytest
tests/test_multigpu.py -m gpu` runs it; the committed output is profiles/r02_multigpu_check_2gpu.txt).

tools/multigpu_check.py asserts, on every rank:
  1. classifier-sharded training: the gathered model == the model one rank builds (bit for bit);
  2. sample-sharded prediction (host API and device helper): concatenated slices == one rank, bit exact;
  3. classifier-sharded prediction + ONE NCCL all-reduce per tile: calls equal, posteriors within
     1e-10 relative of the sequential classifier order (reference src/LibHLA.cpp:2414-2482).
"""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [2])
def test_sharded_training_and_prediction_on_real_ranks(gpu, world):
    if gpu.device_count() < world:
        pytest.skip("needs %d GPUs, this box shows %d" % (world, gpu.device_count()))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tools", "multigpu_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, (out.stdout + out.stderr)[-3000:]
    line = [ln for ln in out.stdout.splitlines() if ln.startswith("{")][-1]
    r = json.loads(line)
    assert r["world"] == world
    assert r["train_shard_equals_single"] and r["sample_sharded_bit_exact"]
    assert r["classifier_sharded_calls_equal"] and r["classifier_sharded_max_rel_err"] <= 1e-10
    assert r["allreduce_bytes"] > 0
