#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "screening or golden_model or many_alleles or on_synthetic or fallback" 2>&1 | tail -25 | tee gpurun_out/pytest_screen.txt
( timeout 300 python tools/train_probe.py 0:1:1:1 0:12:1:1 0:24:1:1
  HIBAG_B200_EM_DENSE=1 timeout 300 python tools/train_probe.py 0:1:1:1 0:12:1:1 0:24:1:1
  HIBAG_B200_EM_DENSE=1 HIBAG_B200_EM_CLUSTER=1 timeout 300 python tools/train_probe.py 0:24:1:1 0:32:1:1
  HIBAG_B200_EM_DENSE=1 HIBAG_B200_EM_CLUSTER=4 timeout 300 python tools/train_probe.py 0:24:1:1 ) 2>&1 | grep -v Warning | tee gpurun_out/probe_screen.txt
HIBAG_B200_EM_DENSE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "golden_model or many_alleles or on_synthetic" 2>&1 | tail -3
