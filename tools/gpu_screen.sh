#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "screening or golden_model or many_alleles or on_synthetic or fallback" 2>&1 | tail -25 | tee gpurun_out/pytest_screen.txt
( timeout 300 python tools/train_probe.py 0:1:1:1 0:6:1:1 0:12:1:1 0:16:1:1
  HIBAG_B200_SCREEN_TAU_LOG2=63 timeout 300 python tools/train_probe.py 0:12:1:1
  HIBAG_B200_SCORE_QUEUES=6 timeout 300 python tools/train_probe.py 0:12:1:1 ) 2>&1 | grep -v Warning | tee gpurun_out/probe_screen.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train_screen.csv python tools/profile_train.py > gpurun_out/ncu_launches_train.log 2>&1
