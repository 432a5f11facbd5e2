#!/bin/bash
mkdir -p gpurun_out
( HIBAG_B200_EM_CLUSTER=1 timeout 300 python tools/train_probe.py 0:16:1:1 0:24:1:1 0:32:1:1
  HIBAG_B200_EM_CLUSTER=2 timeout 300 python tools/train_probe.py 0:16:1:1 0:24:1:1 0:32:1:1
  HIBAG_B200_EM_CLUSTER=1 HIBAG_B200_SCREEN_TAU_LOG2=70 timeout 300 python tools/train_probe.py 0:24:1:1
  HIBAG_B200_SCREEN_TAU_LOG2=70 timeout 300 python tools/train_probe.py 0:12:1:1 ) 2>&1 | grep -v Warning | tee gpurun_out/probe_screen2.txt
nproc; free -g | head -2
