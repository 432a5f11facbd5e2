#!/bin/bash
mkdir -p gpurun_out
( timeout 300 python tools/train_probe.py 0:6:1:1 0:16:1:1 0:20:1:1 0:24:1:1 0:28:1:1
  HIBAG_B200_SCORE_QUEUES=8 timeout 300 python tools/train_probe.py 0:24:1:1
  HIBAG_B200_SCORE_QUEUES=4 timeout 300 python tools/train_probe.py 0:24:1:1
  HIBAG_B200_EM_RINGS=4 timeout 300 python tools/train_probe.py 0:24:1:1 ) 2>&1 | grep -v Warning | tee gpurun_out/probe_screen2.txt
