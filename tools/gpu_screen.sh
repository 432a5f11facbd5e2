#!/bin/bash
# screening bring-up: targeted parity tests, then the config-2 probe with and without screening
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "screening or golden_model or many_alleles or on_synthetic or fallback" 2>&1 | tail -25 | tee gpurun_out/pytest_screen.txt
timeout 600 python tools/train_probe.py 0:6:1:0 0:6:1:1 0:10:1:1 2>&1 | tail -8 | tee gpurun_out/probe_screen.txt
