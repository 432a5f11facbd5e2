#!/bin/bash
# first GPU call of the round: environment facts, pipe peaks, parity tests
mkdir -p gpurun_out
nproc > gpurun_out/nproc.txt; lscpu | head -20 >> gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv >> gpurun_out/nproc.txt
python - <<'PY' > gpurun_out/pipe_peaks.txt 2>&1
import sys; sys.path.insert(0, '.')
from hibag_b200 import api
api.set_device(0)
info = api.device_info(); print(info)
names = ["POPC.32", "LOP3", "DMUL+DADD", "DFMA", "LDS.64 lane-private", "IADD3"]
for w, nm in enumerate(names):
    ops, ms = api.pipe_peak(w)
    print("%-22s %10.3f Gop/s  (%.3f ms)  = %.2f /clk/SM at %d MHz x %d SMs" % (nm, ops / 1e9, ms, ops / (info['clock_khz'] * 1e3 * info['sm_count']), info['clock_khz'] // 1000, info['sm_count']))
PY
cat gpurun_out/pipe_peaks.txt
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/pytest_gpu.txt
