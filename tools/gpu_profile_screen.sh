#!/bin/bash
# ncu evidence for the screened training step (recipe: /opt/skills/guides/B200_PROFILING.md)
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err; tail -2 gpurun_out/bench_1gpu.err
# every launch of one classifier with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train_screen.csv python tools/profile_train.py > gpurun_out/ncu_launches_train.log 2>&1
# the gather form of the pair-scoring kernel, full set: the out-of-bag and the in-bag launch of round ~30
ncu --set full --clock-control none --import-source on -k regex:cell_gather -s 60 -c 2 -o gpurun_out/prof_gather python tools/profile_train.py > gpurun_out/ncu_full_gather.log 2>&1
tail -2 gpurun_out/ncu_full_gather.log
python - <<'PY'
import sys; sys.path.insert(0, '.')
from hibag_b200 import api
api.set_device(0); info = api.device_info()
for w, nm in ((6, "dependent DADD chain"), (7, "dependent DMUL+DADD chain")):
    ops, ms = api.pipe_peak(w)
    print("%s: %.3f Gop/s one warp -> %.1f cycles per op" % (nm, ops / 1e9, 32 * info['clock_khz'] * 1e3 / ops))
PY
ls -la gpurun_out/
