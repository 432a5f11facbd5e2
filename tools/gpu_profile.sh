#!/bin/bash
# ncu evidence for the dominant kernel (recipe: /opt/skills/guides/B200_PROFILING.md)
mkdir -p gpurun_out
python tools/profile_predict.py 65536 12 2>&1 | tail -3
# every launch with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_predict.csv python tools/profile_predict.py 65536 12 > gpurun_out/ncu_launches.log 2>&1
tail -2 gpurun_out/ncu_launches.log
# the top kernel, full set, 2 launches after 3 warm-up launches
ncu --set full --clock-control none --import-source on -k regex:cell_pass -s 3 -c 2 -o gpurun_out/prof_cell python tools/profile_predict.py 65536 6 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out/
