#!/bin/bash
# ncu evidence of the round-2 build (recipe: /opt/skills/guides/B200_PROFILING.md). One GPU.
mkdir -p gpurun_out
# every launch of one classifier with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r02_launches_train.csv python tools/profile_train.py > gpurun_out/r02_ncu_launches.log 2>&1
tail -n 1 gpurun_out/r02_ncu_launches.log | cut -c1-300
# the gather form of the pair-scoring kernel, full set: an out-of-bag and an in-bag launch of a late round
ncu --set full --clock-control none --import-source on -k regex:cell_gather -s 60 -c 2 -o gpurun_out/r02_prof_gather python tools/profile_train.py > gpurun_out/r02_ncu_gather.log 2>&1
tail -n 1 gpurun_out/r02_ncu_gather.log | cut -c1-200
# the shared-memory-resident EM kernel, full set, a late round
ncu --set full --clock-control none --import-source on -k regex:em_resident -s 30 -c 1 -o gpurun_out/r02_prof_em_resident python tools/profile_train.py > gpurun_out/r02_ncu_em.log 2>&1
tail -n 1 gpurun_out/r02_ncu_em.log | cut -c1-200
ls -la gpurun_out/ | grep r02_
