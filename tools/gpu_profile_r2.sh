#!/bin/bash
# ncu evidence of the round-2 build (recipe: /opt/skills/guides/B200_PROFILING.md). One GPU.
mkdir -p gpurun_out
# every launch of one classifier with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r02_launches_train.csv python tools/profile_train.py > gpurun_out/r02_ncu_launches.log 2>&1
tail -n 1 gpurun_out/r02_ncu_launches.log | cut -c1-300
# full sets: the EM kernel of a late round; an out-of-bag and an in-bag gather launch of a late round
ncu --set full --clock-control none --import-source on -k regex:em_chain -s 25 -c 1 -o gpurun_out/r02_prof_em_chain_final python tools/profile_train.py > gpurun_out/r02_ncu_em_final.log 2>&1
tail -n 1 gpurun_out/r02_ncu_em_final.log | cut -c1-200
ncu --set full --clock-control none --import-source on -k regex:cell_gather_kernel -s 60 -c 2 -o gpurun_out/r02_prof_gather_final python tools/profile_train.py > gpurun_out/r02_ncu_gather_final.log 2>&1
tail -n 1 gpurun_out/r02_ncu_gather_final.log | cut -c1-200
ls -la gpurun_out/ | grep "r02_.*final\|r02_launches"
