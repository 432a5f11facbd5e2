#!/bin/bash
# ncu evidence for the training step (recipe: /opt/skills/guides/B200_PROFILING.md)
mkdir -p gpurun_out
# every launch of one classifier with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_train.csv python tools/profile_train.py > gpurun_out/ncu_launches_train.log 2>&1
tail -2 gpurun_out/ncu_launches_train.log
# the batched pair-scoring kernel, full set, one launch in the middle of the classifier
ncu --set full --clock-control none --import-source on -k regex:cell_pass -s 40 -c 1 -o gpurun_out/prof_cell_train python tools/profile_train.py > gpurun_out/ncu_full_train.log 2>&1
tail -2 gpurun_out/ncu_full_train.log
# the EM kernel, full set
ncu --set full --clock-control none --import-source on -k regex:em_kernel -s 30 -c 1 -o gpurun_out/prof_em python tools/profile_train.py > gpurun_out/ncu_em.log 2>&1
tail -2 gpurun_out/ncu_em.log
ls -la gpurun_out/
