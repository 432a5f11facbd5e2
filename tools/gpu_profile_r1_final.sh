#!/bin/bash
# ncu evidence of the round-1 final build (recipe: /opt/skills/guides/B200_PROFILING.md)
mkdir -p gpurun_out
# every launch of one classifier with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_train_final.csv python tools/profile_train.py > gpurun_out/ncu_launches_train.log 2>&1
tail -n 1 gpurun_out/ncu_launches_train.log | cut -c1-300
# the gather form of the pair-scoring kernel, full set: the out-of-bag and the in-bag launch of round ~30
ncu --set full --clock-control none --import-source on -k regex:cell_gather -s 60 -c 2 -o gpurun_out/prof_gather_final python tools/profile_train.py > gpurun_out/ncu_full_gather.log 2>&1
tail -n 1 gpurun_out/ncu_full_gather.log
# the EM kernel in the shape the 24-lane step uses (512 threads, one CTA per candidate), full set
HIBAG_B200_EM_DENSE=1 HIBAG_B200_EM_CLUSTER=1 ncu --set full --clock-control none --import-source on -k regex:em_kernel -s 30 -c 1 -o gpurun_out/prof_em_final python tools/profile_train.py > gpurun_out/ncu_full_em.log 2>&1
tail -n 1 gpurun_out/ncu_full_em.log
# phase clocks of the EM kernel: 24 lanes (as in the bench step) and one lane alone
HIBAG_B200_EM_PROF=1 python tools/train_probe.py 0:24 2>&1 | tail -n 2 | cut -c1-600 > gpurun_out/em_phase_clocks.txt
HIBAG_B200_EM_PROF=1 HIBAG_B200_EM_DENSE=1 HIBAG_B200_EM_CLUSTER=1 python tools/train_probe.py 0:1 2>&1 | tail -n 2 | cut -c1-600 >> gpurun_out/em_phase_clocks.txt
python tools/gather_alone_probe.py 2>&1 | tail -n 1 >> gpurun_out/em_phase_clocks.txt
cat gpurun_out/em_phase_clocks.txt
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
ls -la gpurun_out/ | tail -n 12
