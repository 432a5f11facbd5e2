#!/bin/bash
# ncu evidence of the round-2 build (recipe: /opt/skills/guides/B200_PROFILING.md). One GPU.
mkdir -p gpurun_out
# the gather form of the pair-scoring kernel, full set: an out-of-bag and an in-bag launch of a late round,
# once with the entry-flat out-of-bag form (default) and once with the (cell, positions) form
ncu --set full --clock-control none --import-source on -k regex:cell_gather -s 60 -c 2 -o gpurun_out/r02_prof_gather_flat python tools/profile_train.py > gpurun_out/r02_ncu_gather_flat.log 2>&1
tail -n 1 gpurun_out/r02_ncu_gather_flat.log | cut -c1-200
HIBAG_B200_GATHER_FLAT=0 ncu --set full --clock-control none --import-source on -k regex:cell_gather -s 60 -c 2 -o gpurun_out/r02_prof_gather_cell python tools/profile_train.py > gpurun_out/r02_ncu_gather_cell.log 2>&1
tail -n 1 gpurun_out/r02_ncu_gather_cell.log | cut -c1-200
ls -la gpurun_out/ | grep r02_
