#!/bin/bash
# ncu evidence of the final round-2 build (each distinct genotype scored once). One GPU.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_train_distinct_once.csv python tools/profile_train.py > gpurun_out/r02f_ncu_train.log 2>&1
tail -n 1 gpurun_out/r02f_ncu_train.log | cut -c1-300
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_predict_distinct_once.csv python tools/profile_predict.py 200000 12 > gpurun_out/r02f_ncu_predict.log 2>&1
tail -n 2 gpurun_out/r02f_ncu_predict.log | cut -c1-300
# full sets: the pair-scoring kernel on the distinct genotypes of a 200,000-sample tile, the set insert
ncu --set full --clock-control none --import-source on -k regex:cell_pass_kernel -s 3 -c 1 -o gpurun_out/r02_prof_cell_pass_distinct python tools/profile_predict.py 200000 6 > gpurun_out/r02f_ncu_cp.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:dedup_insert_kernel -s 3 -c 1 -o gpurun_out/r02_prof_dedup_insert python tools/profile_predict.py 200000 6 > gpurun_out/r02f_ncu_dd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:predict_accumulate_dedup_kernel -s 3 -c 1 -o gpurun_out/r02_prof_accumulate_dedup python tools/profile_predict.py 200000 6 > gpurun_out/r02f_ncu_acc.log 2>&1
for f in cell_pass_distinct dedup_insert accumulate_dedup; do ncu -i gpurun_out/r02_prof_$f.ncu-rep --page raw --csv > gpurun_out/r02_${f}_ncu_raw.csv 2>/dev/null; done
ls -la gpurun_out | grep "r02_.*distinct\|r02_.*dedup"
