#!/bin/bash
# round 2, call 15 (1 GPU): ncu captures of the K6 folds and the (f4) parse kernels, to see what bounds them
set -u
OUT=gpurun_out/call15; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'feature_fold_kernel|first_kernel' -s 4 -c 4 -o $OUT/prof_k6 -f python tools/bench_aux.py features > $OUT/ncu_k6.log 2>&1; tail -2 $OUT/ncu_k6.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'snap_' -s 2 -c 2 -o $OUT/prof_snaps -f python tools/bench_aux.py snaps > $OUT/ncu_snaps.log 2>&1; tail -2 $OUT/ncu_snaps.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_features.csv python tools/bench_aux.py features > $OUT/launches_features.log 2>&1; tail -1 $OUT/launches_features.log | cut -c1-300
