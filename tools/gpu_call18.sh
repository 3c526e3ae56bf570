#!/bin/bash
# round 2, call 18 (1 GPU): ncu of the current K6 fold (hot-flow cache, batched staging loads, two-round-trip probe)
set -u
OUT=gpurun_out/call18; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'feature_fold_kernel' -s 4 -c 2 -o $OUT/prof_k6 -f python tools/bench_aux.py features > $OUT/ncu_k6.log 2>&1; tail -2 $OUT/ncu_k6.log
