#!/bin/bash
# round 2, call 21 (1 GPU): full GPU suite with the flow filter; (f4) throughput with a 12-rule filter
set -u
OUT=gpurun_out/call21; mkdir -p $OUT
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/gpu_suite.log 2>&1; tail -4 $OUT/gpu_suite.log
timeout 600 python tools/bench_aux.py snaps 2>&1 | grep "^{" | tee $OUT/aux_snaps.jsonl
