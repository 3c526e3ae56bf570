#!/bin/bash
# round 2, call 25 (1 GPU): K1 identity-line loads with an L2 evict-last keep (on top of the evict-first stream, now default)
set -u
OUT=gpurun_out/call25; mkdir -p $OUT
run() { # lib workload extra...
  lib=$1; w=$2; shift 2
  if [ "$lib" = default ]; then unset FA_LIB_NAME; else export FA_LIB_NAME=$lib; fi
  timeout 300 python bench.py --workload $w --no-cpu --no-e2e --no-verify --batch $((1<<26)) --steps 8 --warmup 4 "$@" 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $w $*: %7.0f Mpkts/s  frac %.3f  ms/step %.3f' % (d['value'], d['roofline']['frac'], d['ms_per_step']))
except Exception as e: print('$lib $w FAILED', e)
"
}
for w in zipf10m zipf1m uniform10m; do
  for l in default libflowagg_keep.so default libflowagg_keep.so; do run $l $w; done
done 2>&1 | tee $OUT/ab.log
echo "== parity with the 512 build"
FA_LIB_NAME=libflowagg_keep.so timeout 900 python -m pytest tests -x -q -m gpu -k "parity or events or sketch" > $OUT/gpu_suite_keep.log 2>&1; tail -3 $OUT/gpu_suite_keep.log
