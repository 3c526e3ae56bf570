#!/bin/bash
# round 2, call 9: K1 build-time experiments (FA_K1_EXP bits) against the default library, same box
set -u
OUT=gpurun_out/call9; mkdir -p $OUT
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
for w in zipf10m uniform10m zipf1m; do
  for l in default libflowagg_x1.so libflowagg_x2.so libflowagg_x3.so libflowagg_x4.so libflowagg_x5.so default; do run $l $w; done
done 2>&1 | tee $OUT/ab.log
for l in default libflowagg_x5.so; do run $l zipf10m --max-batch $((1<<24)); run $l zipf10m --max-batch $((1<<23)); done 2>&1 | tee -a $OUT/ab.log
echo "== parity with the x5 build"
FA_LIB_NAME=libflowagg_x5.so timeout 900 python -m pytest tests -x -q -m gpu -k "not host_cpp" > $OUT/gpu_suite_x5.log 2>&1; tail -4 $OUT/gpu_suite_x5.log
