#!/bin/bash
# round 2, call 20 (1 GPU): K1 with the early TMA re-arm (own values folded in the E phase, reduce does not read the tile) vs default
set -u
OUT=gpurun_out/call20; mkdir -p $OUT
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
  for l in default libflowagg_early.so default libflowagg_early.so; do run $l $w; done
done 2>&1 | tee $OUT/ab.log
echo "== parity with the early build"
FA_LIB_NAME=libflowagg_early.so timeout 900 python -m pytest tests -x -q -m gpu -k "not host_cpp" > $OUT/gpu_suite_early.log 2>&1; tail -3 $OUT/gpu_suite_early.log
FA_LIB_NAME=libflowagg_early.so timeout 600 python bench.py --no-cpu --no-e2e --steps 8 2>/dev/null | python -c "
import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('early default bench: %.0f frac %.3f parity %s %s' % (d['value'], d['roofline']['frac'], d.get('parity_ok'), d.get('parity_checked')))"
