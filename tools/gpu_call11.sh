#!/bin/bash
# round 2, call 11 (1 GPU): two-way cache + packed probe results as the default K1; staging decoupled from max_batch;
# full GPU suite, then bench lines at launch sizes 2^22 / 2^23 and the default line (verify + e2e)
set -u
OUT=gpurun_out/call11; mkdir -p $OUT
echo "== full GPU suite"
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/gpu_suite.log 2>&1; tail -4 $OUT/gpu_suite.log
run() { # workload extra...
  w=$1; shift
  timeout 300 python bench.py --workload $w --no-cpu --no-e2e --no-verify --batch $((1<<26)) --steps 8 --warmup 4 "$@" 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w $*: %7.0f Mpkts/s  frac %.3f  ms/step %.3f' % (d['value'], d['roofline']['frac'], d['ms_per_step']))
except Exception as e: print('$w FAILED', e)
"
}
for w in zipf10m uniform10m zipf1m; do
  run $w --max-batch $((1<<22)); run $w --max-batch $((1<<23)); run $w --max-batch $((1<<24))
done 2>&1 | tee $OUT/ab.log
echo "== bench default (zipf10m, verify, e2e, cpu)"
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 2500 $OUT/bench_default.json; tail -5 $OUT/bench_default.err
