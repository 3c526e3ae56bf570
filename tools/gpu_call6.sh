#!/bin/bash
# round 2, call 6: two-line layout restored + 8 teams x 128 + next-tile L2 prefetch; full GPU suite; default bench
set -u
OUT=gpurun_out/call6; mkdir -p $OUT
echo "== full GPU suite"
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/gpu_suite.log 2>&1; tail -4 $OUT/gpu_suite.log
run() { # lib opt workload
  if [ "$1" = default ]; then unset FA_LIB_NAME; else export FA_LIB_NAME=$1; fi
  FA_K1_OPT=$2 timeout 300 python bench.py --workload $3 --no-cpu --no-e2e --no-verify --batch $((1<<26)) --steps 8 --warmup 4 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1 opt=$2 $3: %7.0f Mpkts/s  frac %.3f  ms/step %.3f' % (d['value'], d['roofline']['frac'], d['ms_per_step']))
except Exception as e: print('$1 opt=$2 $3 FAILED', e)
"
}
for w in zipf10m uniform10m zipf1m; do
  run libflowagg_oldlayout.so 0 $w
  run default 0 $w
  run default 32 $w
done 2>&1 | tee $OUT/ab.log
unset FA_LIB_NAME
echo "== bench default (zipf10m, verify, e2e, cpu)"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 1500 $OUT/bench_default.json; tail -5 $OUT/bench_default.err
echo "== ncu: K1 on zipf10m, a late launch"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 60 -c 1 -o $OUT/prof_k1_zipf10m -f \
    python bench.py --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 4 --warmup 6 > $OUT/ncu_k1.log 2>&1; tail -2 $OUT/ncu_k1.log
