#!/bin/bash
# round 2, call 3: single-line table layout (accumulators inside the identity line) — full GPU suite, K1 vs K1s, ncu
set -u
OUT=gpurun_out/call4; mkdir -p $OUT
echo "== full GPU suite"
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/gpu_suite.log 2>&1; tail -3 $OUT/gpu_suite.log
echo "== K1s parity"; if false; then
FA_K1_OPT=256 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_sketch.py tests/test_gpu_features.py -x -q -m gpu 2>&1 | tail -2
fi
for v in old 0 4 256; do for w in zipf10m uniform10m zipf1m; do
  echo "-- FA_K1_OPT=$v $w"
  if [ $v = old ]; then export FA_LIB_NAME=libflowagg_oldlayout.so; vv=0; else unset FA_LIB_NAME; vv=$v; fi
  FA_K1_OPT=$vv timeout 300 python bench.py --workload $w --no-cpu --no-e2e --no-verify --batch $((1<<26)) --steps 8 --warmup 4 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%7.0f Mpkts/s  frac %.3f  ms/step %.3f  flows %d clocks %s' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['config']['live_flows'], d['clocks']['sm_mhz']))
except Exception as e: print('FAILED', e)
"
done; done 2>&1 | tee $OUT/ab.log; unset FA_LIB_NAME
echo "== bench default (zipf10m, verify, e2e, cpu)"
timeout 900 python bench.py --steps 20 --warmup 5 > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 1200 $OUT/bench_default.json; tail -5 $OUT/bench_default.err
echo "== ncu: K1 on zipf10m, a late launch (table warm)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 60 -c 1 -o $OUT/prof_k1_zipf10m -f \
    python bench.py --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 4 --warmup 6 > $OUT/ncu_k1.log 2>&1; tail -2 $OUT/ncu_k1.log
