#!/bin/bash
# round 2, call 2: memory-pipe microbenchmarks (design input for K1), un-gated KERNEL_MAP tests, the new default bench
set -u
OUT=gpurun_out/call2; mkdir -p $OUT
echo "== microbench (L2-resident table, DRAM-resident table)"
timeout 120 tools/microbench/gather_bench 19 25 | tee $OUT/gather_l2.json
timeout 120 tools/microbench/gather_bench 25 25 | tee $OUT/gather_dram.json
echo "== KERNEL_MAP, un-gated"
timeout 600 python -m pytest tests/test_gpu_kernel_map.py -q -m gpu > $OUT/kmap.log 2>&1; tail -3 $OUT/kmap.log
echo "== bench default (zipf10m, verify)"
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 1500 $OUT/bench_default.json; tail -5 $OUT/bench_default.err
echo "== K1s (FA_K1_OPT=256): parity, then the three workloads against the default K1"
FA_K1_OPT=256 timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py tests/test_gpu_sketch.py tests/test_gpu_features.py -x -q -m gpu 2>&1 | tail -3
for v in 0 256 258; do for w in zipf10m uniform10m zipf1m; do
  echo "-- FA_K1_OPT=$v $w"
  FA_K1_OPT=$v timeout 300 python bench.py --workload $w --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 10 --warmup 3 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%7.0f Mpkts/s  frac %.3f  ms/step %.3f  clocks %s' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['clocks']))
except Exception as e: print('FAILED', e)
"
done; done 2>&1 | tee $OUT/k1s_ab.log
echo "== ncu: K1s on zipf10m"
FA_K1_OPT=256 timeout 600 ncu --set full --clock-control none --import-source on -k regex:aggregate_stream_kernel -s 6 -c 1 -o $OUT/prof_k1s_zipf10m -f \
    python bench.py --no-cpu --no-e2e --no-verify --batch $((1<<24)) --steps 4 --warmup 2 > $OUT/ncu_k1s.log 2>&1; tail -2 $OUT/ncu_k1s.log
echo "== bench reference arm"
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 800 $OUT/bench_reference.json; tail -3 $OUT/bench_reference.err
