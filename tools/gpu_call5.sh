#!/bin/bash
# round 2, call 5: packet-drop / RTT-min GPU tests; K1 tile-size and next-tile-prefetch A/B; K1s steady-state profile
set -u
OUT=gpurun_out/call5; mkdir -p $OUT
echo "== new GPU tests"
timeout 600 python -m pytest tests/test_pktdrop.py tests/test_pbflow.py tests/test_gpu_features.py tests/test_events.py -x -q -m gpu 2>&1 | tail -3
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
  run libflowagg_t128.so 0 $w
  run libflowagg_t128.so 32 $w
done 2>&1 | tee $OUT/ab.log
unset FA_LIB_NAME
echo "== t128 parity"
FA_LIB_NAME=libflowagg_t128.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -x -q -m gpu 2>&1 | tail -2
echo "== ncu: K1s on uniform10m, a late launch"
FA_K1_OPT=256 timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_stream_kernel -s 60 -c 1 -o $OUT/prof_k1s_uniform10m -f \
    python bench.py --workload uniform10m --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 4 --warmup 6 > $OUT/ncu_k1s.log 2>&1; tail -2 $OUT/ncu_k1s.log
