#!/bin/bash
# round 2, call 14 (2 GPUs): (f4) GPU tests; N=2 bench with one round per step (2^27) vs 2^26; uniform10m N=2 with the larger combiner table
set -u
OUT=gpurun_out/call14; mkdir -p $OUT
timeout 600 python -m pytest tests/test_snaps.py -x -q -m gpu > $OUT/gpu_snaps.log 2>&1; tail -3 $OUT/gpu_snaps.log
run() { # name port args...
  name=$1; port=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 2 --no-e2e --no-cpu "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
for ln in open("$OUT/bench_$name.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("$name: %.0f Mpkts/s, %.2f ms/step, timed %.0f ms, parity_ok=%s checked=%s, nvlink/round %.0f MB" % (d["value"], d["ms_per_step"], d["config"]["timed_region_ms"], d.get("parity_ok"), d.get("parity_checked"), d["config"]["nvlink"]["nvlink_bytes_per_round_rank0"]/1e6), d["config"]["nvlink"].get("phase_ms_per_round_rank0"))
PY
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/bench_$name.err | tail -3
}
run n2_r27 29801 --steps 12 --warmup 3
run n2_r26 29802 --steps 12 --warmup 3 --mgpu-round $((1<<26))
run n2_uniform 29803 --steps 8 --warmup 3 --workload uniform10m
run n2_zipf1m 29804 --steps 12 --warmup 3 --workload zipf1m
