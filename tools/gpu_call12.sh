#!/bin/bash
# round 2, call 12 (2 GPUs): tile-sorted route_peer (coalesced peer stores); 2-GPU parity tests; N=2 bench at 2^24 / 2^26 rounds
set -u
OUT=gpurun_out/call12; mkdir -p $OUT
echo "== 2-GPU tests"
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_host_cpp.py -x -q -m gpu > $OUT/gpu_sharded.log 2>&1; tail -3 $OUT/gpu_sharded.log
for r in 24 26; do
  echo "== bench N=2 round=2^$r"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 296$r bench.py --gpus 2 --steps 6 --warmup 3 --mgpu-round $((1<<r)) --no-e2e --no-cpu > $OUT/bench_n2_r$r.json 2> $OUT/bench_n2_r$r.err
  python - <<PY
import json
for ln in open("$OUT/bench_n2_r$r.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("round 2^$r: %.0f Mpkts/s, %.2f ms/step, parity_ok=%s checked=%s, nvlink/round %.0f MB" % (d["value"], d["ms_per_step"], d.get("parity_ok"), d.get("parity_checked"), d["config"]["nvlink"]["nvlink_bytes_per_round_rank0"]/1e6), d["config"]["nvlink"].get("phase_ms_per_round_rank0"))
PY
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/bench_n2_r$r.err | tail -3
done
echo "== uniform10m N=2 (config 4: the combiner does not reduce)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29701 bench.py --gpus 2 --workload uniform10m --steps 6 --warmup 3 --no-e2e --no-cpu > $OUT/bench_n2_uniform.json 2> $OUT/bench_n2_uniform.err
tail -c 1800 $OUT/bench_n2_uniform.json; grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/bench_n2_uniform.err | tail -3
