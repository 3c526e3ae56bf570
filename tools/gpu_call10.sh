#!/bin/bash
# round 2, call 10 (2 GPUs): N=2 bench (parity in the line) at round sizes 2^24 / 2^25 / 2^26 with per-phase timings;
# H2D bandwidth probe; K6 feature folds (tile-staged) parity + throughput
set -u
OUT=gpurun_out/call10; mkdir -p $OUT
python tools/h2d_bench.py 2>&1 | tee $OUT/h2d.log
echo "== feature / pkt-drop / kernel-map GPU tests"
timeout 900 python -m pytest tests -x -q -m gpu -k "feature or pktdrop or kernel_map or events" > $OUT/gpu_feat.log 2>&1; tail -3 $OUT/gpu_feat.log
timeout 600 python tools/bench_aux.py features 2>&1 | tail -6 | tee $OUT/aux_features.log
for r in 24 25 26; do
  echo "== bench N=2 round=2^$r"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 295$r bench.py --gpus 2 --steps 6 --warmup 3 --mgpu-round $((1<<r)) --no-e2e --no-cpu > $OUT/bench_n2_r$r.json 2> $OUT/bench_n2_r$r.err
  python - <<PY
import json
for ln in open("$OUT/bench_n2_r$r.json"):
    if ln.startswith("{"):
        d=json.loads(ln); print("round 2^$r: %.0f Mpkts/s, %.2f ms/step, parity_ok=%s checked=%s, nvlink/round %.0f MB" % (d["value"], d["ms_per_step"], d.get("parity_ok"), d.get("parity_checked"), d["config"]["nvlink"]["nvlink_bytes_per_round_rank0"]/1e6), d["config"]["nvlink"].get("phase_ms_per_round_rank0"))
PY
  tail -2 $OUT/bench_n2_r$r.err
done
