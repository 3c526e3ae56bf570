#!/bin/bash
# round 2, call 16 (8 GPUs): the scaling line at N=8 (and N=4 on the same box), default arguments except fewer steps
set -u
OUT=gpurun_out/call16; mkdir -p $OUT
run() { # name n port args...
  name=$1; n=$2; port=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --no-cpu "$@" > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
for ln in open("$OUT/bench_$name.json"):
    if ln.startswith("{"):
        d=json.loads(ln); e=d.get("e2e") or {}
        print("$name: %.0f Mpkts/s, %.2f ms/step, timed %.0f ms, parity_ok=%s checked=%s, e2e %s, nvlink/round %.0f MB" % (d["value"], d["ms_per_step"], d["config"]["timed_region_ms"], d.get("parity_ok"), d.get("parity_checked"), e.get("value"), d["config"]["nvlink"]["nvlink_bytes_per_round_rank0"]/1e6), d["config"]["nvlink"].get("phase_ms_per_round_rank0"))
PY
  grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/bench_$name.err | tail -3
}
run n8 8 29901 --steps 12 --warmup 3
run n4 4 29902 --steps 12 --warmup 3 --no-e2e
