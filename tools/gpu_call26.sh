#!/bin/bash
# round 2, call 26 (2 GPUs): the N=2 line with the shipped kernel and the NVLink-fraction fields; 2-GPU tests
set -u
OUT=gpurun_out/call26; mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_sharded.py tests/test_gpu_host_cpp.py -x -q -m gpu > $OUT/gpu_sharded.log 2>&1; tail -2 $OUT/gpu_sharded.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29911 bench.py --gpus 2 --no-cpu > $OUT/bench_n2.json 2> $OUT/bench_n2.err
python - <<'PY'
import json
for ln in open("gpurun_out/call26/bench_n2.json"):
    if ln.startswith("{"):
        d=json.loads(ln); e=d.get("e2e") or {}
        print("n2: %.0f Mpkts/s, %.2f ms/step, timed %.0f ms, parity_ok=%s checked=%s, e2e %s" % (d["value"], d["ms_per_step"], d["config"]["timed_region_ms"], d.get("parity_ok"), d.get("parity_checked"), e.get("value")), d["config"]["nvlink"])
PY
grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/bench_n2.err | tail -3
