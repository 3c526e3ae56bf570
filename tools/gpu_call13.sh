#!/bin/bash
# round 2, call 13 (1 GPU): (f4) snapshots on the device; full GPU suite; aux benches (features, snaps); default bench line
set -u
OUT=gpurun_out/call13; mkdir -p $OUT
echo "== full GPU suite"
timeout 1500 python -m pytest tests -x -q -m gpu > $OUT/gpu_suite.log 2>&1; tail -4 $OUT/gpu_suite.log
timeout 900 python tools/bench_aux.py features snaps 2>&1 | grep "^{" | tee $OUT/aux.jsonl
echo "== bench default"
timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; python - <<PY
import json
d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
print("value %.0f frac %.3f ms/step %.3f timed %.0f ms parity %s e2e %.0f events %.0f cpu %.1f clocks %s" % (d["value"], d["roofline"]["frac"], d["ms_per_step"], d["config"]["timed_region_ms"], d.get("parity_ok"), d["e2e"]["value"], d["e2e"]["events_row"]["value"], d["cpu_baseline"]["value"], d["clocks"]))
PY
tail -3 $OUT/bench_default.err
