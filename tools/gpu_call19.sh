#!/bin/bash
# round 2, call 19 (1 GPU): bench lines added after the evidence run: sketch100m (configs[2]), zipf1m with its own launch size,
# K6 with the resident-size grid
set -u
OUT=gpurun_out/call19; mkdir -p $OUT
echo "== bench sketch100m"; timeout 900 python bench.py --workload sketch100m --no-cpu > $OUT/bench_sketch100m.json 2> $OUT/bench_sketch100m.err; tail -c 2600 $OUT/bench_sketch100m.json | head -c 1700; echo; tail -3 $OUT/bench_sketch100m.err
echo "== bench zipf1m"; timeout 600 python bench.py --workload zipf1m --no-cpu > $OUT/bench_zipf1m.json 2> $OUT/bench_zipf1m.err; python -c "
import json;d=json.loads(open('$OUT/bench_zipf1m.json').read().strip().splitlines()[-1]);print(d['value'],d['roofline']['frac'],d.get('parity_ok'),d['e2e']['value'],d['config'].get('records_per_launch'))"
timeout 600 python tools/bench_aux.py features 2>&1 | grep "^{" | tee $OUT/aux_features.jsonl
