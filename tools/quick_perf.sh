#!/bin/bash
# quick perf check of K1 on the three workloads (device-resident inputs); prints Mpkts/s + roofline fraction
for w in ${WORKLOADS:-zipf1m zipf10m uniform10m}; do
  timeout 300 python bench.py --workload $w --no-cpu --no-e2e --steps ${STEPS:-30} "$@" 2>&1 | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%-10s %7.0f Mpkts/s  frac %.3f  ms/step %.3f  clocks %s' % ('$w', d['value'], d['roofline']['frac'], d['ms_per_step'], d['clocks']))
except Exception as e: print('$w FAILED', e)
"
done
