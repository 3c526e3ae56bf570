#!/bin/bash
# round 2, call 29 (1 GPU, the last GPU minutes of the round): default = PIPE + ETAG + EARLY + K6 v2; A/B of a next-tile L2
# prefetch (pf) and of 8 teams x 128-record tiles (t128) on top of it, then GPU suite + bench lines with the winner.
set -u
OUT=gpurun_out/call29; mkdir -p $OUT
DEADLINE=${DEADLINE:-285}
left() { echo $((DEADLINE - SECONDS)); }
run() { # lib workload
  lib=$1; w=$2; shift 2
  if [ "$lib" = default ]; then unset FA_LIB_NAME; else export FA_LIB_NAME=libflowagg_$lib.so; fi
  timeout 100 python bench.py --workload $w --no-cpu --no-e2e --no-verify --batch $((1<<26)) --steps 8 --warmup 4 "$@" 2>>$OUT/ab.err | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $w %.0f Mpkts/s frac %.3f ms/step %.3f' % (d['value'], d['roofline']['frac'], d['ms_per_step']))
except Exception as e: print('$lib $w FAILED', e)
"
  unset FA_LIB_NAME
}
echo "== K1 A/B zipf10m (t=$SECONDS)"
for rep in 1 2; do for l in default pf t128 t128pf; do run $l zipf10m; done; done 2>&1 | tee $OUT/ab_k1.log
W=$(python - <<'PY'
import collections
v=collections.defaultdict(list)
for ln in open("gpurun_out/call29/ab_k1.log"):
    p=ln.split()
    if len(p)>3 and p[1]=="zipf10m" and p[2]!="FAILED": v[p[0]].append(float(p[2]))
m={k:sum(x)/len(x) for k,x in v.items() if x}
best=max(m,key=m.get) if m else "default"
if best!="default" and m.get("default") and m[best] < 1.01*m["default"]: best="default"
print(best)
PY
)
echo "K1 winner: $W" | tee $OUT/winner_k1.txt
if [ "$W" != default ]; then
  echo "== K1 A/B other workloads (t=$SECONDS)"
  for w in uniform10m zipf1m; do for l in default $W; do run $l $w; done; done 2>&1 | tee -a $OUT/ab_k1.log
  export FA_LIB_NAME=libflowagg_$W.so
fi
echo "== GPU suite, library ${FA_LIB_NAME:-libflowagg.so} (t=$SECONDS)"
timeout 300 python -m pytest tests -q -m gpu -x > $OUT/gpu_suite.log 2>&1; tail -3 $OUT/gpu_suite.log
echo "== default bench line (t=$SECONDS)"
timeout 240 python bench.py > $OUT/bench_zipf10m.json 2> $OUT/bench_zipf10m.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/call29/bench_zipf10m.json").read().strip().splitlines()[-1]); e=d.get("e2e") or {}
    print("zipf10m value %.0f frac %.4f parity %s/%s e2e %s events %s" % (d["value"], d["roofline"]["frac"], d.get("parity_ok"), d.get("parity_checked"), e.get("value"), (e.get("events_row") or {}).get("value")))
except Exception as ex: print("ERR", ex)
PY
if [ $(left) -gt 70 ]; then
  echo "== sketch100m line: fused sketches with the re-arm at S2 (t=$SECONDS)"
  timeout 150 python bench.py --workload sketch100m --no-cpu --no-e2e > $OUT/bench_sketch100m.json 2> $OUT/bench_sketch100m.err; python -c "
import json
try:
    d=json.loads(open('$OUT/bench_sketch100m.json').read().strip().splitlines()[-1]); print('sketch100m value %.0f frac %.4f parity %s/%s' % (d['value'], d['roofline']['frac'], d.get('parity_ok'), d.get('parity_checked')), {k:v for k,v in d.items() if 'sketch' in k})
except Exception as ex: print('sketch100m ERR', ex)
"
fi
if [ $(left) -gt 60 ]; then
  echo "== uniform10m / zipf1m lines with parity (t=$SECONDS)"
  for w in uniform10m zipf1m; do timeout 150 python bench.py --workload $w --no-cpu --no-e2e > $OUT/bench_$w.json 2> $OUT/bench_$w.err; python -c "
import json
try:
    d=json.loads(open('$OUT/bench_$w.json').read().strip().splitlines()[-1]); print('$w value %.0f frac %.4f parity %s/%s' % (d['value'], d['roofline']['frac'], d.get('parity_ok'), d.get('parity_checked')))
except Exception as ex: print('$w ERR', ex)
"; [ $(left) -gt 40 ] || break; done
fi
echo "== done (t=$SECONDS)"
