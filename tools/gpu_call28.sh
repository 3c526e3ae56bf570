#!/bin/bash
# round 2, call 28 (1 GPU): call 27 again for the FA_K1_PIPE variants (their first build stored 8 bytes to a 4-byte-aligned
# shared-memory address: "misaligned address" on the device, invisible to the host emulation) with K6 v2 as the default.
set -u
OUT=gpurun_out/call28; mkdir -p $OUT
DEADLINE=${DEADLINE:-420}
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
for rep in 1 2; do for l in default pipe pe pee; do run $l zipf10m; done; done 2>&1 | tee $OUT/ab_k1.log
W=$(python - <<'PY'
import collections
v=collections.defaultdict(list)
for ln in open("gpurun_out/call28/ab_k1.log"):
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
  for w in uniform10m zipf1m; do for l in default $W default $W; do run $l $w; done; done 2>&1 | tee -a $OUT/ab_k1.log
  export FA_LIB_NAME=libflowagg_$W.so
fi
echo "== GPU suite, library ${FA_LIB_NAME:-libflowagg.so} (t=$SECONDS)"
timeout 300 python -m pytest tests -q -m gpu -x > $OUT/gpu_suite.log 2>&1; tail -3 $OUT/gpu_suite.log
echo "== default bench line (t=$SECONDS)"
timeout 240 python bench.py --no-cpu > $OUT/bench_zipf10m.json 2> $OUT/bench_zipf10m.err; python - <<'PY'
import json
try:
    d=json.loads(open("gpurun_out/call28/bench_zipf10m.json").read().strip().splitlines()[-1]); e=d.get("e2e") or {}
    print("zipf10m value %.0f frac %.4f parity %s/%s e2e %s events %s" % (d["value"], d["roofline"]["frac"], d.get("parity_ok"), d.get("parity_checked"), e.get("value"), (e.get("events_row") or {}).get("value")))
except Exception as ex: print("ERR", ex)
PY
if [ $(left) -gt 90 ]; then
  echo "== ncu full: K1 late launch, zipf10m (t=$SECONDS)"
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 30 -c 1 -o $OUT/prof_k1_zipf10m -f python bench.py --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 4 --warmup 6 > $OUT/ncu_k1.log 2>&1; tail -1 $OUT/ncu_k1.log
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
