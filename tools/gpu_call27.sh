#!/bin/bash
# round 2, call 27 (1 GPU, the round's last GPU minutes): same-box A/B of the K1 build variants (FA_K1_PIPE / _ETAG / _EARLY)
# and of K6 v2, then the evidence for the winners: GPU suite with the winning library, default bench line, rttdns line, ncu.
# Every step has its own timeout and the optional ones are skipped when the call runs out of time (DEADLINE seconds).
set -u
OUT=gpurun_out/call27; mkdir -p $OUT
DEADLINE=${DEADLINE:-720}
left() { echo $((DEADLINE - SECONDS)); }
run() { # lib workload extra...
  lib=$1; w=$2; shift 2
  if [ "$lib" = default ]; then unset FA_LIB_NAME; else export FA_LIB_NAME=libflowagg_$lib.so; fi
  timeout 120 python bench.py --workload $w --no-cpu --no-e2e --no-verify --batch $((1<<26)) --steps 8 --warmup 4 "$@" 2>>$OUT/ab.err | python -c "
import sys,json
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$lib $w %.0f Mpkts/s frac %.3f ms/step %.3f' % (d['value'], d['roofline']['frac'], d['ms_per_step']))
except Exception as e: print('$lib $w FAILED', e)
"
  unset FA_LIB_NAME
}
echo "== K1 A/B zipf10m (t=$SECONDS)"
for rep in 1 2; do for l in default pipe etag pe pee; do run $l zipf10m; done; done 2>&1 | tee $OUT/ab_k1.log
W=$(python - <<'PY'
import collections
v=collections.defaultdict(list)
for ln in open("gpurun_out/call27/ab_k1.log"):
    p=ln.split()
    if len(p)>3 and p[1]=="zipf10m" and p[2]!="FAILED": v[p[0]].append(float(p[2]))
m={k:sum(x)/len(x) for k,x in v.items() if x}
best=max(m,key=m.get) if m else "default"
if best!="default" and m.get("default") and m[best] < 1.01*m["default"]: best="default"
print(best)
PY
)
echo "K1 winner: $W" | tee $OUT/winner_k1.txt
echo "== K1 A/B other workloads (t=$SECONDS)"
if [ "$W" != default ]; then for w in uniform10m zipf1m; do for l in default $W default $W; do run $l $w; done; done 2>&1 | tee -a $OUT/ab_k1.log; fi
if [ "$W" != default ]; then
  echo "== GPU suite with libflowagg_$W.so (t=$SECONDS)"
  FA_LIB_NAME=libflowagg_$W.so timeout 420 python -m pytest tests -q -m gpu -x > $OUT/gpu_suite_$W.log 2>&1; tail -3 $OUT/gpu_suite_$W.log
  echo "== default bench line with libflowagg_$W.so (t=$SECONDS)"
  FA_LIB_NAME=libflowagg_$W.so timeout 300 python bench.py --no-cpu > $OUT/bench_zipf10m_$W.json 2> $OUT/bench_zipf10m_$W.err; tail -c 600 $OUT/bench_zipf10m_$W.json | head -c 600; echo
fi
echo "== K6 A/B (t=$SECONDS)"
for l in default k6v2 k6v2b5 k6v2b6; do
  if [ "$l" = default ]; then unset FA_LIB_NAME; else export FA_LIB_NAME=libflowagg_$l.so; fi
  echo -n "$l "; timeout 120 python tools/bench_aux.py features 2>>$OUT/ab.err | tail -1
  unset FA_LIB_NAME
done 2>&1 | tee $OUT/ab_k6.log
W6=$(python - <<'PY'
import json
best,bv,base="default",0.0,0.0
for ln in open("gpurun_out/call27/ab_k6.log"):
    try:
        name,js=ln.split(" ",1); d=json.loads(js); v=d["dns_Msamples_s"]+d["additional_Msamples_s"]
    except Exception: continue
    if name=="default": base=v
    if v>bv: best,bv=name,v
if best!="default" and base and bv<1.02*base: best="default"
print(best)
PY
)
echo "K6 winner: $W6" | tee $OUT/winner_k6.txt
if [ "$W6" != default ] && [ $(left) -gt 150 ]; then
  echo "== K6 GPU tests with libflowagg_$W6.so (t=$SECONDS)"
  FA_LIB_NAME=libflowagg_$W6.so timeout 240 python -m pytest tests -q -m gpu -x -k "features or pktdrop or kernel_map or host_cpp" > $OUT/gpu_suite_$W6.log 2>&1; tail -3 $OUT/gpu_suite_$W6.log
  if [ $(left) -gt 120 ]; then
    echo "== rttdns line with libflowagg_$W6.so (t=$SECONDS)"
    FA_LIB_NAME=libflowagg_$W6.so timeout 200 python bench.py --workload rttdns --steps 10 --warmup 3 --no-cpu > $OUT/bench_rttdns_$W6.json 2> $OUT/bench_rttdns_$W6.err; tail -c 400 $OUT/bench_rttdns_$W6.json; echo
  fi
fi
if [ "$W" != default ] && [ $(left) -gt 100 ]; then
  echo "== ncu full: K1 late launch, zipf10m, libflowagg_$W.so (t=$SECONDS)"
  FA_LIB_NAME=libflowagg_$W.so timeout 240 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 30 -c 1 -o $OUT/prof_k1_zipf10m_$W -f python bench.py --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 4 --warmup 6 > $OUT/ncu_k1.log 2>&1; tail -1 $OUT/ncu_k1.log
fi
if [ "$W6" != default ] && [ $(left) -gt 80 ]; then
  echo "== ncu full: K6 DNS fold, libflowagg_$W6.so (t=$SECONDS)"
  FA_LIB_NAME=libflowagg_$W6.so timeout 200 ncu --set full --clock-control none --import-source on -k regex:feature_fold_kernel -s 4 -c 1 -o $OUT/prof_k6_$W6 -f python tools/bench_aux.py features > $OUT/ncu_k6.log 2>&1; tail -1 $OUT/ncu_k6.log
fi
echo "== done (t=$SECONDS)"
