#!/bin/bash
# round 2, call 17 (1 GPU): K6 with the hot-flow cache across tiles + two-round-trip probe: parity, throughput, launch list, rttdns line
set -u
OUT=gpurun_out/call17; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu -k "feature or pktdrop or kernel_map" > $OUT/gpu_feat.log 2>&1; tail -3 $OUT/gpu_feat.log
timeout 600 python tools/bench_aux.py features 2>&1 | grep "^{" | tee $OUT/aux_features.jsonl
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_features.csv python tools/bench_aux.py features > $OUT/launches_features.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/call17/launches_features.csv')))
hdr=None; agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr and len(r)==len(hdr):
        d=dict(zip(hdr,r))
        try: v=float(d['Metric Value'].replace(',',''))
        except: continue
        agg.setdefault(d['Kernel Name'][:60],[]).append(v)
for k,v in agg.items(): print(f"{k:60s} n={len(v):3d} mean={sum(v)/len(v)/1e3:9.1f} us")
PY
echo "== bench rttdns"; timeout 600 python bench.py --workload rttdns --steps 10 --warmup 3 --no-cpu > $OUT/bench_rttdns.json 2> $OUT/bench_rttdns.err; tail -c 1200 $OUT/bench_rttdns.json; tail -3 $OUT/bench_rttdns.err
