#!/bin/bash
# round 2, evidence refresh after the last K1 changes (L2 evict-first stream, cache admission at 2 records): GPU suite, the
# bench lines of every workload, ncu captures of K1 (Zipf / uniform) + launch list.  Lands in gpurun_out/final2/.
set -u
OUT=gpurun_out/final2; mkdir -p $OUT
echo "== GPU suite"; timeout 1200 python -m pytest tests -q -m gpu > $OUT/gpu_suite.log 2>&1; tail -2 $OUT/gpu_suite.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_zipf10m.json 2> $OUT/bench_zipf10m.err; tail -2 $OUT/bench_zipf10m.err
for w in uniform10m zipf1m sketch100m; do echo "== bench $w"; timeout 600 python bench.py --workload $w --no-cpu > $OUT/bench_$w.json 2> $OUT/bench_$w.err; done
echo "== bench rttdns"; timeout 600 python bench.py --workload rttdns --steps 10 --warmup 3 --no-cpu > $OUT/bench_rttdns.json 2> $OUT/bench_rttdns.err
python - <<'PY'
import json
for w in ("zipf10m","uniform10m","zipf1m","sketch100m","rttdns"):
    try:
        d=json.loads(open(f"gpurun_out/final2/bench_{w}.json").read().strip().splitlines()[-1]); e=d.get("e2e") or {}
        print(w, "value %.0f frac %.4f parity %s/%s e2e %s events %s" % (d["value"], d["roofline"]["frac"], d.get("parity_ok"), d.get("parity_checked"), e.get("value"), (e.get("events_row") or {}).get("value")))
    except Exception as ex: print(w, "ERR", ex)
PY
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 2 --batch $((1<<24)) --no-cpu --no-e2e --no-verify > $OUT/launches.log 2>&1
echo "== ncu full: K1 late launch (zipf10m), then uniform10m"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 30 -c 1 -o $OUT/prof_k1_zipf10m -f python bench.py --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 4 --warmup 6 > $OUT/ncu_k1_zipf10m.log 2>&1; tail -1 $OUT/ncu_k1_zipf10m.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 30 -c 1 -o $OUT/prof_k1_uniform10m -f python bench.py --workload uniform10m --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 4 --warmup 6 > $OUT/ncu_k1_uniform10m.log 2>&1; tail -1 $OUT/ncu_k1_uniform10m.log
