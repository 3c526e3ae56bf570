#!/bin/bash
# round 2, evidence run: GPU suite, the bench lines of every workload, the reference arm, ncu launch list + full captures.
# Everything lands in gpurun_out/final/ and is summarised into profiles/ afterwards (tools/ncu_summary.py).
set -u
OUT=gpurun_out/final; mkdir -p $OUT
echo "== GPU suite"; timeout 1200 python -m pytest tests -q -m gpu > $OUT/gpu_suite.log 2>&1; tail -3 $OUT/gpu_suite.log
echo "== bench default"; timeout 900 python bench.py > $OUT/bench_zipf10m.json 2> $OUT/bench_zipf10m.err; tail -c 300 $OUT/bench_zipf10m.json
for w in uniform10m zipf1m; do echo "== bench $w"; timeout 600 python bench.py --workload $w --no-cpu > $OUT/bench_$w.json 2> $OUT/bench_$w.err; python -c "
import json;d=json.load(open('$OUT/bench_$w.json'));print(d['value'],d['roofline']['frac'],d.get('parity_ok'),d['e2e']['value'])"; done
echo "== bench rttdns"; timeout 600 python bench.py --workload rttdns --steps 10 --warmup 3 > $OUT/bench_rttdns.json 2> $OUT/bench_rttdns.err; tail -c 400 $OUT/bench_rttdns.json; tail -3 $OUT/bench_rttdns.err
echo "== reference arm"; timeout 600 python bench.py --impl reference > $OUT/bench_reference.json 2> $OUT/bench_reference.err; tail -c 300 $OUT/bench_reference.json
echo "== aux: sketch / features / kmap / small cache / protobuf / snapshots"
timeout 900 python tools/bench_aux.py sketch features kmap smallcache pb snaps > $OUT/bench_aux.jsonl 2> $OUT/bench_aux.err; cat $OUT/bench_aux.jsonl; tail -3 $OUT/bench_aux.err
echo "== ncu launch list of the default bench (short)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 3 --warmup 2 --batch $((1<<24)) --no-cpu --no-e2e --no-verify > $OUT/launches.log 2>&1; tail -1 $OUT/launches.log | cut -c1-200
echo "== ncu full: K1 late launch (zipf10m), then uniform10m"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 30 -c 1 -o $OUT/prof_k1_zipf10m -f python bench.py --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 4 --warmup 6 > $OUT/ncu_k1_zipf10m.log 2>&1; tail -1 $OUT/ncu_k1_zipf10m.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 30 -c 1 -o $OUT/prof_k1_uniform10m -f python bench.py --workload uniform10m --no-cpu --no-e2e --no-verify --batch $((1<<25)) --steps 4 --warmup 6 > $OUT/ncu_k1_uniform10m.log 2>&1; tail -1 $OUT/ncu_k1_uniform10m.log
echo "== ncu full: K2 evict, K8 protobuf (K6 and the f4 kernels: tools/gpu_call15.sh)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'evict_kernel|evict_features_kernel|pb_write_kernel|pb_size_kernel' -c 6 -o $OUT/prof_others -f python tools/bench_aux.py features pb > $OUT/ncu_others.log 2>&1; tail -2 $OUT/ncu_others.log
