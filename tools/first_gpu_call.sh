#!/bin/bash
# Everything that was written after the last GPU run of round 1 and is waiting for a B200, in ONE gpurun call
# (~12-15 min of box time):
#   gpurun --timeout 1500 -- 'bash tools/first_gpu_call.sh'
# Logs land in gpurun_out/first_call/.  Order: cheapest + most important first, so a cut-off still leaves data.
set -u
OUT=gpurun_out/first_call; mkdir -p $OUT
echo "== 1. default GPU suite"; timeout 900 python -m pytest tests -x -q -m gpu > $OUT/gpu_suite.log 2>&1; tail -2 $OUT/gpu_suite.log
echo "== 2. KERNEL_MAP mode on the device (tests/test_gpu_kernel_map.py)"
FA_EXPERIMENTAL_KERNEL_MAP=1 timeout 600 python -m pytest tests/test_gpu_kernel_map.py -x -q -m gpu > $OUT/kmap.log 2>&1; tail -3 $OUT/kmap.log
echo "== 3. K1 variants: parity, then same-box perf (0 = default, 128 = 4-lane probes, 256 = K1w, 288 = K1w + L2 prefetch)"
VARIANTS="0 256 4352 2304 768 128" REPS=1 WORKLOADS="zipf1m zipf10m uniform10m" STEPS=20 bash tools/k1_variants_ab.sh > $OUT/k1_variants.log 2>&1; grep -E "==|Mpkts|passed|failed" $OUT/k1_variants.log | head -80
echo "== 4. sketch / multi-engine paths with K1w"
FA_K1_OPT=256 timeout 600 python -m pytest tests/test_gpu_sketch.py tests/test_gpu_features.py tests/test_gpu_host_cpp.py -x -q -m gpu > $OUT/k1w_other.log 2>&1; tail -2 $OUT/k1w_other.log
echo "== 5. KERNEL_MAP throughput + e2e at N=1 (new bench code path)"
FA_EXPERIMENTAL_KERNEL_MAP=1 timeout 300 python tools/bench_aux.py kmap smallcache > $OUT/kmap_bench.jsonl 2>&1; tail -2 $OUT/kmap_bench.jsonl
FA_EXPERIMENTAL_KERNEL_MAP=1 FA_KMAP_IMPL=1 timeout 300 python tools/bench_aux.py kmap > $OUT/kmap_v1_bench.jsonl 2>&1; tail -1 $OUT/kmap_v1_bench.jsonl
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_default.json 2> $OUT/bench_default.err; tail -c 600 $OUT/bench_default.json
echo "== 6. ncu: launch list + full capture of K1w"
FA_K1_OPT=256 timeout 600 ncu --set full --clock-control none --import-source on -k regex:aggregate_warp_kernel -s 3 -c 1 -o $OUT/prof_k1w -f \
    python bench.py --steps 4 --warmup 2 --no-cpu --no-e2e > $OUT/ncu_k1w.log 2>&1; tail -2 $OUT/ncu_k1w.log
echo "== 7. ncu: fused-sketch K1 and the feature folds (SURVEY 8d asks for K4/K5/K6 captures too)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:aggregate_kernel -s 3 -c 1 -o $OUT/prof_k1_sketch -f \
    python tools/bench_aux.py sketch > $OUT/ncu_sketch.log 2>&1; tail -1 $OUT/ncu_sketch.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'dns_fold_kernel|additional_fold_kernel' -c 2 -o $OUT/prof_k6 -f \
    python tools/bench_aux.py features > $OUT/ncu_k6.log 2>&1; tail -1 $OUT/ncu_k6.log
