#!/bin/bash
# Same-box A/B of the K1 experiment variants (aggregate.cu, template parameter kVar; FA_K1_OPT bits 5-7):
#   0 default | 32 next-tile L2 prefetch | 64 branch-free compares | 96 both | 128 4-lane probe groups | 160 4-lane + prefetch
#   256 K1w = aggregate_warp_kernel (warp-independent: no team barriers, no tile-local election, 256-entry cache); 288 = K1w + next-sub-tile L2 prefetch; 320 = K1w + table-line L2 prefetch at hash time; 768 = K1w + warp-aggregated cache folds (match.any + redux); 1280 = K1w + L2 evict-first policy on the record stream; 2304 = K1w with 8-lane probes (1 L1 wavefront per line, twice the probe instructions); 4352 = K1w with 16 warps per CTA, double-buffered sub-tiles, up to 4 probe rounds in flight
# First the parity tests with every variant (a variant that is not bit-exact is not a candidate), then the three
# workloads.  One gpurun call, so that all numbers come from the same GPU (boxes differ by up to 7 %).
#   gpurun --timeout 1500 -- 'bash tools/k1_variants_ab.sh > gpurun_out/k1_variants.log 2>&1'
VARIANTS="${VARIANTS:-0 32 64 128 160 256 288 320 768 1280 2304 4352}"
for v in $VARIANTS; do
  echo "== parity FA_K1_OPT=$v"
  FA_K1_OPT=$v timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -x -q -m gpu 2>&1 | tail -2
done
for rep in $(seq 1 ${REPS:-2}); do
  for v in $VARIANTS; do
    echo "== perf FA_K1_OPT=$v (rep $rep)"
    FA_K1_OPT=$v STEPS=${STEPS:-30} bash tools/quick_perf.sh
  done
done
