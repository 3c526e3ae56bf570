#!/bin/bash
# The gpu-marked tests against the engine emulation (tests/emul/engine_emul.cpp), for a kernel variant, on the CPU.
# Slow (tens of minutes: a kernel launch is ~2000 OS threads) and manual; what it buys is the FULL parity suite for a
# variant before any GPU time is spent on it.  Tests that need torch.cuda or a second GPU are deselected.
#   FA_K1_OPT=256 bash tools/emulated_gpu_suite.sh            # K1w
#   FA_K1_OPT=768 bash tools/emulated_gpu_suite.sh parity     # only test_gpu_parity.py
set -u
WHAT="${1:-all}"
SKIP='not device_pointer and not device_generator and not route_by_hash and not drain_active and not evict_capacity'
export FA_EMULATED_GPU=1
python -m pytest tests/test_gpu_parity.py tests/test_golden_fixtures.py -q -m gpu -k "$SKIP" --durations=5
[ "$WHAT" = "parity" ] && exit 0
python -m pytest tests/test_gpu_features.py tests/test_gpu_sketch.py -q -m gpu --durations=5
