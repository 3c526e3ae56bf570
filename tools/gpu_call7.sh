#!/bin/bash
# round 2, call 7 (2 GPUs): full GPU suite incl. the 2-GPU tests and the C++ fa_sharded client, then the bench at N=2
set -u
OUT=gpurun_out/call7; mkdir -p $OUT
nvidia-smi -L
echo "== full GPU suite (2 GPUs visible)"
timeout 1200 python -m pytest tests -x -q -m gpu > $OUT/gpu_suite.log 2>&1; tail -6 $OUT/gpu_suite.log
echo "== C++ client, 2 GPUs, larger"
timeout 300 netobserv_ebpf_agent_b200/host/test_sharded 2 20000000 2000000 2>&1 | tail -3
echo "== bench N=2 (zipf10m, verify)"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 6 --warmup 3 > $OUT/bench_n2.json 2> $OUT/bench_n2.err; tail -c 1800 $OUT/bench_n2.json; tail -5 $OUT/bench_n2.err
echo "== bench N=1 default for the same box"
timeout 900 python bench.py --steps 10 --warmup 4 --no-cpu > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 600 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
