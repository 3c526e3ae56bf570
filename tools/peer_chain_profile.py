#!/usr/bin/env python
"""Runs the multi-GPU round (combine -> drain -> fused route/exchange -> owner fold) with world_size 1 so that the
per-kernel times of the chain can be listed with `ncu --metrics gpu__time_duration.sum` on a single GPU."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import netobserv_ebpf_agent_b200 as fa  # noqa: E402
from netobserv_ebpf_agent_b200.sharded import PeerShardedAggregator  # noqa: E402

os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29533")
os.environ.setdefault("RANK", "0"); os.environ.setdefault("WORLD_SIZE", "1")
dev = torch.device("cuda", 0)
torch.cuda.set_device(0)
dist.init_process_group("nccl", device_id=dev)
stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream)
B = 1 << 24
eng = fa.FlowAggEngine(1 << 26, device=0, max_batch=B, cuda_stream=stream.cuda_stream, flags=fa.FA_F_NO_FULL_CUT)
agg = PeerShardedAggregator(eng, B, dev)
gp = fa.GenParams(seed=2, n_keys=1_000_000, dist=1, zipf_s_milli=1100, t0_ns=1_000_000, varying_desc=0)
batches = []
for i in range(2):
    t = torch.empty(B * 144, dtype=torch.uint8, device=dev); eng.gen_records(gp, i * B, B, t); batches.append(t)
eng.sync()
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(2):
    agg.ingest(batches[i % 2], B)
agg.flush(); a.record()
for i in range(steps):
    agg.ingest(batches[i % 2], B)
agg.flush(); b.record(); torch.cuda.synchronize()
print("ms per round (world=1):", a.elapsed_time(b) / steps, "flows", eng.live_flows())
agg.close(); eng.close(); dist.destroy_process_group()
