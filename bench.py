#!/usr/bin/env python
"""bench.py — flow-aggregation throughput on B200 (contract: see the task prompt / DESIGN.md §measurement).

A "step" is one pass of the hot path (K1 flow_aggregate, reached through fa_ingest of the C ABI) over one
batch of synthetic 144-byte flow records.  Workloads (BASELINE.json):
  zipf1m    configs[1]: 1 M Zipf-1.1 5-tuples            (default; the config the metric is quoted on)
  zipf10m   north_star headline: 10 M Zipf-1.1 5-tuples
  uniform10m worst case for table locality: 10 M uniform 5-tuples
One JSON line on stdout (rank 0).  --impl reference times the CPU restatement of pkg/flow.Accounter
(the Go reference cannot be built in this image) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

REC = 144
# DRAM bytes per record of K1 from the committed ncu --set full capture (profiles/r1_k1_aggregate_ncu_summary.txt:
# dram__bytes_read.sum 2.863857 GB + dram__bytes_write.sum 70.906 MB for one 16,777,216-record launch of the zipf1m
# workload) -> 174.9 B/record against 144 algorithmic bytes (ratio 1.21: table lines + write-backs, no re-reads).
NCU_DRAM_BYTES_PER_RECORD = {"zipf1m": (2.863857e9 + 70.906112e6) / 16777216}
WORKLOADS = {
    "zipf1m": dict(n_keys=1_000_000, dist=1, seed=2, label="1e9-record stream, 1M Zipf-1.1 5-tuples (BASELINE configs[1])"),
    "zipf10m": dict(n_keys=10_000_000, dist=1, seed=2, label="1e9-record stream, 10M Zipf-1.1 5-tuples (north_star headline)"),
    "uniform10m": dict(n_keys=10_000_000, dist=0, seed=2, label="1e9-record stream, 10M uniform 5-tuples (worst-case locality)"),
}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.p = index, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- CPU legs (oracle = checker / baseline)
def host_sample(wl, n, first=0):
    import netobserv_ebpf_agent_b200 as fa
    p = fa.GenParams(seed=wl["seed"], n_keys=wl["n_keys"], dist=wl["dist"], zipf_s_milli=1100, t0_ns=1_000_000, varying_desc=0)
    return fa.gen_records_host(p, first, n)


def cpu_baseline_port(wl, sample_records, passes=4):
    """Single-thread CPU restatement of pkg/flow.Accounter (one goroutine in the reference)."""
    import oracle_lib as O
    sample = host_sample(wl, sample_records)
    acc = O.Accounter(1 << 26)
    t0 = time.perf_counter()
    for _ in range(passes):                          # ~10 s of CPU work: the same 2^25-record slice folded 4 times
        acc.account(sample)
    flows = len(acc)
    acc.evict()
    dt = time.perf_counter() - t0
    acc.close()
    return {"value": passes * sample_records / dt / 1e6, "unit": "Mpkts/s", "cores": 1, "kind": "port",
            "sample": f"{passes} x {sample_records} records of the same stream, {flows} flows, incl. the final evict "
                      f"(CPU restatement of pkg/flow.Accounter; Go toolchain unavailable)", "seconds": dt}


def run_reference(args, wl):
    """--impl reference: the reference's CPU path for this step (Accounter restatement), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:                                    # under torchrun only rank 0 runs the CPU arm
        return
    import oracle_lib as O
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    n = args.ref_sample
    sample = host_sample(wl, n)
    buf = np.ascontiguousarray(sample).view(np.uint8).reshape(-1)
    out = np.zeros((min(n, wl["n_keys"]) + 1) * REC, dtype=np.uint8)

    def run(threads):
        return O.lib().oracle_accounter_sharded_run(O._p(buf), n, threads, O._p(out), len(out) // REC)
    # "all the host threads it can use": containers often expose more CPUs than they may run on, so pick the
    # thread count that is actually fastest on this box (tried once each, outside the timed region)
    best, cores = None, avail
    run(avail)                                       # touch the sample / warm the allocator first
    for tcount in sorted({avail, max(1, avail // 2), max(1, avail // 4), max(1, avail // 8), min(avail, 16), min(avail, 8)}, reverse=True):
        t0 = time.perf_counter(); run(tcount); dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, cores = dt, tcount

    def step():
        return run(cores)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        flows = step()
    dt = time.perf_counter() - t0
    v = n * args.steps / dt / 1e6
    line = {"impl": "reference", "metric": "Mpkts/s aggregated", "value": v, "unit": "Mpkts/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u64/u32 integer", "data": "synthetic",
            "config": {"workload": wl["label"], "records_per_step": n, "flows": int(flows)},
            "cpu_baseline": {"value": v, "unit": "Mpkts/s", "cores": cores, "kind": "port",
                             "sample": f"{n} records/step; key-sharded over {cores} threads, each a private Accounter "
                                       "(CPU restatement of pkg/flow.Accounter; Go toolchain unavailable)"},
            "e2e": {"value": v, "unit": "Mpkts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm
def run_ours(args, wl):
    import torch
    import torch.distributed as dist
    import netobserv_ebpf_agent_b200 as fa

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    B = args.batch                                   # records per step per GPU
    ring = max(1, min(args.ring, args.steps + args.warmup))
    # Everything (engine kernels, NCCL, timing events) runs on ONE explicit torch stream: the legacy default
    # stream has handle 0, which the engine would take as "create your own".
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    if world > 1:
        args.max_batch = B          # one route -> exchange -> fold round per step: amortises the host-side syncs
    eng = fa.FlowAggEngine(args.max_entries, device=local, max_batch=args.max_batch, cuda_stream=stream.cuda_stream,
                           flags=fa.FA_F_NO_FULL_CUT if (world > 1 and args.exchange == "peer") else 0)
    # one key universe for the whole job; every rank generates its own slice of the record stream
    gp = fa.GenParams(seed=wl["seed"], n_keys=wl["n_keys"], dist=wl["dist"], zipf_s_milli=1100,
                      t0_ns=1_000_000, varying_desc=0)
    batches = []
    for i in range(ring):
        t = torch.empty(B * REC, dtype=torch.uint8, device=dev)
        eng.gen_records(gp, (i * world + rank) * B, B, t)
        batches.append(t)
    eng.sync()

    if world > 1:
        from netobserv_ebpf_agent_b200.sharded import PeerShardedAggregator, ShardedAggregator
        if args.exchange == "peer":
            # local combine (K1+K2) -> K3 fused with the exchange (peer stores over NVLink) -> K1 on the owner
            agg = PeerShardedAggregator(eng, args.max_batch, dev)
        else:
            # local combine (K1+K2) -> K3 route -> NCCL all-to-all -> K1 on the owner
            agg = ShardedAggregator(eng, args.max_batch, dev, combine=not args.no_combine)

    def step(i):
        src = batches[i % ring]
        if world == 1:
            rc, took = eng.ingest(src.data_ptr(), B)
            assert rc == 0 and took == B, (rc, took)
        else:
            agg.ingest(src, B)

    def finish():                                    # N>1: drain + exchange the batch still in a scratch table
        if world > 1:
            agg.flush()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for i in range(args.warmup):
        step(i)
    finish()
    barrier()

    def all_stats():                                 # owner engine + the combiner's scratch engine(s)
        s = eng.stats()
        if world > 1:
            for loc in ([agg.local] if hasattr(agg, "local") else agg.locals):
                ls = loc.stats()
                for k in ("kernel_launches", "order_fixups", "spills"):
                    s[k] += ls[k]
        return s
    st0 = all_stats()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        step(args.warmup + i)
        if i == args.steps - 1:
            finish()
        ev[i + 1].record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[-1])
    step_ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps)]
    st1 = all_stats()
    if world > 1:
        t = torch.tensor([total_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = world * B * args.steps / (total_ms / 1e3) / 1e6
    launches = st1["kernel_launches"] - st0["kernel_launches"]
    flows = eng.live_flows()

    # ---------------------------------------------------------------- end-to-end through the C ABI, host buffers
    # Every rank feeds its own slice from pinned host memory over its own PCIe link (N > 1: through the sharded
    # aggregator, whose local engine does the H2D copy); the timed region has the copies, a per-step read-back and
    # the final lookup-and-delete to host memory.  Preparation failures are agreed on by all ranks before the
    # loop starts, so a rank can never be left alone inside a collective.
    e2e = None
    host_ok = world == 1 or args.exchange == "peer"       # the NCCL variant's scratch streams are device-input only
    if not args.no_e2e and host_ok:
        Be = min(args.e2e_batch, args.max_batch)
        prep_err = None
        hring, out_host = [], None
        try:
            if world == 1:
                eng.evict_into(batches[0].data_ptr(), B)      # reset the cache (flows <= B)
            for i in range(min(4, ring + 1)):
                h = torch.empty(Be * REC, dtype=torch.uint8).pin_memory()
                d = torch.empty(Be * REC, dtype=torch.uint8, device=dev)
                eng.gen_records(gp, ((ring + i) * world + rank) * B, Be, d)
                eng.sync()
                h.copy_(d)
                hring.append(h)
                del d
            out_host = torch.empty(min(args.max_entries, wl["n_keys"]) * REC, dtype=torch.uint8).pin_memory()
        except Exception as ex:                                # noqa: BLE001 - reported in the JSON line
            prep_err = repr(ex)
        ok = torch.tensor([0 if prep_err else 1], device=dev, dtype=torch.int32)
        if world > 1:
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if int(ok.item()) == 1:
            def e2e_step(i):
                h = hring[i % len(hring)]
                if world == 1:
                    rc, took = eng.ingest(h.data_ptr(), Be)   # H2D inside
                    assert rc == 0 and took == Be
                else:
                    agg.ingest(h, Be)                          # H2D inside the local combiner's fa_ingest
            for i in range(2):
                e2e_step(i)
            eng.live_flows()
            barrier()
            t0 = time.perf_counter()
            d2h = 0
            for i in range(args.e2e_steps):
                e2e_step(i)
                eng.live_flows()                               # per-step result read-back (64 B)
                d2h += 64
            finish()
            nfl = eng.evict_into(out_host.data_ptr(), out_host.numel() // REC)   # final lookup-and-delete to host
            d2h += nfl * REC
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                tt = torch.tensor([dt], device=dev, dtype=torch.float64)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                dt = float(tt.item())
            e2e = {"value": world * Be * args.e2e_steps / dt / 1e6, "unit": "Mpkts/s",
                   "h2d_bytes_per_step": world * Be * REC, "d2h_bytes_per_step": d2h // args.e2e_steps,
                   "records_per_step": world * Be, "steps": args.e2e_steps, "flows_evicted_rank0": int(nfl),
                   "note": ("fa_ingest(pinned host records) + fa_live_flows per step, final fa_evict to host inside the "
                            "timed region; wall clock, max over ranks; every rank feeds its slice over its own PCIe link"
                            + ("" if world == 1 else "; d2h bytes are rank 0's"))}
        elif rank == 0:
            e2e = {"value": None, "unit": "Mpkts/s", "error": prep_err or "preparation failed on another rank"}

    if rank == 0:
        peak, peak_src = peaks()
        kern_ms = statistics.mean(step_ms)
        achieved = B * REC / (kern_ms / 1e3) / 1e9
        line = {"metric": "Mpkts/s aggregated", "value": value, "unit": "Mpkts/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "u64/u32 integer (add/or/min/max); no floating point", "data": "synthetic",
                "config": {"workload": wl["label"], "records_per_step_per_gpu": B, "record_bytes": REC,
                           "max_entries": args.max_entries, "max_batch": args.max_batch, "live_flows": int(flows),
                           "input_ring_batches": ring, "l2_policy": f"inputs larger than L2 ({B * REC / 1e6:.0f} MB per step, "
                           f"{ring} distinct batches cycled)",
                           "parallelism": "1 GPU" if world == 1 else
                           f"hash-sharded x{world}: " + ("" if args.no_combine else "per-batch local combine (K1+K2) -> ") +
                           ("K3 fused with the exchange (peer stores over NVLink, device-side counts)" if args.exchange == "peer"
                            else "K3 route -> NCCL all-to-all") + " -> K1 on the owner",
                           "exchanged_records_per_step_rank0": (agg.exchanged_records // (args.steps + args.warmup)) if world > 1 else None},
                "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                             "traffic": (NCU_DRAM_BYTES_PER_RECORD[args.workload] * min(B, args.max_batch)
                                         if args.workload in NCU_DRAM_BYTES_PER_RECORD and world == 1 else None),
                             "algorithmic_bytes_per_launch": REC * min(B, args.max_batch), "peak_source": peak_src,
                             "kernel": "fa::aggregate_kernel (K1); 144 algorithmic bytes per record; duration = CUDA-event "
                                       "time of one step = %d K1 launch(es) + 2 early-exit re-fold kernels each; traffic = "
                                       "ncu dram bytes/record x records per launch" % max(1, B // min(B, args.max_batch))},
                "gpu_launches": int(launches), "clocks": clocks,
                "order_fixups": st1["order_fixups"] - st0["order_fixups"], "spills": st1["spills"]}
        if e2e:
            line["e2e"] = e2e
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline_port(wl, args.cpu_sample)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="zipf1m", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=1 << 24, help="records per step per GPU (2^24 x 60 steps ~ 1e9)")
    ap.add_argument("--max-batch", type=int, default=1 << 24, help="records per K1 launch")
    ap.add_argument("--max-entries", type=int, default=1 << 26,
                    help="flow-cache capacity; >= flows + 3 x max_batch keeps fa_ingest on its zero-sync path")
    ap.add_argument("--ring", type=int, default=8, help="distinct pre-generated input batches cycled through")
    ap.add_argument("--e2e-batch", type=int, default=1 << 22)
    ap.add_argument("--e2e-steps", type=int, default=12)
    ap.add_argument("--cpu-sample", type=int, default=1 << 25)
    ap.add_argument("--ref-sample", type=int, default=1 << 24)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-combine", action="store_true", help="N>1: route raw records instead of per-batch partial flows")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N>1: 'peer' = K3 stores into the owners' buffers over NVLink (no host sync); 'nccl' = all_to_all_single")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
