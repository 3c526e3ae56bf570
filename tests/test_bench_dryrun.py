"""bench.py host logic without a GPU: run_ours() is driven with stand-ins for torch / torch.distributed / the engine
so that a Python-level mistake (wrong name, wrong keyword, a rank-dependent code path) in the measurement script
shows up on the CPU instead of costing a GPU run.  Nothing here measures anything; the JSON line's shape is what
is checked (the keys the bench contract names)."""
import importlib
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeTensor:
    def __init__(self, n=1, value=0, device="cpu"):
        self.n, self.value, self.device = n, value, types.SimpleNamespace(index=0)
    def pin_memory(self): return self
    def data_ptr(self): return 0x10000
    def numel(self): return self.n
    def copy_(self, other): return self
    def item(self): return self.value
    def cpu(self): return self


class FakeEvent:
    def __init__(self, enable_timing=False): pass
    def record(self): pass
    def elapsed_time(self, other): return 1.0


def fake_torch(world):
    t = types.ModuleType("torch")
    t.uint8, t.int32, t.int64, t.float64 = "u8", "i32", "i64", "f64"
    t.device = lambda kind, idx=0: types.SimpleNamespace(type=kind, index=idx)
    t.empty = lambda n, dtype=None, device=None: FakeTensor(n)
    t.zeros = lambda n, dtype=None, device=None: FakeTensor(n)
    t.tensor = lambda v, device=None, dtype=None: FakeTensor(len(v), v[0])
    cuda = types.ModuleType("torch.cuda")
    cuda.is_available = lambda: True
    cuda.set_device = lambda i: None
    cuda.synchronize = lambda: None
    cuda.Event = FakeEvent
    cuda.Stream = lambda device=None: types.SimpleNamespace(cuda_stream=0x1234, synchronize=lambda: None)
    cuda.set_stream = lambda s: None
    cuda.current_stream = lambda: types.SimpleNamespace(cuda_stream=0x1234, synchronize=lambda: None)
    t.cuda = cuda
    dist = types.ModuleType("torch.distributed")
    dist.calls = []
    dist.ReduceOp = types.SimpleNamespace(MAX="max", MIN="min")
    dist.init_process_group = lambda backend, device_id=None: dist.calls.append("init")
    dist.all_reduce = lambda tensor, op=None: dist.calls.append(("all_reduce", op))
    dist.barrier = lambda: dist.calls.append("barrier")
    dist.destroy_process_group = lambda: dist.calls.append("destroy")
    dist.get_world_size = lambda: world
    dist.get_rank = lambda: 0
    t.distributed = dist
    return t, cuda, dist


class FakeEngine:
    created = []
    def __init__(self, max_entries, device=0, max_batch=0, cuda_stream=None, flags=0, **kw):
        assert cuda_stream, "bench must hand the engine an explicit stream"
        self.launches = 0
        FakeEngine.created.append(self)
    def gen_records(self, gp, first, n, dst): pass
    def sync(self): pass
    def ingest(self, ptr, n): self.launches += 3; return 0, n
    def live_flows(self): return 1000
    def evict_into(self, out, cap): return 1000
    def stats(self): return {"kernel_launches": self.launches, "order_fixups": 0, "spills": 0}
    def close(self): pass


class FakeAgg:
    def __init__(self, eng, max_batch, dev, **kw):
        self.eng, self.local, self.exchanged_records = eng, FakeEngine(1, cuda_stream=1), 0
        self.locals = [self.local]
    def ingest(self, records, n): self.eng.launches += 6; return 0
    def flush(self): return 0


def run_bench(monkeypatch, world, argv):
    torch, cuda, dist = fake_torch(world)
    monkeypatch.setitem(sys.modules, "torch", torch)
    monkeypatch.setitem(sys.modules, "torch.cuda", cuda)
    monkeypatch.setitem(sys.modules, "torch.distributed", dist)
    import netobserv_ebpf_agent_b200 as fa
    monkeypatch.setattr(fa, "FlowAggEngine", FakeEngine)
    sharded = types.ModuleType("netobserv_ebpf_agent_b200.sharded")
    sharded.PeerShardedAggregator = FakeAgg
    sharded.ShardedAggregator = FakeAgg
    monkeypatch.setitem(sys.modules, "netobserv_ebpf_agent_b200.sharded", sharded)
    monkeypatch.setenv("WORLD_SIZE", str(world)); monkeypatch.setenv("RANK", "0"); monkeypatch.setenv("LOCAL_RANK", "0")
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    bench = importlib.reload(bench)
    monkeypatch.setattr(bench.ClockSampler, "start", lambda self: None)
    monkeypatch.setattr(bench.ClockSampler, "stop", lambda self: {"sm_mhz": 1965.0, "sm_max_mhz": 1965.0, "reasons": []})
    monkeypatch.setattr(bench, "cpu_baseline_port", lambda wl, n, passes=4: {"value": 1.0, "unit": "Mpkts/s", "cores": 1,
                                                                               "kind": "port", "sample": "stub"})
    buf = io.StringIO()
    with redirect_stdout(buf):
        bench.main()
    lines = [ln for ln in buf.getvalue().splitlines() if ln.startswith("{")]
    assert len(lines) == 1, buf.getvalue()
    return json.loads(lines[0]), dist


CONTRACT_KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                 "vs_baseline", "dtype", "data", "config", "roofline", "gpu_launches", "clocks", "e2e")


def test_single_gpu_line_has_every_contract_key(monkeypatch):
    line, _ = run_bench(monkeypatch, 1, ["--steps", "3", "--warmup", "3", "--e2e-steps", "2"])
    for k in CONTRACT_KEYS + ("cpu_baseline",):
        assert k in line, k
    assert line["n_gpus"] == 1 and line["config"]["workload"].startswith("1e9-record stream, 1M")
    assert set(("bound", "achieved", "peak", "unit", "frac", "traffic")) <= set(line["roofline"])
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["e2e"]["d2h_bytes_per_step"] > 0
    assert line["gpu_launches"] > 0


@pytest.mark.parametrize("exchange", ["peer", "nccl"])
def test_multi_gpu_line_and_matched_collectives(monkeypatch, exchange):
    line, dist = run_bench(monkeypatch, 2, ["--gpus", "2", "--steps", "3", "--warmup", "3", "--e2e-steps", "2",
                                            "--exchange", exchange])
    for k in CONTRACT_KEYS[:-1]:
        assert k in line, k
    assert line["n_gpus"] == 2 and "cpu_baseline" not in line          # CPU baseline only at N=1
    if exchange == "peer":
        assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] == 2 * line["e2e"]["records_per_step"] // 2 * 144
        assert ("all_reduce", "min") in dist.calls                     # go/no-go agreed by all ranks before the e2e loop
    else:
        assert "e2e" not in line
    assert dist.calls[0] == "init" and dist.calls[-1] == "destroy"
