"""2-GPU parity of the sharded path (K3 route -> NCCL all-to-all -> K1 on the owner) against the single-stream oracle.
Skipped unless two CUDA devices are visible (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest
import torch

import oracle_lib as O
from common import gen_host

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q, combine):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    import netobserv_ebpf_agent_b200 as fa
    from netobserv_ebpf_agent_b200.sharded import ShardedAggregator, owner_of
    stream = torch.cuda.Stream(device=dev)             # engine, NCCL and copies share one explicit stream
    torch.cuda.set_stream(stream)
    from netobserv_ebpf_agent_b200.sharded import PeerShardedAggregator
    if combine == "peer":
        eng = fa.FlowAggEngine(1 << 18, device=rank, max_batch=100_000, cuda_stream=stream.cuda_stream, flags=fa.FA_F_NO_FULL_CUT)
        agg = PeerShardedAggregator(eng, 100_000, dev)
    else:
        eng = fa.FlowAggEngine(1 << 18, device=rank, max_batch=30_000, cuda_stream=stream.cuda_stream)
        agg = ShardedAggregator(eng, 30_000, dev, combine=combine)
    keep = []
    for b in range(3):
        local = gen_host(seed=50, n=100_000, n_keys=40_000, dist=1, first=(b * world + rank) * 100_000)
        t = torch.from_numpy(np.ascontiguousarray(local).reshape(-1).copy()).to(dev)
        agg.ingest(t, 100_000)
        keep.append(t)                               # batches stay valid until flush()
    agg.flush()
    out = eng.evict()
    assert (owner_of(out[:, :40], world) == rank).all()
    q.put((rank, out.copy()))
    dist.barrier()
    agg.close()
    eng.close()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("combine", [False, True, "peer"])
def test_two_gpu_sharded_parity(combine):
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q, combine)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    got = O.sort_records(np.concatenate([outs[0], outs[1]]))
    single = O.Accounter(1 << 20)
    for b in range(3):
        for rank in range(world):
            single.account(gen_host(seed=50, n=100_000, n_keys=40_000, dist=1, first=(b * world + rank) * 100_000))
    want = O.sort_records(single.evict())
    assert np.array_equal(got, want)
