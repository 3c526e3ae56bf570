"""world_size-2 gloo test of the multi-GPU host logic (owner hash, count exchange, variable all-to-all) on CPU.
The folding itself needs a GPU; here the received records are folded with the CPU oracle only to check that the
sharded result equals the single-stream result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_lib as O
from common import gen_host


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from netobserv_ebpf_agent_b200 import sharded
    acc = O.Accounter(1 << 20)
    for b in range(3):
        local = gen_host(seed=40, n=20_000, n_keys=3_000, dist=1, first=(b * world + rank) * 20_000)
        grouped, counts = sharded.route_host(local, world)
        send = torch.from_numpy(np.ascontiguousarray(grouped).reshape(-1).copy())
        recv, out_counts = sharded.exchange(send, counts)
        got = recv.numpy().reshape(-1, 144)
        assert len(got) == sum(out_counts)
        assert (sharded.owner_of(got[:, :40], world) == rank).all()      # only keys this rank owns arrive
        acc.account(got)
    q.put((rank, acc.evict().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_equals_single_stream():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    outs = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    sharded_result = O.sort_records(np.concatenate([outs[0], outs[1]]))
    keys0 = {bytes(k) for k in outs[0][:, :40]}
    keys1 = {bytes(k) for k in outs[1][:, :40]}
    assert not (keys0 & keys1)                                           # disjoint shards: evict needs no collective
    single = O.Accounter(1 << 20)
    for b in range(3):
        for rank in range(world):
            single.account(gen_host(seed=40, n=20_000, n_keys=3_000, dist=1, first=(b * world + rank) * 20_000))
    want = O.sort_records(single.evict())
    # per-key-constant descriptors + commutative counters: the sharded fold is bit-identical to the single stream
    assert np.array_equal(sharded_result, want)


def test_owner_hash_matches_oracle_spec():
    recs = gen_host(seed=41, n=2_000, n_keys=500)
    from netobserv_ebpf_agent_b200 import sharded
    own = sharded.owner_of(recs[:, :40], 8)
    want = np.array([O.lib().oracle_owner_hash(O._p(np.ascontiguousarray(r[:40]))) % 8 for r in recs])
    assert np.array_equal(own, want)
    assert len(set(own.tolist())) == 8
