"""Hash-sharded multi-GPU aggregation: owner = fa_owner_hash(flow_id) % world.

One process per GPU.  Per batch: K3 (`fa_route`) groups the local records by owner -> the per-owner counts are
exchanged -> one variable-size all-to-all moves every record to its owning rank (NCCL over NVLink on GPUs, gloo in the
CPU tests) -> the owner folds what it received with K1.  Eviction needs no collective: the key sets are disjoint.
"""
import ctypes as C

import numpy as np

from ._lib import REC_BYTES, lib


def owner_of(keys40, world):
    """Owning rank of each 40-byte key (host side; same function K3 uses on the device)."""
    k = np.ascontiguousarray(keys40).view(np.uint8).reshape(-1, 40)
    out = np.empty(len(k), dtype=np.int64)
    f = lib().fa_owner_hash
    for i in range(len(k)):
        out[i] = f(C.c_void_p(k[i].ctypes.data)) % world
    return out


def route_host(records, world):
    """CPU twin of K3: stable partition of (n,144) records by owner -> (grouped records, counts[world])."""
    r = np.ascontiguousarray(records).view(np.uint8).reshape(-1, REC_BYTES)
    own = owner_of(r[:, :40], world)
    order = np.argsort(own, kind="stable")
    return r[order], np.bincount(own, minlength=world).astype(np.int64)


def exchange(send, send_counts, group=None):
    """All-to-all of fixed-size records.

    send: 1-D uint8 torch tensor holding the records grouped by destination rank; send_counts: records per
    destination (len == world).  Returns (recv tensor, recv_counts list).  Works for CUDA tensors over NCCL and for
    CPU tensors over gloo."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    cin = torch.as_tensor(np.asarray(send_counts, dtype=np.int64), device=send.device)
    cout = torch.empty(world, dtype=torch.int64, device=send.device)
    dist.all_to_all_single(cout, cin, group=group)
    out_counts = [int(x) for x in cout.cpu().tolist()]
    in_counts = [int(x) for x in send_counts]
    recv = torch.empty(sum(out_counts) * REC_BYTES, dtype=torch.uint8, device=send.device)
    dist.all_to_all_single(recv, send[: sum(in_counts) * REC_BYTES],
                           output_split_sizes=[c * REC_BYTES for c in out_counts],
                           input_split_sizes=[c * REC_BYTES for c in in_counts], group=group)
    return recv, out_counts


class ShardedAggregator:
    """GPU path: [combine locally ->] route on the device, exchange over NCCL, fold on the owner.

    combine=True puts a local combiner in front of the exchange (the reference does the same thing with its
    per-CPU maps folded in user space, pkg/tracer/tracer.go:1159-1187): each batch is first aggregated into a
    scratch flow table on the source GPU (K1), drained (K2 in lookup-and-reset mode) into partial 144-byte flow
    records, and only those partials are routed — the owner folds partials exactly like single-packet records
    (AccumulateBase).  On heavy-tailed traffic this cuts the bytes crossing NVLink by an order of magnitude.

    pipeline=True (with combine) ping-pongs two scratch tables on their own streams: while the host drains,
    routes and exchanges the partials of batch k (several small kernels and host synchronisations), the GPU is
    already folding batch k+1 into the other scratch table.  Call flush() before reading results."""

    def __init__(self, engine, max_batch, device, combine=True, pipeline=True):
        """The owner engine must have been created on the CURRENT torch stream (cuda_stream=torch.cuda.
        current_stream().cuda_stream of a non-default stream): route, the NCCL exchange and the fold are then
        stream-ordered."""
        import torch
        import torch.distributed as dist
        from .engine import FlowAggEngine
        assert torch.cuda.current_stream().cuda_stream != 0, "use an explicit torch.cuda.Stream (see bench.py)"
        self.eng, self.max_batch, self.world = engine, max_batch, dist.get_world_size()
        self.send = torch.empty(max_batch * REC_BYTES, dtype=torch.uint8, device=device)
        self.locals, self.parts, self.streams = [], [], []
        self.pending, self.turn = None, 0
        if combine:
            for _ in range(2 if pipeline else 1):
                st = torch.cuda.Stream(device=device) if pipeline else torch.cuda.current_stream()
                # scratch flow table of the combiner; flows stay cached across batches (drain, not evict)
                self.locals.append(FlowAggEngine(2 * max_batch, device=device.index, max_batch=max_batch,
                                                 cuda_stream=st.cuda_stream))
                self.parts.append(torch.empty(max_batch * REC_BYTES, dtype=torch.uint8, device=device))
                self.streams.append(st)
        self.exchanged_records = 0

    def ingest(self, records, n):
        """records: device tensor / address of n local records (they must stay valid until flush())."""
        import torch
        folded, done = 0, 0
        base = records.data_ptr() if hasattr(records, "data_ptr") else int(records)
        while done < n:
            c = min(self.max_batch, n - done)
            src = base + done * REC_BYTES
            done += c
            if not self.locals:
                folded += self._exchange_and_fold(src, c)
                continue
            i = self.turn
            self.turn = (self.turn + 1) % len(self.locals)
            if self.pending == i:                    # single scratch table: finish the previous batch first
                folded += self._finish(i)
            self.streams[i].wait_stream(torch.cuda.current_stream())   # the batch may have been produced on the main stream
            rc, took = self.locals[i].ingest(src, c)                   # asynchronous on the scratch table's stream
            if self.pending is not None and self.pending != i:
                folded += self._finish(self.pending)                   # overlaps with the fold just launched
            if rc != 0:                              # scratch table full: flush it completely, then fold the rest
                live = self.locals[i].live_flows()
                big = torch.empty(max(live, 1) * REC_BYTES, dtype=torch.uint8, device=self.send.device)
                k = self.locals[i].evict_into(big, live)
                for off in range(0, k, self.max_batch):
                    folded += self._exchange_and_fold(big.data_ptr() + off * REC_BYTES, min(self.max_batch, k - off))
                del big
                rc, took2 = self.locals[i].ingest(src + took * REC_BYTES, c - took)
                assert rc == 0 and took + took2 == c, (rc, took, took2)
            self.pending = i
        return folded

    def _finish(self, i):
        # partial flow records of the batch folded into scratch table i; the flows stay cached there
        c = self.locals[i].drain_active(self.parts[i], self.max_batch)     # host-synchronous on that table's stream
        self.pending = None
        return self._exchange_and_fold(self.parts[i].data_ptr(), c)

    def flush(self):
        """Drain + exchange whatever is still sitting in a scratch table."""
        return self._finish(self.pending) if self.pending is not None else 0

    def _exchange_and_fold(self, src, c):
        self.exchanged_records += c
        counts = self.eng.route(src, c, self.world, self.send)
        recv, out_counts = exchange(self.send, counts)
        tot = sum(out_counts)
        rc, took = self.eng.ingest(recv.data_ptr(), tot)
        assert rc == 0 and took == tot, (rc, took)
        return tot

    def reset_local(self):
        """See PeerShardedAggregator.reset_local."""
        import torch
        self.flush()
        for loc in self.locals:
            live = loc.live_flows()
            if live:
                tmp = torch.empty(live * REC_BYTES, dtype=torch.uint8, device=self.send.device)
                loc.evict_into(tmp, live)
                del tmp

    def close(self):
        for l in self.locals:
            l.close()
        self.locals = []


class PeerShardedAggregator:
    """Sharded aggregation with the exchange fused into K3: no NCCL data movement, no host synchronisation.

    Per batch, all enqueued on one stream without waiting for anything:
      K1 into the scratch table (local combine) -> K2 drain (count stays on the device) ->
      K3 `fa_route_peer`: partition the partials by owner and store them straight into the owners' receive
      buffers over NVLink (CUDA IPC mappings, one remote atomicAdd per CTA and shard reserves the room) ->
      a 4-byte NCCL all-reduce as stream-ordered barrier -> the owner folds what it received
      (`fa_ingest_counted`, count read from its own device memory) and re-arms the buffer.
    Receive buffers are double-buffered so that batch k+1 may be delivered while batch k is being folded."""

    def __init__(self, engine, max_batch, device, recv_cap=None, local_entries=None, profile=False):
        """`local_entries` sizes the combiner's scratch table: it has to hold the DISTINCT flows of one round (default:
        2 x max_batch, enough for any round; a smaller table that overflows makes flush() raise).  `profile` brackets
        the phases of every round with CUDA events (no synchronisation), read back by exchange_stats()."""
        import torch
        import torch.distributed as dist
        from ._lib import FA_F_NO_FULL_CUT, check
        from .engine import FlowAggEngine
        assert torch.cuda.current_stream().cuda_stream != 0, "use an explicit torch.cuda.Stream (see bench.py)"
        self.eng, self.max_batch = engine, max_batch
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.profile, self._ev = profile, []
        local_entries = local_entries or 2 * max_batch
        # a round's partials are distinct flows of the scratch table: neither the drain buffer nor (in expectation, owners
        # being a uniform hash) a receive buffer needs more room than that table has slots; what does not fit is counted
        # (overflow[0]) and makes flush() raise
        need, slots = (4 * local_entries + 2) // 3, 1024
        while slots < need:
            slots <<= 1                                  # the engine's table sizing: a drain can never emit more than this
        self.part_cap = min(max_batch, slots)
        self.recv_cap = recv_cap or self.part_cap
        self.local = FlowAggEngine(local_entries, device=device.index, max_batch=max_batch, flags=FA_F_NO_FULL_CUT,
                                   cuda_stream=torch.cuda.current_stream().cuda_stream)
        self.part = torch.empty(self.part_cap * REC_BYTES, dtype=torch.uint8, device=device)
        self.part_n = torch.zeros(1, dtype=torch.int64, device=device)
        self.token = torch.zeros(1, dtype=torch.int32, device=device)
        L = lib()

        def dalloc(nbytes):
            p = C.c_void_p()
            check(L.fa_device_alloc(engine._h, nbytes, C.byref(p)))
            return p.value
        self.mine = [(dalloc(self.recv_cap * REC_BYTES), dalloc(8)) for _ in range(2)]     # (buffer, counter) x 2
        self.overflow = dalloc(16)                      # [0] receive-buffer overflow, [1] records stored into other GPUs
        # exchange the IPC handles of the four allocations and map every peer's
        hbuf = np.zeros((4, 64), dtype=np.uint8)
        for i, p in enumerate([self.mine[0][0], self.mine[0][1], self.mine[1][0], self.mine[1][1]]):
            check(L.fa_ipc_export(engine._h, C.c_void_p(p), C.c_void_p(hbuf[i].ctypes.data)))
        mine_t = torch.from_numpy(hbuf.reshape(-1).copy()).to(device)
        all_t = [torch.empty_like(mine_t) for _ in range(self.world)]
        dist.all_gather(all_t, mine_t)
        self.mapped = []
        self.bufs = [(C.c_void_p * 16)() for _ in range(2)]
        self.cnts = [(C.c_void_p * 16)() for _ in range(2)]
        for r in range(self.world):
            hs = all_t[r].cpu().numpy().reshape(4, 64)
            ptrs = []
            for i in range(4):
                if r == self.rank:
                    ptrs.append([self.mine[0][0], self.mine[0][1], self.mine[1][0], self.mine[1][1]][i])
                else:
                    q = C.c_void_p()
                    hh = np.ascontiguousarray(hs[i])
                    check(L.fa_ipc_open(engine._h, C.c_void_p(hh.ctypes.data), C.byref(q)))
                    self.mapped.append(q.value)
                    ptrs.append(q.value)
            self.bufs[0][r], self.cnts[0][r], self.bufs[1][r], self.cnts[1][r] = ptrs
        dist.barrier()
        self.step = 0
        self.exchanged_records = 0      # not known on the host in this mode

    def ingest(self, records, n):
        import torch.distributed as dist
        from ._lib import check
        L = lib()
        base = records.data_ptr() if hasattr(records, "data_ptr") else int(records)
        done = 0
        while done < n:
            c = min(self.max_batch, n - done)
            b = self.step & 1
            mark = self._mark if self.profile else (lambda: None)
            mark()
            rc, took = self.local.ingest(base + done * REC_BYTES, c)
            assert rc == 0 and took == c, (rc, took)
            mark()
            check(L.fa_drain_active_counted(self.local._h, C.c_void_p(self.part.data_ptr()), self.part_cap,
                                            C.c_void_p(self.part_n.data_ptr())))
            mark()
            check(L.fa_route_peer(self.eng._h, C.c_void_p(self.part.data_ptr()), C.c_void_p(self.part_n.data_ptr()),
                                  self.part_cap, self.world, self.rank, self.bufs[b], self.cnts[b], self.recv_cap,
                                  C.c_void_p(self.overflow)))
            mark()
            dist.all_reduce(self.token)                    # stream-ordered barrier: every rank has delivered batch `step`
            mark()
            check(L.fa_ingest_counted(self.eng._h, C.c_void_p(self.mine[b][0]), C.c_void_p(self.mine[b][1]),
                                      self.recv_cap, 1))
            mark()
            self.step += 1
            done += c
        return 0

    def _mark(self):
        import torch
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self._ev.append(e)

    def phase_ms(self, last_rounds=None):
        """Mean milliseconds per round of (combine K1, drain K2, route K3 + peer stores, barrier, owner fold K1) over the
        profiled rounds (the last `last_rounds` of them); the stream must be idle."""
        import torch
        torch.cuda.current_stream().synchronize()
        ev = self._ev[-6 * last_rounds:] if last_rounds else self._ev
        names = ("combine", "drain", "route", "barrier", "fold")
        tot = dict.fromkeys(names, 0.0)
        rounds = len(ev) // 6
        for r in range(rounds):
            for i, k in enumerate(names):
                tot[k] += ev[6 * r + i].elapsed_time(ev[6 * r + i + 1])
        return {k: v / max(rounds, 1) for k, v in tot.items()}

    def _counters(self):
        import ctypes
        import torch
        torch.cuda.current_stream().synchronize()
        ov = np.zeros(2, dtype=np.uint64)
        cudart = ctypes.CDLL("libcudart.so")
        cudart.cudaMemcpy(ctypes.c_void_p(ov.ctypes.data), ctypes.c_void_p(self.overflow), 16, 2)
        return int(ov[0]), int(ov[1])

    def flush(self):
        overflow, _ = self._counters()
        if overflow:
            raise RuntimeError(f"peer receive buffer overflow: {overflow} records did not fit (raise recv_cap)")
        spills = self.local.stats()["spills"] + self.eng.stats()["spills"]
        if spills:                                   # a scratch / owner table that is physically full drops records
            raise RuntimeError(f"{spills} records found a flow table physically full (raise max_entries / evict more often)")
        return 0

    def reset_local(self):
        """To be called whenever the owner tables are evicted: the combiner's scratch table caches keys WITH their start
        mirror, so a flow that outlives its owner-side eviction would send partials without a start ("already covered").
        Lookup-and-delete of the scratch table, output discarded."""
        import torch
        live = self.local.live_flows()
        if live:
            tmp = torch.empty(live * REC_BYTES, dtype=torch.uint8, device=self.part.device)
            self.local.evict_into(tmp, live)
            del tmp

    def exchange_stats(self):
        """What crossed NVLink so far: records this rank stored into other GPUs' receive buffers (x 144 B)."""
        _, remote = self._counters()
        return {"nvlink_records_rank0": remote, "nvlink_bytes_rank0": remote * REC_BYTES, "rounds": self.step,
                "nvlink_bytes_per_round_rank0": remote * REC_BYTES // max(self.step, 1)}

    def close(self):
        import torch.distributed as dist
        from ._lib import check
        torch_sync = __import__("torch").cuda.synchronize
        torch_sync()
        dist.barrier()
        L = lib()
        for q in self.mapped:
            L.fa_ipc_close(self.eng._h, C.c_void_p(q))
        self.mapped = []
        dist.barrier()
        for bufp, cntp in self.mine:
            L.fa_device_free(self.eng._h, C.c_void_p(bufp)); L.fa_device_free(self.eng._h, C.c_void_p(cntp))
        L.fa_device_free(self.eng._h, C.c_void_p(self.overflow))
        self.local.close()
