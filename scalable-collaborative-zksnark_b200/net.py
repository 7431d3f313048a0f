"""The reference's star collectives (mpc-net/src/lib.rs:64-286, typed by
dist-primitive/src/utils/serializing_net.rs:8-142) carried by torch.distributed
-- NCCL over NVLink on the GPU box, gloo in the CPU tests -- in place of TCP.

`TorchDistNet` plugs into libscz through the `scz_net_vtable` callbacks: the
library hands over raw buffers (device pointers on the CUDA path) in its own
layout, nothing is serialised; `wire_bytes` only feeds the get_comm() counters.
One rank = one party (rank 0 = the leader, `MPCNet::is_leader`).
"""
import ctypes as C

import torch
import torch.distributed as dist

from .binding import NetVTable


class _CudaBuf:
    """minimal __cuda_array_interface__ view of a raw device pointer"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 2}


def _tensor(ptr, nbytes, device):
    """uint8 tensor aliasing `nbytes` at `ptr` (CUDA pointer when device is cuda, host pointer otherwise)"""
    if device.type == "cuda":
        return torch.as_tensor(_CudaBuf(ptr, nbytes), device=device)
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return torch.frombuffer(buf, dtype=torch.uint8)


def _on_stream(device, stream):
    """Context manager: torch ops inside run on the CUDA stream libscz passed to the callback (ctx->stream), whatever the
    calling thread's current torch stream is -- the collectives must be ordered after the kernels that produced
    d_send and before the ones that consume d_recv, and those run on ctx->stream (which differs from torch's current
    stream after scz_ctx_own_stream or inside a `with torch.cuda.stream(..)` block)."""
    import contextlib
    if device.type == "cuda" and stream:
        return torch.cuda.stream(torch.cuda.ExternalStream(int(stream), device=device))
    return contextlib.nullcontext()


class TorchDistNet:
    """gather / scatter / all_gather / sync of the MPCSerializeNet seam on a torch.distributed group."""

    def __init__(self, device, group=None):
        self.device = torch.device(device)
        self.group = group
        self.rank = dist.get_rank(group)
        self.n_parties = dist.get_world_size(group)
        self.calls = {"gather": 0, "scatter": 0, "all_gather": 0, "sync": 0}
        self._keep = None

    # -- the four collectives on tensors (also used directly by hosts that run several parties per rank)
    def gather_t(self, send, recv):
        """worker_send_or_leader_receive_element: leader's `recv` (n_parties * len(send)) is party-major"""
        self.calls["gather"] += 1
        if self.rank == 0:
            dist.gather(send, list(recv.view(self.n_parties, -1).unbind(0)), dst=0, group=self.group)
        else:
            dist.gather(send, None, dst=0, group=self.group)

    def scatter_t(self, send, recv):
        """worker_receive_or_leader_send_element: party j receives slice j of the leader's `send`"""
        self.calls["scatter"] += 1
        if self.rank == 0:
            dist.scatter(recv, list(send.view(self.n_parties, -1).unbind(0)), src=0, group=self.group)
        else:
            dist.scatter(recv, None, src=0, group=self.group)

    def all_gather_t(self, send, recv):
        """the N hub rounds of dhyperplonk.rs:271-294 as one exchange"""
        self.calls["all_gather"] += 1
        dist.all_gather_into_tensor(recv, send, group=self.group)

    def sync_t(self):
        self.calls["sync"] += 1
        dist.barrier(group=self.group)

    def gather_to_t(self, root, send, recv):
        """dynamic_worker_send_or_leader_receive_element (serializing_net.rs:41-74): hub = `root`"""
        self.calls["gather"] += 1
        if self.rank == root:
            dist.gather(send, list(recv.view(self.n_parties, -1).unbind(0)), dst=root, group=self.group)
        else:
            dist.gather(send, None, dst=root, group=self.group)

    def scatter_from_t(self, root, send, recv):
        """dynamic_worker_receive_or_worker_send_element (serializing_net.rs:98-126)"""
        self.calls["scatter"] += 1
        if self.rank == root:
            dist.scatter(recv, list(send.view(self.n_parties, -1).unbind(0)), src=root, group=self.group)
        else:
            dist.scatter(recv, None, src=root, group=self.group)

    # -- C callbacks
    def vtable(self):
        dev = self.device

        def _gather_to(user, root, d_send, d_recv, nbytes, wire, stream):
            try:
                with _on_stream(dev, stream):
                    send = _tensor(d_send, nbytes, dev)
                    recv = _tensor(d_recv, nbytes * self.n_parties, dev) if self.rank == root else None
                    self.gather_to_t(root, send, recv)
                return 0
            except Exception as e:
                print(f"[scz net] gather_to failed: {e!r}", flush=True)
                return 1

        def _scatter_from(user, root, d_send, d_recv, nbytes, wire, stream):
            try:
                with _on_stream(dev, stream):
                    recv = _tensor(d_recv, nbytes, dev)
                    send = _tensor(d_send, nbytes * self.n_parties, dev) if self.rank == root else None
                    self.scatter_from_t(root, send, recv)
                return 0
            except Exception as e:
                print(f"[scz net] scatter_from failed: {e!r}", flush=True)
                return 1

        def _gather(user, d_send, d_recv, nbytes, wire, stream):
            try:
                with _on_stream(dev, stream):
                    send = _tensor(d_send, nbytes, dev)
                    recv = _tensor(d_recv, nbytes * self.n_parties, dev) if self.rank == 0 else None
                    self.gather_t(send, recv)
                return 0
            except Exception as e:   # never let an exception cross the C boundary
                print(f"[scz net] gather failed: {e!r}", flush=True)
                return 1

        def _scatter(user, d_send, d_recv, nbytes, wire, stream):
            try:
                with _on_stream(dev, stream):
                    recv = _tensor(d_recv, nbytes, dev)
                    send = _tensor(d_send, nbytes * self.n_parties, dev) if self.rank == 0 else None
                    self.scatter_t(send, recv)
                return 0
            except Exception as e:
                print(f"[scz net] scatter failed: {e!r}", flush=True)
                return 1

        def _all_gather(user, d_send, d_recv, nbytes, wire, stream):
            try:
                with _on_stream(dev, stream):
                    self.all_gather_t(_tensor(d_send, nbytes, dev), _tensor(d_recv, nbytes * self.n_parties, dev))
                return 0
            except Exception as e:
                print(f"[scz net] all_gather failed: {e!r}", flush=True)
                return 1

        def _sync(user, stream):
            try:
                with _on_stream(dev, stream):
                    self.sync_t()
                return 0
            except Exception as e:
                print(f"[scz net] sync failed: {e!r}", flush=True)
                return 1

        vt = NetVTable()
        vt.user = None
        vt.gather = NetVTable._COLL(_gather)
        vt.scatter = NetVTable._COLL(_scatter)
        vt.all_gather = NetVTable._COLL(_all_gather)
        vt.sync = NetVTable._SYNC(_sync)
        vt.gather_to = NetVTable._ROOTED(_gather_to)
        vt.scatter_from = NetVTable._ROOTED(_scatter_from)
        self._keep = vt   # the callbacks must outlive the ctx
        return vt


class LocalTestNet:
    """All parties in ONE process on ONE GPU, one host thread per party -- the analogue of the reference's
    `LocalTestNet` (mpc-net/src/multi.rs:268-362: n parties as tokio tasks over loopback TCP).  Every party's
    ctx runs on the same CUDA stream, so a host-side barrier is all the ordering the star collectives need:
    whatever a party enqueued before the barrier precedes the copies the receiver enqueues after it."""

    def __init__(self, n_parties, device):
        import threading
        self.n = n_parties
        self.device = torch.device(device)
        self.barrier = threading.Barrier(n_parties)
        self.slots = [None] * n_parties
        self.leader_send = None
        cuda = self.device.type == "cuda"
        self.ev_ready = [torch.cuda.Event() if cuda else None for _ in range(n_parties)]
        self.ev_copied = [torch.cuda.Event() if cuda else None for _ in range(n_parties)]

    def party(self, party_id):
        return _LocalParty(self, party_id)

    def simulate_network_round(self, fn):
        """run fn(party_id, net_for_that_party) on n threads (multi.rs:329-352); returns the results by party"""
        import threading
        out, err = [None] * self.n, [None] * self.n

        def body(j):
            try:
                torch.cuda.set_device(self.device)
                out[j] = fn(j, self.party(j))
            except BaseException as e:   # noqa: BLE001 - re-raised on the caller's thread
                err[j] = e
                self.barrier.abort()
        ts = [threading.Thread(target=body, args=(j,)) for j in range(self.n)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        for e in err:
            if e is not None and not isinstance(e, __import__("threading").BrokenBarrierError):
                raise e
        for e in err:
            if e is not None:
                raise e
        return out


class _LocalParty:
    """One party of a LocalTestNet.  The torch copies of a callback run on the stream libscz passed in (ctx->stream);
    parties whose ctxs sit on different streams are fenced with events: a sender's `ready` event is recorded before the
    host barrier and waited for by every reader, a reader's `copied` event is waited for by everybody after the second
    barrier (so send buffers and stream-ordered temporaries are not reused under a pending copy).  On one shared
    stream the fences are no-ops."""

    def __init__(self, hub, party_id):
        self.hub, self.rank, self.n_parties = hub, party_id, hub.n
        self._keep = None

    def vtable(self):
        hub, me, dev = self.hub, self.rank, self.hub.device
        cuda = dev.type == "cuda"

        def ready():
            if cuda:
                hub.ev_ready[me].record()

        def wait_ready(who):
            if cuda:
                cur = torch.cuda.current_stream()
                for j in who:
                    cur.wait_event(hub.ev_ready[j])

        def copied_then_wait():
            if cuda:
                hub.ev_copied[me].record()
            hub.barrier.wait()
            if cuda:
                cur = torch.cuda.current_stream()
                for e in hub.ev_copied:
                    cur.wait_event(e)

        def _gather_to(user, root, d_send, d_recv, nbytes, wire, stream):
            try:
                with _on_stream(dev, stream):
                    hub.slots[me] = d_send
                    ready()
                    hub.barrier.wait()
                    if me == root:
                        wait_ready(range(hub.n))
                        recv = _tensor(d_recv, nbytes * hub.n, dev).view(hub.n, nbytes)
                        for j in range(hub.n):
                            recv[j].copy_(_tensor(hub.slots[j], nbytes, dev))
                    copied_then_wait()
                return 0
            except Exception as e:
                print(f"[scz local net] gather failed: {e!r}", flush=True)
                return 1

        def _scatter_from(user, root, d_send, d_recv, nbytes, wire, stream):
            try:
                with _on_stream(dev, stream):
                    if me == root:
                        hub.leader_send = d_send
                    ready()
                    hub.barrier.wait()
                    wait_ready([root])
                    src = _tensor(hub.leader_send, nbytes * hub.n, dev).view(hub.n, nbytes)
                    _tensor(d_recv, nbytes, dev).copy_(src[me])
                    copied_then_wait()
                return 0
            except Exception as e:
                print(f"[scz local net] scatter failed: {e!r}", flush=True)
                return 1

        def _gather(user, d_send, d_recv, nbytes, wire, stream):
            return _gather_to(user, 0, d_send, d_recv, nbytes, wire, stream)

        def _scatter(user, d_send, d_recv, nbytes, wire, stream):
            return _scatter_from(user, 0, d_send, d_recv, nbytes, wire, stream)

        def _all_gather(user, d_send, d_recv, nbytes, wire, stream):
            try:
                with _on_stream(dev, stream):
                    hub.slots[me] = d_send
                    ready()
                    hub.barrier.wait()
                    wait_ready(range(hub.n))
                    recv = _tensor(d_recv, nbytes * hub.n, dev).view(hub.n, nbytes)
                    for j in range(hub.n):
                        recv[j].copy_(_tensor(hub.slots[j], nbytes, dev))
                    copied_then_wait()
                return 0
            except Exception as e:
                print(f"[scz local net] all_gather failed: {e!r}", flush=True)
                return 1

        def _sync(user, stream):
            try:
                hub.barrier.wait()
                return 0
            except Exception:
                return 1

        vt = NetVTable()
        vt.user = None
        vt.gather = NetVTable._COLL(_gather)
        vt.scatter = NetVTable._COLL(_scatter)
        vt.all_gather = NetVTable._COLL(_all_gather)
        vt.sync = NetVTable._SYNC(_sync)
        vt.gather_to = NetVTable._ROOTED(_gather_to)
        vt.scatter_from = NetVTable._ROOTED(_scatter_from)
        self._keep = vt
        return vt


class HybridNet:
    """Fewer GPUs than parties: every rank hosts `per_rank` parties (one host thread and one ctx each, all on the
    rank's GPU), party id = rank * per_rank + local index; party 0 is the leader.
    A collective is a host barrier among the local parties, ONE torch.distributed collective per rank issued by
    local party 0 on the concatenated payloads, and a second host barrier.  With per_rank = 1 this is
    TorchDistNet; with world = 1 it is LocalTestNet.

    By default all hosted parties' ctxs share the caller's CUDA stream.  SCZ_PARTY_STREAMS=1 gives every hosted party
    its OWN high-priority stream (`party_stream`); the host barriers then order the ENQUEUEING and CUDA events order
    the EXECUTION across the parties' streams: every party fences its stream before local party 0 touches its buffers
    (`ready`), waits for party 0's work (`done`), and the owner of a source buffer waits for the readers (`copied`).
    Measured on 2 GPUs x 4 parties: bit-exact, but SLOWER than the shared stream (806 vs 714 ms per round of 8
    proofs) -- the star rounds keep the parties of one proof in the same phase, so there is nothing to overlap and
    four concurrent MSM sequences only contend; independent proofs are where separate streams pay (bench.py,
    `pipelined`).  The fences are exercised either way (they are no-ops on one stream)."""

    def __init__(self, device, per_rank, group=None):
        import os
        import threading
        self.device = torch.device(device)
        self.per_rank = per_rank
        self.group = group
        self.dist = dist.is_available() and dist.is_initialized()
        self.rank = dist.get_rank(group) if self.dist else 0
        self.world = dist.get_world_size(group) if self.dist else 1
        self.n = self.world * per_rank
        self.barrier = threading.Barrier(per_rank)
        self.slots = [None] * per_rank
        self.stage = None
        self.calls = {"gather": 0, "scatter": 0, "all_gather": 0, "sync": 0}
        self.cuda = self.device.type == "cuda"
        self.streams = None
        if self.cuda:
            self.ev_ready = [torch.cuda.Event() for _ in range(per_rank)]
            self.ev_copied = [torch.cuda.Event() for _ in range(per_rank)]
            self.ev_done = torch.cuda.Event()
            if per_rank > 1 and os.environ.get("SCZ_PARTY_STREAMS", "0") == "1":
                self.streams = [torch.cuda.Stream(device=self.device, priority=-1) for _ in range(per_rank)]

    def party(self, local_index):
        return _HybridParty(self, local_index)

    def close(self):
        pass

    def party_stream(self, p):
        """the CUDA stream of local party p (None: the caller's current stream)"""
        return self.streams[p] if self.streams else None

    def adopt(self, p, ctx):
        """bind a ctx that was created outside run_parties to local party p's stream"""
        if self.streams:
            with torch.cuda.stream(self.streams[p]):
                ctx.use_torch_stream()

    def _buf(self, nbytes):
        return torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    # -- stream fences of a collective (no-ops on CPU)
    def _on(self, stream):
        """torch ops of a callback run on the ctx's own stream, whatever the calling thread's current stream is"""
        import contextlib
        if self.cuda and stream:
            return torch.cuda.stream(torch.cuda.ExternalStream(int(stream), device=self.device))
        return contextlib.nullcontext()

    def _ready(self, p):
        if self.cuda:
            self.ev_ready[p].record()

    def _wait_ready(self):
        if self.cuda:
            cur = torch.cuda.current_stream()
            for e in self.ev_ready:
                cur.wait_event(e)

    def _done(self):
        if self.cuda:
            self.ev_done.record()

    def _wait_done(self):
        if self.cuda:
            torch.cuda.current_stream().wait_event(self.ev_done)

    def _copied(self, p):
        if self.cuda:
            self.ev_copied[p].record()

    def _wait_copied(self):
        if self.cuda:
            cur = torch.cuda.current_stream()
            for e in self.ev_copied:
                cur.wait_event(e)

    def run_parties(self, fn):
        """fn(party_id, local_index, net) on one thread per hosted party; returns the results by local index.
        With party streams the call still behaves like work enqueued on the caller's current stream: the party
        streams start after what the caller enqueued before, and the caller's stream continues after all of them."""
        import threading
        out, err = [None] * self.per_rank, [None] * self.per_rank
        start = end = None
        caller = torch.cuda.current_stream() if self.cuda else None   # a new thread would start on the default stream
        if self.streams:
            start = torch.cuda.Event()
            start.record()
            end = [torch.cuda.Event() for _ in range(self.per_rank)]

        def body(p):
            try:
                if self.cuda:
                    torch.cuda.set_device(self.device)
                if self.streams:
                    with torch.cuda.stream(self.streams[p]):
                        self.streams[p].wait_event(start)
                        out[p] = fn(self.rank * self.per_rank + p, p, self.party(p))
                        end[p].record()
                elif self.cuda:
                    with torch.cuda.stream(caller):
                        out[p] = fn(self.rank * self.per_rank + p, p, self.party(p))
                else:
                    out[p] = fn(self.rank * self.per_rank + p, p, self.party(p))
            except BaseException as e:   # noqa: BLE001
                err[p] = e
                self.barrier.abort()
                if hasattr(self, "abort"):
                    self.abort()         # native hub: release the parties waiting at its host barrier
        if self.per_rank == 1:
            body(0)
        else:
            ts = [threading.Thread(target=body, args=(p,)) for p in range(self.per_rank)]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        for e in err:
            if e is not None and not isinstance(e, __import__("threading").BrokenBarrierError):
                raise e
        for e in err:
            if e is not None:
                raise e
        if self.streams:
            cur = torch.cuda.current_stream()
            for e in end:
                cur.wait_event(e)
        return out


class _HybridParty:
    def __init__(self, hub, p):
        self.hub, self.p = hub, p
        self.rank = hub.rank * hub.per_rank + p      # party id
        self.n_parties = hub.n
        self._keep = None

    def vtable(self):
        hub, p, dev, P, W = self.hub, self.p, self.hub.device, self.hub.per_rank, self.hub.world

        def _gather_to(user, root, d_send, d_recv, nbytes, wire, stream):
            # the root party lives on rank root // P as local party root % P; local party 0 of every rank drives
            try:
                with hub._on(stream):
                    rr, rp = root // P, root % P
                    hub.slots[p] = d_send
                    if hub.rank == rr and p == rp:
                        hub.root_recv = d_recv
                    hub._ready(p)
                    hub.barrier.wait()
                    if p == 0:
                        hub._wait_ready()
                        hub.calls["gather"] += 1
                        if W == 1:
                            recv = _tensor(hub.root_recv, nbytes * hub.n, dev).view(P, nbytes)
                            for q in range(P):
                                recv[q].copy_(_tensor(hub.slots[q], nbytes, dev))
                        else:
                            loc = hub._buf(P * nbytes).view(P, nbytes)
                            for q in range(P):
                                loc[q].copy_(_tensor(hub.slots[q], nbytes, dev))
                            if hub.rank == rr:
                                recv = _tensor(hub.root_recv, nbytes * hub.n, dev).view(W, P * nbytes)
                                dist.gather(loc.view(-1), list(recv.unbind(0)), dst=rr, group=hub.group)
                            else:
                                dist.gather(loc.view(-1), None, dst=rr, group=hub.group)
                        hub._done()
                    hub.barrier.wait()
                    hub._wait_done()    # the root's buffer is filled, every send buffer may be reused
                return 0
            except Exception as e:
                print(f"[scz hybrid net] gather failed: {e!r}", flush=True)
                return 1

        def _scatter_from(user, root, d_send, d_recv, nbytes, wire, stream):
            try:
                with hub._on(stream):
                    rr, rp = root // P, root % P
                    if hub.rank == rr and p == rp:
                        hub.root_send = d_send
                    hub._ready(p)
                    hub.barrier.wait()
                    if p == 0:
                        hub._wait_ready()
                        if W == 1:
                            hub.stage = _tensor(hub.root_send, nbytes * hub.n, dev).view(P, nbytes)
                        else:
                            hub.calls["scatter"] += 1
                            loc = hub._buf(P * nbytes)
                            if hub.rank == rr:
                                send = _tensor(hub.root_send, nbytes * hub.n, dev).view(W, P * nbytes)
                                dist.scatter(loc, list(send.unbind(0)), src=rr, group=hub.group)
                            else:
                                dist.scatter(loc, None, src=rr, group=hub.group)
                            hub.stage = loc.view(P, nbytes)
                        hub._done()
                    hub.barrier.wait()
                    hub._wait_done()
                    _tensor(d_recv, nbytes, dev).copy_(hub.stage[p])
                    hub._copied(p)
                    hub.barrier.wait()
                    hub._wait_copied()  # the staging buffer / the root's send buffer is free again
                return 0
            except Exception as e:
                print(f"[scz hybrid net] scatter failed: {e!r}", flush=True)
                return 1

        def _gather(user, d_send, d_recv, nbytes, wire, stream):
            return _gather_to(user, 0, d_send, d_recv, nbytes, wire, stream)

        def _scatter(user, d_send, d_recv, nbytes, wire, stream):
            return _scatter_from(user, 0, d_send, d_recv, nbytes, wire, stream)

        def _all_gather(user, d_send, d_recv, nbytes, wire, stream):
            try:
                with hub._on(stream):
                    hub.slots[p] = d_send
                    hub._ready(p)
                    hub.barrier.wait()
                    if p == 0:
                        hub._wait_ready()
                        hub.calls["all_gather"] += 1
                        full = _tensor(d_recv, nbytes * hub.n, dev)
                        if W == 1:
                            for q in range(P):
                                full.view(P, nbytes)[q].copy_(_tensor(hub.slots[q], nbytes, dev))
                        else:
                            loc = hub._buf(P * nbytes).view(P, nbytes)
                            for q in range(P):
                                loc[q].copy_(_tensor(hub.slots[q], nbytes, dev))
                            dist.all_gather_into_tensor(full, loc.view(-1), group=hub.group)
                        hub.stage = full
                        hub._done()
                    hub.barrier.wait()
                    hub._wait_done()
                    if p != 0:
                        _tensor(d_recv, nbytes * hub.n, dev).copy_(hub.stage)
                    hub._copied(p)
                    hub.barrier.wait()
                    hub._wait_copied()  # party 0's buffer has been read by everybody
                return 0
            except Exception as e:
                print(f"[scz hybrid net] all_gather failed: {e!r}", flush=True)
                return 1

        def _sync(user, stream):
            try:
                with hub._on(stream):
                    hub._ready(p)
                    hub.barrier.wait()
                    if p == 0:
                        hub._wait_ready()
                        if W > 1:
                            hub.calls["sync"] += 1
                            dist.barrier(group=hub.group)
                        hub._done()
                    hub.barrier.wait()
                    hub._wait_done()
                return 0
            except Exception:
                return 1

        vt = NetVTable()
        vt.user = None
        vt.gather = NetVTable._COLL(_gather)
        vt.scatter = NetVTable._COLL(_scatter)
        vt.all_gather = NetVTable._COLL(_all_gather)
        vt.sync = NetVTable._SYNC(_sync)
        vt.gather_to = NetVTable._ROOTED(_gather_to)
        vt.scatter_from = NetVTable._ROOTED(_scatter_from)
        self._keep = vt
        return vt


class NativeNcclNet(HybridNet):
    """The same party layout as HybridNet with the data plane inside libscz (csrc/nccl_net.cu, include/scz.h "native
    data plane"): this class only creates the hub -- NCCL's unique id travels from rank 0 through torch.distributed, the
    channel a Python host happens to have; a Rust host would use its own -- and hands out parties whose ctxs are made
    by scz_ctx_create_on_hub.  No collective of a proof ever calls back into Python."""

    def __init__(self, device, per_rank, group=None):
        import os
        os.environ.setdefault("SCZ_PARTY_STREAMS", "0")
        super().__init__(device, per_rank, group)
        self.streams = None
        from .binding import lib
        self.L = lib()
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.world > 1:
            if self.rank == 0:
                buf = (C.c_uint8 * 128)()
                rc = self.L.scz_nccl_unique_id(buf)
                if rc != 0:
                    raise RuntimeError(f"scz_nccl_unique_id failed ({rc})")
                uid = torch.tensor(list(buf), dtype=torch.uint8)
            t = uid.to(self.device) if dist.get_backend(group) == "nccl" else uid
            dist.broadcast(t, src=0, group=group)
            uid = t.cpu()
        raw = (C.c_uint8 * 128)(*uid.tolist())
        hub = C.c_void_p()
        rc = self.L.scz_nccl_hub_create(C.c_int32(self.device.index or 0), C.c_uint32(self.rank), C.c_uint32(self.world),
                                        C.c_uint32(per_rank), raw, C.byref(hub))
        if rc != 0:
            raise RuntimeError(f"scz_nccl_hub_create failed ({rc})")
        self.hub = hub
        self._ctxs = 0

    @property
    def calls(self):
        out = (C.c_uint64 * 4)()
        if getattr(self, "hub", None):
            self.L.scz_nccl_hub_calls(self.hub, out)
        return {"gather": int(out[0]), "scatter": int(out[1]), "all_gather": int(out[2]), "sync": int(out[3])}

    @calls.setter
    def calls(self, _):   # HybridNet.__init__ assigns its own counter dict; the hub counts in C
        pass

    def party(self, local_index):
        return _NativeParty(self, local_index)

    def adopt(self, p, ctx):
        pass

    def abort(self):
        self.L.scz_nccl_hub_abort(self.hub)

    def close(self):
        """after every ctx created on the hub has been closed"""
        if getattr(self, "hub", None):
            self.L.scz_nccl_hub_destroy.restype = None
            self.L.scz_nccl_hub_destroy(self.hub)
            self.hub = None


class _NativeParty:
    def __init__(self, hub, p):
        self.hub, self.p = hub, p
        self.rank = hub.rank * hub.per_rank + p      # party id
        self.n_parties = hub.n

    def native_create(self, L, out_handle):
        return L.scz_ctx_create_on_hub(self.hub.hub, C.c_uint32(self.p), C.byref(out_handle))
