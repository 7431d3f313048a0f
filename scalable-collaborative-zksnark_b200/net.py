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

    # -- C callbacks
    def vtable(self):
        dev = self.device

        def _gather(user, d_send, d_recv, nbytes, wire, stream):
            try:
                send = _tensor(d_send, nbytes, dev)
                recv = _tensor(d_recv, nbytes * self.n_parties, dev) if self.rank == 0 else None
                self.gather_t(send, recv)
                return 0
            except Exception as e:   # never let an exception cross the C boundary
                print(f"[scz net] gather failed: {e!r}", flush=True)
                return 1

        def _scatter(user, d_send, d_recv, nbytes, wire, stream):
            try:
                recv = _tensor(d_recv, nbytes, dev)
                send = _tensor(d_send, nbytes * self.n_parties, dev) if self.rank == 0 else None
                self.scatter_t(send, recv)
                return 0
            except Exception as e:
                print(f"[scz net] scatter failed: {e!r}", flush=True)
                return 1

        def _all_gather(user, d_send, d_recv, nbytes, wire, stream):
            try:
                self.all_gather_t(_tensor(d_send, nbytes, dev), _tensor(d_recv, nbytes * self.n_parties, dev))
                return 0
            except Exception as e:
                print(f"[scz net] all_gather failed: {e!r}", flush=True)
                return 1

        def _sync(user, stream):
            try:
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
        self._keep = vt   # the callbacks must outlive the ctx
        return vt
