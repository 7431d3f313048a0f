"""Byte-level `MPCNet` (mpc-net/src/lib.rs:34-286) over torch.distributed -- NCCL on the GPU box, gloo on CPU --
for callers of the reference that still want raw `send_to` / `recv_from` and the star helpers built on them, in
place of the multiplexed TCP streams of mpc-net/src/multi.rs.

Same names, argument meaning, return shapes and error behaviour as the trait:
  * `worker_send_or_leader_receive(bytes, sid)` -> list of n byte strings on the leader (its own first-hand copy at
    index 0, lib.rs:78-104), `None` on workers; the `dynamic_` form takes the receiver (lib.rs:111-162);
  * `worker_receive_or_leader_send(bytes_out, sid)` -> this party's slice; passing bytes when not the leader, or
    nothing when the leader, is `MPCNetError::BadInput` (lib.rs:178-205); the `dynamic_` form also enforces equal
    lengths (`MPCNetError::Protocol`, lib.rs:231-237);
  * `leader_compute(bytes, sid, f)` and `sync()` are the trait's default compositions (lib.rs:259-285);
  * `get_comm()` / `add_comm()`: upload grows by the payload length on every successful send, download on every
    receive (multi.rs:389-417) -- the 8-byte length frame below is transport, not payload, and is not counted.

`sid` (MultiplexedStreamID Zero / One / Two, lib.rs:28-32) becomes the point-to-point tag, so messages of
different streams between the same two parties cannot be confused (gloo honours tags; NCCL ignores them and
delivers in call order per pair, which is the order the reference's protocols use anyway).  A message travels as an 8-byte length followed
by the payload, both as uint8 tensors on the net's device.
"""
import torch
import torch.distributed as dist


class MPCNetError(RuntimeError):
    """mpc-net/src/lib.rs:15-26 (Generic / Protocol / BadInput)"""

    def __init__(self, kind, err, party=None):
        super().__init__(f"{kind}: {err}" + (f" (party {party})" if party is not None else ""))
        self.kind, self.party = kind, party


class MultiplexedStreamID:
    Zero, One, Two = 0, 1, 2


class TorchDistMPCNet:
    """one rank = one party; rank 0 is the leader (`is_leader`, lib.rs:38-40)"""

    def __init__(self, device="cpu", group=None):
        self.device = torch.device(device)
        self.group = group
        self._init = dist.is_available() and dist.is_initialized()
        if not self._init:
            raise MPCNetError("Generic", "torch.distributed is not initialised")
        self._id = dist.get_rank(group)
        self._n = dist.get_world_size(group)
        self.upload = 0
        self.download = 0

    # -- the trait's required methods
    def is_leader(self):
        return self._id == 0

    def n_parties(self):
        return self._n

    def party_id(self):
        return self._id

    def is_init(self):
        return self._init

    def get_comm(self):
        return (self.upload, self.download)

    def add_comm(self, up, down):
        self.upload += up
        self.download += down

    def _peer(self, pid):
        if not (0 <= pid < self._n) or pid == self._id:
            raise MPCNetError("Generic", f"Peer {pid} not found")   # multi.rs:391-393
        return pid if self.group is None else dist.get_global_rank(self.group, pid)

    def send_to(self, pid, data, sid=MultiplexedStreamID.Zero):
        peer = self._peer(pid)
        data = bytes(data)
        head = torch.tensor(list(len(data).to_bytes(8, "little")), dtype=torch.uint8, device=self.device)
        dist.send(head, dst=peer, group=self.group, tag=sid)
        if data:
            body = torch.frombuffer(bytearray(data), dtype=torch.uint8).to(self.device)
            dist.send(body, dst=peer, group=self.group, tag=sid)
        self.upload += len(data)

    def recv_from(self, pid, sid=MultiplexedStreamID.Zero):
        peer = self._peer(pid)
        head = torch.empty(8, dtype=torch.uint8, device=self.device)
        dist.recv(head, src=peer, group=self.group, tag=sid)
        n = int.from_bytes(bytes(head.cpu().tolist()), "little")
        data = b""
        if n:
            body = torch.empty(n, dtype=torch.uint8, device=self.device)
            dist.recv(body, src=peer, group=self.group, tag=sid)
            data = bytes(body.cpu().numpy().tobytes())
        self.download += n
        return data

    # -- the trait's provided methods
    def worker_send_or_leader_receive(self, data, sid=MultiplexedStreamID.Zero):
        return self.dynamic_worker_send_or_leader_receive(data, 0, sid)

    def dynamic_worker_send_or_leader_receive(self, data, receiver, sid=MultiplexedStreamID.Zero):
        data = bytes(data)
        if receiver == self._id:
            return [data if j == self._id else self.recv_from(j, sid) for j in range(self._n)]
        self.send_to(receiver, data, sid)
        return None

    def worker_receive_or_leader_send(self, bytes_out, sid=MultiplexedStreamID.Zero):
        if bytes_out is not None:
            if not self.is_leader():
                raise MPCNetError("BadInput", "recv_from_leader called with bytes_out when not leader")
            for j in range(self._n):
                if j != self._id:
                    self.send_to(j, bytes_out[j], sid)
            return bytes(bytes_out[self._id])
        if self.is_leader():
            raise MPCNetError("BadInput", "recv_from_leader called with no bytes_out when leader")
        return self.recv_from(0, sid)

    def dynamic_worker_receive_or_leader_send(self, bytes_out, sender, sid=MultiplexedStreamID.Zero):
        if bytes_out is not None:
            if self._id != sender:
                raise MPCNetError("BadInput", "recv_from_leader called with bytes_out when not leader")
            m = len(bytes_out[0])
            for j in range(self._n):
                if j == self._id:
                    continue
                if len(bytes_out[j]) != m:
                    raise MPCNetError("Protocol", f"The leader sent wrong number of bytes to Peer {j}", party=j)
                self.send_to(j, bytes_out[j], sid)
            return bytes(bytes_out[self._id])
        if self._id == sender:
            raise MPCNetError("BadInput", "recv_from_leader called with no bytes_out when leader")
        return self.recv_from(sender, sid)

    def leader_compute(self, data, sid, f):
        got = self.worker_send_or_leader_receive(data, sid)
        return self.worker_receive_or_leader_send(f(got) if got is not None else None, sid)

    def sync(self):
        got = self.worker_send_or_leader_receive(b"\x87", MultiplexedStreamID.Zero)   # vec![135u8; 1], lib.rs:274
        back = self.worker_receive_or_leader_send(got, MultiplexedStreamID.Zero)
        assert back == b"\x87"
