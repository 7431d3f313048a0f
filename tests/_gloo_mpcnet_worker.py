"""worker of tests/test_mpcnet_gloo.py: the byte-level MPCNet (scz_b200.mpcnet) over gloo -- the semantics of the
trait's provided methods (mpc-net/src/lib.rs:64-285) and of the comm counters (multi.rs:389-417)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist  # noqa: E402

from scz_b200.mpcnet import MPCNetError, MultiplexedStreamID as Sid, TorchDistMPCNet  # noqa: E402


def main():
    backend = os.environ.get("SCZ_TEST_BACKEND", "gloo")   # "nccl" on the GPU box: one GPU per rank
    dev = "cpu"
    if backend == "nccl":
        import torch
        dev = f"cuda:{os.environ.get('LOCAL_RANK', '0')}"
        torch.cuda.set_device(dev)
    dist.init_process_group(backend)
    net = TorchDistMPCNet(dev)
    me, n = net.party_id(), net.n_parties()
    assert net.is_init() and net.is_leader() == (me == 0) and n == dist.get_world_size()
    mine = bytes([me + 1]) * (5 + me)                       # ragged lengths
    got = net.worker_send_or_leader_receive(mine, Sid.One)
    if me == 0:
        assert got == [bytes([j + 1]) * (5 + j) for j in range(n)]
        assert net.get_comm() == (0, sum(5 + j for j in range(1, n)))
    else:
        assert got is None and net.get_comm() == (5 + me, 0)
    out = [bytes([100 + j]) * 3 for j in range(n)] if me == 0 else None
    back = net.worker_receive_or_leader_send(out, Sid.One)
    assert back == bytes([100 + me]) * 3
    # leader_compute: the leader reverses every party's message
    r = net.leader_compute(mine, Sid.Two, lambda v: [b[::-1] + bytes([len(v)]) for b in v])
    assert r == mine[::-1] + bytes([n])
    # moving hub: every party takes a turn
    for hub in range(n):
        g = net.dynamic_worker_send_or_leader_receive(mine, hub, Sid.Zero)
        assert (g == [bytes([j + 1]) * (5 + j) for j in range(n)]) if me == hub else g is None
        s = net.dynamic_worker_receive_or_leader_send([bytes([hub, j]) for j in range(n)] if me == hub else None, hub, Sid.Zero)
        assert s == bytes([hub, me])
    # empty payloads travel too
    e = net.worker_send_or_leader_receive(b"", Sid.Zero)
    assert e == ([b""] * n if me == 0 else None)
    net.sync()
    # error behaviour
    for bad in ((lambda: net.worker_receive_or_leader_send([b"x"] * n if me != 0 else None, Sid.Zero)),
                (lambda: net.send_to(me, b"x")), (lambda: net.recv_from(n + 3))):
        try:
            bad()
            raise SystemExit("expected MPCNetError")
        except MPCNetError as ex:
            assert ex.kind in ("BadInput", "Generic")
    if me == 0:
        try:
            net.dynamic_worker_receive_or_leader_send([b"ab"] + [b"a"] * (n - 1), 0, Sid.Zero)
            raise SystemExit("expected Protocol error")
        except MPCNetError as ex:
            assert ex.kind == "Protocol" and ex.party == 1
    up, down = net.get_comm()
    net.add_comm(7, 9)
    assert net.get_comm() == (up + 7, down + 9)
    dist.barrier()
    if me == 0:
        print("GLOO_MPCNET_OK", net.get_comm())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
