"""worker of tests/test_net_gloo.py: world_size 2 over gloo on CPU buffers.  Drives the scz_net_vtable callbacks of
TorchDistNet (one party per rank) and HybridNet (two parties per rank) exactly as libscz does -- raw pointers, byte
counts -- and checks the star-collective semantics of serializing_net.rs:11-39, 76-96 and the hub rounds of
dhyperplonk.rs:271-294."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch.distributed as dist  # noqa: E402

from scz_b200.net import HybridNet, TorchDistNet  # noqa: E402


def ptr(a):
    return a.ctypes.data if a is not None else None


def check_party(vt, pid, n, leader):
    nb = 24
    mine = np.full(nb, 10 + pid, dtype=np.uint8)
    recv = np.zeros(n * nb, dtype=np.uint8) if leader else None
    assert vt.gather(None, ptr(mine), ptr(recv), nb, nb, None) == 0
    if leader:
        assert all((recv[j * nb:(j + 1) * nb] == 10 + j).all() for j in range(n)), recv
    send = np.repeat(np.arange(n, dtype=np.uint8) + 50, nb) if leader else None
    got = np.zeros(nb, dtype=np.uint8)
    assert vt.scatter(None, ptr(send), ptr(got), nb, nb, None) == 0
    assert (got == 50 + pid).all(), (pid, got)
    full = np.zeros(n * nb, dtype=np.uint8)
    assert vt.all_gather(None, ptr(mine), ptr(full), nb, nb, None) == 0
    assert all((full[j * nb:(j + 1) * nb] == 10 + j).all() for j in range(n)), (pid, full)
    # moving hub (serializing_net.rs:41-74, 98-126): every party takes a turn as the root
    for root in range(n):
        recv = np.zeros(n * nb, dtype=np.uint8) if pid == root else None
        assert vt.gather_to(None, root, ptr(mine), ptr(recv), nb, nb, None) == 0
        if pid == root:
            assert all((recv[j * nb:(j + 1) * nb] == 10 + j).all() for j in range(n)), (root, recv)
        send = np.repeat(np.arange(n, dtype=np.uint8) + 100 + root, nb) if pid == root else None
        got = np.zeros(nb, dtype=np.uint8)
        assert vt.scatter_from(None, root, ptr(send), ptr(got), nb, nb, None) == 0
        assert (got == 100 + root + pid).all(), (root, pid, got)
    assert vt.sync(None, None) == 0
    return True


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    net = TorchDistNet("cpu")
    assert check_party(net.vtable(), rank, world, rank == 0)
    hub = HybridNet("cpu", per_rank=2)
    res = hub.run_parties(lambda pid, p, pnet: check_party(pnet.vtable(), pid, hub.n, pid == 0))
    assert res == [True, True] and hub.n == 2 * world
    dist.barrier()
    if rank == 0:
        print("GLOO_NET_OK", hub.calls)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
