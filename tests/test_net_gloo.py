"""CPU (no GPU): the N > 1 host path -- TorchDistNet and HybridNet over torch.distributed/gloo, world_size 2 --
plus HybridNet with one rank (the LocalTestNet shape: all parties in one process)."""
import os
import socket
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_star_collectives_world_size_2_gloo():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "_gloo_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "GLOO_NET_OK" in r.stdout


def test_hybrid_net_single_rank_four_parties():
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _gloo_worker import check_party
    from scz_b200.net import HybridNet
    hub = HybridNet("cpu", per_rank=4)
    res = hub.run_parties(lambda pid, p, pnet: check_party(pnet.vtable(), pid, 4, pid == 0))
    assert res == [True] * 4
