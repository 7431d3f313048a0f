"""-m gpu, needs >= 2 visible GPUs (skipped otherwise): the N = 8 l parties of one collaborative proof spread over
all GPUs of the box, one rank per GPU over NCCL, every party's proof bit for bit against the oracle's N-party run
(tests/multi_gpu_parity.py; reference: hyperplonk/examples/bench_hyperplonk.rs:32-56, one process per party)."""
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _run(world, l, nv, extra_env=None):
    env = dict(os.environ, SCZ_PARITY_L=str(l), SCZ_PARITY_NV=str(nv))
    env.update(extra_env or {})
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "multi_gpu_parity.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-6000:]
    assert f"MULTI_GPU_PARITY_OK world={world} l={l}" in r.stdout, r.stdout[-3000:]
    return r.stdout


@pytest.mark.parametrize("l,nv", [(1, 6), (2, 6)])
def test_parties_over_nccl_all_gpus(l, nv):
    g = _gpus()
    if g < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 8 if g >= 8 else 4 if g >= 4 else 2
    _run(world, l, nv)


def test_parties_over_nccl_native_net():
    """the same proof with the collectives issued by libscz.so itself (ncclSend / ncclRecv on the ctx stream,
    csrc/nccl_net.cu) instead of the Python torch.distributed callbacks: one party per rank needs 8 GPUs"""
    g = _gpus()
    if g < 8:
        pytest.skip("needs 8 GPUs (one party per rank)")
    out = _run(8, 1, 6, {"SCZ_PARITY_NET": "native"})
    assert "net=native" in out
