"""Multi-GPU parity check (run under torchrun on a box with W = 2, 4 or 8 GPUs; `-m gpu` launches it through
tests/test_gpu_multi.py):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P tests/multi_gpu_parity.py

The N = 8 l parties of a run (l = SCZ_PARITY_L, default 1) are spread over the W ranks (N / W per GPU); every party
proves a 2^SCZ_PARITY_NV-constraint circuit (default 2^5) with dhyperplonk on seeded inputs, rank 0 re-runs the
oracle's N-party restatement and compares every party's outputs bit for bit (canonical affine for points).
SCZ_PARITY_NET = python (HybridNet: torch.distributed callbacks, default) | native (libscz's own NCCL hub)."""
import os
import sys

os.environ.setdefault("SCZ_MSM_STREAM", "1")   # exercise the side-stream MSM path too

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from tests.parity_util import nccl_parity_check  # noqa: E402


def main():
    dist.init_process_group("nccl")
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    l = int(os.environ.get("SCZ_PARITY_L", "1"))
    nv = int(os.environ.get("SCZ_PARITY_NV", "5"))
    kind = os.environ.get("SCZ_PARITY_NET", "python")
    info = nccl_parity_check(local, nv=nv, l=l, net_kind=kind)
    if dist.get_rank() == 0:
        assert info["result"] == "ok", info
        print(f"MULTI_GPU_PARITY_OK world={info['world']} l={l} parties={info['parties']} log2_constraints={nv} "
              f"parties_per_gpu={info['parties_per_gpu']} net={kind} collectives={info['collectives']}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
