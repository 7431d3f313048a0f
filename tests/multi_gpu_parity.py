"""Multi-GPU parity check (run under torchrun on a box with W = 2, 4 or 8 GPUs; not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node W --master-addr 127.0.0.1 --master-port P tests/multi_gpu_parity.py

The N = 8 l parties of a run (l = SCZ_PARITY_L, default 1) are spread over the W ranks (N / W per GPU, HybridNet over
NCCL); every party proves a 2^SCZ_PARITY_NV-constraint circuit (default 2^5) with dhyperplonk on seeded inputs, rank 0
re-runs the oracle's N-party restatement and compares every party's outputs bit for bit (canonical affine for
points)."""
import os
import sys

os.environ.setdefault("SCZ_MSM_STREAM", "1")   # exercise the side-stream MSM path too

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import scz_b200 as scz  # noqa: E402
from oracle import hyperplonk as ohp  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from scz_b200.net import HybridNet  # noqa: E402
from tests.gpu_util import oracle_affine  # noqa: E402
from tests.test_gpu_hyperplonk import _same_proof, _tables_for_product  # noqa: E402

L_PACK = int(os.environ.get("SCZ_PARITY_L", "1"))
N, NV = 8 * L_PACK, int(os.environ.get("SCZ_PARITY_NV", "5"))


def all_inputs():
    """every rank regenerates ALL parties' inputs from the same seed (host side); SRS points from seeded scalars"""
    rng = np.random.default_rng(2024)
    csz, dsz = ohp.srs_level_sizes(NV, L_PACK, N)
    pks, scal = [], []
    for j in range(N):
        ks = ([orc.random_fr(rng, m) for m in csz], [orc.random_fr(rng, m) for m in dsz])
        scal.append(ks)
        pks.append(ohp.random_pk(rng, NV, L_PACK, N, None, None, shared=pks[0] if pks else None))
    return pks, scal


def main():
    dist.init_process_group("nccl")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    P = N // world
    pks, scal = all_inputs()
    hub = HybridNet(dev, P)
    seed_ctx = scz.Context(device=local, n_parties=N)
    srs_dev = {}
    for p in range(P):
        j = rank * P + p
        srs_dev[p] = ([seed_ctx.g1_generator_mul(seed_ctx.to_device(k, 4)) for k in scal[j][0]],
                      [seed_ctx.g1_generator_mul(seed_ctx.to_device(k, 4)) for k in scal[j][1]])

    def party(pid, p, net):
        c = scz.Context(device=local, party_id=pid, n_parties=N, net=net)
        pp = scz.PackedSharingParams(c, L_PACK)
        c_srs = scz.PolynomialCommitment(c, srs_dev[p][0]).precompute()
        d_srs = scz.PolynomialCommitment(c, srs_dev[p][1])
        pk = scz.PackedProvingParameters(c, NV, L_PACK, _tables_for_product(pks[pid]), c_srs, d_srs)
        out = scz.dhyperplonk(c, NV, pk, pp).nested()
        c.sync()
        c.close()
        return out

    mine = hub.run_parties(party)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    if rank == 0:
        # oracle side: the SRS points as the devices made them (rank 0 rebuilds all of them on its GPU)
        for j in range(N):
            pks[j]["c_commitment"] = orc.Srs.from_levels(
                [oracle_affine(seed_ctx.to_host(seed_ctx.g1_generator_mul(seed_ctx.to_device(k, 4)))) for k in scal[j][0]])
            pks[j]["d_commitment"] = orc.Srs.from_levels(
                [oracle_affine(seed_ctx.to_host(seed_ctx.g1_generator_mul(seed_ctx.to_device(k, 4)))) for k in scal[j][1]])
        want = ohp.dhyperplonk(NV, pks, orc.pp_new(L_PACK), orc.PARTIES, N)
        flat = [x for per_rank in gathered for x in per_rank]
        for j in range(N):
            _same_proof(orc, flat[j], want[j], f"party {j} (rank {j // P})")
        print(f"MULTI_GPU_PARITY_OK world={world} l={L_PACK} parties={N} log2_constraints={NV} parties_per_gpu={P} collectives={hub.calls}", flush=True)
    dist.barrier()
    seed_ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
