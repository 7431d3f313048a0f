"""Checker shared by tests/multi_gpu_parity.py (torchrun, `-m gpu` via tests/test_gpu_multi.py) and bench.py's
untimed `parity_check` leg at N > 1: the N = 8 l parties of one collaborative proof spread over the live
torch.distributed world (one rank per GPU), every party's `dhyperplonk` output bit for bit against the oracle's
N-party restatement (oracle/hyperplonk.py) on the same seeded inputs.  Test infrastructure: imports the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def all_inputs(nv, l, n_parties, seed=2024):
    """every rank regenerates ALL parties' inputs from the same seed (host side); SRS points from seeded scalars"""
    import numpy as np
    from oracle import hyperplonk as ohp
    from oracle import oracle as orc
    rng = np.random.default_rng(seed)
    csz, dsz = ohp.srs_level_sizes(nv, l, n_parties)
    pks, scal = [], []
    for _ in range(n_parties):
        scal.append(([orc.random_fr(rng, m) for m in csz], [orc.random_fr(rng, m) for m in dsz]))
        pks.append(ohp.random_pk(rng, nv, l, n_parties, None, None, shared=pks[0] if pks else None))
    return pks, scal


def nccl_parity_check(local_device, nv=5, l=1, net_kind="python", precompute=True):
    """Needs an initialised torch.distributed NCCL group with world in {1, 2, 4, 8} ranks.  net_kind: "python" =
    HybridNet (torch.distributed callbacks), "native" = libscz's own NCCL hub (csrc/nccl_net.cu).
    Returns (on rank 0) a dict describing the check; info["result"] is "ok" or "MISMATCH: ...".  None on other ranks."""
    import torch
    import torch.distributed as dist
    import scz_b200 as scz
    from oracle import hyperplonk as ohp
    from oracle import oracle as orc
    from scz_b200.net import HybridNet, NativeNcclNet
    from tests.gpu_util import oracle_affine
    from tests.test_gpu_hyperplonk import _same_proof, _tables_for_product

    n_parties = 8 * l
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda", local_device)
    per_rank = n_parties // world
    pks, scal = all_inputs(nv, l, n_parties)
    hub = NativeNcclNet(dev, per_rank) if net_kind == "native" else HybridNet(dev, per_rank)
    seed_ctx = scz.Context(device=local_device, n_parties=n_parties)
    srs_dev = {}
    for p in range(per_rank):
        j = rank * per_rank + p
        srs_dev[p] = ([seed_ctx.g1_generator_mul(seed_ctx.to_device(k, 4)) for k in scal[j][0]],
                      [seed_ctx.g1_generator_mul(seed_ctx.to_device(k, 4)) for k in scal[j][1]])

    def party(pid, p, net):
        c = scz.Context(device=local_device, party_id=pid, n_parties=n_parties, net=net)
        pp = scz.PackedSharingParams(c, l)
        c_srs = scz.PolynomialCommitment(c, srs_dev[p][0])
        if precompute:
            c_srs.precompute()
        d_srs = scz.PolynomialCommitment(c, srs_dev[p][1])
        pk = scz.PackedProvingParameters(c, nv, l, _tables_for_product(pks[pid]), c_srs, d_srs)
        out = scz.dhyperplonk(c, nv, pk, pp).nested()
        c.sync()
        c.close()
        return out

    mine = hub.run_parties(party)
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(mine, gathered, dst=0)
    info = None
    if rank == 0:
        # oracle side: the SRS points as the devices made them (rank 0 rebuilds all of them on its GPU)
        for j in range(n_parties):
            pks[j]["c_commitment"] = orc.Srs.from_levels(
                [oracle_affine(seed_ctx.to_host(seed_ctx.g1_generator_mul(seed_ctx.to_device(k, 4)))) for k in scal[j][0]])
            pks[j]["d_commitment"] = orc.Srs.from_levels(
                [oracle_affine(seed_ctx.to_host(seed_ctx.g1_generator_mul(seed_ctx.to_device(k, 4)))) for k in scal[j][1]])
        orc.set_msm_threads(os.cpu_count() or 1)
        want = ohp.dhyperplonk(nv, pks, orc.pp_new(l), orc.PARTIES, n_parties)
        orc.set_msm_threads(1)
        flat = [x for per in gathered for x in per]
        info = {"result": "ok", "world": world, "l": l, "parties": n_parties, "log2_constraints": nv,
                "parties_per_gpu": per_rank, "net": net_kind, "collectives": dict(hub.calls),
                "checked": "every party's sumcheck triples, commitments, opening values and proofs bit for bit "
                           "against oracle/hyperplonk.py (canonical affine for points)"}
        try:
            for j in range(n_parties):
                _same_proof(orc, flat[j], want[j], f"party {j} (rank {j // per_rank})")
        except AssertionError as e:   # reported, not raised: the other ranks are waiting at the barrier below
            info["result"] = f"MISMATCH: {e!r}"[:400]
    dist.barrier()
    seed_ctx.close()
    hub.close()
    return info
