"""-m gpu: the collaborative (PSS) permutation check -- c_acc_product_and_share (dacc_product.rs:66-292) and cpermcheck
(dhyperplonk.rs:1249-1385) -- through the C ABI against the oracle's restatement, leader mode and 8-party mode."""
import numpy as np
import pytest

from tests.gpu_util import oracle_affine

pytestmark = pytest.mark.gpu

TABLES = ("V", "sid", "ssigma", "eq_r1", "mask", "unmask0", "unmask1", "unmask2")


def _srs(ctx, orc, rng, sizes):
    dev, host = [], []
    for m in sizes:
        b = ctx.g1_generator_mul(ctx.to_device(orc.random_fr(rng, m), 4))
        dev.append(b)
        host.append(oracle_affine(ctx.to_host(b)))
    return dev, orc.Srs.from_levels(host)


def _pk(orc, rng, n, l, shared=None):
    L = (4 << n) // l
    pk = {k: orc.random_fr(rng, L) for k in TABLES}
    for k, m in (("challenge_r1", n + 2), ("alpha", 1), ("beta", 1)):
        pk[k] = shared[k] if shared is not None else orc.random_fr(rng, m)
    return pk


def _product_tables(pk):
    t = {k: pk[k] for k in TABLES + ("challenge_r1",)}
    t["alpha_beta"] = np.concatenate([pk["alpha"], pk["beta"]])
    return t


def _same(orc, got, want, who):
    (gp, gc), (wp, wc, wo) = got
    assert gp == [] and gc == []
    assert len(wp) == len(want["wiring_proofs"]) == 6 and len(wc) == len(want["wiring_commits"]) == 10
    assert len(wo) == len(want["wiring_opens"]) == 12
    for k, (a, b) in enumerate(zip(wp, want["wiring_proofs"])):
        assert np.array_equal(a, b), (who, "proof", k)
    for k, (a, b) in enumerate(zip(wc, want["wiring_commits"])):
        assert orc.canon_g1(a) == orc.canon_g1(b), (who, "commit", k)
    for k, ((v, p), (ov, op)) in enumerate(zip(wo, want["wiring_opens"])):
        assert np.array_equal(v, ov) and orc.canon_g1(p) == orc.canon_g1(op), (who, "open", k)


@pytest.mark.parametrize("l,n", [(1, 4), (1, 7), (2, 6)])
def test_c_acc_product_and_share_leader_mode(orc, l, n):
    import scz_b200 as scz
    from oracle import hyperplonk as ohp
    rng = np.random.default_rng(700 + 10 * l + n)
    ctx = scz.Context(device=0, n_parties=8 * l)
    pp, opp = scz.PackedSharingParams(ctx, l), orc.pp_new(l)
    L = (4 << n) // l
    args = [orc.random_fr(rng, L) for _ in range(5)]
    got = scz.c_acc_product_and_share(ctx, pp, *args)
    want = ohp.c_acc_product_and_share(opp, orc.LEADER_SIM, *[[a] for a in args])[0]
    for k in range(3):
        assert np.array_equal(got[k], want[k]), k
    up, down = ctx.get_comm()
    assert up > 0 and down > 0
    ctx.close()


def test_cpermcheck_leader_mode(orc):
    import scz_b200 as scz
    from oracle import hyperplonk as ohp
    n, l, N = 4, 1, 8
    rng = np.random.default_rng(720)
    ctx = scz.Context(device=0, n_parties=N)
    pp, opp = scz.PackedSharingParams(ctx, l), orc.pp_new(l)
    cdev, csrs = _srs(ctx, orc, rng, [max(1, (1 << i) // l) for i in range(n + 3)])
    pk = _pk(orc, rng, n, l)
    pk["c_commitment"] = csrs
    want = ohp.cpermcheck(n, [pk], opp, orc.LEADER_SIM, N)[0]
    got = scz.cpermcheck(ctx, n, _product_tables(pk), scz.PolynomialCommitment(ctx, cdev).precompute(), pp).nested()
    _same(orc, got, want, "leader")
    ctx.close()


def test_cpermcheck_parties_mode(orc):
    """N = 8 parties on one GPU under LocalTestNet: the moving-hub rounds (gather_to / scatter_from) for real"""
    import scz_b200 as scz
    from oracle import hyperplonk as ohp
    from scz_b200.net import LocalTestNet
    n, l, N = 4, 1, 8
    rng = np.random.default_rng(740)
    opp = orc.pp_new(l)
    seed_ctx = scz.Context(device=0, n_parties=N)
    pks, devs = [], []
    for j in range(N):
        cdev, csrs = _srs(seed_ctx, orc, rng, [1 << i for i in range(n + 3)])
        pk = _pk(orc, rng, n, l, shared=pks[0] if pks else None)
        pk["c_commitment"] = csrs
        pks.append(pk)
        devs.append(cdev)
    want = ohp.cpermcheck(n, pks, opp, orc.PARTIES, N)
    want_sh = ohp.c_acc_product_and_share(opp, orc.PARTIES, *[[pk[k] for pk in pks] for k in ("V", "mask", "unmask0", "unmask1", "unmask2")])

    def party(j, net):
        c = scz.Context(device=0, party_id=j, n_parties=N, net=net)
        pp = scz.PackedSharingParams(c, l)
        sh = scz.c_acc_product_and_share(c, pp, *[pks[j][k] for k in ("V", "mask", "unmask0", "unmask1", "unmask2")])
        got = scz.cpermcheck(c, n, _product_tables(pks[j]), scz.PolynomialCommitment(c, devs[j]), pp).nested()
        c.sync()
        c.close()
        return sh, got

    res = LocalTestNet(N, "cuda:0").simulate_network_round(party)
    for j in range(N):
        for k in range(3):
            assert np.array_equal(res[j][0][k], want_sh[j][k]), (j, k)
        _same(orc, res[j][1], want[j], f"party {j}")
    seed_ctx.close()
