"""-m gpu: the Fr-table kernels and the dist-primitive protocols through the C ABI, bit-exact against the oracle.

Leader mode (the reference's build without `comm`, serializing_net.rs:144-264) runs on one ctx; parties mode runs
N = 8 ctxs on one GPU under `LocalTestNet`, the analogue of the reference's own LocalTestNet (multi.rs:268-362)."""
import numpy as np
import pytest

from oracle import py_twin as tw
from tests.gpu_util import oracle_affine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import scz_b200 as scz
    c = scz.Context(device=0, n_parties=8)
    yield c
    c.close()


@pytest.fixture(scope="module")
def ctx_l():
    """leader-mode contexts for l > 1: N = 8 l parties (pss.rs:38-41)"""
    import scz_b200 as scz
    made = {}

    def get(l):
        if l not in made:
            made[l] = scz.Context(device=0, n_parties=8 * l)
        return made[l]
    yield get
    for c in made.values():
        c.close()


def _levels(ctx, orc, rng, sizes):
    """random bases per level (the shape of new_single / new_random, dpoly_comm.rs:197-234): device + oracle copies"""
    dev, host = [], []
    for n in sizes:
        b = ctx.g1_generator_mul(ctx.to_device(orc.random_fr(rng, n), 4))
        dev.append(b)
        host.append(oracle_affine(ctx.to_host(b)))
    return dev, host


# ------------------------------------------------------------------------------------------- kernels
# 19, 20: the first rounds run in the lazy kernel (unreduced 17-limb round sums, poly.cu k_sumcheck_round), which takes
# over from ~4 pairs per thread of a one-CTA-per-SM grid; below that every round is k_sumcheck_round_direct
@pytest.mark.parametrize("logn", [0, 1, 2, 5, 11, 12, 13, 15, 19, 20])
def test_sumcheck_product_matches_oracle(orc, ctx, logn):
    import scz_b200 as scz
    rng = np.random.default_rng(400 + logn)
    n = 1 << logn
    f, g = orc.random_fr(rng, n), orc.random_fr(rng, n)
    ch = orc.random_fr(rng, max(logn, 1))
    got = scz.sumcheck_product(ctx, f, g, ch)
    assert np.array_equal(got, orc.sumcheck_product(f, g, ch))
    fd, gd = ctx.to_device(f, 4), ctx.to_device(g, 4)
    scz.sumcheck_product(ctx, fd, gd, ctx.to_device(ch, 4))
    assert np.array_equal(ctx.to_host(fd), f) and np.array_equal(ctx.to_host(gd), g)      # inputs untouched


def test_sumcheck_round_identity_at_2p20(orc, ctx):
    """BASELINE config 3 size (c_sumcheck_product tables of 2^20): verifier identity of check_sumcheck_product
    (dsumcheck.rs:558-588): the degree-2 round polynomial through (p(0), p(1), p(2)) evaluated at the challenge
    equals p(0) + p(1) of the next round; round 0 sums to <f, g>."""
    import scz_b200 as scz
    rng = np.random.default_rng(420)
    logn = 20
    n = 1 << logn
    import torch
    gen = torch.Generator(device="cuda").manual_seed(5)
    f = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=gen)
    g = torch.randint(-2**63, 2**63 - 1, (n, 4), dtype=torch.int64, device="cuda", generator=gen)
    f[:, 3] &= (1 << 62) - 1
    g[:, 3] &= (1 << 62) - 1
    ch = orc.random_fr(rng, logn)
    out = ctx.to_host(scz.sumcheck_product(ctx, f, g, ctx.to_device(ch, 4)))
    R = tw.R_MOD
    tri = [[orc.fr_to_ints(out[i, j:j + 1])[0] for j in range(3)] for i in range(logn + 1)]
    chi = orc.fr_to_ints(ch)
    inv2 = pow(2, R - 2, R)
    for i in range(logn):
        p0, p1, p2 = tri[i]
        r = chi[i]
        # Lagrange through x = 0, 1, 2
        val = (p0 * (r - 1) * (r - 2) * inv2 - p1 * r * (r - 2) + p2 * r * (r - 1) * inv2) % R
        nxt = (tri[i + 1][0] + tri[i + 1][1]) % R
        assert val == nxt, i
    # <f, g> on the device equals p0 + p1 of round 0
    prod = ctx.fr_op("mul", f, g)
    while len(prod) > 1:
        h = len(prod) // 2
        prod = ctx.fr_op("add", prod[:h].contiguous(), prod[h:].contiguous())
    assert orc.fr_to_ints(ctx.to_host(prod))[0] == (tri[0][0] + tri[0][1]) % R


@pytest.mark.parametrize("logn,k", [(1, 1), (4, 2), (13, 2), (14, 5), (3, 7)])
def test_fix_variable(orc, ctx, logn, k):
    import scz_b200 as scz
    rng = np.random.default_rng(430 + logn)
    e, pts = orc.random_fr(rng, 1 << logn), orc.random_fr(rng, k)
    assert np.array_equal(scz.fix_variable(ctx, e, pts), orc.fix_variable(e, pts))


@pytest.mark.parametrize("logm", [0, 1, 2, 8, 9, 10, 14])
def test_acc_product_tree(orc, ctx, logm):
    import scz_b200 as scz
    rng = np.random.default_rng(440 + logm)
    x = orc.random_fr(rng, 1 << logm)
    assert np.array_equal(scz.acc_product_tree(ctx, x), orc.acc_product_tree(x))


def test_acc_product_known_answer(orc, ctx):
    """dacc_product.rs:450-466 (vectors valid for input 1..=4)"""
    import scz_b200 as scz
    t = scz.acc_product_tree(ctx, orc.fr_from_ints([1, 2, 3, 4]))
    ints = orc.fr_to_ints(t)
    assert ints[0::2] == [1, 3, 2, 24] and ints[1::2] == [2, 4, 12, 0] and ints[4:] == [2, 12, 24, 0]


def test_pointwise_maps(orc, ctx):
    import scz_b200 as scz
    rng = np.random.default_rng(450)
    n = 5000
    a, b = orc.random_fr(rng, n), orc.random_fr(rng, n)
    k = orc.random_fr(rng, 2)
    assert np.array_equal(scz.fr_pointwise(ctx, "add", a, b), orc.fr_add(a, b))
    assert np.array_equal(scz.fr_pointwise(ctx, "rsub", a, b), orc.fr_sub(b, a))
    want = orc.fr_add(orc.fr_add(a, orc.fr_mul(np.repeat(k[0:1], n, axis=0), b)), np.repeat(k[1:2], n, axis=0))
    assert np.array_equal(scz.fr_pointwise(ctx, "axpb", a, b, k), want)
    b[n - 1] = orc.fr_from_ints([1])[0]
    inv = orc.fr_inv(b)
    assert np.array_equal(scz.fr_pointwise(ctx, "div", a, b), orc.fr_mul(a, inv))
    for m in (1, 3, 127, 128, 129, 1025):
        assert np.array_equal(scz.fr_pointwise(ctx, "div", a[:m], b[:m]), orc.fr_mul(a[:m], inv[:m])), m
    # a zero denominator: arkworks panics (dhyperplonk.rs:338-339).  The device path writes 0 for that element, keeps
    # the rest of the batch right and raises the ctx's sticky SCZ_STATUS_DIV_BY_ZERO bit; the host path raises.
    assert ctx.take_status() == 0
    b[7] = 0
    b[4096] = 0
    inv[7] = 0
    inv[4096] = 0
    with pytest.raises(ZeroDivisionError):
        scz.fr_pointwise(ctx, "div", a, b)
    assert ctx.take_status() == 0                      # taken (and cleared) by the failed call
    got = scz.fr_pointwise(ctx, "div", ctx.to_device(a, 4), ctx.to_device(b, 4))
    assert np.array_equal(ctx.to_host(got), orc.fr_mul(a, inv))
    assert ctx.take_status() == 1 and ctx.take_status() == 0


# ------------------------------------------------------------------------------------------- leader mode
@pytest.mark.parametrize("l", [1, 2, 4])
def test_resharing_rounds_leader_mode(orc, ctx_l, l):
    import scz_b200 as scz
    ctx = ctx_l(l)
    rng = np.random.default_rng(460 + l)
    pp, opp = scz.PackedSharingParams(ctx, l), orc.pp_new(l)
    x = orc.random_fr(rng, 1)
    assert np.array_equal(scz.pss2ss(ctx, pp, x), orc.pss2ss(opp, orc.LEADER_SIM, x)[0])
    assert np.array_equal(scz.degree_reduce(ctx, pp, x), orc.degree_reduce(opp, orc.LEADER_SIM, x))
    if l == 1:   # closed forms of BASELINE.md 4
        xi = orc.fr_to_ints(x)[0]
        assert orc.fr_to_ints(scz.pss2ss(ctx, pp, x))[0] == xi * tw.MU0 % tw.R_MOD
        assert orc.fr_to_ints(scz.degree_reduce(ctx, pp, x))[0] == xi * tw.LAMBDA0 % tw.R_MOD


@pytest.mark.parametrize("l,logn", [(1, 6), (1, 13), (2, 5), (4, 12)])
def test_c_sumcheck_product_leader_mode(orc, ctx_l, l, logn):
    import scz_b200 as scz
    ctx = ctx_l(l)
    rng = np.random.default_rng(470 + 10 * l + logn)
    pp, opp = scz.PackedSharingParams(ctx, l), orc.pp_new(l)
    n = 1 << logn
    f, g, ch = orc.random_fr(rng, n), orc.random_fr(rng, n), orc.random_fr(rng, logn + 3)
    got = scz.c_sumcheck_product(ctx, pp, f, g, ch)
    assert np.array_equal(got, orc.c_sumcheck_product(opp, orc.LEADER_SIM, [f], [g], ch)[0])


@pytest.mark.parametrize("logn", [2, 7, 13])
def test_d_sumcheck_product_leader_mode(orc, ctx, logn):
    import scz_b200 as scz
    rng = np.random.default_rng(480 + logn)
    n = 1 << logn
    f, g, ch = orc.random_fr(rng, n), orc.random_fr(rng, n), orc.random_fr(rng, logn + 3)
    got = scz.d_sumcheck_product(ctx, f, g, ch)
    assert np.array_equal(got, orc.d_sumcheck_product(orc.LEADER_SIM, 8, [f], [g], ch))


@pytest.mark.parametrize("logn", [0, 1, 3, 8, 9, 13, 16, 20])
def test_single_mle_sumcheck_family(orc, ctx, logn):
    """sumcheck / c_sumcheck / d_sumcheck (dsumcheck.rs:6-26, 92-146, 287-357), leader mode, against the oracle"""
    import scz_b200 as scz
    rng = np.random.default_rng(485 + logn)
    n = 1 << logn
    f, ch = orc.random_fr(rng, n), orc.random_fr(rng, logn + 3)
    assert np.array_equal(scz.sumcheck(ctx, f, ch), orc.sumcheck(f, ch))
    pp, opp = scz.PackedSharingParams(ctx, 1), orc.pp_new(1)
    assert np.array_equal(scz.c_sumcheck(ctx, pp, f, ch), orc.c_sumcheck(opp, orc.LEADER_SIM, [f], ch)[0])
    assert np.array_equal(scz.d_sumcheck(ctx, f, ch), orc.d_sumcheck(orc.LEADER_SIM, 8, [f], ch))


def test_d_acc_product_leader_mode(orc, ctx):
    import scz_b200 as scz
    rng = np.random.default_rng(490)
    x = orc.random_fr(rng, 1 << 9)
    sub, top = scz.d_acc_product(ctx, x)
    osub, otop = orc.d_acc_product(orc.LEADER_SIM, 8, [x])
    assert np.array_equal(sub, osub[0]) and np.array_equal(top, otop)


def test_commit_open_trapdoor_srs(orc, ctx):
    """dpoly_comm.rs:511-531 (should_commit_and_open) without pairings: with the real SRS new(g, _, s) the
    commitment is [p(s)] g; plus open() against the oracle, and the reference's length / level asserts."""
    import scz_b200 as scz
    rng = np.random.default_rng(500)
    nv = 6
    s = orc.random_fr(rng, nv)
    g = orc.g1_from_affine(orc.g1_generator())
    osrs = orc.Srs.new(g, s)
    pc = scz.PolynomialCommitment(ctx, [osrs.level(i)[:, :12] for i in range(osrs.levels)])
    p = orc.random_fr(rng, 1 << nv)
    com = pc.commit(p)
    assert orc.canon_g1(com) == orc.canon_g1(orc.commit(osrs, p))
    # p(s): the SRS orders variables so that fix_variable on s reproduces the evaluation
    ps = orc.fr_to_ints(orc.fix_variable(p, s))[0]
    exp = tw.g1_mul((tw.G1_X, tw.G1_Y), ps)
    assert orc.canon_g1(com) == [(exp[0], exp[1], 0)]
    u = orc.random_fr(rng, nv)
    val, proofs = pc.open(p, u)
    oval, oproofs = orc.open_(osrs, p, u)
    assert np.array_equal(val, oval) and orc.canon_g1(proofs) == orc.canon_g1(oproofs)
    assert np.array_equal(val, orc.fix_variable(p, u))
    with pytest.raises(scz.SczError) as e:
        pc.commit(p[:48])
    assert e.value.code == -3                      # assert!(peval.len() == 1 << level), dpoly_comm.rs:240
    with pytest.raises(scz.SczError) as e:
        pc.commit(np.concatenate([p, p]))
    assert e.value.code == -4                      # assert!(level < powers_of_g.len()), dpoly_comm.rs:239


@pytest.mark.parametrize("l", [1, 2])
def test_c_commit_c_open_leader_mode(orc, ctx_l, l):
    import scz_b200 as scz
    ctx = ctx_l(l)
    rng = np.random.default_rng(510 + l)
    pp, opp = scz.PackedSharingParams(ctx, l), orc.pp_new(l)
    nv = 7                                          # shares of 2^nv entries; level = log2(len * l)
    top = nv + (l.bit_length() - 1)
    dev, host = _levels(ctx, orc, rng, [max(1, (1 << i) // l) for i in range(top + 1)])   # new_single, :208-217
    pc, osrs = scz.PolynomialCommitment(ctx, dev), orc.Srs.from_levels(host)
    ps = [orc.random_fr(rng, 1 << nv), orc.random_fr(rng, 1 << (nv - 2)), orc.random_fr(rng, 1)]
    got = pc.c_commit(pp, ps)
    want = orc.c_commit([osrs], opp, orc.LEADER_SIM, [ps])[0]
    assert orc.canon_g1(got) == orc.canon_g1(want)
    u = orc.random_fr(rng, nv + 2)
    val, proofs = pc.c_open(pp, ps[0], u)
    oval, oproofs = orc.c_open([osrs], opp, orc.LEADER_SIM, [ps[0]], u)
    assert np.array_equal(val[0], oval[0]) and orc.canon_g1(proofs) == orc.canon_g1(oproofs[0])


def test_d_commit_d_open_leader_mode(orc, ctx):
    import scz_b200 as scz
    rng = np.random.default_rng(520)
    nv = 6
    dev, host = _levels(ctx, orc, rng, [1 << i for i in range(nv + 1)])                    # new_random, :222-231
    pc, osrs = scz.PolynomialCommitment(ctx, dev), orc.Srs.from_levels(host)
    p = orc.random_fr(rng, 1 << nv)
    up0, down0 = ctx.get_comm()
    assert orc.canon_g1(pc.d_commit(p)) == orc.canon_g1(orc.d_commit([osrs], orc.LEADER_SIM, 8, [p]))
    u = orc.random_fr(rng, nv + 3)
    val, proofs = pc.d_open(p, u)
    oval, oproofs = orc.d_open([osrs], orc.LEADER_SIM, 8, [p], u)
    assert np.array_equal(val, oval) and orc.canon_g1(proofs) == orc.canon_g1(oproofs)
    # get_comm() in the reference's serialised bytes, leader simulator (serializing_net.rs:158, :210), N = 8:
    #   d_commit  gather of one G1 (48 B) from 7 parties, scatter of one G1 to 7 parties       (dpoly_comm.rs:285-295)
    #   d_open    gather of (Fr, Vec<G1> of nv) = 32 + 8 + 48 nv, scatter of (0, []) = 32 + 8  (:368-391)
    assert ctx.get_comm() == (up0 + 7 * 48 + 7 * 40, down0 + 7 * 48 + 7 * (32 + 8 + 48 * nv))


def test_fixed_base_tables_same_results(orc, ctx):
    """scz_srs_precompute: commit / open / d_commit / d_open / c_open over levels of 2^0 .. 2^11 points (window sizes
    4 .. 11 bits, 24 .. 64 windows), an infinity base included, against the oracle and against the plain path"""
    import scz_b200 as scz
    rng = np.random.default_rng(525)
    nv = 11
    dev, host = _levels(ctx, orc, rng, [1 << i for i in range(nv + 1)])
    dev[5][3] = 0                                   # a point at infinity inside a level
    host[5][3] = 0
    host[5][3, 12] = 1
    pc, osrs = scz.PolynomialCommitment(ctx, dev).precompute(), orc.Srs.from_levels(host)
    plain = scz.PolynomialCommitment(ctx, dev)
    pp, opp = scz.PackedSharingParams(ctx, 1), orc.pp_new(1)
    for lv in (0, 1, 5, 8, nv):
        p = orc.random_fr(rng, 1 << lv)
        want = orc.canon_g1(orc.commit(osrs, p))
        assert orc.canon_g1(pc.commit(p)) == want == orc.canon_g1(plain.commit(p)), lv
    p = orc.random_fr(rng, 1 << nv)
    p[7] = 0                                        # zero scalar
    p[9] = orc.fr_from_ints([tw.R_MOD - 1])[0]      # largest scalar: every signed digit carries
    u = orc.random_fr(rng, nv + 3)
    val, proofs = pc.open(p, u)
    oval, oproofs = orc.open_(osrs, p, u)
    assert np.array_equal(val, oval) and orc.canon_g1(proofs) == orc.canon_g1(oproofs)
    assert orc.canon_g1(pc.d_commit(p)) == orc.canon_g1(orc.d_commit([osrs], orc.LEADER_SIM, 8, [p]))
    val, proofs = pc.d_open(p, u)
    oval, oproofs = orc.d_open([osrs], orc.LEADER_SIM, 8, [p], u)
    assert np.array_equal(val, oval) and orc.canon_g1(proofs) == orc.canon_g1(oproofs)
    val, proofs = pc.c_open(pp, p, u)
    oval, oproofs = orc.c_open([osrs], opp, orc.LEADER_SIM, [p], u)
    assert np.array_equal(val[0], oval[0]) and orc.canon_g1(proofs) == orc.canon_g1(oproofs[0])


# ------------------------------------------------------------------------------------------- parties mode
def test_parties_mode_local_test_net(orc):
    """N = 8 parties (l = 1), one ctx each, star collectives through LocalTestNet: d_msm, pss2ss,
    c_sumcheck_product, d_sumcheck_product, c_open, d_commit, d_open, d_acc_product -- each party's output
    against the oracle's N-party simulation (ORC_PARTIES)."""
    import scz_b200 as scz
    from scz_b200.net import LocalTestNet
    N, l, nv = 8, 1, 5
    rng = np.random.default_rng(530)
    opp = orc.pp_new(l)
    n = 1 << nv
    f = [orc.random_fr(rng, n) for _ in range(N)]
    g = [orc.random_fr(rng, n) for _ in range(N)]
    ch = orc.random_fr(rng, nv + 3)
    x1 = orc.random_fr(rng, N)
    seed_ctx = scz.Context(device=0, n_parties=N)
    srs_dev, srs_host = [], []
    for j in range(N):                               # every party draws its own random bases
        d, h = _levels(seed_ctx, orc, rng, [1 << i for i in range(nv + 1)])
        srs_dev.append([ctx_t.clone() for ctx_t in d])
        srs_host.append(orc.Srs.from_levels(h))
    msm_b = [seed_ctx.to_host(srs_dev[j][nv]) for j in range(N)]

    want_msm = orc.d_msm(opp, orc.PARTIES, [[oracle_affine(msm_b[j])] for j in range(N)], [[f[j]] for j in range(N)])
    want_p2s = orc.pss2ss(opp, orc.PARTIES, x1)
    want_cs = orc.c_sumcheck_product(opp, orc.PARTIES, f, g, ch)
    want_ds = orc.d_sumcheck_product(orc.PARTIES, N, f, g, ch)
    want_cs1 = orc.c_sumcheck(opp, orc.PARTIES, f, ch)
    want_ds1 = orc.d_sumcheck(orc.PARTIES, N, g, ch)
    want_co = orc.c_open(srs_host, opp, orc.PARTIES, f, ch)
    want_dc = orc.d_commit(srs_host, orc.PARTIES, N, g)
    want_do = orc.d_open(srs_host, orc.PARTIES, N, g, ch)
    want_sub, want_top = orc.d_acc_product(orc.PARTIES, N, f)

    def party(j, net):
        c = scz.Context(device=0, party_id=j, n_parties=N, net=net)
        pp = scz.PackedSharingParams(c, l)
        pc = scz.PolynomialCommitment(c, srs_dev[j])
        r = {}
        r["msm"] = scz.d_msm(c, pp, [msm_b[j]], [f[j]])
        r["p2s"] = scz.pss2ss(c, pp, x1[j:j + 1])
        r["cs"] = scz.c_sumcheck_product(c, pp, f[j], g[j], ch)
        r["ds"] = scz.d_sumcheck_product(c, f[j], g[j], ch)
        r["cs1"] = scz.c_sumcheck(c, pp, f[j], ch)
        r["ds1"] = scz.d_sumcheck(c, g[j], ch)
        r["co"] = pc.c_open(pp, f[j], ch)
        r["dc"] = pc.d_commit(g[j])
        r["do"] = pc.d_open(g[j], ch)
        r["acc"] = scz.d_acc_product(c, f[j])
        r["comm"] = c.get_comm()
        c.sync()
        c.close()
        return r

    res = LocalTestNet(N, "cuda:0").simulate_network_round(party)
    for j in range(N):
        r = res[j]
        assert orc.canon_g1(r["msm"]) == orc.canon_g1(want_msm[j]), j
        assert np.array_equal(r["p2s"], want_p2s[j]), j
        assert np.array_equal(r["cs"], want_cs[j]), j
        assert np.array_equal(r["cs1"], want_cs1[j]), j
        assert np.array_equal(r["ds1"], want_ds1) if j == 0 else len(r["ds1"]) == 0, j
        if j == 0:
            assert np.array_equal(r["ds"], want_ds)
            assert np.array_equal(r["do"][0], want_do[0]) and orc.canon_g1(r["do"][1]) == orc.canon_g1(want_do[1])
            assert np.array_equal(r["acc"][1], want_top)
        else:
            assert len(r["ds"]) == 0 and len(r["do"][1]) == 0 and not r["do"][0].any()      # dsumcheck.rs:507-509, dpoly_comm.rs:387
            assert r["acc"][1] is None
        assert np.array_equal(r["co"][0][0], want_co[0][j]) and orc.canon_g1(r["co"][1]) == orc.canon_g1(want_co[1][j]), j
        assert orc.canon_g1(r["dc"]) == orc.canon_g1(want_dc), j
        assert np.array_equal(r["acc"][0], want_sub[j]), j
    # star topology byte accounting: workers upload what the leader downloads
    assert sum(res[j]["comm"][0] for j in range(1, N)) == res[0]["comm"][1]
    assert sum(res[j]["comm"][1] for j in range(1, N)) == res[0]["comm"][0]
    seed_ctx.close()


@pytest.mark.parametrize("l", [1, 2])
def test_real_srs_and_packed_srs_end_to_end(orc, ctx_l, l):
    """PolynomialCommitmentCub::new / to_packed (dpoly_comm.rs:37-67, 164-194) on the device, and the collaborative
    pipeline's MEANING for N = 8 l parties: with the SRS and the polynomial both PSS-packed, the per-party MSMs unpack2
    to l points whose sum is the plain commitment [p(s)] g (the reference's pack_unpack2_test, dmsm.rs:104-138, at the
    protocol level), and the full c_commit protocol run by N parties hands every party a share that unpacks to it."""
    import torch
    import scz_b200 as scz
    from scz_b200.api import msm_batched
    from scz_b200.net import LocalTestNet
    ctx = ctx_l(l)
    N, nv = 8 * l, 5
    rng = np.random.default_rng(560 + l)
    s = orc.random_fr(rng, nv)
    g = orc.g1_from_affine(orc.g1_generator())
    osrs = orc.Srs.new(g, s)
    srs = scz.PolynomialCommitment.new(ctx, g, s)
    for i in range(nv + 1):
        assert np.array_equal(ctx.to_host(srs.level(i)), osrs.level(i)[:, :12]), i      # bit-exact affine points
    pp, opp = scz.PackedSharingParams(ctx, l), orc.pp_new(l)
    p = orc.random_fr(rng, 1 << nv)
    want = orc.canon_g1(orc.commit(osrs, p))
    assert orc.canon_g1(srs.commit(p)) == want
    # PSS-pack the polynomial in chunks of l (what a delegator does, examples/poly_comm.rs:52-55 uses random shares)
    shares = np.stack([orc.pack_from_public(opp, p[c:c + l]) for c in range(0, len(p), l)], axis=1)   # (N, 2^nv / l, 4)
    packed = [srs.to_packed(pp, j) for j in range(N)]
    local = torch.cat([msm_batched(ctx, [packed[j].level(nv)], [ctx.to_device(shares[j], 4)]) for j in range(N)])
    secrets = pp.unpack2(local, kind="g1").reshape(l, 18)
    total = secrets[0:1].contiguous()
    for k in range(1, l):
        total = ctx.g1_add(total, secrets[k:k + 1].contiguous())
    assert orc.canon_g1(ctx.to_host(total)) == want
    # the protocol itself, N parties in one process
    levels = [[packed[j].level(i) for i in range(nv + 1)] for j in range(N)]

    def party(j, net):
        c = scz.Context(device=0, party_id=j, n_parties=N, net=net)
        ppj = scz.PackedSharingParams(c, l)
        out = scz.PolynomialCommitment(c, levels[j]).c_commit(ppj, [shares[j]])
        c.sync()
        c.close()
        return out

    res = LocalTestNet(N, "cuda:0").simulate_network_round(party)
    got = pp.unpack(ctx.to_device(np.concatenate(res), 18), kind="g1").reshape(l, 18)
    for k in range(l):
        assert orc.canon_g1(ctx.to_host(got[k:k + 1].contiguous())) == want, k
