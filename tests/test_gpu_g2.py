"""-m gpu: MSM and d_msm over BLS12-381 G2 (csrc/g2.cuh, csrc/msm_g2.cu) through the C ABI.  `d_msm` is generic over
`G: CurveGroup` in the reference (dist-primitive/src/dmsm.rs:9-15); no reference caller uses G2, so the checker is the Python
big-int twin (oracle/py_twin.py: Fq2 and G2 from the public curve constants) plus the closed forms the G1 tests use:
msm(k_i G, s_i) = (sum k_i s_i) G, leader-mode d_msm = (4/7) MSM, and the leader closure as the Fr-linear map it is."""
import numpy as np
import pytest
import torch

from oracle import py_twin as tw

pytestmark = pytest.mark.gpu
G2 = (tw.G2_X, tw.G2_Y)


@pytest.fixture(scope="module")
def ctx():
    import scz_b200 as scz
    c = scz.Context(device=0, n_parties=8)
    yield c
    c.close()


def canon(orc, jac):
    """(n, 36) Jacobian Montgomery limbs -> list of affine ((x0, x1), (y0, y1)) or None"""
    j = np.ascontiguousarray(jac, dtype=np.uint64).reshape(-1, 36)
    out = []
    for row in j:
        v = orc.fq_to_ints(row.reshape(6, 6))
        X, Y, Z = (v[0], v[1]), (v[2], v[3]), (v[4], v[5])
        if Z == (0, 0):
            out.append(None)
            continue
        zi = tw.f2_inv(Z)
        zi2 = tw.f2_mul(zi, zi)
        out.append((tw.f2_mul(X, zi2), tw.f2_mul(Y, tw.f2_mul(zi2, zi))))
    return out


def affine_limbs(orc, pts):
    """list of affine points (or None) -> (n, 24) Montgomery limbs, identity = all zero"""
    vals = []
    for p in pts:
        vals += [0, 0, 0, 0] if p is None else [p[0][0], p[0][1], p[1][0], p[1][1]]
    return orc.fq_from_ints(vals).reshape(len(pts), 24)


def gen_mul(scz, ctx, k_fr):
    """k[i] * G2 on the device: (n, 36) Jacobian tensor"""
    n = len(k_fr)
    g = torch.from_numpy(scz.g2_affine_to_jac(np.repeat(scz.G2_GENERATOR_AFFINE, n, axis=0)).view(np.int64)).to(ctx.device)
    return scz.g2_op(ctx, "mul", g, ctx.to_device(k_fr, 4))


def to_affine_dev(orc, ctx, jac):
    """device Jacobian -> device affine (n, 24) through the host twin (test helper, small n)"""
    pts = canon(orc, ctx.to_host(jac))
    return ctx.to_device(affine_limbs(orc, pts).reshape(-1, 24), 24)


def test_generator_and_unit_ops(orc, ctx):
    import scz_b200 as scz
    assert tw.g2_on_curve(G2) and tw.g2_mul(G2, tw.R_MOD) is None
    rng = np.random.default_rng(900)
    ks = [1, 2, 3, tw.R_MOD - 1, 0] + [int(x) for x in rng.integers(1, 1 << 62, 6)]
    k = orc.fr_from_ints(ks)
    P = gen_mul(scz, ctx, k)
    assert canon(orc, ctx.to_host(P)) == [tw.g2_mul(G2, v) for v in ks]
    Q = P.roll(1, 0).contiguous()
    assert canon(orc, ctx.to_host(scz.g2_op(ctx, "add", P, Q))) == [tw.g2_add(tw.g2_mul(G2, a), tw.g2_mul(G2, b))
                                                                      for a, b in zip(ks, ks[-1:] + ks[:-1])]
    assert canon(orc, ctx.to_host(scz.g2_op(ctx, "double", P))) == [tw.g2_mul(G2, 2 * v) for v in ks]
    assert canon(orc, ctx.to_host(scz.g2_op(ctx, "add", P, P))) == [tw.g2_mul(G2, 2 * v) for v in ks]     # P + P -> doubling
    neg = scz.g2_op(ctx, "mul", P, ctx.to_device(orc.fr_from_ints([tw.R_MOD - 1] * len(ks)), 4))
    assert canon(orc, ctx.to_host(scz.g2_op(ctx, "add", P, neg))) == [None] * len(ks)                     # P + (-P)


@pytest.mark.parametrize("n", [0, 1, 2, 33, 300])
def test_msm_g2_small_against_the_twin(orc, ctx, n):
    import scz_b200 as scz
    rng = np.random.default_rng(910 + n)
    k = [int(x) for x in rng.integers(1, 1 << 40, n)]
    pts = [tw.g2_mul(G2, v) for v in k]
    s = orc.random_fr(rng, n)
    si = orc.fr_to_ints(s) if n else []
    got = scz.msm_g2(ctx, affine_limbs(orc, pts) if n else np.zeros((0, 24), dtype=np.uint64), s)     # host path
    want = tw.g2_mul(G2, sum(a * b for a, b in zip(k, si)) % tw.R_MOD) if n else None
    assert canon(orc, got) == [want]
    if n:
        dev = scz.msm_g2(ctx, ctx.to_device(affine_limbs(orc, pts), 24), ctx.to_device(s, 4))           # device path
        assert canon(orc, ctx.to_host(dev)) == [want]


def test_msm_g2_degenerate_inputs(orc, ctx):
    """dmsm.rs:97-104 shape: one point many times with all-one scalars (every bucket addition is a doubling), an identity
    base through the mask and through the encoding, zero scalars, P / -P cancellation, every window size"""
    import scz_b200 as scz
    rng = np.random.default_rng(920)
    P = tw.g2_mul(G2, 77)
    M = 500
    one = orc.fr_from_ints([1])
    got = scz.msm_g2(ctx, affine_limbs(orc, [P] * M), np.repeat(one, M, axis=0))
    assert canon(orc, got) == [tw.g2_mul(G2, 77 * M)]
    k = [int(x) for x in rng.integers(1, 1 << 30, 40)]
    pts = [tw.g2_mul(G2, v) for v in k]
    s = orc.random_fr(rng, 40)
    s[9] = 0
    s[3] = orc.fr_from_ints([tw.R_MOD - 1])[0]
    si = orc.fr_to_ints(s)
    mask = np.zeros(40, dtype=np.uint8)
    mask[5] = 1
    want = tw.g2_mul(G2, sum(a * b for j, (a, b) in enumerate(zip(k, si)) if j != 5) % tw.R_MOD)
    assert canon(orc, scz.msm_g2(ctx, affine_limbs(orc, pts), s, inf_mask=mask)) == [want]
    pts2 = list(pts)
    pts2[5] = None
    assert canon(orc, scz.msm_g2(ctx, affine_limbs(orc, pts2), s)) == [want]
    for c in (1, 3, 8, 12, 16):
        ctx.msm_set_window(c)
        try:
            assert canon(orc, scz.msm_g2(ctx, affine_limbs(orc, pts2), s)) == [want], c
        finally:
            ctx.msm_set_window(0)
    negP = (P[0], tw.f2_sub((0, 0), P[1]))
    assert canon(orc, scz.msm_g2(ctx, affine_limbs(orc, [P, negP]), np.repeat(s[:1], 2, axis=0))) == [None]
    with pytest.raises(scz.SczError) as e:
        scz.msm_g2(ctx, affine_limbs(orc, pts), s[:39])
    assert e.value.code == -2


def test_msm_g2_2p14_trapdoor_batch_and_linearity(orc, ctx):
    """bases k_i G2 made on the device (512 distinct points, tiled to 2^14): msm = (sum k_i s_i) G2; linearity; a ragged batch"""
    import scz_b200 as scz
    from tests.gpu_util import fr_dot
    rng = np.random.default_rng(930)
    sub, reps = 512, 32
    k = orc.random_fr(rng, sub)
    aff = to_affine_dev(orc, ctx, gen_mul(scz, ctx, k))                      # (512, 24) affine on the device
    n = sub * reps
    big = aff.repeat(reps, 1).contiguous()
    kk = np.tile(k, (reps, 1))
    s = orc.random_fr(rng, n)
    r1 = scz.msm_g2(ctx, big, ctx.to_device(s, 4))
    assert canon(orc, ctx.to_host(r1)) == [tw.g2_mul(G2, orc.fr_to_ints(fr_dot(orc, kk, s))[0])]
    s2 = orc.random_fr(rng, n)
    r2 = scz.msm_g2(ctx, big, ctx.to_device(s2, 4))
    r12 = scz.msm_g2(ctx, big, ctx.to_device(orc.fr_add(s, s2), 4))
    assert canon(orc, ctx.to_host(scz.g2_op(ctx, "add", r1, r2))) == canon(orc, ctx.to_host(r12))
    lens = [128, 64, 1, 0, 63]
    off, bl, sl, exp = 0, [], [], []
    for ln in lens:
        bl.append(aff[off:off + ln].contiguous())
        sl.append(ctx.to_device(s[off:off + ln], 4))
        exp.append(tw.g2_mul(G2, orc.fr_to_ints(fr_dot(orc, k[off:off + ln], s[off:off + ln]))[0]) if ln else None)
        off += ln
    assert canon(orc, ctx.to_host(scz.msm_g2_batched(ctx, bl, sl))) == exp


def test_d_msm_g2_leader_mode_closed_form(orc, ctx):
    """leader simulator, l = 1: out = lambda_0 * MSM = (4/7) * (sum k_i s_i) * G2; comm counters in the
    reference's serialised sizes: Vec<G2> of one element = 8 + 96 B to / from 7 peers"""
    import scz_b200 as scz
    from tests.gpu_util import fr_dot
    rng = np.random.default_rng(940)
    n = 200
    k = orc.random_fr(rng, n)
    aff = to_affine_dev(orc, ctx, gen_mul(scz, ctx, k))
    s = orc.random_fr(rng, n)
    pp = scz.PackedSharingParams(ctx, 1)
    up0, down0 = ctx.get_comm()
    got = scz.d_msm_g2(ctx, pp, [aff], [ctx.to_device(s, 4)])
    dot = orc.fr_to_ints(fr_dot(orc, k, s))[0]
    assert canon(orc, ctx.to_host(got)) == [tw.g2_mul(G2, dot * tw.LAMBDA0 % tw.R_MOD)]
    assert ctx.get_comm() == (up0 + 7 * 104, down0 + 7 * 104)


@pytest.mark.parametrize("l", [1, 2])
def test_d_msm_g2_leader_closure_is_the_pss_map(orc, ctx, l):
    """dmsm.rs:31-38 with N real parties: the closure (unpack2 -> sum of the l secrets -> replicate -> pack) is Fr-linear,
    so on inputs k_j * G2 its outputs are map(k)_o * G2 with the map evaluated by the oracle's Fr PSS (pss.rs:93-171)"""
    import scz_b200 as scz
    rng = np.random.default_rng(950 + l)
    pp, opp = scz.PackedSharingParams(ctx, l), orc.pp_new(l)
    n, batch = 8 * l, 2
    k = orc.random_fr(rng, n * batch).reshape(n, batch, 4)
    pts = ctx.to_host(gen_mul(scz, ctx, k.reshape(-1, 4))).reshape(n, batch, 36)      # [party][entry]
    got = scz.d_msm_g2_leader(ctx, pp, pts)
    for e in range(batch):
        sec = orc.unpack2(opp, k[:, e, :], kind=0)
        tot = sec[0:1]
        for i in range(1, l):
            tot = orc.fr_add(tot, sec[i:i + 1])
        shares = orc.pack_from_public(opp, np.repeat(tot, l, axis=0), kind=0)
        want = [tw.g2_mul(G2, v) for v in orc.fr_to_ints(shares)]
        assert canon(orc, got[:, e, :]) == want, (l, e)
