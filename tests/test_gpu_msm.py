"""-m gpu: Pippenger MSM and d_msm through the C ABI against the oracle.

Configs (BASELINE.json): #1 leader-mode d_msm, 2^16 bases, bit-exact vs the CPU restatement of the
arkworks path and vs the closed form (4/7)*MSM; #2 d_msm at 2^20 bases checked through the
size-independent trapdoor property  msm(k_i*G, s_i) == (sum k_i s_i)*G  and linearity."""
import numpy as np
import pytest
import torch

from oracle import py_twin as tw
from tests.gpu_util import fr_dot, packed_affine

pytestmark = pytest.mark.gpu


# every test of this file runs with both bucket accumulation variants: "xyzz" = mixed additions only (the path small
# sequences take by default), "affine" = the batched-affine levels forced on (csrc/msm_affine.cu: 3 levels, slabs of
# 4096 entries so that several slabs and cut runs occur even in small inputs)
@pytest.fixture(scope="module", params=["xyzz", "affine"])
def ctx(request):
    import scz_b200 as scz
    c = scz.Context(device=0, n_parties=8)
    if request.param == "affine":
        c.msm_set_affine(1, 3, 4096)
    else:
        c.msm_set_affine(2)
    c.variant = request.param
    yield c
    if request.param == "affine":
        assert c.msm_affine_sequences() > 0
    else:
        assert c.msm_affine_sequences() == 0
    c.close()


@pytest.fixture(scope="module")
def pool(orc):
    """2^12 random bases/scalars shared by the small cases"""
    rng = np.random.default_rng(300)
    n = 1 << 12
    return orc.random_g1(rng, n), orc.random_fr(rng, n)


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 1000, 4096])
def test_msm_matches_oracle(orc, ctx, pool, n):
    import scz_b200 as scz
    bases, scalars = pool[0][:n], pool[1][:n]
    got = scz.msm(ctx, packed_affine(bases), scalars)                       # host path
    want = orc.msm(bases, scalars, "ark") if n else None
    if n == 0:
        assert orc.canon_g1(got) == [(0, 0, 1)]
    else:
        assert orc.canon_g1(got) == orc.canon_g1(want)
    if n in (33, 1000):
        assert orc.canon_g1(got) == orc.canon_g1(orc.msm(bases, scalars, "naive"))
    dev = scz.msm(ctx, ctx.to_device(packed_affine(bases), 12), ctx.to_device(scalars, 4))   # device path
    assert orc.canon_g1(ctx.to_host(dev)) == orc.canon_g1(got)


@pytest.mark.parametrize("c", [1, 2, 3, 5, 8, 11, 13, 15, 16])
def test_every_window_size_gives_the_same_point(orc, ctx, pool, c):
    import scz_b200 as scz
    n = 700
    bases, scalars = pool[0][:n], pool[1][:n].copy()
    scalars[0] = 0
    scalars[1] = orc.fr_from_ints([1])[0]
    scalars[2] = orc.fr_from_ints([tw.R_MOD - 1])[0]
    scalars[3] = orc.fr_from_ints([(1 << 254) + 12345])[0]
    want = orc.canon_g1(orc.msm(bases, scalars, "ark"))
    ctx.msm_set_window(c)
    try:
        got = scz.msm(ctx, packed_affine(bases), scalars)
    finally:
        ctx.msm_set_window(0)
    assert orc.canon_g1(got) == want


def test_msm_degenerate_inputs(orc, ctx, pool):
    """dmsm.rs:97-104: one point repeated M times, scalars all one (every bucket add hits P+P);
    plus infinity bases (mask and x=y=0), zero scalars, and P / -P cancellation."""
    import scz_b200 as scz
    M = 5000
    one = orc.fr_from_ints([1])
    bases = np.repeat(pool[0][:1], M, axis=0)
    scalars = np.repeat(one, M, axis=0)
    got = scz.msm(ctx, packed_affine(bases), scalars)
    want = orc.g1_mul(orc.g1_from_affine(pool[0][:1]), orc.fr_from_ints([M]))
    assert orc.canon_g1(got) == orc.canon_g1(want)
    # same scalar everywhere (one heavy bucket per window)
    n2 = 4096
    scalars = np.repeat(pool[1][:1], n2, axis=0)
    got = scz.msm(ctx, packed_affine(pool[0][:n2]), scalars)
    want = orc.msm(pool[0][:n2], scalars, "ark")
    assert orc.canon_g1(got) == orc.canon_g1(want)
    # infinity through the mask and through the encoding, zero scalars
    n = 300
    b = pool[0][:n].copy()
    s = pool[1][:n].copy()
    mask = np.zeros(n, dtype=np.uint8)
    mask[5] = mask[77] = 1
    b_ref = b.copy()
    b_ref[5, 12] = b_ref[77, 12] = 1
    s[9] = 0
    got = scz.msm(ctx, b[:, :12].copy(), s, inf_mask=mask)
    assert orc.canon_g1(got) == orc.canon_g1(orc.msm(b_ref, s, "ark"))
    got2 = scz.msm(ctx, packed_affine(b_ref), s)
    assert orc.canon_g1(got2) == orc.canon_g1(got)
    # P and -P with the same scalar cancel to the identity
    pm = np.concatenate([pool[0][:1], pool[0][:1]])
    pm[1, 6:12] = orc.fq_sub(np.zeros((1, 6), dtype=np.uint64), pm[1:2, 6:12])[0]
    got = scz.msm(ctx, packed_affine(pm), np.repeat(pool[1][:1], 2, axis=0))
    assert orc.canon_g1(got) == [(0, 0, 1)]


def test_msm_length_mismatch_is_an_error(ctx, pool):
    import scz_b200 as scz
    with pytest.raises(scz.SczError) as e:
        scz.msm(ctx, packed_affine(pool[0][:10]), pool[1][:9])
    assert e.value.code == -2      # the reference panics via unwrap() (dmsm.rs:23)


def test_batched_msm_ragged(orc, ctx, pool):
    """the c_open shape (dpoly_comm.rs:436): halving lengths 2^11 .. 1, plus an empty entry"""
    from scz_b200.api import msm_batched
    lens = [1 << k for k in range(11, -1, -1)] + [0, 3]
    off, bl, sl, want = 0, [], [], []
    for ln in lens:
        b, s = pool[0][off % 1024: off % 1024 + ln], pool[1][off % 1024: off % 1024 + ln]
        bl.append(ctx.to_device(packed_affine(b), 12))
        sl.append(ctx.to_device(s, 4))
        want.append(orc.canon_g1(orc.msm(b, s, "ark"))[0] if ln else (0, 0, 1))
        off += 97
    got = msm_batched(ctx, bl, sl)
    assert orc.canon_g1(ctx.to_host(got)) == want


def test_config1_d_msm_leader_mode_2p16(orc, ctx):
    """BASELINE config 1: leader-mode d_msm, 2^16 bases, l = 1, bit-exact vs the CPU arkworks restatement"""
    import scz_b200 as scz
    rng = np.random.default_rng(301)
    n = 1 << 16
    k = orc.random_fr(rng, n)
    bases_dev = ctx.g1_generator_mul(ctx.to_device(k, 4))
    bases = ctx.to_host(bases_dev)
    scalars = orc.random_fr(rng, n)
    pp = scz.PackedSharingParams(ctx, 1)
    got = scz.d_msm(ctx, pp, [bases], [scalars])                                        # host path
    from tests.gpu_util import oracle_affine
    opp = orc.pp_new(1)
    want = orc.d_msm(opp, orc.LEADER_SIM, [[oracle_affine(bases)]], [[scalars]])
    assert orc.canon_g1(got) == orc.canon_g1(want[0])
    # closed form: (4/7) * (sum k_i s_i) * G
    dot = orc.fr_to_ints(fr_dot(orc, k, scalars))[0]
    G = (tw.G1_X, tw.G1_Y)
    exp = tw.g1_mul(G, dot * tw.LAMBDA0 % tw.R_MOD)
    assert orc.canon_g1(got) == [(exp[0], exp[1], 0)]
    up, down = ctx.get_comm()
    assert (up, down) == (7 * 56, 7 * 56)      # Vec<G1> of one element = 8 + 48 B, N - 1 = 7 peers


def test_config2_msm_2p20_trapdoor_and_linearity(orc, ctx):
    """BASELINE config 2 size: checked through properties that do not need a 2^20 CPU MSM"""
    import scz_b200 as scz
    rng = np.random.default_rng(302)
    n = 1 << 20
    k = orc.random_fr(rng, n)
    kd = ctx.to_device(k, 4)
    bases = ctx.g1_generator_mul(kd)
    s1, s2 = orc.random_fr(rng, n), orc.random_fr(rng, n)
    d1, d2 = ctx.to_device(s1, 4), ctx.to_device(s2, 4)
    G = (tw.G1_X, tw.G1_Y)
    r1 = scz.msm(ctx, bases, d1)
    exp = tw.g1_mul(G, orc.fr_to_ints(fr_dot(orc, k, s1))[0])
    assert orc.canon_g1(ctx.to_host(r1)) == [(exp[0], exp[1], 0)]
    r2 = scz.msm(ctx, bases, d2)
    r12 = scz.msm(ctx, bases, ctx.fr_op("add", d1, d2))
    assert orc.canon_g1(ctx.to_host(ctx.g1_add(r1, r2))) == orc.canon_g1(ctx.to_host(r12))
    st = ctx.msm_last_stats()
    assert st["bucket_adds"] == n * st["windows"]
    if ctx.variant == "affine":   # deep trees, one slab and the automatic choice: all the same point
        for levels, slab in ((6, 0), (1, 1 << 20), (0, 0)):
            ctx.msm_set_affine(1, levels, slab)
            assert orc.canon_g1(ctx.to_host(scz.msm(ctx, bases, d1))) == [(exp[0], exp[1], 0)], (levels, slab)
        # 8-bit windows: runs of ~8 k entries per bucket, 6 affine levels, every chunk boundary cuts a run
        ctx.msm_set_window(8)
        ctx.msm_set_affine(1, 6, 1 << 22)
        try:
            assert orc.canon_g1(ctx.to_host(scz.msm(ctx, bases, d1))) == [(exp[0], exp[1], 0)]
            # all scalars equal to one (dmsm.rs:103): ONE bucket holds the whole stream, every addition inside the
            # affine levels is of distinct points; then the same point 2^20 times: every addition is a doubling
            ones = ctx.to_device(np.repeat(orc.fr_from_ints([1]), n, axis=0), 4)
            ksum = orc.fr_to_ints(fr_dot(orc, k, np.repeat(orc.fr_from_ints([1]), n, axis=0)))[0]
            e1 = tw.g1_mul(G, ksum)
            assert orc.canon_g1(ctx.to_host(scz.msm(ctx, bases, ones))) == [(e1[0], e1[1], 0)]
            same = bases[:1].repeat(n, 1).contiguous()
            e2 = tw.g1_mul(G, orc.fr_to_ints(k[:1])[0] * n % tw.R_MOD)
            assert orc.canon_g1(ctx.to_host(scz.msm(ctx, same, ones))) == [(e2[0], e2[1], 0)]
        finally:
            ctx.msm_set_window(0)
            ctx.msm_set_affine(1, 3, 4096)


@pytest.mark.parametrize("l", [1, 2, 4, 8])
def test_d_msm_leader_closure_parties_mode(orc, ctx, l):
    """dmsm.rs:31-38 with N real parties: per batch entry unpack2 -> sum of the l secrets -> [sum; l] -> pack.
    The device runs the closure as the rank-one map it is (n scalar multiplications into the sum, n out of it, in
    two launches); the oracle runs the FFT pairs step by step."""
    import scz_b200 as scz
    rng = np.random.default_rng(310 + l)
    pp = scz.PackedSharingParams(ctx, l)
    opp = orc.pp_new(l)
    n, batch = 8 * l, 3
    pts = orc.g1_from_affine(orc.random_g1(rng, n * batch)).reshape(n, batch, 18)   # [party][k]
    got = scz.d_msm_leader(ctx, pp, pts)
    for k in range(batch):
        sec = orc.unpack2(opp, pts[:, k, :], kind=1)
        tot = sec[0:1]
        for i in range(1, l):
            tot = orc.g1_add(tot, sec[i:i + 1])
        want = orc.pack_from_public(opp, np.repeat(tot, l, axis=0), kind=1)
        assert orc.canon_g1(got[:, k, :]) == orc.canon_g1(want), (l, k)
