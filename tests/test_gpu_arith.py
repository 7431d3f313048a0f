"""-m gpu: device field / group arithmetic through the C ABI, bit-exact against the oracle."""
import numpy as np
import pytest
import torch

from oracle import py_twin as tw
from tests.gpu_util import oracle_affine, packed_affine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import scz_b200 as scz
    c = scz.Context(device=0, n_parties=8)
    yield c
    c.close()


def _edge_fr(orc):
    m = tw.R_MOD
    return orc.fr_from_ints([0, 1, 2, m - 1, m - 2, (m - 1) // 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1, 1 << 128])


def test_fr_ops(orc, ctx):
    rng = np.random.default_rng(200)
    e = _edge_fr(orc)
    a = np.concatenate([orc.random_fr(rng, 1 << 16), np.repeat(e, len(e), axis=0)])
    b = np.concatenate([orc.random_fr(rng, 1 << 16), np.tile(e, (len(e), 1))])
    da, db = ctx.to_device(a, 4), ctx.to_device(b, 4)
    for op, ref in (("mul", orc.fr_mul), ("add", orc.fr_add), ("sub", orc.fr_sub)):
        assert np.array_equal(ctx.to_host(ctx.fr_op(op, da, db)), ref(a, b)), op
    assert np.array_equal(ctx.to_host(ctx.fr_inv(da[:4096])), orc.fr_inv(a[:4096]))
    can = ctx.to_host(ctx.fr_to_canonical(da))
    assert [orc.limbs_to_int(r) for r in can[:512]] == orc.fr_to_ints(a[:512])
    assert np.array_equal(ctx.to_host(ctx.fr_from_canonical(ctx.to_device(can, 4))), a)


def test_fq_ops(orc, ctx):
    rng = np.random.default_rng(201)
    m = tw.P_MOD
    n = 1 << 14
    a = orc.fq_from_ints([int.from_bytes(rng.bytes(64), "little") for _ in range(n)] + [0, 1, m - 1, m - 2, 2, (1 << 380)])
    b = orc.fq_from_ints([int.from_bytes(rng.bytes(64), "little") for _ in range(n)] + [m - 1, m - 1, m - 1, 1, m - 2, (1 << 380) + 5])
    da, db = ctx.to_device(a, 6), ctx.to_device(b, 6)
    for op, ref in (("mul", orc.fq_mul), ("add", orc.fq_add), ("sub", orc.fq_sub)):
        assert np.array_equal(ctx.to_host(ctx.fq_op(op, da, db)), ref(a, b)), op


def _inf_jac(orc):
    inf = np.zeros((1, 18), dtype=np.uint64)
    inf[0, 0:6] = orc.fq_from_ints([1])[0]
    inf[0, 6:12] = orc.fq_from_ints([1])[0]
    return inf


def test_g1_ops_with_exceptional_cases(orc, ctx):
    rng = np.random.default_rng(202)
    n = 200
    pa, pb = orc.random_g1(rng, n), orc.random_g1(rng, n)
    ja = orc.g1_double(orc.g1_from_affine(pa))                         # Z != 1
    jb = orc.g1_add(orc.g1_from_affine(pb), orc.g1_from_affine(pa))
    inf = _inf_jac(orc)
    # mixed add: generic, acc == P (doubling), acc == -P (identity), acc identity, P identity
    acc = np.concatenate([ja, orc.g1_from_affine(pb[:1]), orc.g1_from_affine(pb[1:2]), inf, ja[:1]])
    aff = np.concatenate([pb, pb[:1], pb[1:2], pb[2:3], np.zeros((1, 13), dtype=np.uint64)])
    aff[-1, 12] = 1
    neg = np.zeros(len(acc), dtype=np.uint8)
    neg[n + 1] = 1
    neg[3] = 1
    aff_ref = aff.copy()
    for i in np.nonzero(neg)[0]:
        aff_ref[i, 6:12] = orc.fq_sub(np.zeros((1, 6), dtype=np.uint64), aff_ref[i:i + 1, 6:12])[0]
    got = ctx.g1_add_affine(ctx.to_device(acc, 18), ctx.to_device(packed_affine(aff), 12),
                            torch.from_numpy(neg).to(ctx.device))
    want = orc.canon_g1(orc.g1_add_mixed(acc, aff_ref))
    assert orc.canon_g1(ctx.to_host(got)) == want
    assert want[n] != (0, 0, 1) and want[n + 1] == (0, 0, 1)
    # full add incl. P+P, P+(-P), identities; double
    negja = ja[:1].copy()
    negja[0, 6:12] = orc.fq_sub(np.zeros((1, 6), dtype=np.uint64), ja[:1, 6:12])[0]
    A = np.concatenate([ja, ja[:1], ja[:1], inf, ja[:1], inf])
    B = np.concatenate([jb, ja[:1], negja, jb[:1], inf, inf])
    dA, dB = ctx.to_device(A, 18), ctx.to_device(B, 18)
    assert orc.canon_g1(ctx.to_host(ctx.g1_add(dA, dB))) == orc.canon_g1(orc.g1_add(A, B))
    assert orc.canon_g1(ctx.to_host(ctx.g1_double(dA))) == orc.canon_g1(orc.g1_double(A))
    # scalar multiplication and affine normalisation
    k = orc.random_fr(rng, 16)
    k[0] = 0
    k[1] = orc.fr_from_ints([1])[0]
    k[2] = orc.fr_from_ints([tw.R_MOD - 1])[0]
    got = ctx.g1_mul(ctx.to_device(ja[:16], 18), ctx.to_device(k, 4))
    assert orc.canon_g1(ctx.to_host(got)) == orc.canon_g1(orc.g1_mul(ja[:16], k))
    aff_dev = ctx.to_host(ctx.g1_to_affine(dA))
    ref_aff = orc.g1_to_affine(A)
    assert np.array_equal(oracle_affine(aff_dev)[:, :12], packed_affine(ref_aff))
    assert np.array_equal(oracle_affine(aff_dev)[:, 12] != 0, (ref_aff[:, 12] & np.uint64(0xFFFFFFFF)) != 0)


def test_generator_mul(orc, ctx):
    rng = np.random.default_rng(203)
    k = orc.random_fr(rng, 64)
    k[0] = 0
    k[1] = orc.fr_from_ints([1])[0]
    got = ctx.to_host(ctx.g1_generator_mul(ctx.to_device(k, 4)))
    assert np.array_equal(got, packed_affine(orc.g1_gen_mul(k)))
    assert not got[0].any()


def test_g1_compressed_wire_format(orc):
    """scz_g1_serialize_compressed_dev / _deserialize_ against the big-int twin and the public golden vectors
    (generator = 97f1d3a7...c6bb, -G = b7..., infinity = c0 00..00); malformed and wrong-subgroup encodings are flagged"""
    import torch
    import scz_b200 as scz
    from oracle import py_twin as tw
    ctx = scz.Context(device=0, n_parties=8)
    rng = np.random.default_rng(31)
    k = orc.random_fr(rng, 40)
    aff = ctx.g1_generator_mul(ctx.to_device(k, 4))                       # random points, affine
    jac = ctx.g1_mul(_affine_to_jac(ctx, aff), ctx.to_device(orc.random_fr(rng, 40), 4))   # non-trivial Z
    gen = orc.g1_from_affine(orc.g1_generator())
    g13 = orc.g1_generator()
    neg = np.concatenate([g13[:, :6], orc.fq_sub(np.zeros((1, 6), dtype=np.uint64), g13[:, 6:12]), np.zeros((1, 1), dtype=np.uint64)], axis=1)
    special = np.concatenate([gen, orc.g1_from_affine(neg), np.zeros((1, 18), dtype=np.uint64)])
    special[2, 0] = 1                                                        # (1, *, 0): infinity whatever X, Y
    allj = torch.cat([jac, ctx.to_device(special, 18)])
    ser = ctx.g1_serialize_compressed(allj)
    host = ser.cpu().numpy()
    canon = orc.canon_g1(ctx.to_host(allj))
    for i, (x, y, inf) in enumerate(canon):
        want = tw.g1_serialize_compressed(tw.INF if inf else (x, y))
        assert bytes(host[i]) == want, i
    gen_hex = "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    assert bytes(host[40]).hex() == gen_hex and bytes(host[41]).hex() == "b7" + gen_hex[2:]
    assert bytes(host[42]).hex() == "c0" + "00" * 47
    back, st = ctx.g1_deserialize_compressed(ser)
    assert not st.any().item()
    assert orc.canon_g1(ctx.to_host(back)) == canon
    # malformed inputs
    bad = np.zeros((4, 48), dtype=np.uint8)
    bad[1] = np.frombuffer(bytes([0xC0]) + bytes(46) + b"\x01", dtype=np.uint8)
    bad[2] = np.frombuffer(bytes([0x9F]) + b"\xff" * 47, dtype=np.uint8)
    x = 0
    while True:
        rhs = (x ** 3 + 4) % tw.P_MOD
        y = pow(rhs, (tw.P_MOD + 1) // 4, tw.P_MOD)
        if y * y % tw.P_MOD == rhs:
            break
        x += 1
    bad[3] = np.frombuffer(tw.g1_serialize_compressed((x, y)), dtype=np.uint8)
    _, st = ctx.g1_deserialize_compressed(torch.from_numpy(bad).to("cuda"))
    assert st.cpu().tolist() == [1, 1, 1, 2]
    # Fr: canonical little-endian bytes; values >= r are rejected
    fr = orc.random_fr(rng, 50)
    can = ctx.fr_to_canonical(ctx.to_device(fr, 4))
    assert [orc.limbs_to_int(r) for r in ctx.to_host(can)] == orc.fr_to_ints(fr)
    back, st = ctx.fr_deserialize(can)
    assert np.array_equal(ctx.to_host(back), fr) and not st.any().item()
    over = orc.ints_to_arr([tw.R_MOD, tw.R_MOD + 5, (1 << 256) - 1, tw.R_MOD - 1], 4)
    _, st = ctx.fr_deserialize(ctx.to_device(over, 4))
    assert st.cpu().tolist() == [1, 1, 1, 0]
    ctx.close()


def _affine_to_jac(ctx, aff):
    import torch
    one = torch.tensor(np.array([0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745,
                                 0x5c071a97a256ec6d, 0x15f65ec3fa80e493], dtype=np.uint64).view(np.int64), device=aff.device)
    return torch.cat([aff, one.repeat(len(aff), 1)], dim=1).contiguous()
