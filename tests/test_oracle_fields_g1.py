"""Pins the C oracle (oracle/liboracle.so) against the Python big-int twin and
public BLS12-381 constants.  CPU only."""
import numpy as np

from oracle import py_twin as tw


def _rand_ints(rng, n, mod):
    return [int.from_bytes(rng.bytes(64), "little") % mod for _ in range(n)]


def test_montgomery_constants(orc):
    import ctypes as C
    L = orc.lib()
    r1 = orc.limbs_to_int(np.ctypeslib.as_array((C.c_uint64 * 4).in_dll(L, "fr_R1")))
    assert r1 == (1 << 256) % tw.R_MOD == 0x1824b159acc5056f998c4fefecbc4ff55884b7fa0003480200000001fffffffe
    r2 = orc.limbs_to_int(np.ctypeslib.as_array((C.c_uint64 * 4).in_dll(L, "fr_R2")))
    assert r2 == pow(2, 512, tw.R_MOD) == 0x0748d9d99f59ff1105d314967254398f2b6cedcb87925c23c999e990f3f29c6d
    assert C.c_uint64.in_dll(L, "fr_INV").value == (-pow(tw.R_MOD, -1, 1 << 64)) % (1 << 64) == 0xfffffffeffffffff
    assert C.c_uint64.in_dll(L, "fq_INV").value == (-pow(tw.P_MOD, -1, 1 << 64)) % (1 << 64) == 0x89f3fffcfffcfffd
    q1 = orc.limbs_to_int(np.ctypeslib.as_array((C.c_uint64 * 6).in_dll(L, "fq_R1")))
    assert q1 == (1 << 384) % tw.P_MOD


def test_field_ops_vs_bigint(orc):
    rng = np.random.default_rng(1)
    for mod, frm, to, mul, add, sub in (
        (tw.R_MOD, orc.fr_from_ints, orc.fr_to_ints, orc.fr_mul, orc.fr_add, orc.fr_sub),
        (tw.P_MOD, orc.fq_from_ints, orc.fq_to_ints, orc.fq_mul, orc.fq_add, orc.fq_sub),
    ):
        a = _rand_ints(rng, 200, mod) + [0, 1, mod - 1, mod - 2, 2]
        b = _rand_ints(rng, 200, mod) + [mod - 1, mod - 1, mod - 1, 0, mod - 2]
        A, B = frm(a), frm(b)
        assert to(A) == a
        assert to(mul(A, B)) == [x * y % mod for x, y in zip(a, b)]
        assert to(add(A, B)) == [(x + y) % mod for x, y in zip(a, b)]
        assert to(sub(A, B)) == [(x - y) % mod for x, y in zip(a, b)]


def test_fr_inverse(orc):
    rng = np.random.default_rng(2)
    a = _rand_ints(rng, 50, tw.R_MOD) + [1, 2, tw.R_MOD - 1]
    inv = orc.fr_to_ints(orc.fr_inv(orc.fr_from_ints(a)))
    assert inv == [pow(x, -1, tw.R_MOD) for x in a]


def test_random_fr_is_reduced(orc):
    rng = np.random.default_rng(3)
    a = orc.random_fr(rng, 4096)
    assert all(orc.limbs_to_int(r) < tw.R_MOD for r in a)


def test_generator(orc):
    g = orc.g1_generator()
    assert orc.g1_on_curve(g) == [True]
    assert orc.fq_to_ints(g[:, 0:6]) == [tw.G1_X] and orc.fq_to_ints(g[:, 6:12]) == [tw.G1_Y]
    assert tw.on_curve((tw.G1_X, tw.G1_Y))
    # r * G = infinity
    import ctypes as C
    out = np.zeros((1, 18), dtype=np.uint64)
    gj = orc.g1_from_affine(g)
    k = np.array(orc.int_to_limbs(tw.R_MOD, 4), dtype=np.uint64)
    orc.lib().g1j_mul_bits(C.c_void_p(out.ctypes.data), C.c_void_p(gj.ctypes.data), C.c_void_p(k.ctypes.data), 4)
    assert orc.canon_g1(out) == [(0, 0, 1)]


def test_group_law_vs_twin(orc):
    rng = np.random.default_rng(4)
    ks = _rand_ints(rng, 12, tw.R_MOD) + [1, 2, tw.R_MOD - 1]
    G = (tw.G1_X, tw.G1_Y)
    pts = orc.g1_gen_mul(orc.fr_from_ints(ks))
    exp = [tw.g1_mul(G, k) for k in ks]
    jac = orc.g1_from_affine(pts)
    assert orc.canon_g1(jac) == [(p[0], p[1], 0) for p in exp]
    assert all(orc.g1_on_curve(pts))
    # add (incl. P+P, P+(-P)), mixed add, double
    a_idx = [0, 1, 2, 12, 12, 13, 3]
    b_idx = [1, 2, 2, 14, 12, 12, 3]     # 12: G, 13: 2G, 14: -G ; (2,2),(12,12),(3,3): doubling ; (12,14): G + (-G)
    A, B = jac[a_idx], jac[b_idx]
    want = [tw.g1_add(exp[i], exp[j]) for i, j in zip(a_idx, b_idx)]
    want = [(0, 0, 1) if w is None else (w[0], w[1], 0) for w in want]
    assert orc.canon_g1(orc.g1_add(A, B)) == want
    assert orc.canon_g1(orc.g1_add_mixed(A, pts[b_idx])) == want
    assert orc.canon_g1(orc.g1_double(jac[:5])) == [(p[0], p[1], 0) for p in (tw.g1_add(e, e) for e in exp[:5])]
    # identity handling
    inf = np.zeros((1, 18), dtype=np.uint64)
    inf[0, 0:6] = orc.fq_from_ints([1])[0]
    inf[0, 6:12] = orc.fq_from_ints([1])[0]
    assert orc.canon_g1(orc.g1_add(inf, jac[:1])) == orc.canon_g1(jac[:1])
    assert orc.canon_g1(orc.g1_add_mixed(inf, pts[:1])) == orc.canon_g1(jac[:1])
    assert orc.canon_g1(orc.g1_double(inf)) == [(0, 0, 1)]
    assert orc.g1_eq(orc.g1_add(A, B), orc.g1_add(B, A)) == [True] * len(a_idx)


def test_msm_ark_equals_naive_equals_twin(orc):
    rng = np.random.default_rng(5)
    G = (tw.G1_X, tw.G1_Y)
    for n in (1, 5, 31, 32, 100):
        ks = _rand_ints(rng, n, tw.R_MOD)
        ss = _rand_ints(rng, n, tw.R_MOD)
        if n >= 5:
            ss[0], ss[1], ss[2] = 0, 1, tw.R_MOD - 1
        bases = orc.g1_gen_mul(orc.fr_from_ints(ks))
        S = orc.fr_from_ints(ss)
        want = tw.g1_mul(G, sum(k * s for k, s in zip(ks, ss)) % tw.R_MOD)
        want = (0, 0, 1) if want is None else (want[0], want[1], 0)
        assert orc.canon_g1(orc.msm(bases, S, "naive")) == [want]
        assert orc.canon_g1(orc.msm(bases, S, "ark")) == [want]
        assert orc.canon_g1(orc.msm(bases, S, "ark", threads=4)) == [want]
    # degenerate distribution of dmsm.rs:92-138 (pack_unpack2_test): one point repeated, scalars all 1
    n = 256
    bases = np.repeat(orc.g1_gen_mul(orc.fr_from_ints([ks[0]])), n, axis=0)
    S = orc.fr_from_ints([1] * n)
    want = tw.g1_mul(G, ks[0] * n % tw.R_MOD)
    assert orc.canon_g1(orc.msm(bases, S, "ark")) == [(want[0], want[1], 0)]


def test_g1_compressed_wire_format_golden_vectors():
    """ark-bls12-381 0.4.0 serialises G1 in the Zcash / IETF compressed form.  Public known answers: the generator is
    97f1d3a7...c6bb (flag byte 0x80 | 0x17: its y is the smaller root), -G flips the sort bit (b7...), infinity is c0 00..00."""
    from oracle import py_twin as tw
    g = (tw.G1_X, tw.G1_Y)
    gen_hex = ("97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac58"
               "6c55e83ff97a1aeffb3af00adb22c6bb")
    assert tw.g1_serialize_compressed(g).hex() == gen_hex
    assert tw.g1_serialize_compressed(tw.g1_neg(g)).hex() == "b7" + gen_hex[2:]
    assert tw.g1_serialize_compressed(tw.INF).hex() == "c0" + "00" * 47
    for pt in (g, tw.g1_neg(g), tw.g1_mul(g, 0xDEADBEEF), tw.INF):
        back, st = tw.g1_deserialize_compressed(tw.g1_serialize_compressed(pt))
        assert st == 0 and back == pt
    assert tw.g1_deserialize_compressed(bytes(48))[1] == 1                                     # compression bit missing
    assert tw.g1_deserialize_compressed(bytes([0xC0]) + bytes(46) + b"\x01")[1] == 1           # infinity with payload
    assert tw.g1_deserialize_compressed(bytes([0x9F]) + b"\xff" * 47)[1] == 1                  # x >= p
    # x = 4: x^3 + 4 = 68 is a square mod p? find one on-curve x whose point is outside the r-torsion (cofactor != 1)
    x = 0
    while True:
        rhs = (x ** 3 + 4) % tw.P_MOD
        y = pow(rhs, (tw.P_MOD + 1) // 4, tw.P_MOD)
        if y * y % tw.P_MOD == rhs:
            break
        x += 1
    assert tw.g1_deserialize_compressed(tw.g1_serialize_compressed((x, y)))[1] == 2            # on the curve, wrong subgroup
