"""Host build of the DEVICE arithmetic headers (csrc/field.cuh, g1.cuh, g2.cuh, msm_digits.cuh)
checked against the oracle.  This exercises the limb/carry logic the CUDA kernels inline;
it is a CI aid for the GPU-less authoring box, not a product path (tests/emu/emu.cpp)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import py_twin as tw

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def emu():
    so = os.path.join(HERE, "emu", "libemu.so")
    src = os.path.join(HERE, "emu", "emu.cpp")
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", so, src], check=True)
    return C.CDLL(so)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _edge(orc, mod, frm, n):
    vals = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (1 << 32) - 1, 1 << 32, (1 << 64) - 1]
    return frm(vals)


def test_field_ops(orc, emu):
    rng = np.random.default_rng(100)
    for name, mod, cols, rnd, frm, omul, oadd, osub in (
        ("fr", tw.R_MOD, 4, lambda n: orc.random_fr(rng, n), orc.fr_from_ints, orc.fr_mul, orc.fr_add, orc.fr_sub),
        ("fq", tw.P_MOD, 6, lambda n: orc.fq_from_ints([int.from_bytes(rng.bytes(64), "little") for _ in range(n)]),
         orc.fq_from_ints, orc.fq_mul, orc.fq_add, orc.fq_sub),
    ):
        e = _edge(orc, mod, frm, 0)
        a = np.concatenate([rnd(3000), np.repeat(e, len(e), axis=0)])
        b = np.concatenate([rnd(3000), np.tile(e, (len(e), 1))])
        for op, ref in (("mul", omul), ("add", oadd), ("sub", osub)):
            out = np.zeros_like(a)
            getattr(emu, f"emu_{name}_{op}")(_p(a), _p(b), _p(out), C.c_size_t(len(a)))
            assert np.array_equal(out, ref(a, b)), (name, op)


def test_dot_wide(orc, emu):
    """sum a_i b_i through the unreduced (2N + 1)-limb accumulator == the sum of the Montgomery products"""
    rng = np.random.default_rng(105)
    for name, mod, rnd, frm, omul, oadd in (
        ("fr", tw.R_MOD, lambda n: orc.random_fr(rng, n), orc.fr_from_ints, orc.fr_mul, orc.fr_add),
        ("fq", tw.P_MOD, lambda n: orc.fq_from_ints([int.from_bytes(rng.bytes(64), "little") for _ in range(n)]),
         orc.fq_from_ints, orc.fq_mul, orc.fq_add),
    ):
        e = _edge(orc, mod, frm, 0)
        top = frm([mod - 1])
        cases = [
            (rnd(1), rnd(1)), (rnd(700), rnd(700)), (np.repeat(e, len(e), axis=0), np.tile(e, (len(e), 1))),
            (np.repeat(top, 5000, axis=0), np.repeat(top, 5000, axis=0)),   # the accumulator's 2N-th limb fills up
            (np.concatenate([np.repeat(top, 3000, axis=0), rnd(500)]), np.concatenate([np.repeat(top, 3000, axis=0), rnd(500)])),
        ]
        for a, b in cases:
            a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
            out = np.zeros_like(a[:1])
            getattr(emu, f"emu_{name}_dot_wide")(_p(a), _p(b), C.c_size_t(len(a)), _p(out))
            prod = omul(a, b)
            want = prod[:1]
            # tree-sum with the oracle's vector add
            acc = prod
            while len(acc) > 1:
                if len(acc) % 2:
                    acc = np.concatenate([acc, np.zeros_like(acc[:1])])
                acc = oadd(acc[0::2], acc[1::2])
            assert np.array_equal(out, acc), (name, len(a))


def test_fq_dot2_sub(orc, emu):
    """a*b - c*d with one shared Montgomery reduction (field.cuh fp_dot2_sub): the Y coordinate of every G1 formula"""
    rng = np.random.default_rng(102)
    rnd = lambda n: orc.fq_from_ints([int.from_bytes(rng.bytes(64), "little") for _ in range(n)])   # noqa: E731
    e = _edge(orc, tw.P_MOD, orc.fq_from_ints, 0)
    k = len(e)
    cols = [np.concatenate([rnd(2000), e[(np.arange(k ** 2) // (k ** s)) % k][: k * k]]) for s in (0, 1, 0, 1)]
    cols[2] = np.concatenate([rnd(2000), np.roll(e, 3, axis=0)[np.arange(k * k) % k]])
    cols[3] = np.concatenate([rnd(2000), np.roll(e, 5, axis=0)[(np.arange(k * k) // k) % k]])
    a, b, c, d = cols
    out = np.zeros_like(a)
    emu.emu_fq_dot2_sub(_p(a), _p(b), _p(c), _p(d), _p(out), C.c_size_t(len(a)))
    assert np.array_equal(out, orc.fq_sub(orc.fq_mul(a, b), orc.fq_mul(c, d)))
    # worst case for the accumulator bound: all four operands p - 1
    m = orc.fq_from_ints([tw.P_MOD - 1] * 4)
    o1 = np.zeros_like(m)
    emu.emu_fq_dot2_sub(_p(m), _p(m), _p(orc.fq_from_ints([1] * 4)), _p(m), _p(o1), C.c_size_t(4))
    assert np.array_equal(o1, orc.fq_sub(orc.fq_mul(m, m), orc.fq_mul(orc.fq_from_ints([1] * 4), m)))


def test_fq_mul_f64(orc, emu):
    """Fq product on 24-bit double limbs (fq_f64.cuh): same residues as the integer multiplier, both m-chain variants"""
    rng = np.random.default_rng(104)
    rnd = lambda n: orc.fq_from_ints([int.from_bytes(rng.bytes(64), "little") for _ in range(n)])   # noqa: E731
    e = _edge(orc, tw.P_MOD, orc.fq_from_ints, 0)
    raw = orc.ints_to_arr([tw.P_MOD - 1, tw.P_MOD - 2, (1 << 380) - 1, (1 << 380), (1 << 381) - 1 - (1 << 200),
                           int("ffffff" * 15, 16) + (0x19ffff << 360), int("800000" * 15, 16), int("7fffff" * 15, 16), 0, 1], 6)
    e = np.concatenate([e, raw])
    a = np.concatenate([rnd(4000), np.repeat(e, len(e), axis=0)])
    b = np.concatenate([rnd(4000), np.tile(e, (len(e), 1))])
    want = orc.fq_mul(a, b)
    for mchain in (0, 1):
        out = np.zeros_like(a)
        emu.emu_fq_mul_f64(_p(a), _p(b), _p(out), C.c_size_t(len(a)), C.c_int(mchain))
        assert np.array_equal(out, want), mchain
        # the interleaved pair: (a*b on integer rows, b*a' on FP64 rows)
        a2 = np.ascontiguousarray(a[::-1])
        out2 = np.zeros((2 * len(a), a.shape[1]), dtype=a.dtype)
        emu.emu_fq_mul_dual(_p(a), _p(b), _p(b), _p(a2), _p(out2), C.c_size_t(len(a)), C.c_int(mchain))
        assert np.array_equal(out2[0::2], want), mchain
        assert np.array_equal(out2[1::2], orc.fq_mul(b, a2)), mchain


def test_fq_sqr_sos(orc, emu):
    """dedicated Fq squaring (field.cuh fq_sqr_sos): symmetric product + separate Montgomery reduction"""
    rng = np.random.default_rng(103)
    a = np.concatenate([orc.fq_from_ints([int.from_bytes(rng.bytes(64), "little") for _ in range(4000)]),
                        _edge(orc, tw.P_MOD, orc.fq_from_ints, 0)])
    # raw limb patterns below p that stress the carry chains (these are Montgomery representatives of something)
    raw = orc.ints_to_arr([tw.P_MOD - 1, tw.P_MOD - 2, (1 << 380) - 1, (1 << 380), int("ffffffff00000000" * 5 + "0fffffff00000000", 16) % tw.P_MOD,
                           int("00000000ffffffff" * 6, 16) % tw.P_MOD, 1, 0, (1 << 381) - 1 - (1 << 200)], 6)
    a = np.concatenate([a, raw])
    out = np.zeros_like(a)
    emu.emu_fq_sqr_sos(_p(a), _p(out), C.c_size_t(len(a)))
    assert np.array_equal(out, orc.fq_mul(a, a))


def test_fr_inv_and_canon(orc, emu):
    rng = np.random.default_rng(101)
    a = orc.random_fr(rng, 64)
    out = np.zeros_like(a)
    emu.emu_fr_inv(_p(a), _p(out), C.c_size_t(len(a)))
    assert np.array_equal(out, orc.fr_inv(a))
    emu.emu_fr_canon(_p(a), _p(out), C.c_size_t(len(a)), 1)
    assert [orc.limbs_to_int(r) for r in out] == orc.fr_to_ints(a)
    back = np.zeros_like(a)
    emu.emu_fr_canon(_p(out), _p(back), C.c_size_t(len(a)), 0)
    assert np.array_equal(back, a)


def _inf_jac(orc):
    inf = np.zeros((1, 18), dtype=np.uint64)
    inf[0, 0:6] = orc.fq_from_ints([1])[0]
    inf[0, 6:12] = orc.fq_from_ints([1])[0]
    return inf


def test_g1_ops(orc, emu):
    rng = np.random.default_rng(102)
    n = 24
    pa = orc.random_g1(rng, n)
    pb = orc.random_g1(rng, n)
    ja = orc.g1_mul(orc.g1_from_affine(pa), orc.random_fr(rng, n)) * 0 + orc.g1_double(orc.g1_from_affine(pa))  # non-trivial Z
    jb = orc.g1_add(orc.g1_from_affine(pb), orc.g1_from_affine(pa))
    # exceptional cases: acc == P (doubling), acc == -P (-> identity), acc identity, P identity
    inf = _inf_jac(orc)
    neg_flags = np.zeros(n + 4, dtype=np.uint8)
    acc = np.concatenate([ja, orc.g1_from_affine(pb[:1]), orc.g1_from_affine(pb[1:2]), inf, ja[:1]])
    aff = np.concatenate([pb, pb[:1], pb[1:2], pb[2:3], np.zeros((1, 13), dtype=np.uint64)])
    neg_flags[n + 1] = 1
    neg_flags[3] = 1
    aff_packed = np.ascontiguousarray(aff[:, :12])
    aff_ref = aff.copy()
    aff_ref[-1, 12] = 1                       # oracle flags infinity explicitly
    for i in np.nonzero(neg_flags)[0]:
        aff_ref[i, 6:12] = orc.fq_sub(np.zeros((1, 6), dtype=np.uint64), aff_ref[i:i + 1, 6:12])[0]
    out = np.zeros_like(acc)
    emu.emu_g1_add_affine(_p(acc), _p(aff_packed), _p(neg_flags), _p(out), C.c_size_t(len(acc)))
    want = orc.canon_g1(orc.g1_add_mixed(acc, aff_ref))
    assert orc.canon_g1(out) == want
    assert want[n] != (0, 0, 1) and want[n + 1] == (0, 0, 1)
    # full add incl. P+P, P+(-P), identities
    A = np.concatenate([ja, ja[:1], ja[:1], inf, ja[:1], inf])
    negja = ja[:1].copy()
    negja[0, 6:12] = orc.fq_sub(np.zeros((1, 6), dtype=np.uint64), ja[:1, 6:12])[0]
    B = np.concatenate([jb, ja[:1], negja, jb[:1], inf, inf])
    out = np.zeros_like(A)
    emu.emu_g1_add(_p(A), _p(B), _p(out), C.c_size_t(len(A)))
    assert orc.canon_g1(out) == orc.canon_g1(orc.g1_add(A, B))
    out = np.zeros_like(A)
    emu.emu_g1_double(_p(A), _p(out), C.c_size_t(len(A)))
    assert orc.canon_g1(out) == orc.canon_g1(orc.g1_double(A))
    k = orc.random_fr(rng, 6)
    k[0] = 0
    k[1] = orc.fr_from_ints([1])[0]
    out = np.zeros_like(ja[:6])
    emu.emu_g1_mul_fr(_p(np.ascontiguousarray(ja[:6])), _p(k), _p(out), C.c_size_t(6))
    assert orc.canon_g1(out) == orc.canon_g1(orc.g1_mul(ja[:6], k))


@pytest.mark.parametrize("c", [1, 2, 3, 5, 8, 13, 15, 16, 17, 20])
def test_msm_digits_recompose(orc, emu, c):
    rng = np.random.default_rng(103 + c)
    ks = [int.from_bytes(rng.bytes(40), "little") % tw.R_MOD for _ in range(200)] + [0, 1, tw.R_MOD - 1, (1 << 254) - 1,
                                                                                       1 << 254, (1 << 255) - 19 - (1 << 255) + tw.R_MOD - 2]
    k = orc.ints_to_arr(ks, 4)
    W = (256 + c - 1) // c
    digits = np.zeros((len(ks), W), dtype=np.int32)
    got_w = emu.emu_msm_digits(_p(k), C.c_size_t(len(ks)), c, _p(digits))
    assert got_w == W
    half = 1 << (c - 1)
    assert digits.max() <= half and digits.min() > -half - (1 if c == 1 else 0)
    for i, kv in enumerate(ks):
        assert sum(int(d) << (c * w) for w, d in enumerate(digits[i])) == kv


def test_g1_batch_affine_add(orc, emu):
    """affine additions with one shared inversion per batch (g1_batch_affine.cuh, the round-2 bucket kernel's group
    law): every exceptional case in the same batch as ordinary additions"""
    from tests.gpu_util import oracle_affine, packed_affine
    rng = np.random.default_rng(105)
    n = 96
    pa, qa = orc.random_g1(rng, n), orc.random_g1(rng, n)
    inf = np.zeros((1, 13), dtype=np.uint64)
    inf[0, 12] = 1
    neg = qa.copy()
    neg[:, 6:12] = orc.fq_sub(np.zeros((n, 6), dtype=np.uint64), qa[:, 6:12])      # -Q
    qa[5], pa[9] = pa[5], inf[0]            # P + P, inf + Q
    qa[13] = inf[0]                         # P + inf
    pa[20], qa[20] = inf[0], inf[0]         # inf + inf
    pa[31] = neg[31]                        # (-Q) + Q = inf
    qa[40], qa[41] = pa[40], pa[41]         # two doublings next to each other
    pa[95] = neg[95]                        # last of its batch
    qa[64] = pa[64]                         # first of a batch of 32
    p, q = packed_affine(pa), packed_affine(qa)
    want = orc.canon_g1(orc.g1_add(orc.g1_from_affine(pa), orc.g1_from_affine(qa)))
    for batch in (1, 7, 32, 96, 200):
        out = np.zeros_like(p)
        emu.emu_g1_batch_affine_add(_p(p), _p(q), _p(out), C.c_size_t(n), C.c_size_t(batch))
        got = orc.canon_g1(orc.g1_from_affine(oracle_affine(out)))
        assert got == want, batch
    assert want[31] == (0, 0, 1) and want[20] == (0, 0, 1) and want[95] == (0, 0, 1)


def test_inverse_by_binary_gcd(orc, emu):
    """field.cuh fp_inv_bingcd (the one inversion at the top of the batched-affine product tree, msm_affine.cu):
    a * inv(a) == 1 in Montgomery form for random and edge values of Fq and Fr, and equality with the oracle's Fr inverse"""
    rng = np.random.default_rng(140)
    fq = np.concatenate([orc.fq_from_ints([int.from_bytes(rng.bytes(64), "little") for _ in range(300)]),
                         orc.fq_from_ints([1, 2, 3, tw.P_MOD - 1, tw.P_MOD - 2, (tw.P_MOD - 1) // 2, (1 << 380) + 5, 1 << 64])])
    out = np.zeros_like(fq)
    emu.emu_fq_inv_bingcd(_p(fq), _p(out), C.c_size_t(len(fq)))
    one = orc.fq_from_ints([1])
    assert np.array_equal(orc.fq_mul(fq, out), np.repeat(one, len(fq), axis=0))
    z = np.zeros((1, 6), dtype=np.uint64)
    emu.emu_fq_inv_bingcd(_p(z), _p(out[:1]), C.c_size_t(1))
    assert not out[:1].any()
    fr = np.concatenate([orc.random_fr(rng, 300), orc.fr_from_ints([1, 2, tw.R_MOD - 1, (tw.R_MOD + 1) // 2])])
    out = np.zeros_like(fr)
    emu.emu_fr_inv_bingcd(_p(fr), _p(out), C.c_size_t(len(fr)))
    assert np.array_equal(out, orc.fr_inv(fr))


# ---------------------------------------------------------------- G2 (csrc/g2.cuh) against the Python big-int twin
def _f2_limbs(orc, vals):
    """list of Fq2 (c0, c1) ints -> (n, 12) uint64 Montgomery limbs"""
    flat = []
    for a in vals:
        flat += [a[0], a[1]]
    return np.ascontiguousarray(orc.fq_from_ints(flat).reshape(len(vals), 12))


def _f2_ints(orc, limbs):
    v = orc.fq_to_ints(np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1, 6))
    return [(v[2 * i], v[2 * i + 1]) for i in range(len(v) // 2)]


def _g2_jac(orc, pts, z=None):
    """affine twin points (None = identity) -> (n, 36) Jacobian limbs; `z`: per-point Fq2 to scale the representative with"""
    rows = []
    for i, p in enumerate(pts):
        if p is None:
            rows += [(1, 0), (1, 0), (0, 0)]
            continue
        zz = (1, 0) if z is None else z[i]
        z2 = tw.f2_mul(zz, zz)
        rows += [tw.f2_mul(p[0], z2), tw.f2_mul(p[1], tw.f2_mul(z2, zz)), zz]
    return _f2_limbs(orc, rows).reshape(len(pts), 36)


def _g2_canon(orc, jac):
    out = []
    for row in np.ascontiguousarray(jac, dtype=np.uint64).reshape(-1, 36):
        X, Y, Z = _f2_ints(orc, row)
        if Z == (0, 0):
            out.append(None)
            continue
        zi = tw.f2_inv(Z)
        zi2 = tw.f2_mul(zi, zi)
        out.append((tw.f2_mul(X, zi2), tw.f2_mul(Y, tw.f2_mul(zi2, zi))))
    return out


def test_fq2_mul_sqr(orc, emu):
    import random
    r = random.Random(31)
    P = tw.P_MOD
    edge = [(0, 0), (1, 0), (0, 1), (P - 1, P - 1), (P - 1, 0), (0, P - 1), (1, 1), ((P - 1) // 2, 2)]
    a = edge + [(r.randrange(P), r.randrange(P)) for _ in range(40)]
    b = list(reversed(edge)) + [(r.randrange(P), r.randrange(P)) for _ in range(40)]
    A, B = _f2_limbs(orc, a), _f2_limbs(orc, b)
    out = np.zeros_like(A)
    emu.emu_fq2_mul(_p(A), _p(B), _p(out), C.c_size_t(len(a)))
    assert _f2_ints(orc, out) == [tw.f2_mul(x, y) for x, y in zip(a, b)]
    out = np.zeros_like(A)
    emu.emu_fq2_sqr(_p(A), _p(out), C.c_size_t(len(a)))
    assert _f2_ints(orc, out) == [tw.f2_mul(x, x) for x in a]


def test_g2_group_law(orc, emu):
    """XYZZ addition / doubling / mixed addition / scalar multiplication of g2.cuh incl. every exceptional case (P + P, P - P,
    identities on either side), on non-trivial Jacobian representatives"""
    import random
    r = random.Random(32)
    G = (tw.G2_X, tw.G2_Y)
    assert tw.g2_on_curve(G)
    ks = [r.randrange(1, tw.R_MOD) for _ in range(6)]
    pts = [tw.g2_mul(G, k) for k in ks]
    zs = [(r.randrange(1, tw.P_MOD), r.randrange(tw.P_MOD)) for _ in pts]
    neg = lambda p: (p[0], tw.f2_sub((0, 0), p[1]))   # noqa: E731
    a_pts = pts + [pts[0], pts[0], None, pts[1], None]
    b_pts = pts[1:] + pts[:1] + [pts[0], neg(pts[0]), pts[2], None, None]
    za = zs + [zs[1], zs[2], None, zs[3], None]
    zb = zs[1:] + zs[:1] + [zs[4], zs[5], zs[0], None, None]
    A, B = _g2_jac(orc, a_pts, za), _g2_jac(orc, b_pts, zb)
    out = np.zeros_like(A)
    emu.emu_g2_add(_p(A), _p(B), _p(out), C.c_size_t(len(A)))
    want = [tw.g2_add(p, q) for p, q in zip(a_pts, b_pts)]
    assert _g2_canon(orc, out) == want
    assert want[len(pts) + 1] is None and want[-1] is None and all(tw.g2_on_curve(p) for p in want)
    out = np.zeros_like(A)
    emu.emu_g2_double(_p(A), _p(out), C.c_size_t(len(A)))
    assert _g2_canon(orc, out) == [tw.g2_add(p, p) for p in a_pts]
    # mixed addition: acc (Jacobian) += affine, optionally negated; acc == P, acc == -P, acc identity, P identity
    acc_pts = pts + [pts[0], pts[1], None, pts[2]]
    aff_pts = pts[2:] + pts[:2] + [pts[0], pts[1], pts[3], None]
    flags = np.zeros(len(acc_pts), dtype=np.uint8)
    flags[1] = flags[len(pts) + 1] = 1
    ACC = _g2_jac(orc, acc_pts, zs + [zs[3], zs[4], None, zs[5]])
    AFF = _f2_limbs(orc, [c for p in aff_pts for c in (((0, 0), (0, 0)) if p is None else p)]).reshape(len(aff_pts), 24)
    out = np.zeros_like(ACC)
    emu.emu_g2_add_affine(_p(ACC), _p(AFF), _p(flags), _p(out), C.c_size_t(len(ACC)))
    want = [tw.g2_add(p, (neg(q) if f and q is not None else q)) for p, q, f in zip(acc_pts, aff_pts, flags)]
    assert _g2_canon(orc, out) == want
    assert want[len(pts)] == tw.g2_add(pts[0], pts[0]) and want[len(pts) + 1] is None
    # k * P, canonical scalars: 0, 1, r - 1, random
    sc = [0, 1, tw.R_MOD - 1] + [r.randrange(tw.R_MOD) for _ in range(3)]
    K = np.array([[(k >> (64 * j)) & ((1 << 64) - 1) for j in range(4)] for k in sc], dtype=np.uint64)
    PJ = _g2_jac(orc, pts, zs)
    out = np.zeros_like(PJ)
    emu.emu_g2_mul_bits(_p(PJ), _p(K), _p(out), C.c_size_t(len(sc)))
    assert _g2_canon(orc, out) == [tw.g2_mul(p, k) for p, k in zip(pts, sc)]
