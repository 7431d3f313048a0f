"""ORACLE (test infrastructure) -- ctypes front-end of oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
``--impl reference`` legs may import this module; the product package never
does.  PARITY UNPINNED at the arkworks byte boundary (see oracle/src/oracle.h
and DESIGN.md): pinned by public constants, the big-int twin
(oracle/py_twin.py) and the reference's own property tests.

Array conventions (numpy, dtype uint64, C-contiguous):
  Fr   (n, 4)   Montgomery limbs, R = 2^256  (= ark-ff Fp.0.0)
  Fq   (n, 6)   Montgomery limbs, R = 2^384
  G1 affine   (n, 13)  x[6] | y[6] | infinity flag   (oracle struct g1a_t)
  G1 Jacobian (n, 18)  X[6] | Y[6] | Z[6]            (ark-ec Projective)
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

LEADER_SIM, PARTIES = 0, 1
R_MOD = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
P_MOD = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        _LIB.orc_init()
        for name in ("orc_sumcheck_product", "orc_c_sumcheck_product", "orc_d_sumcheck_product", "orc_c_open",
                     "orc_d_open", "orc_sumcheck", "orc_c_sumcheck", "orc_d_sumcheck"):
            getattr(_LIB, name).restype = C.c_size_t
    return _LIB


class PP(C.Structure):
    _fields_ = [("t", C.c_size_t), ("l", C.c_size_t), ("n", C.c_size_t), ("share_gen", C.c_uint64 * 4),
                ("secret_gen", C.c_uint64 * 4), ("secret2_gen", C.c_uint64 * 4), ("coset", C.c_uint64 * 4)]


class SRS(C.Structure):
    _fields_ = [("levels", C.c_size_t), ("powers_of_g", C.POINTER(C.c_void_p)), ("level_len", C.POINTER(C.c_size_t))]


def _u64(a, cols=None):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ptr_array(arrs):
    return (C.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


# ---------------------------------------------------------------- integers <-> limbs
def int_to_limbs(v, n):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)]


def limbs_to_int(row):
    return sum(int(x) << (64 * i) for i, x in enumerate(row))


def ints_to_arr(vals, n):
    return np.array([int_to_limbs(v, n) for v in vals], dtype=np.uint64).reshape(-1, n)


def fr_from_ints(vals):
    """canonical python ints -> Montgomery Fr array"""
    a = ints_to_arr([v % R_MOD for v in vals], 4)
    out = np.empty_like(a)
    lib().orc_fr_from_canon_vec(_p(a), _p(out), C.c_size_t(len(a)))
    return out


def fr_to_ints(a):
    a = _u64(a, 4)
    out = np.empty_like(a)
    lib().orc_fr_to_canon_vec(_p(a), _p(out), C.c_size_t(len(a)))
    return [limbs_to_int(r) for r in out]


def fq_from_ints(vals):
    a = ints_to_arr([v % P_MOD for v in vals], 6)
    out = np.empty_like(a)
    lib().orc_fq_from_canon_vec(_p(a), _p(out), C.c_size_t(len(a)))
    return out


def fq_to_ints(a):
    a = _u64(a, 6)
    out = np.empty_like(a)
    lib().orc_fq_to_canon_vec(_p(a), _p(out), C.c_size_t(len(a)))
    return [limbs_to_int(r) for r in out]


def _vec2(name, a, b, cols):
    a, b = _u64(a, cols), _u64(b, cols)
    out = np.empty_like(a)
    getattr(lib(), name)(_p(a), _p(b), _p(out), C.c_size_t(len(a)))
    return out


def fr_mul(a, b): return _vec2("orc_fr_mul_vec", a, b, 4)
def fr_add(a, b): return _vec2("orc_fr_add_vec", a, b, 4)
def fr_sub(a, b): return _vec2("orc_fr_sub_vec", a, b, 4)
def fq_mul(a, b): return _vec2("orc_fq_mul_vec", a, b, 6)
def fq_add(a, b): return _vec2("orc_fq_add_vec", a, b, 6)
def fq_sub(a, b): return _vec2("orc_fq_sub_vec", a, b, 6)


def fr_inv(a):
    a = _u64(a, 4)
    out = np.empty_like(a)
    lib().orc_fr_inv_vec(_p(a), _p(out), C.c_size_t(len(a)))
    return out


# ---------------------------------------------------------------- random inputs
def random_fr(rng, n):
    """Uniform Montgomery limbs below r: a uniformly random representation IS a
    uniformly random field element, so no conversion is needed (the reference
    draws witnesses with F::rand, dist-primitive/src/lib.rs:13-18)."""
    a = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 63) - 1)
    mod = np.array(int_to_limbs(R_MOD, 4), dtype=np.uint64)
    # lexicographic a >= r  -> subtract r once (2^255 < 2r)
    ge = np.zeros(n, dtype=bool)
    decided = np.zeros(n, dtype=bool)
    for i in (3, 2, 1, 0):
        gt = (a[:, i] > mod[i]) & ~decided
        lt = (a[:, i] < mod[i]) & ~decided
        ge |= gt
        decided |= gt | lt
    ge |= ~decided
    idx = np.nonzero(ge)[0]
    for k in idx:
        v = limbs_to_int(a[k]) - R_MOD
        a[k] = int_to_limbs(v, 4)
    return a


def g1_gen_mul(k_fr):
    """affine (n,13) = k[i] * G1 generator"""
    k = _u64(k_fr, 4)
    out = np.zeros((len(k), 13), dtype=np.uint64)
    lib().orc_g1_gen_mul_vec(_p(k), _p(out), C.c_size_t(len(k)))
    return out


def random_g1(rng, n):
    return g1_gen_mul(random_fr(rng, n))


def g1_generator():
    return np.ctypeslib.as_array((C.c_uint64 * 13).in_dll(lib(), "G1_GEN")).copy().reshape(1, 13)


# ---------------------------------------------------------------- G1
def g1_to_affine(j):
    j = _u64(j, 18)
    out = np.zeros((len(j), 13), dtype=np.uint64)
    lib().orc_g1j_to_affine_vec(_p(j), _p(out), C.c_size_t(len(j)))
    return out


def g1_from_affine(a):
    a = _u64(a, 13)
    out = np.zeros((len(a), 18), dtype=np.uint64)
    for i in range(len(a)):
        lib().g1j_from_affine(C.c_void_p(out[i].ctypes.data), C.c_void_p(a[i].ctypes.data))
    return out


def g1_add(a, b):
    a, b = _u64(a, 18), _u64(b, 18)
    out = np.zeros_like(a)
    lib().orc_g1_add_vec(_p(a), _p(b), _p(out), C.c_size_t(len(a)))
    return out


def g1_add_mixed(a, b):
    a, b = _u64(a, 18), _u64(b, 13)
    out = np.zeros_like(a)
    lib().orc_g1_add_mixed_vec(_p(a), _p(b), _p(out), C.c_size_t(len(a)))
    return out


def g1_double(a):
    a = _u64(a, 18)
    out = np.zeros_like(a)
    lib().orc_g1_double_vec(_p(a), _p(out), C.c_size_t(len(a)))
    return out


def g1_mul(j, k_fr):
    j, k = _u64(j, 18), _u64(k_fr, 4)
    out = np.zeros_like(j)
    for i in range(len(j)):
        lib().g1j_mul_fr(C.c_void_p(out[i].ctypes.data), C.c_void_p(j[i].ctypes.data), C.c_void_p(k[i].ctypes.data))
    return out


def g1_eq(a, b):
    a, b = _u64(a, 18), _u64(b, 18)
    return [bool(lib().g1j_eq(C.c_void_p(a[i].ctypes.data), C.c_void_p(b[i].ctypes.data))) for i in range(len(a))]


def g1_on_curve(a):
    a = _u64(a, 13)
    return [bool(lib().g1a_on_curve(C.c_void_p(a[i].ctypes.data))) for i in range(len(a))]


def canon_g1(j):
    """Canonical comparison form of Jacobian points: list of (x, y, inf) python ints."""
    aff = g1_to_affine(j)
    xs, ys = fq_to_ints(aff[:, 0:6]), fq_to_ints(aff[:, 6:12])
    return [(0, 0, 1) if int(aff[i, 12]) & 0xFFFFFFFF else (xs[i], ys[i], 0) for i in range(len(aff))]


def set_msm_threads(t):
    """threads used by the ark-style MSM inside every protocol function (default 1, like the reference)"""
    lib().orc_set_msm_threads(C.c_int(t))


def msm(bases, scalars, algo="ark", threads=1):
    b, s = _u64(bases, 13), _u64(scalars, 4)
    assert len(b) == len(s)
    out = np.zeros((1, 18), dtype=np.uint64)
    if algo == "naive":
        lib().g1_msm_naive(_p(out), _p(b), _p(s), C.c_size_t(len(b)))
    else:
        lib().g1_msm_ark_mt(_p(out), _p(b), _p(s), C.c_size_t(len(b)), C.c_int(threads))
    return out


# ---------------------------------------------------------------- PSS
_KCOLS = {0: 4, 1: 18}


def pp_new(l):
    pp = PP()
    lib().orc_pp_new(C.byref(pp), C.c_size_t(l))
    return pp


def pack_from_public(pp, secrets, kind=0):
    s = _u64(secrets, _KCOLS[kind])
    out = np.zeros((pp.n, _KCOLS[kind]), dtype=np.uint64)
    lib().orc_pack_from_public(C.byref(pp), kind, _p(s), C.c_size_t(len(s)), _p(out))
    return out


def pack_single(pp, secret, kind=0):
    s = _u64(secret, _KCOLS[kind])
    out = np.zeros((pp.n, _KCOLS[kind]), dtype=np.uint64)
    lib().orc_pack_single(C.byref(pp), kind, _p(s), _p(out))
    return out


def unpack(pp, shares, kind=0):
    s = _u64(shares, _KCOLS[kind])
    assert len(s) == pp.n
    out = np.zeros((pp.l, _KCOLS[kind]), dtype=np.uint64)
    lib().orc_unpack(C.byref(pp), kind, _p(s), _p(out))
    return out


def unpack2(pp, shares, kind=0):
    s = _u64(shares, _KCOLS[kind])
    assert len(s) == pp.n
    out = np.zeros((pp.l, _KCOLS[kind]), dtype=np.uint64)
    lib().orc_unpack2(C.byref(pp), kind, _p(s), _p(out))
    return out


# ---------------------------------------------------------------- protocols
def _nparties(pp, mode):
    return pp.n if mode == PARTIES else 1


def d_msm(pp, mode, bases, scalars, algo="ark"):
    """bases/scalars: [party][k] nested lists of arrays (LEADER_SIM: one party). -> (P, batch, 18)"""
    P = _nparties(pp, mode)
    assert len(bases) == P and len(scalars) == P
    batch = len(bases[0])
    bs = [_u64(bases[p][k], 13) for p in range(P) for k in range(batch)]
    ss = [_u64(scalars[p][k], 4) for p in range(P) for k in range(batch)]
    lens = (C.c_size_t * batch)(*[len(ss[k]) for k in range(batch)])
    for p in range(P):
        for k in range(batch):
            assert len(bs[p * batch + k]) == len(ss[p * batch + k]) == lens[k]
    out = np.zeros((P, batch, 18), dtype=np.uint64)
    lib().orc_d_msm(C.byref(pp), mode, C.c_size_t(batch), lens, _ptr_array(bs), _ptr_array(ss), _p(out),
                    1 if algo == "ark" else 0)
    return out


def sumcheck_product(f, g, challenge):
    f, g, ch = _u64(f, 4), _u64(g, 4), _u64(challenge, 4)
    n = len(f).bit_length() - 1
    out = np.zeros((n + 1, 3, 4), dtype=np.uint64)
    lib().orc_sumcheck_product(_p(f), _p(g), C.c_size_t(len(f)), _p(ch), _p(out))
    return out


def sumcheck(f, challenge):
    """dsumcheck.rs:6-26 -> (n + 1, 2, 4)"""
    f, ch = _u64(f, 4), _u64(challenge, 4)
    n = len(f).bit_length() - 1
    out = np.zeros((n + 1, 2, 4), dtype=np.uint64)
    lib().orc_sumcheck(_p(f), C.c_size_t(len(f)), _p(ch), _p(out))
    return out


def c_sumcheck(pp, mode, f, challenge):
    """dsumcheck.rs:92-146 -> (P, n + log2 l + 1, 2, 4)"""
    P = _nparties(pp, mode)
    fs = [_u64(x, 4) for x in f]
    assert len(fs) == P
    ch = _u64(challenge, 4)
    cnt = (len(fs[0]).bit_length() - 1) + (pp.l.bit_length() - 1) + 1
    out = np.zeros((P, cnt, 2, 4), dtype=np.uint64)
    lib().orc_c_sumcheck(C.byref(pp), mode, _ptr_array(fs), C.c_size_t(len(fs[0])), _p(ch), _p(out))
    return out


def d_sumcheck(mode, nparties, f, challenge):
    """dsumcheck.rs:287-357 -> the leader's (n + log2 N, 2, 4)"""
    fs = [_u64(x, 4) for x in f]
    ch = _u64(challenge, 4)
    cnt = (len(fs[0]).bit_length() - 1) + (nparties.bit_length() - 1)
    out = np.zeros((cnt, 2, 4), dtype=np.uint64)
    lib().orc_d_sumcheck(mode, C.c_size_t(nparties), _ptr_array(fs), C.c_size_t(len(fs[0])), _p(ch), _p(out))
    return out


def c_sumcheck_product(pp, mode, f, g, challenge):
    P = _nparties(pp, mode)
    fs, gs = [_u64(x, 4) for x in f], [_u64(x, 4) for x in g]
    assert len(fs) == P
    ch = _u64(challenge, 4)
    n = len(fs[0]).bit_length() - 1
    cnt = n + (pp.l.bit_length() - 1) + 1
    out = np.zeros((P, cnt, 3, 4), dtype=np.uint64)
    lib().orc_c_sumcheck_product(C.byref(pp), mode, _ptr_array(fs), _ptr_array(gs), C.c_size_t(len(fs[0])), _p(ch),
                                 _p(out))
    return out


def d_sumcheck_product(mode, nparties, f, g, challenge):
    fs, gs = [_u64(x, 4) for x in f], [_u64(x, 4) for x in g]
    ch = _u64(challenge, 4)
    n = len(fs[0]).bit_length() - 1
    s = nparties.bit_length() - 1
    out = np.zeros((n + s, 3, 4), dtype=np.uint64)
    lib().orc_d_sumcheck_product(mode, C.c_size_t(nparties), _ptr_array(fs), _ptr_array(gs), C.c_size_t(len(fs[0])),
                                 _p(ch), _p(out))
    return out


def pss2ss(pp, mode, shares):
    s = _u64(shares, 4)
    P = _nparties(pp, mode)
    assert len(s) == P
    out = np.zeros((P, pp.l, 4), dtype=np.uint64)
    lib().orc_pss2ss(C.byref(pp), mode, _p(s), _p(out))
    return out


def degree_reduce(pp, mode, shares):
    s = _u64(shares, 4)
    P = _nparties(pp, mode)
    out = np.zeros((P, 4), dtype=np.uint64)
    lib().orc_degree_reduce(C.byref(pp), mode, _p(s), _p(out))
    return out


def fix_variable(evals, points):
    e, pts = _u64(evals, 4), _u64(points, 4)
    n = len(e).bit_length() - 1
    k = min(n, len(pts))
    out = np.zeros((len(e) >> k, 4), dtype=np.uint64)
    lib().orc_fix_variable(_p(e), C.c_size_t(len(e)), _p(pts), C.c_size_t(len(pts)), _p(out))
    return out


def sub_index(i):
    a, b = C.c_size_t(), C.c_size_t()
    lib().orc_sub_index(C.c_size_t(i), C.byref(a), C.byref(b))
    return a.value, b.value


def acc_product_tree(x):
    x = _u64(x, 4)
    out = np.zeros((2 * len(x), 4), dtype=np.uint64)
    lib().orc_acc_product_tree(_p(x), C.c_size_t(len(x)), _p(out))
    return out


def acc_product(x):
    """dacc_product.rs:30-57 -> (v(x,0), v(x,1), v(1,x))"""
    t = acc_product_tree(x)
    return t[0::2].copy(), t[1::2].copy(), t[len(t) // 2:].copy()


def d_acc_product(mode, nparties, inputs):
    xs = [_u64(x, 4) for x in inputs]
    ln = len(xs[0])
    subs = [np.zeros((2 * ln, 4), dtype=np.uint64) for _ in xs]
    top = np.zeros((2 * nparties, 4), dtype=np.uint64)
    lib().orc_d_acc_product(mode, C.c_size_t(nparties), _ptr_array(xs), C.c_size_t(ln), _ptr_array(subs), _p(top))
    return subs, top


class Srs:
    """PolynomialCommitment (dpoly_comm.rs:30-34): powers_of_g[level] affine arrays (n,13)."""

    def __init__(self, c_srs, keep=None):
        self.c = c_srs
        self._keep = keep

    @classmethod
    def new(cls, g_jac, s_fr):
        """PolynomialCommitmentCub::new(g, _, s).mature()  -- real trapdoor SRS"""
        srs = SRS()
        g, s = _u64(g_jac, 18), _u64(s_fr, 4)
        lib().orc_srs_new(C.byref(srs), _p(g), _p(s), C.c_size_t(len(s)))
        return cls(srs)

    @classmethod
    def from_levels(cls, levels):
        """wrap caller-made levels (new_single / new_random shapes, dpoly_comm.rs:197-234)"""
        srs = SRS()
        arrs = [_u64(a, 13) for a in levels]
        lens = (C.c_size_t * len(arrs))(*[len(a) for a in arrs])
        lib().orc_srs_from_levels(C.byref(srs), C.c_size_t(len(arrs)), _ptr_array(arrs), lens)
        return cls(srs)

    def level(self, i):
        n = self.c.level_len[i]
        buf = (C.c_uint64 * (13 * n)).from_address(self.c.powers_of_g[i])
        return np.ctypeslib.as_array(buf).reshape(n, 13).copy()

    @property
    def levels(self):
        return self.c.levels


def _srs_ptrs(srs_list):
    return (C.POINTER(SRS) * len(srs_list))(*[C.pointer(s.c) for s in srs_list])


def commit(srs, peval, algo="ark"):
    p = _u64(peval, 4)
    out = np.zeros((1, 18), dtype=np.uint64)
    lib().orc_commit(C.byref(srs.c), _p(p), C.c_size_t(len(p)), _p(out), 1 if algo == "ark" else 0)
    return out


def open_(srs, peval, point, algo="ark"):
    p, pt = _u64(peval, 4), _u64(point, 4)
    n = len(p).bit_length() - 1
    val = np.zeros((1, 4), dtype=np.uint64)
    proofs = np.zeros((max(n, 1), 18), dtype=np.uint64)
    lib().orc_open(C.byref(srs.c), _p(p), C.c_size_t(len(p)), _p(pt), _p(val), _p(proofs), 1 if algo == "ark" else 0)
    return val, proofs[:n]


def c_commit(srs_list, pp, mode, pevals, algo="ark"):
    P = _nparties(pp, mode)
    batch = len(pevals[0])
    ps = [_u64(pevals[p][k], 4) for p in range(P) for k in range(batch)]
    lens = (C.c_size_t * batch)(*[len(ps[k]) for k in range(batch)])
    out = np.zeros((P, batch, 18), dtype=np.uint64)
    lib().orc_c_commit(_srs_ptrs(srs_list), C.byref(pp), mode, C.c_size_t(batch), lens, _ptr_array(ps), _p(out),
                       1 if algo == "ark" else 0)
    return out


def c_open(srs_list, pp, mode, pevals, point, algo="ark"):
    P = _nparties(pp, mode)
    ps = [_u64(x, 4) for x in pevals]
    pt = _u64(point, 4)
    n = len(ps[0]).bit_length() - 1
    cnt = n + pp.l.bit_length() - 1
    val = np.zeros((P, 4), dtype=np.uint64)
    proofs = np.zeros((P, max(cnt, 1), 18), dtype=np.uint64)
    lib().orc_c_open(_srs_ptrs(srs_list), C.byref(pp), mode, _ptr_array(ps), C.c_size_t(len(ps[0])), _p(pt), _p(val),
                     _p(proofs), 1 if algo == "ark" else 0)
    return val, proofs[:, :cnt]


def d_commit(srs_list, mode, nparties, pevals, algo="ark"):
    ps = [_u64(x, 4) for x in pevals]
    out = np.zeros((1, 18), dtype=np.uint64)
    lib().orc_d_commit(_srs_ptrs(srs_list), mode, C.c_size_t(nparties), _ptr_array(ps), C.c_size_t(len(ps[0])),
                       _p(out), 1 if algo == "ark" else 0)
    return out


def d_open(srs_list, mode, nparties, pevals, point, algo="ark"):
    ps = [_u64(x, 4) for x in pevals]
    pt = _u64(point, 4)
    n = len(ps[0]).bit_length() - 1
    pl = nparties.bit_length() - 1
    val = np.zeros((1, 4), dtype=np.uint64)
    proofs = np.zeros((pl + n, 18), dtype=np.uint64)
    lib().orc_d_open(_srs_ptrs(srs_list), mode, C.c_size_t(nparties), _ptr_array(ps), C.c_size_t(len(ps[0])), _p(pt),
                     C.c_size_t(len(pt)), _p(val), _p(proofs), 1 if algo == "ark" else 0)
    return val, proofs
