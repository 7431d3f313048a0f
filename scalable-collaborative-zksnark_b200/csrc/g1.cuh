// BLS12-381 G1 group law for the MSM / PSS kernels (sm_100a).
//
// Bases arrive as packed affine (x | y), 96 B per point, Montgomery coordinates
// (the reference's `G::Affine`, dist-primitive/src/dmsm.rs:10, with the
// `infinity` flag folded into the otherwise impossible encoding x = y = 0).
// Accumulators are extended Jacobian "XYZZ" (X, Y, ZZ, ZZZ) with x = X/ZZ,
// y = Y/ZZZ, ZZ^3 = ZZZ^2: a mixed add costs 8M + 2S, and no inversion is
// ever needed on the device.  Results leave as Jacobian (X, Y, Z), the layout
// of ark-ec's `Projective` (what `d_msm` returns, dmsm.rs:15).
//
// Every exceptional case is handled (identity operands, P + P, P + (-P)): the
// reference's own d_msm test feeds 256 copies of one point with all-one scalars
// (dmsm.rs:97-104), which drives every addition of a bucket through them.
#pragma once
#include "field.cuh"

namespace scz {

struct G1Affine {   // 96 B in HBM
    Fq x, y;
    SCZ_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
struct G1Jac {      // 144 B in HBM, ark-ec Projective
    Fq x, y, z;
};
struct G1X {        // XYZZ
    Fq x, y, zz, zzz;
    SCZ_HD bool is_inf() const { return zz.is_zero(); }
    SCZ_HD static G1X inf() {
        G1X r;
        r.x = Fq::zero();
        r.y = Fq::zero();
        r.zz = Fq::zero();
        r.zzz = Fq::zero();
        return r;
    }
};

SCZ_HD G1X g1x_from_affine(const G1Affine &p) {
    if (p.is_inf()) return G1X::inf();
    G1X r;
    r.x = p.x;
    r.y = p.y;
    r.zz = Fq::one();
    r.zzz = Fq::one();
    return r;
}
SCZ_HD G1X g1x_from_jac(const G1Jac &p) {
    if (p.z.is_zero()) return G1X::inf();
    G1X r;
    r.x = p.x;
    r.y = p.y;
    r.zz = fp_sqr(p.z);
    r.zzz = fp_mul(r.zz, p.z);
    return r;
}
// (X, Y, ZZ, ZZZ) -> Jacobian with Z = ZZZ: x = X*ZZ^2 / ZZZ^2 (ZZ^3 = ZZZ^2), y = Y*ZZZ^2 / ZZZ^3
SCZ_HD G1Jac g1x_to_jac(const G1X &p) {
    G1Jac r;
    if (p.is_inf()) {   // ark-ec's identity: (1, 1, 0)
        r.x = Fq::one();
        r.y = Fq::one();
        r.z = Fq::zero();
        return r;
    }
    Fq zz2 = fp_sqr(p.zz);
    Fq zzz2 = fp_sqr(p.zzz);
    r.x = fp_mul(p.x, zz2);
    r.y = fp_mul(p.y, zzz2);
    r.z = p.zzz;
    return r;
}
SCZ_HD G1X g1x_neg(const G1X &p) {
    G1X r = p;
    r.y = fp_neg(p.y);
    return r;
}
// Multiplier policy of the group law.  The throughput kernels inline every field product (MulInline); the
// latency-bound tail kernels (bucket tree, fix-up) call ONE out-of-line copy of each (a general addition is 14
// products: inlined it is ~90 KB of code, which, with the doubling beside it, no longer fits the instruction cache).
struct MulInline {
    SCZ_HD static Fq mul(const Fq &a, const Fq &b) { return fp_mul(a, b); }
    SCZ_HD static Fq sqr(const Fq &a) { return fp_sqr(a); }
    SCZ_HD static Fq dot2_sub(const Fq &a, const Fq &b, const Fq &c, const Fq &d) { return fp_dot2_sub(a, b, c, d); }
};
// dbl-2008-s-1 (a = 0)
template <class M = MulInline>
SCZ_HD G1X g1x_double(const G1X &p) {
    if (p.is_inf()) return p;
    Fq u = fp_dbl(p.y);
    Fq v = M::sqr(u);
    Fq w = M::mul(u, v);
    Fq s = M::mul(p.x, v);
    Fq xx = M::sqr(p.x);
    Fq m = fp_add(fp_dbl(xx), xx);
    G1X r;
    r.x = fp_sub(fp_sub(M::sqr(m), s), s);
    r.y = M::dot2_sub(m, fp_sub(s, r.x), w, p.y);          // m (s - x3) - w y, one reduction
    r.zz = M::mul(v, p.zz);
    r.zzz = M::mul(w, p.zzz);
    return r;
}
// mdbl-2008-s-1: double an affine point
template <class M = MulInline>
SCZ_HD G1X g1x_double_affine(const Fq &x, const Fq &y) {
    Fq u = fp_dbl(y);
    Fq v = M::sqr(u);
    Fq w = M::mul(u, v);
    Fq s = M::mul(x, v);
    Fq xx = M::sqr(x);
    Fq m = fp_add(fp_dbl(xx), xx);
    G1X r;
    r.x = fp_sub(fp_sub(M::sqr(m), s), s);
    r.y = M::dot2_sub(m, fp_sub(s, r.x), w, y);
    r.zz = v;
    r.zzz = w;
    return r;
}
// madd-2008-s: acc += (x2, y2) affine, not infinity
template <class M = MulInline>
SCZ_HD void g1x_add_affine(G1X &acc, const Fq &x2, const Fq &y2) {
    if (acc.is_inf()) {
        acc.x = x2;
        acc.y = y2;
        acc.zz = Fq::one();
        acc.zzz = Fq::one();
        return;
    }
    Fq u2 = M::mul(x2, acc.zz);
    Fq s2 = M::mul(y2, acc.zzz);
    Fq p = fp_sub(u2, acc.x);
    Fq r = fp_sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = g1x_double_affine<M>(x2, y2);
        else acc = G1X::inf();
        return;
    }
    Fq pp = M::sqr(p);
    Fq ppp = M::mul(p, pp);
    Fq q = M::mul(acc.x, pp);
    Fq x3 = fp_sub(fp_sub(fp_sub(M::sqr(r), ppp), q), q);
    Fq y3 = M::dot2_sub(r, fp_sub(q, x3), acc.y, ppp);      // r (q - x3) - y1 ppp, one reduction
    acc.x = x3;
    acc.y = y3;
    acc.zz = M::mul(acc.zz, pp);
    acc.zzz = M::mul(acc.zzz, ppp);
}
template <class M = MulInline>
SCZ_HD void g1x_add_affine(G1X &acc, const G1Affine &p, bool negate) {
    if (p.is_inf()) return;
    Fq y = negate ? fp_neg(p.y) : p.y;
    g1x_add_affine<M>(acc, p.x, y);
}
// add-2008-s
template <class M = MulInline>
SCZ_HD G1X g1x_add(const G1X &a, const G1X &b) {
    if (a.is_inf()) return b;
    if (b.is_inf()) return a;
    Fq u1 = M::mul(a.x, b.zz);
    Fq u2 = M::mul(b.x, a.zz);
    Fq s1 = M::mul(a.y, b.zzz);
    Fq s2 = M::mul(b.y, a.zzz);
    Fq p = fp_sub(u2, u1);
    Fq r = fp_sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) return g1x_double<M>(a);
        return G1X::inf();
    }
    Fq pp = M::sqr(p);
    Fq ppp = M::mul(p, pp);
    Fq q = M::mul(u1, pp);
    G1X o;
    o.x = fp_sub(fp_sub(fp_sub(M::sqr(r), ppp), q), q);
    o.y = M::dot2_sub(r, fp_sub(q, o.x), s1, ppp);
    o.zz = M::mul(M::mul(a.zz, b.zz), pp);
    o.zzz = M::mul(M::mul(a.zzz, b.zzz), ppp);
    return o;
}
// k * P, k = canonical little-endian 32-bit limbs (double-and-add, MSB first)
template <int KL>
SCZ_HD G1X g1x_mul_bits(const G1X &p, const uint32_t (&k)[KL]) {
    G1X acc = G1X::inf();
    for (int i = KL * 32 - 1; i >= 0; i--) {
        acc = g1x_double(acc);
        if ((k[i >> 5] >> (i & 31)) & 1) acc = g1x_add(acc, p);
    }
    return acc;
}
// k in Montgomery form
SCZ_HD G1X g1x_mul_fr(const G1X &p, const Fr &k) {
    Fr c = fp_to_canon(k);
    return g1x_mul_bits(p, c.l);
}

#if defined(__CUDACC__)
SCZ_D G1Affine g1a_load(const void *base, size_t idx) {
    G1Affine p;
    const char *b = reinterpret_cast<const char *>(base) + idx * 96;
    p.x = fp_load<FqP>(b, 0);
    p.y = fp_load<FqP>(b, 1);
    return p;
}
// gather of a base point that will not be touched again by this SM: bypass L1 (ld.global.cg) so that the streamed
// entry lists keep their L1 lines.  (Measured, no effect: the L2 fetch-size qualifiers on these loads -- ld.global.cg.L2::64B
// = LDG.E.LTC64B, L1::no_allocate.L2::64B, L2::256B -- give 119.9 / 120.0 / 120.5 ms of bucket accumulation per 2^20 proof
// against 120.0: the DRAM traffic of the random 96 B gathers stays at 128 B lines whatever the hint says.)
SCZ_D G1Affine g1a_load_stream(const void *base, size_t idx) {
    G1Affine p;
    const uint4 *q = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(base) + idx * 96);
    uint32_t *w = &p.x.l[0];
#pragma unroll
    for (int i = 0; i < 6; i++) {
        uint4 v = __ldcg(q + i);
        if (i == 3) w = &p.y.l[0] - 12;
        w[4 * i] = v.x, w[4 * i + 1] = v.y, w[4 * i + 2] = v.z, w[4 * i + 3] = v.w;
    }
    return p;
}
SCZ_D void g1a_store(void *base, size_t idx, const G1Affine &p) {
    char *b = reinterpret_cast<char *>(base) + idx * 96;
    fp_store<FqP>(b, 0, p.x);
    fp_store<FqP>(b, 1, p.y);
}
SCZ_D G1Jac g1j_load(const void *base, size_t idx) {
    G1Jac p;
    const char *b = reinterpret_cast<const char *>(base) + idx * 144;
    p.x = fp_load_rw<FqP>(b, 0);
    p.y = fp_load_rw<FqP>(b, 1);
    p.z = fp_load_rw<FqP>(b, 2);
    return p;
}
SCZ_D void g1j_store(void *base, size_t idx, const G1Jac &p) {
    char *b = reinterpret_cast<char *>(base) + idx * 144;
    fp_store<FqP>(b, 0, p.x);
    fp_store<FqP>(b, 1, p.y);
    fp_store<FqP>(b, 2, p.z);
}
SCZ_D G1X g1x_load(const void *base, size_t idx) {
    G1X p;
    const char *b = reinterpret_cast<const char *>(base) + idx * 192;
    p.x = fp_load_rw<FqP>(b, 0);
    p.y = fp_load_rw<FqP>(b, 1);
    p.zz = fp_load_rw<FqP>(b, 2);
    p.zzz = fp_load_rw<FqP>(b, 3);
    return p;
}
SCZ_D void g1x_store(void *base, size_t idx, const G1X &p) {
    char *b = reinterpret_cast<char *>(base) + idx * 192;
    fp_store<FqP>(b, 0, p.x);
    fp_store<FqP>(b, 1, p.y);
    fp_store<FqP>(b, 2, p.zz);
    fp_store<FqP>(b, 3, p.zzz);
}
#endif

}   // namespace scz
