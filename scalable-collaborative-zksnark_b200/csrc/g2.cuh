// BLS12-381 G2 for `d_msm` over G2 (the reference's `d_msm` is generic over `G: CurveGroup`,
// dist-primitive/src/dmsm.rs:9-15; G2 appears in the SRS, dpoly_comm.rs:27,59-62 -- no caller instantiates d_msm with
// it, so this path is built for completeness of the API, not tuned like G1).
//
// Fq2 = Fq[u] / (u^2 + 1), element c0 + c1 u stored c0 | c1 (ark-ff Fp2: 96 B).  G2: y^2 = x^3 + 4 (1 + u), a = 0, so
// every formula of g1.cuh carries over with Fq2 in place of Fq.
//   G2 affine    192 B  x | y, the identity is x = y = 0 (host entry points take ark-ec's `infinity` mask)
//   G2 Jacobian  288 B  X | Y | Z = ark-ec short_weierstrass::Projective<g2::Config>, identity has Z = 0
//   accumulators XYZZ (X, Y, ZZ, ZZZ) as in g1.cuh
#pragma once
#include "field.cuh"

namespace scz {

struct Fq2 {
    Fq c0, c1;
    SCZ_HD static Fq2 zero() {
        Fq2 r;
        r.c0 = Fq::zero();
        r.c1 = Fq::zero();
        return r;
    }
    SCZ_HD static Fq2 one() {
        Fq2 r;
        r.c0 = Fq::one();
        r.c1 = Fq::zero();
        return r;
    }
    SCZ_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    SCZ_HD bool operator==(const Fq2 &o) const { return c0 == o.c0 && c1 == o.c1; }
};
SCZ_HD Fq2 f2_add(const Fq2 &a, const Fq2 &b) {
    Fq2 r;
    r.c0 = fp_add(a.c0, b.c0);
    r.c1 = fp_add(a.c1, b.c1);
    return r;
}
SCZ_HD Fq2 f2_sub(const Fq2 &a, const Fq2 &b) {
    Fq2 r;
    r.c0 = fp_sub(a.c0, b.c0);
    r.c1 = fp_sub(a.c1, b.c1);
    return r;
}
SCZ_HD Fq2 f2_neg(const Fq2 &a) {
    Fq2 r;
    r.c0 = fp_neg(a.c0);
    r.c1 = fp_neg(a.c1);
    return r;
}
SCZ_HD Fq2 f2_dbl(const Fq2 &a) { return f2_add(a, a); }
// On the device every Fq product of the G2 formulas is ONE out-of-line function: a G2 addition is 14 Fq2 = 42 Fq
// products, inlined it would be a quarter of a megabyte of code and 255 registers with kilobytes of spills.
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ Fq g2_fq_mul(const Fq &a, const Fq &b) { return fp_mul(a, b); }
#else
inline Fq g2_fq_mul(const Fq &a, const Fq &b) { return fp_mul(a, b); }
#endif
// (a0 + a1 u)(b0 + b1 u) = (a0 b0 - a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) u : three Fq products
SCZ_HD Fq2 f2_mul(const Fq2 &a, const Fq2 &b) {
    Fq t0 = g2_fq_mul(a.c0, b.c0), t1 = g2_fq_mul(a.c1, b.c1);
    Fq t2 = g2_fq_mul(fp_add(a.c0, a.c1), fp_add(b.c0, b.c1));
    Fq2 r;
    r.c0 = fp_sub(t0, t1);
    r.c1 = fp_sub(fp_sub(t2, t0), t1);
    return r;
}
// (a0 + a1 u)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 u : two Fq products
SCZ_HD Fq2 f2_sqr(const Fq2 &a) {
    Fq2 r;
    r.c0 = g2_fq_mul(fp_add(a.c0, a.c1), fp_sub(a.c0, a.c1));
    r.c1 = fp_dbl(g2_fq_mul(a.c0, a.c1));
    return r;
}

struct G2Affine {   // 192 B
    Fq2 x, y;
    SCZ_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }
};
struct G2Jac {      // 288 B, ark-ec Projective
    Fq2 x, y, z;
};
struct G2X {        // XYZZ
    Fq2 x, y, zz, zzz;
    SCZ_HD bool is_inf() const { return zz.is_zero(); }
    SCZ_HD static G2X inf() {
        G2X r;
        r.x = Fq2::zero();
        r.y = Fq2::zero();
        r.zz = Fq2::zero();
        r.zzz = Fq2::zero();
        return r;
    }
};

SCZ_HD G2X g2x_from_jac(const G2Jac &p) {
    if (p.z.is_zero()) return G2X::inf();
    G2X r;
    r.x = p.x;
    r.y = p.y;
    r.zz = f2_sqr(p.z);
    r.zzz = f2_mul(r.zz, p.z);
    return r;
}
// (X, Y, ZZ, ZZZ) -> Jacobian with Z = ZZZ (g1.cuh g1x_to_jac)
SCZ_HD G2Jac g2x_to_jac(const G2X &p) {
    G2Jac r;
    if (p.is_inf()) {   // ark-ec's identity: (1, 1, 0)
        r.x = Fq2::one();
        r.y = Fq2::one();
        r.z = Fq2::zero();
        return r;
    }
    Fq2 zz2 = f2_sqr(p.zz), zzz2 = f2_sqr(p.zzz);
    r.x = f2_mul(p.x, zz2);
    r.y = f2_mul(p.y, zzz2);
    r.z = p.zzz;
    return r;
}
// dbl-2008-s-1 (a = 0)
SCZ_HD G2X g2x_double(const G2X &p) {
    if (p.is_inf()) return p;
    Fq2 u = f2_dbl(p.y);
    Fq2 v = f2_sqr(u);
    Fq2 w = f2_mul(u, v);
    Fq2 s = f2_mul(p.x, v);
    Fq2 xx = f2_sqr(p.x);
    Fq2 m = f2_add(f2_dbl(xx), xx);
    G2X r;
    r.x = f2_sub(f2_sub(f2_sqr(m), s), s);
    r.y = f2_sub(f2_mul(m, f2_sub(s, r.x)), f2_mul(w, p.y));
    r.zz = f2_mul(v, p.zz);
    r.zzz = f2_mul(w, p.zzz);
    return r;
}
// mdbl-2008-s-1: double an affine point
SCZ_HD G2X g2x_double_affine(const Fq2 &x, const Fq2 &y) {
    Fq2 u = f2_dbl(y);
    Fq2 v = f2_sqr(u);
    Fq2 w = f2_mul(u, v);
    Fq2 s = f2_mul(x, v);
    Fq2 xx = f2_sqr(x);
    Fq2 m = f2_add(f2_dbl(xx), xx);
    G2X r;
    r.x = f2_sub(f2_sub(f2_sqr(m), s), s);
    r.y = f2_sub(f2_mul(m, f2_sub(s, r.x)), f2_mul(w, y));
    r.zz = v;
    r.zzz = w;
    return r;
}
// madd-2008-s: acc += (x2, y2) affine, not the identity; every exceptional case handled
SCZ_HD void g2x_add_affine(G2X &acc, const Fq2 &x2, const Fq2 &y2) {
    if (acc.is_inf()) {
        acc.x = x2;
        acc.y = y2;
        acc.zz = Fq2::one();
        acc.zzz = Fq2::one();
        return;
    }
    Fq2 u2 = f2_mul(x2, acc.zz);
    Fq2 s2 = f2_mul(y2, acc.zzz);
    Fq2 p = f2_sub(u2, acc.x);
    Fq2 r = f2_sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = g2x_double_affine(x2, y2);
        else acc = G2X::inf();
        return;
    }
    Fq2 pp = f2_sqr(p);
    Fq2 ppp = f2_mul(p, pp);
    Fq2 q = f2_mul(acc.x, pp);
    Fq2 x3 = f2_sub(f2_sub(f2_sub(f2_sqr(r), ppp), q), q);
    Fq2 y3 = f2_sub(f2_mul(r, f2_sub(q, x3)), f2_mul(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = f2_mul(acc.zz, pp);
    acc.zzz = f2_mul(acc.zzz, ppp);
}
SCZ_HD void g2x_add_affine(G2X &acc, const G2Affine &p, bool negate) {
    if (p.is_inf()) return;
    g2x_add_affine(acc, p.x, negate ? f2_neg(p.y) : p.y);
}
// add-2008-s
SCZ_HD G2X g2x_add(const G2X &a, const G2X &b) {
    if (a.is_inf()) return b;
    if (b.is_inf()) return a;
    Fq2 u1 = f2_mul(a.x, b.zz);
    Fq2 u2 = f2_mul(b.x, a.zz);
    Fq2 s1 = f2_mul(a.y, b.zzz);
    Fq2 s2 = f2_mul(b.y, a.zzz);
    Fq2 p = f2_sub(u2, u1);
    Fq2 r = f2_sub(s2, s1);
    if (p.is_zero()) {
        if (r.is_zero()) return g2x_double(a);
        return G2X::inf();
    }
    Fq2 pp = f2_sqr(p);
    Fq2 ppp = f2_mul(p, pp);
    Fq2 q = f2_mul(u1, pp);
    G2X o;
    o.x = f2_sub(f2_sub(f2_sub(f2_sqr(r), ppp), q), q);
    o.y = f2_sub(f2_mul(r, f2_sub(q, o.x)), f2_mul(s1, ppp));
    o.zz = f2_mul(f2_mul(a.zz, b.zz), pp);
    o.zzz = f2_mul(f2_mul(a.zzz, b.zzz), ppp);
    return o;
}
// k * P, k = canonical little-endian 32-bit limbs (double-and-add, MSB first)
SCZ_HD G2X g2x_mul_bits(const G2X &p, const uint32_t (&k)[8]) {
    G2X acc = G2X::inf();
    for (int i = 255; i >= 0; i--) {
        acc = g2x_double(acc);
        if ((k[i >> 5] >> (i & 31)) & 1) acc = g2x_add(acc, p);
    }
    return acc;
}

#if defined(__CUDACC__)
SCZ_D Fq2 f2_load(const void *base, size_t idx) {   // idx in units of 96 B
    Fq2 r;
    const char *b = reinterpret_cast<const char *>(base) + idx * 96;
    r.c0 = fp_load_rw<FqP>(b, 0);
    r.c1 = fp_load_rw<FqP>(b, 1);
    return r;
}
SCZ_D void f2_store(void *base, size_t idx, const Fq2 &v) {
    char *b = reinterpret_cast<char *>(base) + idx * 96;
    fp_store<FqP>(b, 0, v.c0);
    fp_store<FqP>(b, 1, v.c1);
}
SCZ_D G2Affine g2a_load(const void *base, size_t idx) {
    G2Affine p;
    p.x = f2_load(base, 2 * idx);
    p.y = f2_load(base, 2 * idx + 1);
    return p;
}
SCZ_D void g2a_store(void *base, size_t idx, const G2Affine &p) {
    f2_store(base, 2 * idx, p.x);
    f2_store(base, 2 * idx + 1, p.y);
}
SCZ_D G2Jac g2j_load(const void *base, size_t idx) {
    G2Jac p;
    p.x = f2_load(base, 3 * idx);
    p.y = f2_load(base, 3 * idx + 1);
    p.z = f2_load(base, 3 * idx + 2);
    return p;
}
SCZ_D void g2j_store(void *base, size_t idx, const G2Jac &p) {
    f2_store(base, 3 * idx, p.x);
    f2_store(base, 3 * idx + 1, p.y);
    f2_store(base, 3 * idx + 2, p.z);
}
SCZ_D G2X g2x_load(const void *base, size_t idx) {
    G2X p;
    p.x = f2_load(base, 4 * idx);
    p.y = f2_load(base, 4 * idx + 1);
    p.zz = f2_load(base, 4 * idx + 2);
    p.zzz = f2_load(base, 4 * idx + 3);
    return p;
}
SCZ_D void g2x_store(void *base, size_t idx, const G2X &p) {
    f2_store(base, 4 * idx, p.x);
    f2_store(base, 4 * idx + 1, p.y);
    f2_store(base, 4 * idx + 2, p.zz);
    f2_store(base, 4 * idx + 3, p.zzz);
}
#endif

}   // namespace scz
