// Lane-cooperative G1 arithmetic for the serial chains of the path.
//
// Two places on the hot path are inherently sequential in the group law: the Horner
// recombination of the Pippenger windows (255 doublings, msm.cu step 8) and the
// scalar multiplications of the PSS maps over G1 (the leader closure of d_msm,
// dist-primitive/src/dmsm.rs:31-38, which the reference runs as FFTs over points).
// A single thread spends ~0.8 us per Fq product (an IMAD.WIDE holds the multiply
// pipe for 4 cycles whatever the number of active lanes), so the only cheap
// parallelism left is ACROSS LANES: a group of 4 adjacent lanes keeps one Jacobian
// point replicated in registers, each lane evaluates a different field product of
// the addition formula in the same instruction, and the products are exchanged by
// warp shuffles.  A doubling is 3 product levels instead of 7 products, a general
// addition 5 levels instead of 16.
//
// All 4 lanes of a group must call these functions together (control flow inside
// depends only on replicated values, so a group never diverges; different groups of
// a warp may).
#pragma once
#include "g1.cuh"

namespace scz {

struct Coop {
    unsigned mask;
    int base, role;
    __device__ __forceinline__ Coop() {
        int lane = threadIdx.x & 31;
        base = lane & ~3;
        role = lane & 3;
        mask = 0xFu << base;
    }
    __device__ __forceinline__ Fq bcast(const Fq &v, int src) const {
        Fq r;
#pragma unroll
        for (int i = 0; i < 12; i++) r.l[i] = __shfl_sync(mask, v.l[i], base + src);
        return r;
    }
};

__device__ __forceinline__ Fq fq_sel(bool c, const Fq &a, const Fq &b) {
    Fq r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = c ? a.l[i] : b.l[i];
    return r;
}
__device__ __forceinline__ G1Jac g1j_inf() {   // ark-ec's identity (1, 1, 0)
    G1Jac r;
    r.x = Fq::one();
    r.y = Fq::one();
    r.z = Fq::zero();
    return r;
}

// dbl-2009-l (a = 0): A=X^2 B=Y^2 C=B^2 D=2((X+B)^2-A-C) E=3A F=E^2  X3=F-2D  Y3=E(D-X3)-8C  Z3=2YZ
static __device__ __noinline__ void coop_double(const Coop &g, G1Jac &p) {
    if (p.z.is_zero()) return;
    // level 1: X*X | Y*Y | Y*Z | (Y*Z)
    Fq a1 = fq_sel(g.role == 0, p.x, p.y);
    Fq b1 = g.role == 0 ? p.x : (g.role == 1 ? p.y : p.z);
    Fq r1 = fp_mul(a1, b1);
    Fq A = g.bcast(r1, 0), B = g.bcast(r1, 1), YZ = g.bcast(r1, 2);
    // level 2: (X+B)^2 | B^2 | E^2 | (E^2)
    Fq E = fp_add(fp_dbl(A), A);
    Fq t = fp_add(p.x, B);
    Fq a2 = g.role == 0 ? t : (g.role == 1 ? B : E);
    Fq r2 = fp_mul(a2, a2);
    Fq T = g.bcast(r2, 0), C = g.bcast(r2, 1), F = g.bcast(r2, 2);
    Fq D = fp_dbl(fp_sub(fp_sub(T, A), C));
    Fq X3 = fp_sub(F, fp_dbl(D));
    // level 3 (every lane computes it: the state stays replicated without a shuffle)
    Fq C8 = fp_dbl(fp_dbl(fp_dbl(C)));
    p.y = fp_sub(fp_mul(E, fp_sub(D, X3)), C8);
    p.x = X3;
    p.z = fp_dbl(YZ);
}

// add-2007-bl.  a += b.
static __device__ __noinline__ void coop_add(const Coop &g, G1Jac &a, const G1Jac &b) {
    if (b.z.is_zero()) return;
    if (a.z.is_zero()) {
        a = b;
        return;
    }
    // level 1: Z1^2 | Z2^2 | (Z1+Z2)^2 | (Z1^2)
    Fq zs = fp_add(a.z, b.z);
    Fq o1 = g.role == 1 ? b.z : (g.role == 2 ? zs : a.z);
    Fq r1 = fp_mul(o1, o1);
    Fq Z1Z1 = g.bcast(r1, 0), Z2Z2 = g.bcast(r1, 1), ZS2 = g.bcast(r1, 2);
    // level 2: X1*Z2Z2 | X2*Z1Z1 | Z2*Z2Z2 | Z1*Z1Z1
    Fq a2 = g.role == 0 ? a.x : (g.role == 1 ? b.x : (g.role == 2 ? b.z : a.z));
    Fq b2 = (g.role == 0 || g.role == 2) ? Z2Z2 : Z1Z1;
    Fq r2 = fp_mul(a2, b2);
    Fq U1 = g.bcast(r2, 0), U2 = g.bcast(r2, 1), Z2C = g.bcast(r2, 2), Z1C = g.bcast(r2, 3);
    Fq H = fp_sub(U2, U1);
    // level 3: Y1*Z2^3 | Y2*Z1^3 | (2H)^2 | ((Z1+Z2)^2 - Z1Z1 - Z2Z2) * H
    Fq H2 = fp_dbl(H);
    Fq ZZ = fp_sub(fp_sub(ZS2, Z1Z1), Z2Z2);
    Fq a3 = g.role == 0 ? a.y : (g.role == 1 ? b.y : (g.role == 2 ? H2 : ZZ));
    Fq b3 = g.role == 0 ? Z2C : (g.role == 1 ? Z1C : (g.role == 2 ? H2 : H));
    Fq r3 = fp_mul(a3, b3);
    Fq S1 = g.bcast(r3, 0), S2 = g.bcast(r3, 1), I = g.bcast(r3, 2), Z3 = g.bcast(r3, 3);
    if (H.is_zero()) {   // same x: P + P or P + (-P)
        if (S1 == S2) coop_double(g, a);
        else a = g1j_inf();
        return;
    }
    Fq rr = fp_dbl(fp_sub(S2, S1));
    // level 4: H*I | U1*I | r^2 | (r^2)
    Fq a4 = g.role == 0 ? H : (g.role == 1 ? U1 : rr);
    Fq b4 = g.role <= 1 ? I : rr;
    Fq r4 = fp_mul(a4, b4);
    Fq J = g.bcast(r4, 0), V = g.bcast(r4, 1), R2 = g.bcast(r4, 2);
    Fq X3 = fp_sub(fp_sub(R2, J), fp_dbl(V));
    // level 5: r*(V - X3) | S1*J
    Fq a5 = g.role == 1 ? S1 : rr;
    Fq b5 = g.role == 1 ? J : fp_sub(V, X3);
    Fq r5 = fp_mul(a5, b5);
    Fq Y3a = g.bcast(r5, 0), Y3b = g.bcast(r5, 1);
    a.x = X3;
    a.y = fp_sub(Y3a, fp_dbl(Y3b));
    a.z = Z3;
}

// XYZZ (x = X/ZZ, y = Y/ZZZ) -> Jacobian with Z = ZZZ: X' = X*ZZ^2, Y' = Y*ZZZ^2  (ZZ^3 = ZZZ^2).  Per thread.
__device__ __forceinline__ G1Jac g1x_to_jac_nl(const G1X &p) { return g1x_to_jac(p); }

// acc = k * P for a canonical 256-bit scalar k (little-endian 32-bit limbs), 4-bit fixed windows.
// tab: 16 Jacobian slots private to the group (shared or global memory); every lane of the group writes the
// same values.  ~64 * (4 doublings + 1 addition) = 1.1k product levels instead of ~4k serial products.
// The operand is taken from ONE copy in the group's (otherwise unused) slot 0: the cooperative routines rely on the
// group's lanes holding a bit-identical operand, and a value each lane loaded by itself from global memory was seen
// to break that on sm_100a / nvcc 12.9 in k_pss_dmsm_multi<2> (pss.cu; root cause not isolated, memcheck / synccheck
// clean), so every entry point goes through the same canonicalisation (tests/test_gpu_msm.py::test_d_msm_closure_*).
__device__ __forceinline__ G1Jac coop_mul_bits(const Coop &g, const G1Jac &p_in, const uint32_t (&k)[8], G1Jac *tab) {
    if (g.role == 0) tab[0] = p_in;
    __syncwarp(g.mask);
    const G1Jac p = tab[0];
    G1Jac t = p;
    tab[1] = t;
    for (int j = 2; j < 16; j++) {
        if (j & 1) {
            t = tab[j - 1];
            coop_add(g, t, p);
        } else {
            t = tab[j >> 1];
            coop_double(g, t);
        }
        tab[j] = t;
        __syncwarp(g.mask);
    }
    G1Jac acc = g1j_inf();
    for (int w = 63; w >= 0; w--) {
        if (!acc.z.is_zero()) {
            coop_double(g, acc);
            coop_double(g, acc);
            coop_double(g, acc);
            coop_double(g, acc);
        }
        uint32_t d = (k[w >> 3] >> ((w & 7) * 4)) & 15u;
        if (d) {
            G1Jac q = tab[d];
            coop_add(g, acc, q);
        }
    }
    return acc;
}

}   // namespace scz
