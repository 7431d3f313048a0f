// Prime-field arithmetic for BLS12-381 Fr (8 x u32) and Fq (12 x u32) on sm_100a.
//
// Elements live in HBM exactly as arkworks keeps them in host memory
// (ark-ff 0.4.2 Fp<MontBackend>: little-endian u64 limbs, Montgomery form,
// R = 2^256 / 2^384), so a `&[Fr]` / coordinate array from the reference's Rust
// side is consumed zero-copy.  Inside a thread an element is N 32-bit limbs in
// registers.
//
// The multiplier is a CIOS Montgomery product on split even/odd accumulator
// columns: every (mad.lo.cc, madc.hi.cc) pair lands on an aligned register pair
// so ptxas emits one IMAD.WIDE.U32(.X) per 32x32 partial product and the carry
// chains never need a separate propagate pass.  The kernels on this path are
// bound by that integer pipe, not by HBM.
//
// Every carry-chain primitive has a host emulation behind `#ifndef
// __CUDA_ARCH__` so tests/emu can check the limb logic on the CPU box that has
// no GPU; the product library only ever runs the device side.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SCZ_HD __host__ __device__ __forceinline__
#define SCZ_D __device__ __forceinline__
#else
#define SCZ_HD inline
#define SCZ_D inline
#endif

namespace scz {

// carry flag: lives in CC.CF on the device, in this struct on the host
struct CF {
    uint32_t v;
};

SCZ_HD uint32_t add_cc(CF &c, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t t = (uint64_t)a + b;
    c.v = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
SCZ_HD uint32_t addc_cc(CF &c, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t t = (uint64_t)a + b + c.v;
    c.v = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
SCZ_HD uint32_t addc(CF &c, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    return a + b + c.v;
#endif
}
SCZ_HD uint32_t sub_cc(CF &c, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t t = (uint64_t)a - b;
    c.v = (uint32_t)(t >> 63);   // borrow
    return (uint32_t)t;
#endif
}
SCZ_HD uint32_t subc_cc(CF &c, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#else
    uint64_t t = (uint64_t)a - b - c.v;
    c.v = (uint32_t)(t >> 63);
    return (uint32_t)t;
#endif
}
// 0 - borrow  -> 0xffffffff when the chain borrowed, else 0
SCZ_HD uint32_t borrow_mask(CF &c) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("subc.u32 %0, 0, 0;" : "=r"(r));
    return r;
#else
    return 0u - c.v;
#endif
}
SCZ_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
SCZ_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
SCZ_HD uint32_t mad_lo_cc(CF &c, uint32_t a, uint32_t b, uint32_t d) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(d));
    return r;
#else
    uint64_t t = (uint64_t)(uint32_t)(a * b) + d;
    c.v = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
SCZ_HD uint32_t madc_lo_cc(CF &c, uint32_t a, uint32_t b, uint32_t d) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(d));
    return r;
#else
    uint64_t t = (uint64_t)(uint32_t)(a * b) + d + c.v;
    c.v = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
SCZ_HD uint32_t madc_hi_cc(CF &c, uint32_t a, uint32_t b, uint32_t d) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(d));
    return r;
#else
    uint64_t t = (((uint64_t)a * b) >> 32) + d + c.v;
    c.v = (uint32_t)(t >> 32);
    return (uint32_t)t;
#endif
}
SCZ_HD uint32_t madc_hi(CF &c, uint32_t a, uint32_t b, uint32_t d) {
#ifdef __CUDA_ARCH__
    uint32_t r;
    asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(d));
    return r;
#else
    return (uint32_t)(((uint64_t)a * b) >> 32) + d + c.v;
#endif
}

// ---------------------------------------------------------------- field parameters
// BLS12-381 scalar field r (ark-bls12-381 0.4.0 Fr), 8 x u32 little endian
struct FrP {
    static constexpr int N = 8;
    static constexpr uint32_t INV = 0xffffffffu;   // -r^{-1} mod 2^32
    // The row multiplier m = e0 * INV mod 2^32 of a Montgomery step.  INV = -1 here, and when ptxas sees m = -e0 it
    // splits every m * p_j of the row into IMAD.X (low half) + IMAD.HI.U32.X (high half) -- 176 multiply-pipe
    // instructions per Fr product instead of the 128 IMAD.WIDE.U32(.X) it emits for Fq (seen in the SASS of every Fr
    // kernel).  An opaque subtraction keeps the fused form.
    SCZ_HD static uint32_t mont_m(uint32_t e0) {
#ifdef __CUDA_ARCH__
        uint32_t m;
        asm volatile("sub.u32 %0, 0, %1;" : "=r"(m) : "r"(e0));
        return m;
#else
        return e0 * INV;
#endif
    }
    SCZ_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t M[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u,
                                   0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
        return M[i];
    }
    SCZ_HD static constexpr uint32_t one(int i) {   // R mod r
        constexpr uint32_t M[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau,
                                   0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
        return M[i];
    }
    SCZ_HD static constexpr uint32_t r2(int i) {    // R^2 mod r
        constexpr uint32_t M[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu,
                                   0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
        return M[i];
    }
};
// BLS12-381 base field p, 12 x u32
struct FqP {
    static constexpr int N = 12;
    static constexpr uint32_t INV = 0xfffcfffdu;   // -p^{-1} mod 2^32
    SCZ_HD static constexpr uint32_t mont_m(uint32_t e0) { return e0 * INV; }
    SCZ_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t M[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                    0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
        return M[i];
    }
    SCZ_HD static constexpr uint32_t one(int i) {   // R mod p
        constexpr uint32_t M[12] = {0x0002fffdu, 0x76090000u, 0xc40c0002u, 0xebf4000bu, 0x53c758bau, 0x5f489857u,
                                    0x70525745u, 0x77ce5853u, 0xa256ec6du, 0x5c071a97u, 0xfa80e493u, 0x15f65ec3u};
        return M[i];
    }
    SCZ_HD static constexpr uint32_t r2(int i) {    // R^2 mod p
        constexpr uint32_t M[12] = {0x1c341746u, 0xf4df1f34u, 0x09d104f1u, 0x0a76e6a6u, 0x4c95b6d5u, 0x8de5476cu,
                                    0x939d83c0u, 0x67eb88a9u, 0xb519952du, 0x9a793e85u, 0x92cae3aau, 0x11988fe5u};
        return M[i];
    }
};

// ---------------------------------------------------------------- element type
template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t l[N];

    SCZ_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    SCZ_HD static Fp one() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::one(i);
        return r;
    }
    SCZ_HD static Fp rsquared() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r;
    }
    SCZ_HD bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i];
        return acc == 0;
    }
    SCZ_HD bool operator==(const Fp &o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i] ^ o.l[i];
        return acc == 0;
    }
    SCZ_HD bool operator!=(const Fp &o) const { return !(*this == o); }
};

// r = a - p if a >= p  (a < 2p)
template <class P>
SCZ_HD void fp_final_sub(Fp<P> &a) {
    constexpr int N = P::N;
    CF c{0};
    uint32_t t[N];
    t[0] = sub_cc(c, a.l[0], P::mod(0));
#pragma unroll
    for (int i = 1; i < N; i++) t[i] = subc_cc(c, a.l[i], P::mod(i));
    uint32_t borrow = borrow_mask(c);   // all-ones when a < p
#pragma unroll
    for (int i = 0; i < N; i++) a.l[i] = borrow ? a.l[i] : t[i];
}

template <class P>
SCZ_HD Fp<P> fp_add(const Fp<P> &a, const Fp<P> &b) {
    constexpr int N = P::N;
    Fp<P> r;
    CF c{0};
    r.l[0] = add_cc(c, a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(c, a.l[i], b.l[i]);
    r.l[N - 1] = addc(c, a.l[N - 1], b.l[N - 1]);   // p < 2^(32N-1): no carry out
    fp_final_sub(r);
    return r;
}
template <class P>
SCZ_HD Fp<P> fp_sub(const Fp<P> &a, const Fp<P> &b) {
    constexpr int N = P::N;
    Fp<P> r;
    CF c{0};
    r.l[0] = sub_cc(c, a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i++) r.l[i] = subc_cc(c, a.l[i], b.l[i]);
    uint32_t m = borrow_mask(c);
    r.l[0] = add_cc(c, r.l[0], P::mod(0) & m);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(c, r.l[i], P::mod(i) & m);
    r.l[N - 1] = addc(c, r.l[N - 1], P::mod(N - 1) & m);
    return r;
}
template <class P>
SCZ_HD Fp<P> fp_neg(const Fp<P> &a) {
    return fp_sub(Fp<P>::zero(), a);
}
template <class P>
SCZ_HD Fp<P> fp_dbl(const Fp<P> &a) {
    return fp_add(a, a);
}

// ---------------------------------------------------------------- Montgomery product
namespace detail {
template <int N>
SCZ_HD void mul_n(uint32_t *acc, const uint32_t *a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        // one 32 x 32 -> 64 multiply (mul.wide.u32 = a single IMAD.WIDE) instead of a lo / hi pair of instructions
        uint64_t w = (uint64_t)a[j] * bi;
        acc[j] = (uint32_t)w;
        acc[j + 1] = (uint32_t)(w >> 32);
    }
}
// acc[j],acc[j+1] += a[j]*bi over even j; leaves the carry out in CF
template <int N>
SCZ_HD void cmad_n(CF &c, uint32_t *acc, const uint32_t *a, uint32_t bi) {
    acc[0] = mad_lo_cc(c, a[0], bi, acc[0]);
    acc[1] = madc_hi_cc(c, a[0], bi, acc[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = madc_lo_cc(c, a[j], bi, acc[j]);
        acc[j + 1] = madc_hi_cc(c, a[j], bi, acc[j + 1]);
    }
}
// odd[j],odd[j+1] = a[j]*bi + odd[j+2],odd[j+3] + carry-in   (shift right by two limbs while accumulating)
template <int N>
SCZ_HD void madc_n_rshift(CF &c, uint32_t *odd, const uint32_t *a, uint32_t bi) {
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) {
        odd[j] = madc_lo_cc(c, a[j], bi, odd[j + 2]);
        odd[j + 1] = madc_hi_cc(c, a[j], bi, odd[j + 3]);
    }
    odd[N - 2] = madc_lo_cc(c, a[N - 2], bi, 0);
    odd[N - 1] = madc_hi(c, a[N - 2], bi, 0);
}
// same as cmad_n but the multiplicand is the (compile-time) modulus, offset `off` in {0,1}
template <class P, int off>
SCZ_HD void cmad_mod(CF &c, uint32_t *acc, uint32_t mi) {
    constexpr int N = P::N;
    if constexpr (P::mod(off) == 1u) {   // Fr: p_0 = 1, the product is mi itself (ptxas would still spend an IMAD.HI on its zero high half)
        acc[0] = add_cc(c, acc[0], mi);
        acc[1] = addc_cc(c, acc[1], 0);
    } else {
        acc[0] = mad_lo_cc(c, P::mod(off), mi, acc[0]);
        acc[1] = madc_hi_cc(c, P::mod(off), mi, acc[1]);
    }
#pragma unroll
    for (int j = 2; j < N; j += 2) {
        acc[j] = madc_lo_cc(c, P::mod(j + off), mi, acc[j]);
        acc[j + 1] = madc_hi_cc(c, P::mod(j + off), mi, acc[j + 1]);
    }
}
template <class P>
SCZ_HD void mad_n_redc(uint32_t *even, uint32_t *odd, const uint32_t *a, uint32_t bi, bool first) {
    constexpr int N = P::N;
    CF c{0};
    if (first) {
        mul_n<N>(odd, a + 1, bi);
        mul_n<N>(even, a, bi);
    } else {
        even[0] = add_cc(c, even[0], odd[1]);
        madc_n_rshift<N>(c, odd, a + 1, bi);
        cmad_n<N>(c, even, a, bi);
        odd[N - 1] = addc(c, odd[N - 1], 0);
    }
    uint32_t mi = P::mont_m(even[0]);
    cmad_mod<P, 1>(c, odd, mi);
    cmad_mod<P, 0>(c, even, mi);
    odd[N - 1] = addc(c, odd[N - 1], 0);
}
}   // namespace detail

template <class P>
SCZ_HD Fp<P> fp_mul(const Fp<P> &a, const Fp<P> &b) {
    constexpr int N = P::N;
    uint32_t even[N], odd[N];
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        detail::mad_n_redc<P>(even, odd, a.l, b.l[i], i == 0);
        detail::mad_n_redc<P>(odd, even, a.l, b.l[i + 1], false);
    }
    Fp<P> r;
    CF c{0};
    r.l[0] = add_cc(c, even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(c, even[i], odd[i + 1]);
    r.l[N - 1] = addc(c, even[N - 1], 0);
    fp_final_sub(r);
    return r;
}
// a1*b1 + a2*b2 (Montgomery) with ONE reduction pass: every row accumulates both partial products before the
// shared Montgomery step, 2*N^2 + N^2 wide multiplies instead of 4*N^2.  The running sum stays below 3p * 2^32,
// which needs p < 2^(32N - 2) (true for Fq: 381 of 384 bits; NOT for Fr: 255 of 256), so Fq only.
namespace detail {
template <class P>
SCZ_HD void mad2_n_redc(uint32_t *even, uint32_t *odd, const uint32_t *a1, uint32_t b1i, const uint32_t *a2, uint32_t b2i,
                        bool first) {
    constexpr int N = P::N;
    CF c{0};
    if (first) {
        mul_n<N>(odd, a1 + 1, b1i);
        mul_n<N>(even, a1, b1i);
    } else {
        even[0] = add_cc(c, even[0], odd[1]);
        madc_n_rshift<N>(c, odd, a1 + 1, b1i);
        cmad_n<N>(c, even, a1, b1i);
        odd[N - 1] = addc(c, odd[N - 1], 0);
    }
    cmad_n<N>(c, odd, a2 + 1, b2i);          // position 32N+32 and up stays empty: the sum is < 3p * 2^32 < 2^(32(N+1))
    cmad_n<N>(c, even, a2, b2i);
    odd[N - 1] = addc(c, odd[N - 1], 0);
    uint32_t mi = P::mont_m(even[0]);
    cmad_mod<P, 1>(c, odd, mi);
    cmad_mod<P, 0>(c, even, mi);
    odd[N - 1] = addc(c, odd[N - 1], 0);
}
}   // namespace detail
template <class P>
SCZ_HD Fp<P> fp_dot2(const Fp<P> &a1, const Fp<P> &b1, const Fp<P> &a2, const Fp<P> &b2) {
    constexpr int N = P::N;
    static_assert(P::N == 12, "fp_dot2 needs two spare bits above the modulus");
    uint32_t even[N], odd[N];
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        detail::mad2_n_redc<P>(even, odd, a1.l, b1.l[i], a2.l, b2.l[i], i == 0);
        detail::mad2_n_redc<P>(odd, even, a1.l, b1.l[i + 1], a2.l, b2.l[i + 1], false);
    }
    Fp<P> r;
    CF c{0};
    r.l[0] = add_cc(c, even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(c, even[i], odd[i + 1]);
    r.l[N - 1] = addc(c, even[N - 1], 0);
    fp_final_sub(r);   // r < 3p
    fp_final_sub(r);
    return r;
}
// a1*b1 - a2*b2
template <class P>
SCZ_HD Fp<P> fp_dot2_sub(const Fp<P> &a1, const Fp<P> &b1, const Fp<P> &a2, const Fp<P> &b2) {
    return fp_dot2(a1, b1, fp_neg(a2), b2);
}

// ---------------------------------------------------------------- squaring (Fq only)
// Operand scanning with the symmetric half of the partial products: 66 off-diagonal products accumulated on even / odd
// columns (so that every mad.lo.cc / madc.hi.cc pair is again one IMAD.WIDE), doubled, plus the 12 squares on the
// diagonal, then one Montgomery reduction of the 24-limb result: 78 + 144 = 222 wide multiplies instead of 288.
// (A squaring cannot share CIOS's short accumulator: the high products arrive early, the running sum needs all 2N limbs.)
namespace detail {
// t[0..2N) = a^2
template <int N>
SCZ_HD void sqr_wide(uint32_t *t, const uint32_t *a) {
    uint32_t ev[2 * N], od[2 * N];   // value = ev + (od << 32); od[k] sits at limb k + 1
#pragma unroll
    for (int k = 0; k < 2 * N; k++) ev[k] = od[k] = 0;
#pragma unroll
    for (int i = 0; i < N - 1; i++) {
        CF c{0};
        // products a_i * a_j with i + j odd (j = i+1, i+3, ...): limb i+j is od[i+j-1]
        {
            bool started = false;
            int top = 0;
#pragma unroll
            for (int j = i + 1; j < N; j += 2) {
                int k = i + j - 1;
                od[k] = started ? madc_lo_cc(c, a[i], a[j], od[k]) : mad_lo_cc(c, a[i], a[j], od[k]);
                od[k + 1] = madc_hi_cc(c, a[i], a[j], od[k + 1]);
                started = true;
                top = k + 2;
            }
            if (started && top < 2 * N) od[top] = addc(c, od[top], 0);
        }
        // products with i + j even (j = i+2, i+4, ...): limb i+j is ev[i+j]
        {
            bool started = false;
            int top = 0;
#pragma unroll
            for (int j = i + 2; j < N; j += 2) {
                int k = i + j;
                ev[k] = started ? madc_lo_cc(c, a[i], a[j], ev[k]) : mad_lo_cc(c, a[i], a[j], ev[k]);
                ev[k + 1] = madc_hi_cc(c, a[i], a[j], ev[k + 1]);
                started = true;
                top = k + 2;
            }
            if (started && top < 2 * N) ev[top] = addc(c, ev[top], 0);
        }
    }
    // t = ev + (od << 32)
    {
        CF c{0};
        t[0] = ev[0];
        t[1] = add_cc(c, ev[1], od[0]);
#pragma unroll
        for (int k = 2; k < 2 * N - 1; k++) t[k] = addc_cc(c, ev[k], od[k - 1]);
        t[2 * N - 1] = addc(c, ev[2 * N - 1], od[2 * N - 2]);
    }
    // t = 2 t (the off-diagonal sum is below 2^(64N - 1))
#pragma unroll
    for (int k = 2 * N - 1; k > 0; k--) t[k] = (t[k] << 1) | (t[k - 1] >> 31);
    t[0] <<= 1;
    // + the diagonal a_i^2 at limb 2i: one carry chain over all 2N limbs
    {
        CF c{0};
        t[0] = mad_lo_cc(c, a[0], a[0], t[0]);
        t[1] = madc_hi_cc(c, a[0], a[0], t[1]);
#pragma unroll
        for (int i = 1; i < N; i++) {
            t[2 * i] = madc_lo_cc(c, a[i], a[i], t[2 * i]);
            if (i < N - 1) t[2 * i + 1] = madc_hi_cc(c, a[i], a[i], t[2 * i + 1]);
            else t[2 * i + 1] = madc_hi(c, a[i], a[i], t[2 * i + 1]);
        }
    }
}
// one Montgomery step on the (even, odd) pair without a multiplicand row: the value is divided by 2^32
template <class P>
SCZ_HD void redc_row(uint32_t *even, uint32_t *odd, bool first) {
    constexpr int N = P::N;
    CF c{0};
    if (!first) {
        even[0] = add_cc(c, even[0], odd[1]);
#pragma unroll
        for (int j = 0; j < N - 2; j++) odd[j] = addc_cc(c, odd[j + 2], 0);   // shift right by two limbs, carry rides along
        odd[N - 2] = addc(c, 0, 0);
        odd[N - 1] = 0;
    }
    uint32_t mi = P::mont_m(even[0]);
    cmad_mod<P, 1>(c, odd, mi);
    cmad_mod<P, 0>(c, even, mi);
    odd[N - 1] = addc(c, odd[N - 1], 0);
}
}   // namespace detail
// REDC of a 2N-limb value t < p * 2^(32N): t / 2^(32N) mod p
template <class P>
SCZ_HD Fp<P> fp_redc_wide(const uint32_t *t) {
    constexpr int N = P::N;
    uint32_t even[N], odd[N];
#pragma unroll
    for (int k = 0; k < N; k++) {
        even[k] = t[k];
        odd[k] = 0;
    }
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        detail::redc_row<P>(even, odd, i == 0);
        detail::redc_row<P>(odd, even, false);
    }
    // low half reduced (< 2p after merging), plus the untouched high half
    Fp<P> r;
    CF c{0};
    r.l[0] = add_cc(c, even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(c, even[i], odd[i + 1]);
    r.l[N - 1] = addc(c, even[N - 1], 0);
    r.l[0] = add_cc(c, r.l[0], t[N]);
#pragma unroll
    for (int i = 1; i < N - 1; i++) r.l[i] = addc_cc(c, r.l[i], t[N + i]);
    r.l[N - 1] = addc(c, r.l[N - 1], t[2 * N - 1]);
    fp_final_sub(r);   // < 2p + p/8
    fp_final_sub(r);
    return r;
}
// ---------------------------------------------------------------- unreduced sums of products
// A sum of Montgomery products sum_i a_i * b_i needs ONE Montgomery reduction, not one per term: the plain 2N-limb
// integer products are added up in a (2N + 1)-limb accumulator (N^2 wide multiplies per term instead of 2 N^2) and the
// accumulator is reduced once at the end.  Used by the product sumcheck rounds (poly.cu), whose three round sums are
// exactly such sums (dsumcheck.rs:37-85).  Operands < p, at most 2^32 terms.
namespace detail {
// ev + (od << 32) = a * b: products a_j b_i with i + j even go to the `ev` columns, the others to `od` (od[k] sits at
// limb k + 1), so that every mad.lo.cc / madc.hi.cc pair is one IMAD.WIDE on an aligned register pair (as in the CIOS rows)
template <int N>
SCZ_HD void mul_wide(uint32_t *ev, uint32_t *od, const uint32_t *a, const uint32_t *b) {
    static_assert(N % 2 == 0, "even limb count");
#pragma unroll
    for (int k = 0; k < 2 * N; k++) ev[k] = od[k] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) {
        {
            CF c{0};
            int top = 0;
#pragma unroll
            for (int j = (i & 1); j < N; j += 2) {
                const int k = i + j;
                ev[k] = j == (i & 1) ? mad_lo_cc(c, a[j], b[i], ev[k]) : madc_lo_cc(c, a[j], b[i], ev[k]);
                ev[k + 1] = madc_hi_cc(c, a[j], b[i], ev[k + 1]);
                top = k + 2;
            }
            if (top < 2 * N) ev[top] = addc(c, ev[top], 0);
        }
        {
            CF c{0};
            int top = 0;
#pragma unroll
            for (int j = 1 - (i & 1); j < N; j += 2) {
                const int k = i + j - 1;
                od[k] = j == 1 - (i & 1) ? mad_lo_cc(c, a[j], b[i], od[k]) : madc_lo_cc(c, a[j], b[i], od[k]);
                od[k + 1] = madc_hi_cc(c, a[j], b[i], od[k + 1]);
                top = k + 2;
            }
            if (top < 2 * N) od[top] = addc(c, od[top], 0);
        }
    }
}
}   // namespace detail
// acc[0 .. 2N] += a * b
template <class P>
SCZ_HD void fp_mul_acc_wide(uint32_t *acc, const Fp<P> &a, const Fp<P> &b) {
    constexpr int N = P::N;
    uint32_t ev[2 * N], od[2 * N];
    detail::mul_wide<N>(ev, od, a.l, b.l);
    CF c{0};
    acc[0] = add_cc(c, acc[0], ev[0]);
#pragma unroll
    for (int k = 1; k < 2 * N; k++) acc[k] = addc_cc(c, acc[k], ev[k]);
    acc[2 * N] = addc(c, acc[2 * N], 0);
    acc[1] = add_cc(c, acc[1], od[0]);
#pragma unroll
    for (int k = 1; k < 2 * N - 1; k++) acc[k + 1] = addc_cc(c, acc[k + 1], od[k]);   // od[2N - 1] is never written
    acc[2 * N] = addc(c, acc[2 * N], 0);
}
// the accumulator as a field element: acc / 2^(32N) mod p (= the sum of the Montgomery products).  acc = H 2^(64N) + L:
// the high half of L is brought below p by conditional subtractions (2^(32N) / p < 3 for Fr, < 10 for Fq) so that
// fp_redc_wide applies; H 2^(64N) / 2^(32N) = H 2^(32N) = the Montgomery form of the integer H.
template <class P>
SCZ_HD Fp<P> fp_acc_wide_reduce(const uint32_t *acc) {
    constexpr int N = P::N;
    uint32_t t[2 * N];
    Fp<P> hi;
#pragma unroll
    for (int k = 0; k < N; k++) {
        t[k] = acc[k];
        hi.l[k] = acc[N + k];
    }
    // hi < 2^(32N): floor(2^(32N) / p) conditional subtractions (full-width compare: hi may have its top bit set)
    constexpr int REPS = (int)(0xffffffffu / P::mod(N - 1));
#pragma unroll
    for (int rep = 0; rep < REPS; rep++) {
        CF c{0};
        uint32_t d[N];
        d[0] = sub_cc(c, hi.l[0], P::mod(0));
#pragma unroll
        for (int k = 1; k < N; k++) d[k] = subc_cc(c, hi.l[k], P::mod(k));
        uint32_t borrow = borrow_mask(c);
#pragma unroll
        for (int k = 0; k < N; k++) hi.l[k] = borrow ? hi.l[k] : d[k];
    }
#pragma unroll
    for (int k = 0; k < N; k++) t[N + k] = hi.l[k];
    Fp<P> lo = fp_redc_wide<P>(t);
    Fp<P> h = Fp<P>::zero();
    h.l[0] = acc[2 * N];
    return fp_add(lo, fp_mul(h, Fp<P>::rsquared()));
}

SCZ_HD Fp<FqP> fq_sqr_sos(const Fp<FqP> &a) {
    uint32_t t[24];
    detail::sqr_wide<12>(t, a.l);
    return fp_redc_wide<FqP>(t);
}

template <class P>
SCZ_HD Fp<P> fp_sqr(const Fp<P> &a) {
    return fp_mul(a, a);
}
// Measured on B200 (dhyperplonk 2^20): with fq_sqr_sos in the group law the bucket kernel takes 132.9 ms instead of
// 128.9 ms -- 66 fewer wide multiplies per squaring do not pay for the longer carry / shift chains and 14 more
// registers -- so squarings stay fp_mul(a, a).  -DSCZ_FQ_SQR_SOS switches it on for experiments.
#ifdef SCZ_FQ_SQR_SOS
SCZ_HD Fp<FqP> fp_sqr(const Fp<FqP> &a) { return fq_sqr_sos(a); }   // preferred over the template for Fq
#endif
// Montgomery form -> canonical integer (ark-ff into_bigint)
template <class P>
SCZ_HD Fp<P> fp_to_canon(const Fp<P> &a) {
    Fp<P> one = Fp<P>::zero();
    one.l[0] = 1;
    return fp_mul(a, one);
}
template <class P>
SCZ_HD Fp<P> fp_from_canon(const Fp<P> &a) {
    return fp_mul(a, Fp<P>::rsquared());
}
// a^e for a public exponent given as canonical 32-bit limbs (square-and-multiply, MSB first)
template <class P, int EL>
SCZ_HD Fp<P> fp_pow(const Fp<P> &a, const uint32_t (&e)[EL]) {
    Fp<P> acc = Fp<P>::one();
    for (int i = EL * 32 - 1; i >= 0; i--) {
        acc = fp_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fp_mul(acc, a);
    }
    return acc;
}
// Fermat inverse a^(p-2); 0 -> 0
template <class P>
SCZ_HD Fp<P> fp_inv(const Fp<P> &a) {
    constexpr int N = P::N;
    uint32_t e[N];
    uint32_t borrow = 2;   // e = p - 2 with borrow propagation (r ends in ...00000001)
#pragma unroll
    for (int i = 0; i < N; i++) {
        uint32_t m = P::mod(i);
        e[i] = m - borrow;
        borrow = m < borrow ? 1u : 0u;
    }
    Fp<P> acc = Fp<P>::one();
    for (int i = N * 32 - 1; i >= 0; i--) {
        acc = fp_sqr(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fp_mul(acc, a);
    }
    return acc;
}

// Inverse by the binary extended Euclidean algorithm (shifts, additions, subtractions: no multiplier), for the places
// where ONE inversion sits on the critical path of a whole launch sequence (the top of the batched-affine product
// tree, msm_affine.cu): a Fermat chain is ~570 dependent Montgomery products (~0.6 ms on one thread), this is ~760
// halving / subtraction steps of 12-limb integer arithmetic (~0.1 ms).  a in Montgomery form, a != 0; 0 -> 0.
//   invariants: x1 * v == u, x2 * v == w (mod p), v = the stored integer a R; ends with u == 1 or w == 1
//   result a^-1 R = (v^-1) * R^2 = montmul(montmul(v^-1, R^2), R^2)
template <class P>
SCZ_HD Fp<P> fp_inv_bingcd(const Fp<P> &a) {
    constexpr int N = P::N;
    if (a.is_zero()) return a;
    uint32_t u[N], w[N], x1[N], x2[N];
#pragma unroll
    for (int i = 0; i < N; i++) {
        u[i] = a.l[i];
        w[i] = P::mod(i);
        x1[i] = i == 0 ? 1u : 0u;
        x2[i] = 0;
    }
    auto is_one = [](const uint32_t *t) {
        uint32_t acc = t[0] ^ 1u;
#pragma unroll
        for (int i = 1; i < N; i++) acc |= t[i];
        return acc == 0;
    };
    auto halve = [](uint32_t *t, uint32_t *x) {   // t even: t /= 2, x /= 2 (mod p)
#pragma unroll
        for (int i = 0; i < N - 1; i++) t[i] = (t[i] >> 1) | (t[i + 1] << 31);
        t[N - 1] >>= 1;
        uint32_t m = 0u - (x[0] & 1u);            // odd: add p first (x + p < 2^(32 N): p < 2^(32 N - 1))
        CF c{0};
        x[0] = add_cc(c, x[0], P::mod(0) & m);
#pragma unroll
        for (int i = 1; i < N - 1; i++) x[i] = addc_cc(c, x[i], P::mod(i) & m);
        x[N - 1] = addc(c, x[N - 1], P::mod(N - 1) & m);
#pragma unroll
        for (int i = 0; i < N - 1; i++) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
        x[N - 1] >>= 1;
    };
    auto sub_mod = [](uint32_t *x, const uint32_t *y) {   // x = x - y mod p, x, y < p
        CF c{0};
        x[0] = sub_cc(c, x[0], y[0]);
#pragma unroll
        for (int i = 1; i < N; i++) x[i] = subc_cc(c, x[i], y[i]);
        uint32_t m = borrow_mask(c);
        x[0] = add_cc(c, x[0], P::mod(0) & m);
#pragma unroll
        for (int i = 1; i < N - 1; i++) x[i] = addc_cc(c, x[i], P::mod(i) & m);
        x[N - 1] = addc(c, x[N - 1], P::mod(N - 1) & m);
    };
    while (!is_one(u) && !is_one(w)) {
        while (!(u[0] & 1u)) halve(u, x1);
        while (!(w[0] & 1u)) halve(w, x2);
        // u >= w ?
        uint32_t t[N];
        CF c{0};
        t[0] = sub_cc(c, u[0], w[0]);
#pragma unroll
        for (int i = 1; i < N; i++) t[i] = subc_cc(c, u[i], w[i]);
        uint32_t lt = borrow_mask(c);
        if (!lt) {
#pragma unroll
            for (int i = 0; i < N; i++) u[i] = t[i];
            sub_mod(x1, x2);
        } else {
            CF c2{0};
            w[0] = sub_cc(c2, w[0], u[0]);
#pragma unroll
            for (int i = 1; i < N; i++) w[i] = subc_cc(c2, w[i], u[i]);
            sub_mod(x2, x1);
        }
    }
    Fp<P> r;
    const bool from_u = is_one(u);
#pragma unroll
    for (int i = 0; i < N; i++) r.l[i] = from_u ? x1[i] : x2[i];
    r = fp_mul(r, Fp<P>::rsquared());
    return fp_mul(r, Fp<P>::rsquared());
}

using Fr = Fp<FrP>;
using Fq = Fp<FqP>;

// ---------------------------------------------------------------- HBM access (128-bit loads/stores)
#if defined(__CUDACC__)
template <class P>
SCZ_D Fp<P> fp_load(const void *base, size_t idx) {
    constexpr int N = P::N;
    const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(base) + idx * (N * 4));
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < N / 4; i++) {
        uint4 v = __ldg(p + i);
        r.l[4 * i] = v.x;
        r.l[4 * i + 1] = v.y;
        r.l[4 * i + 2] = v.z;
        r.l[4 * i + 3] = v.w;
    }
    return r;
}
template <class P>
SCZ_D Fp<P> fp_load_rw(const void *base, size_t idx) {   // plain (coherent) load for buffers written in-kernel
    constexpr int N = P::N;
    const uint4 *p = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(base) + idx * (N * 4));
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < N / 4; i++) {
        uint4 v = p[i];
        r.l[4 * i] = v.x;
        r.l[4 * i + 1] = v.y;
        r.l[4 * i + 2] = v.z;
        r.l[4 * i + 3] = v.w;
    }
    return r;
}
template <class P>
SCZ_D void fp_store(void *base, size_t idx, const Fp<P> &a) {
    constexpr int N = P::N;
    uint4 *p = reinterpret_cast<uint4 *>(reinterpret_cast<char *>(base) + idx * (N * 4));
#pragma unroll
    for (int i = 0; i < N / 4; i++) p[i] = make_uint4(a.l[4 * i], a.l[4 * i + 1], a.l[4 * i + 2], a.l[4 * i + 3]);
}
#endif

}   // namespace scz
