// Fq Montgomery product on the FP64 pipe (sm_100a: DFMA issues at ~2x the rate of IMAD.WIDE and the pipe is idle
// in every kernel of this path).  Same contract as fp_mul<FqP> (csrc/field.cuh): inputs and output are 12 x u32
// Montgomery residues with R = 2^384, output fully reduced, so the two multipliers can be mixed freely inside one
// group-law formula (g1.cuh, MulHybrid) and the bucket kernel keeps both pipes busy.
//
// Representation inside the product: 16 limbs of 24 bits held as doubles (16 * 24 = 384: the same R).  A partial
// product is < 2^48 and a column collects at most 16 (a*b) + 16 (m*p) of them plus a carry < 2^29: every column sum
// is an integer below 2^53, so plain DFMA chains are exact -- no hi / lo splitting, no rounding-mode tricks.
// Word-serial Montgomery: after row i, m_i = (column_i mod 2^24) * (-p^-1) mod 2^24 clears column i, whose upper part
// is carried into column i + 1 with one more DFMA.  512 DFMA + 16 carries per product; the m_i chain runs either on the
// FP64 pipe (round-to-multiple-of-2^24 by adding and subtracting 1.5 * 2^76: balanced digits, 7 DADD / DMUL per row)
// or through the conversion unit (F2I, one 32-bit IMAD, I2F: nothing on the FP64 pipe).
//
// Host emulation: all arithmetic is exact in IEEE double, so the g++ build of tests/emu runs the very same code.
#pragma once
#include <math.h>

#include "../../scalable-collaborative-zksnark_b200/csrc/field.cuh"

namespace scz {
namespace f64 {

constexpr double TWO24 = 16777216.0;
constexpr double INV24 = 1.0 / 16777216.0;
constexpr double ROUND24 = 1.5 * 4503599627370496.0 * 16777216.0;   // 1.5 * 2^76: ulp = 2^24
constexpr uint32_t NPRIME24 = FqP::INV & 0xffffffu;                  // -p^-1 mod 2^24

// 24-bit limb i of a 12-word little-endian integer
SCZ_HD constexpr uint32_t limb24_const(int i) {
    int w = (24 * i) / 32, s = (24 * i) % 32;
    uint64_t lo = FqP::mod(w);
    uint64_t hi = (w + 1 < 12) ? FqP::mod(w + 1) : 0;
    return (uint32_t)((((hi << 32) | lo) >> s) & 0xffffffu);
}
template <int I>
SCZ_HD uint32_t limb24(const uint32_t *l) {
    constexpr int w = (24 * I) / 32, s = (24 * I) % 32;
    if (s == 0) return l[w] & 0xffffffu;
    if (s == 8) return l[w] >> 8;
#ifdef __CUDA_ARCH__
    return __funnelshift_r(l[w], l[w + 1], s) & 0xffffffu;
#else
    return (uint32_t)(((((uint64_t)l[w + 1]) << 32) | l[w]) >> s) & 0xffffffu;
#endif
}
SCZ_HD double u2d(uint32_t x) {
#ifdef __CUDA_ARCH__
    return __uint2double_rn(x);
#else
    return (double)x;
#endif
}
SCZ_HD double i2d(int32_t x) {
#ifdef __CUDA_ARCH__
    return __int2double_rn(x);
#else
    return (double)x;
#endif
}
SCZ_HD int64_t d2ll(double x) {
#ifdef __CUDA_ARCH__
    return __double2ll_rn(x);
#else
    return (int64_t)x;
#endif
}
SCZ_HD double dfma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
SCZ_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);   // never contracted, never re-associated
#else
    volatile double r = a + b;
    return r;
#endif
}

// 16 doubles = the 24-bit limbs of a 12-word element
struct FqD {
    double d[16];
};
namespace detail {
template <int I>
SCZ_HD void to_d(const uint32_t *l, double *d) {
    if constexpr (I < 16) {
        d[I] = u2d(limb24<I>(l));
        to_d<I + 1>(l, d);
    }
}
// the row multiplier that clears column value x (an integer, |x| < 2^53)
template <int MCHAIN>
SCZ_HD double mont_m(double x) {
    if (MCHAIN == 0) {
        double hi = dadd(dadd(x, ROUND24), -ROUND24);
        double lo = dadd(x, -hi);                       // balanced: |lo| <= 2^23
        double q = lo * (double)NPRIME24;               // exact, |q| < 2^47
        double qh = dadd(dadd(q, ROUND24), -ROUND24);
        return dadd(q, -qh);                            // |m| <= 2^23
    } else {
        uint32_t lo = (uint32_t)(uint64_t)d2ll(x);      // x >= 0 on this variant
        return u2d((lo * NPRIME24) & 0xffffffu);
    }
}
// one row of the word-serial product: columns I .. I+15 += a_I * b + m_I * p, column I carried into I + 1
template <int MCHAIN, int I>
SCZ_HD void row1(double *acc, const uint32_t *a, const double *bd) {
    double ai = u2d(limb24<I>(a));
#pragma unroll
    for (int j = 0; j < 16; j++) acc[I + j] = dfma(ai, bd[j], acc[I + j]);
    double m = mont_m<MCHAIN>(acc[I]);
#pragma unroll
    for (int j = 0; j < 16; j++) acc[I + j] = dfma(m, (double)limb24_const(j), acc[I + j]);
    acc[I + 1] = dfma(acc[I], INV24, acc[I + 1]);   // column I is now a multiple of 2^24
}
template <int MCHAIN, int I>
SCZ_HD void rows(double *acc, const uint32_t *a, const double *bd) {
    if constexpr (I < 16) {
        row1<MCHAIN, I>(acc, a, bd);
        rows<MCHAIN, I + 1>(acc, a, bd);
    }
}
// columns 16..31 -> 12 words, fully reduced
template <int MCHAIN>
SCZ_HD Fq finish(const double *acc) {
    // 24-bit limbs (signed carries: with balanced m_i the total lies in (-p/2, p))
    uint32_t limb[16];
    int64_t c = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int64_t t = d2ll(acc[16 + k]) + c;
        limb[k] = (uint32_t)t & 0xffffffu;
        c = t >> 24;
    }
    Fq r;
#pragma unroll
    for (int g = 0; g < 4; g++) {   // 4 limbs -> 3 words
        uint32_t l0 = limb[4 * g], l1 = limb[4 * g + 1], l2 = limb[4 * g + 2], l3 = limb[4 * g + 3];
        r.l[3 * g] = l0 | (l1 << 24);
        r.l[3 * g + 1] = (l1 >> 8) | (l2 << 16);
        r.l[3 * g + 2] = (l2 >> 16) | (l3 << 8);
    }
    if (MCHAIN == 0) {
        uint32_t neg = (uint32_t)(c >> 32);   // all ones when the total is negative: add p
        CF cf{0};
        r.l[0] = add_cc(cf, r.l[0], FqP::mod(0) & neg);
#pragma unroll
        for (int i = 1; i < 11; i++) r.l[i] = addc_cc(cf, r.l[i], FqP::mod(i) & neg);
        r.l[11] = addc(cf, r.l[11], FqP::mod(11) & neg);
    } else {
        fp_final_sub(r);
    }
    return r;
}
// one row of the integer CIOS product (field.cuh fp_mul), K = 0 .. 11
template <int K>
SCZ_HD void irow(uint32_t *even, uint32_t *odd, const uint32_t *a, const uint32_t *b) {
    if (K % 2 == 0) scz::detail::mad_n_redc<FqP>(even, odd, a, b[K], K == 0);
    else scz::detail::mad_n_redc<FqP>(odd, even, a, b[K], false);
}
SCZ_HD Fq ifinish(const uint32_t *even, const uint32_t *odd) {
    Fq r;
    CF c{0};
    r.l[0] = add_cc(c, even[0], odd[1]);
#pragma unroll
    for (int i = 1; i < 11; i++) r.l[i] = addc_cc(c, even[i], odd[i + 1]);
    r.l[11] = addc(c, even[11], 0);
    fp_final_sub(r);
    return r;
}
// rows of both products in one instruction stream: integer row K, then FP rows 4K/3 .. 4(K+1)/3 - 1
template <int MCHAIN, int K>
SCZ_HD void dual_rows(uint32_t *even, uint32_t *odd, const uint32_t *ai, const uint32_t *bi, double *acc, const uint32_t *af,
                      const double *bd) {
    if constexpr (K < 12) {
        irow<K>(even, odd, ai, bi);
        constexpr int f0 = 4 * K / 3, f1 = 4 * (K + 1) / 3;
        row1<MCHAIN, f0>(acc, af, bd);
        if constexpr (f1 - f0 == 2) row1<MCHAIN, f0 + 1>(acc, af, bd);
        dual_rows<MCHAIN, K + 1>(even, odd, ai, bi, acc, af, bd);
    }
}
}   // namespace detail

// Montgomery product a * b / 2^384 mod p, fully reduced.  MCHAIN: 0 = m_i on the FP64 pipe, 1 = through F2I / I2F.
template <int MCHAIN = 0>
SCZ_HD Fq fq_mul_f64(const Fq &a, const Fq &b) {
    double bd[16];
    detail::to_d<0>(b.l, bd);
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; k++) acc[k] = 0.0;
    detail::rows<MCHAIN, 0>(acc, a.l, bd);
    return detail::finish<MCHAIN>(acc);
}
// Two independent products in one interleaved instruction stream: ri = ai * bi on the integer pipe, rf = af * bf on
// the FP64 pipe.  Same results as fp_mul for both.
template <int MCHAIN = 1>
SCZ_HD void fq_mul_dual(Fq &ri, const Fq &ai, const Fq &bi, Fq &rf, const Fq &af, const Fq &bf) {
    double bd[16];
    detail::to_d<0>(bf.l, bd);
    double acc[32];
#pragma unroll
    for (int k = 0; k < 32; k++) acc[k] = 0.0;
    uint32_t even[12], odd[12];
    detail::dual_rows<MCHAIN, 0>(even, odd, ai.l, bi.l, acc, af.l, bd);
    Fq t = detail::ifinish(even, odd);
    rf = detail::finish<MCHAIN>(acc);
    ri = t;
}

}   // namespace f64
}   // namespace scz
