// TEST-ONLY host build of the device arithmetic headers (csrc/field.cuh, g1.cuh, g2.cuh,
// msm_digits.cuh).  The authoring container has no GPU, so the limb logic that the
// CUDA kernels inline is compiled here with g++ (carry-chain primitives fall back
// to their host emulation) and compared with the oracle by tests/test_emu_arith.py.
// This file is NEVER linked into libscz.so: the product has no CPU path.
#include <cstring>
#include "../../scalable-collaborative-zksnark_b200/csrc/field.cuh"
#include "../../scalable-collaborative-zksnark_b200/csrc/g1.cuh"
#include "../../scalable-collaborative-zksnark_b200/csrc/g2.cuh"
#include "../../tools/ubench/fq_f64.cuh"
#include "../../scalable-collaborative-zksnark_b200/csrc/g1_batch_affine.cuh"
#include <vector>
#include "../../scalable-collaborative-zksnark_b200/csrc/msm_digits.cuh"
using namespace scz;

template <class P> static Fp<P> ld(const uint32_t *p, size_t i) {
    Fp<P> r;
    memcpy(r.l, p + i * P::N, sizeof r.l);
    return r;
}
template <class P> static void st(uint32_t *p, size_t i, const Fp<P> &v) { memcpy(p + i * P::N, v.l, sizeof v.l); }

#define VEC2(name, P, op)                                                                  \
    extern "C" void name(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n) {   \
        for (size_t i = 0; i < n; i++) st<P>(r, i, op(ld<P>(a, i), ld<P>(b, i)));         \
    }
VEC2(emu_fr_mul, FrP, fp_mul)
VEC2(emu_fr_add, FrP, fp_add)
VEC2(emu_fr_sub, FrP, fp_sub)
VEC2(emu_fq_mul, FqP, fp_mul)
VEC2(emu_fq_add, FqP, fp_add)
VEC2(emu_fq_sub, FqP, fp_sub)
// the FP64-pipe Fq product (fq_f64.cuh), both row-multiplier variants
extern "C" void emu_fq_mul_f64(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n, int mchain) {
    for (size_t i = 0; i < n; i++)
        st<FqP>(r, i, mchain ? f64::fq_mul_f64<1>(ld<FqP>(a, i), ld<FqP>(b, i)) : f64::fq_mul_f64<0>(ld<FqP>(a, i), ld<FqP>(b, i)));
}
// two products in one interleaved stream (fq_mul_dual): r[2i] = a*b (integer rows), r[2i+1] = c*d (FP64 rows)
extern "C" void emu_fq_mul_dual(const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d, uint32_t *r, size_t n,
                                int mchain) {
    for (size_t i = 0; i < n; i++) {
        Fq ri, rf;
        if (mchain) f64::fq_mul_dual<1>(ri, ld<FqP>(a, i), ld<FqP>(b, i), rf, ld<FqP>(c, i), ld<FqP>(d, i));
        else f64::fq_mul_dual<0>(ri, ld<FqP>(a, i), ld<FqP>(b, i), rf, ld<FqP>(c, i), ld<FqP>(d, i));
        st<FqP>(r, 2 * i, ri);
        st<FqP>(r, 2 * i + 1, rf);
    }
}
extern "C" void emu_fq_sqr_sos(const uint32_t *a, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) st<FqP>(r, i, fq_sqr_sos(ld<FqP>(a, i)));
}
// r = a*b - c*d with one reduction (fp_dot2_sub)
extern "C" void emu_fq_dot2_sub(const uint32_t *a, const uint32_t *b, const uint32_t *c, const uint32_t *d, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) st<FqP>(r, i, fp_dot2_sub(ld<FqP>(a, i), ld<FqP>(b, i), ld<FqP>(c, i), ld<FqP>(d, i)));
}
// sum_i a_i * b_i with ONE reduction (fp_mul_acc_wide / fp_acc_wide_reduce: the round sums of the product sumcheck)
extern "C" void emu_fr_dot_wide(const uint32_t *a, const uint32_t *b, size_t n, uint32_t *r) {
    uint32_t acc[17] = {0};
    for (size_t i = 0; i < n; i++) fp_mul_acc_wide<FrP>(acc, ld<FrP>(a, i), ld<FrP>(b, i));
    st<FrP>(r, 0, fp_acc_wide_reduce<FrP>(acc));
}
extern "C" void emu_fq_dot_wide(const uint32_t *a, const uint32_t *b, size_t n, uint32_t *r) {
    uint32_t acc[25] = {0};
    for (size_t i = 0; i < n; i++) fp_mul_acc_wide<FqP>(acc, ld<FqP>(a, i), ld<FqP>(b, i));
    st<FqP>(r, 0, fp_acc_wide_reduce<FqP>(acc));
}
extern "C" void emu_fq_inv_bingcd(const uint32_t *a, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) st<FqP>(r, i, fp_inv_bingcd(ld<FqP>(a, i)));
}
extern "C" void emu_fr_inv_bingcd(const uint32_t *a, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) st<FrP>(r, i, fp_inv_bingcd(ld<FrP>(a, i)));
}
extern "C" void emu_fr_inv(const uint32_t *a, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) st<FrP>(r, i, fp_inv(ld<FrP>(a, i)));
}
extern "C" void emu_fr_canon(const uint32_t *a, uint32_t *r, size_t n, int to) {
    for (size_t i = 0; i < n; i++) st<FrP>(r, i, to ? fp_to_canon(ld<FrP>(a, i)) : fp_from_canon(ld<FrP>(a, i)));
}
static G1Jac ldj(const uint32_t *p, size_t i) {
    G1Jac r;
    r.x = ld<FqP>(p, 3 * i);
    r.y = ld<FqP>(p, 3 * i + 1);
    r.z = ld<FqP>(p, 3 * i + 2);
    return r;
}
static void stj(uint32_t *p, size_t i, const G1Jac &v) {
    st<FqP>(p, 3 * i, v.x);
    st<FqP>(p, 3 * i + 1, v.y);
    st<FqP>(p, 3 * i + 2, v.z);
}
// acc (Jacobian) += affine (x|y, 24 words; all-zero = infinity), optionally negated
extern "C" void emu_g1_add_affine(const uint32_t *acc, const uint32_t *aff, const uint8_t *neg, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        G1X a = g1x_from_jac(ldj(acc, i));
        G1Affine p;
        p.x = ld<FqP>(aff, 2 * i);
        p.y = ld<FqP>(aff, 2 * i + 1);
        g1x_add_affine(a, p, neg[i] != 0);
        stj(out, i, g1x_to_jac(a));
    }
}
extern "C" void emu_g1_add(const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) stj(out, i, g1x_to_jac(g1x_add(g1x_from_jac(ldj(a, i)), g1x_from_jac(ldj(b, i)))));
}
extern "C" void emu_g1_double(const uint32_t *a, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) stj(out, i, g1x_to_jac(g1x_double(g1x_from_jac(ldj(a, i)))));
}
extern "C" void emu_g1_mul_fr(const uint32_t *a, const uint32_t *k, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) stj(out, i, g1x_to_jac(g1x_mul_fr(g1x_from_jac(ldj(a, i)), ld<FrP>(k, i))));
}
// digits[i*W + w] for canonical scalars
extern "C" uint32_t emu_msm_digits(const uint32_t *k, size_t n, uint32_t c, int32_t *digits) {
    uint32_t W = msm_num_windows(c);
    for (size_t i = 0; i < n; i++) {
        uint32_t s[8], carry = 0;
        memcpy(s, k + 8 * i, sizeof s);
        for (uint32_t w = 0; w < W; w++) digits[i * W + w] = msm_signed_digit(s, c, w, carry);
        if (carry) return 0xffffffffu;   // must never happen
    }
    return W;
}

// n affine additions out[i] = p[i] + q[i] with ONE inversion per `batch` additions (g1_batch_affine.cuh); packed affine
// rows of 24 words, all-zero = infinity
extern "C" void emu_g1_batch_affine_add(const uint32_t *p, const uint32_t *q, uint32_t *out, size_t n, size_t batch) {
    std::vector<G1Affine> P(n), Q(n), R(n);
    for (size_t i = 0; i < n; i++) {
        P[i].x = ld<FqP>(p, 2 * i), P[i].y = ld<FqP>(p, 2 * i + 1);
        Q[i].x = ld<FqP>(q, 2 * i), Q[i].y = ld<FqP>(q, 2 * i + 1);
    }
    std::vector<Fq> prefix(batch);
    for (size_t lo = 0; lo < n; lo += batch) {
        int m = (int)(n - lo < batch ? n - lo : batch);
        Fq total = g1a_batch_phase1(&P[lo], &Q[lo], m, prefix.data());
        g1a_batch_phase2(&P[lo], &Q[lo], m, prefix.data(), fp_inv(total), &R[lo]);
    }
    for (size_t i = 0; i < n; i++) {
        st<FqP>(out, 2 * i, R[i].x);
        st<FqP>(out, 2 * i + 1, R[i].y);
    }
}

// ---- G2 (g2.cuh): Fq2 = c0 | c1 (24 words), Jacobian X | Y | Z (72 words), packed affine x | y (48 words, all-zero = identity)
static Fq2 ld2(const uint32_t *p, size_t i) {
    Fq2 r;
    r.c0 = ld<FqP>(p, 2 * i);
    r.c1 = ld<FqP>(p, 2 * i + 1);
    return r;
}
static void st2(uint32_t *p, size_t i, const Fq2 &v) {
    st<FqP>(p, 2 * i, v.c0);
    st<FqP>(p, 2 * i + 1, v.c1);
}
static G2Jac ldj2(const uint32_t *p, size_t i) {
    G2Jac r;
    r.x = ld2(p, 3 * i);
    r.y = ld2(p, 3 * i + 1);
    r.z = ld2(p, 3 * i + 2);
    return r;
}
static void stj2(uint32_t *p, size_t i, const G2Jac &v) {
    st2(p, 3 * i, v.x);
    st2(p, 3 * i + 1, v.y);
    st2(p, 3 * i + 2, v.z);
}
extern "C" void emu_fq2_mul(const uint32_t *a, const uint32_t *b, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) st2(r, i, f2_mul(ld2(a, i), ld2(b, i)));
}
extern "C" void emu_fq2_sqr(const uint32_t *a, uint32_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) st2(r, i, f2_sqr(ld2(a, i)));
}
extern "C" void emu_g2_add(const uint32_t *a, const uint32_t *b, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) stj2(out, i, g2x_to_jac(g2x_add(g2x_from_jac(ldj2(a, i)), g2x_from_jac(ldj2(b, i)))));
}
extern "C" void emu_g2_double(const uint32_t *a, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) stj2(out, i, g2x_to_jac(g2x_double(g2x_from_jac(ldj2(a, i)))));
}
extern "C" void emu_g2_add_affine(const uint32_t *acc, const uint32_t *aff, const uint8_t *neg, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        G2X a = g2x_from_jac(ldj2(acc, i));
        G2Affine p;
        p.x = ld2(aff, 2 * i);
        p.y = ld2(aff, 2 * i + 1);
        g2x_add_affine(a, p, neg[i] != 0);
        stj2(out, i, g2x_to_jac(a));
    }
}
// k = canonical 256-bit scalars (8 words each)
extern "C" void emu_g2_mul_bits(const uint32_t *a, const uint32_t *k, uint32_t *out, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint32_t kk[8];
        memcpy(kk, k + 8 * i, sizeof kk);
        stj2(out, i, g2x_to_jac(g2x_mul_bits(g2x_from_jac(ldj2(a, i)), kk)));
    }
}
