// Element-wise field / group kernels: the unit-level surface that pins the device
// arithmetic against ark-ff / ark-ec semantics (tests/test_gpu_arith.py), plus the
// small utilities the rest of the path needs (affine normalisation, synthetic bases).
// One element per thread, 128-bit coalesced loads and stores.
#include "ctx.h"
#include "g1.cuh"

namespace scz {

constexpr int EW_THREADS = 256;

template <class P, int OP>
__global__ void __launch_bounds__(EW_THREADS) k_fp_vec_op(const void *a, const void *b, void *out, size_t n) {
    size_t i = blockIdx.x * (size_t)EW_THREADS + threadIdx.x;
    if (i >= n) return;
    Fp<P> x = fp_load<P>(a, i), y = fp_load<P>(b, i), r;
    if (OP == 0) r = fp_add(x, y);
    else if (OP == 1) r = fp_sub(x, y);
    else r = fp_mul(x, y);
    fp_store<P>(out, i, r);
}
// MODE 0: inverse, 1: to canonical, 2: from canonical
template <int MODE>
__global__ void __launch_bounds__(EW_THREADS) k_fr_unary(const void *a, void *out, size_t n) {
    size_t i = blockIdx.x * (size_t)EW_THREADS + threadIdx.x;
    if (i >= n) return;
    Fr x = fp_load<FrP>(a, i), r;
    if (MODE == 0) r = fp_inv(x);
    else if (MODE == 1) r = fp_to_canon(x);
    else r = fp_from_canon(x);
    fp_store<FrP>(out, i, r);
}
__global__ void __launch_bounds__(128) k_g1_add_affine(const void *acc, const void *aff, const uint8_t *neg, void *out,
                                                       size_t n) {
    size_t i = blockIdx.x * (size_t)128 + threadIdx.x;
    if (i >= n) return;
    G1X a = g1x_from_jac(g1j_load(acc, i));
    G1Affine p = g1a_load(aff, i);
    g1x_add_affine(a, p, neg ? neg[i] != 0 : false);
    g1j_store(out, i, g1x_to_jac(a));
}
template <int OP>
__global__ void __launch_bounds__(128) k_g1_vec_op(const void *a, const void *b, void *out, size_t n) {
    size_t i = blockIdx.x * (size_t)128 + threadIdx.x;
    if (i >= n) return;
    G1X x = g1x_from_jac(g1j_load(a, i));
    G1X r = OP == 0 ? g1x_add(x, g1x_from_jac(g1j_load(b, i))) : g1x_double(x);
    g1j_store(out, i, g1x_to_jac(r));
}
__global__ void __launch_bounds__(128) k_g1_mul_fr(const void *a, const void *k, void *out, size_t n) {
    size_t i = blockIdx.x * (size_t)128 + threadIdx.x;
    if (i >= n) return;
    G1X x = g1x_from_jac(g1j_load(a, i));
    g1j_store(out, i, g1x_to_jac(g1x_mul_fr(x, fp_load<FrP>(k, i))));
}
__device__ __forceinline__ G1Affine g1x_to_affine(const G1X &p) {
    G1Affine r;
    if (p.is_inf()) {
        r.x = Fq::zero();
        r.y = Fq::zero();
        return r;
    }
    // ZZ = z^2, ZZZ = z^3: one inversion gives 1/z^3, then 1/z = ZZ/ZZZ and 1/z^2 = (1/z)^2
    Fq iz3 = fp_inv(p.zzz);
    Fq iz = fp_mul(p.zz, iz3);
    Fq iz2 = fp_sqr(iz);
    r.x = fp_mul(p.x, iz2);
    r.y = fp_mul(p.y, iz3);
    return r;
}
__global__ void __launch_bounds__(128) k_g1_to_affine(const void *a, void *out, size_t n) {
    size_t i = blockIdx.x * (size_t)128 + threadIdx.x;
    if (i >= n) return;
    g1a_store(out, i, g1x_to_affine(g1x_from_jac(g1j_load(a, i))));
}
__global__ void __launch_bounds__(128) k_g1_generator_mul(const void *k, void *out, size_t n) {
    size_t i = blockIdx.x * (size_t)128 + threadIdx.x;
    if (i >= n) return;
    // BLS12-381 G1 generator, Montgomery form (x, y)
    constexpr uint32_t GX[12] = {0xfd530c16u, 0x5cb38790u, 0x9976fff5u, 0x7817fc67u, 0x143ba1c1u, 0x154f95c7u,
                                 0xf3d0e747u, 0xf0ae6acdu, 0x21dbf440u, 0xedce6eccu, 0x9e0bfb75u, 0x12017741u};
    constexpr uint32_t GY[12] = {0x0ce72271u, 0xbaac93d5u, 0x7918fd8eu, 0x8c22631au, 0x570725ceu, 0xdd595f13u,
                                 0x50405194u, 0x51ac5829u, 0xad0059c0u, 0x0e1c8c3fu, 0x5008a26au, 0x0bbc3efcu};
    G1X g;
#pragma unroll
    for (int j = 0; j < 12; j++) {
        g.x.l[j] = GX[j];
        g.y.l[j] = GY[j];
    }
    g.zz = Fq::one();
    g.zzz = Fq::one();
    g1a_store(out, i, g1x_to_affine(g1x_mul_fr(g, fp_load<FrP>(k, i))));
}
__global__ void k_g1_apply_inf_mask(void *bases, const uint8_t *mask, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n || !mask[i]) return;
    G1Affine z;
    z.x = Fq::zero();
    z.y = Fq::zero();
    g1a_store(bases, i, z);
}

}   // namespace scz

using namespace scz;

#define CHECK_ARGS(h, cond) \
    if (!(h)) return SCZ_ERR_BAD_ARG; \
    if (!(cond)) return (h)->c.fail(SCZ_ERR_BAD_ARG, "%s: bad argument", __func__)

extern "C" {

int32_t scz_fr_vec_op_dev(scz_ctx *h, int32_t op, const void *a, const void *b, void *out, size_t n) {
    scz::DeviceGuard dg__(h);
    CHECK_ARGS(h, a && b && out && op >= 0 && op <= 2);
    if (!n) return SCZ_OK;
    Ctx *c = &h->c;
    dim3 g(ceil_div_u32(n, EW_THREADS));
    if (op == 0) k_fp_vec_op<FrP, 0><<<g, EW_THREADS, 0, c->stream>>>(a, b, out, n);
    else if (op == 1) k_fp_vec_op<FrP, 1><<<g, EW_THREADS, 0, c->stream>>>(a, b, out, n);
    else k_fp_vec_op<FrP, 2><<<g, EW_THREADS, 0, c->stream>>>(a, b, out, n);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}
int32_t scz_fq_vec_op_dev(scz_ctx *h, int32_t op, const void *a, const void *b, void *out, size_t n) {
    scz::DeviceGuard dg__(h);
    CHECK_ARGS(h, a && b && out && op >= 0 && op <= 2);
    if (!n) return SCZ_OK;
    Ctx *c = &h->c;
    dim3 g(ceil_div_u32(n, EW_THREADS));
    if (op == 0) k_fp_vec_op<FqP, 0><<<g, EW_THREADS, 0, c->stream>>>(a, b, out, n);
    else if (op == 1) k_fp_vec_op<FqP, 1><<<g, EW_THREADS, 0, c->stream>>>(a, b, out, n);
    else k_fp_vec_op<FqP, 2><<<g, EW_THREADS, 0, c->stream>>>(a, b, out, n);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}
static int32_t fr_unary(scz_ctx *h, int mode, const void *a, void *out, size_t n) {
    CHECK_ARGS(h, a && out);
    if (!n) return SCZ_OK;
    Ctx *c = &h->c;
    dim3 g(ceil_div_u32(n, EW_THREADS));
    if (mode == 0) k_fr_unary<0><<<g, EW_THREADS, 0, c->stream>>>(a, out, n);
    else if (mode == 1) k_fr_unary<1><<<g, EW_THREADS, 0, c->stream>>>(a, out, n);
    else k_fr_unary<2><<<g, EW_THREADS, 0, c->stream>>>(a, out, n);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}
int32_t scz_fr_inv_dev(scz_ctx *h, const void *a, void *out, size_t n) { return fr_unary(h, 0, a, out, n); }
int32_t scz_fr_to_canonical_dev(scz_ctx *h, const void *a, void *out, size_t n) { return fr_unary(h, 1, a, out, n); }
int32_t scz_fr_from_canonical_dev(scz_ctx *h, const void *a, void *out, size_t n) { return fr_unary(h, 2, a, out, n); }

int32_t scz_g1_add_affine_dev(scz_ctx *h, const void *acc, const void *aff, const uint8_t *neg, void *out, size_t n) {
    scz::DeviceGuard dg__(h);
    CHECK_ARGS(h, acc && aff && out);
    if (!n) return SCZ_OK;
    Ctx *c = &h->c;
    k_g1_add_affine<<<ceil_div_u32(n, 128), 128, 0, c->stream>>>(acc, aff, neg, out, n);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}
int32_t scz_g1_vec_op_dev(scz_ctx *h, int32_t op, const void *a, const void *b, void *out, size_t n) {
    scz::DeviceGuard dg__(h);
    CHECK_ARGS(h, a && out && (op == 1 || (op == 0 && b)));
    if (!n) return SCZ_OK;
    Ctx *c = &h->c;
    if (op == 0) k_g1_vec_op<0><<<ceil_div_u32(n, 128), 128, 0, c->stream>>>(a, b, out, n);
    else k_g1_vec_op<1><<<ceil_div_u32(n, 128), 128, 0, c->stream>>>(a, b, out, n);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}
int32_t scz_g1_mul_fr_dev(scz_ctx *h, const void *a, const void *k, void *out, size_t n) {
    scz::DeviceGuard dg__(h);
    CHECK_ARGS(h, a && k && out);
    if (!n) return SCZ_OK;
    Ctx *c = &h->c;
    k_g1_mul_fr<<<ceil_div_u32(n, 128), 128, 0, c->stream>>>(a, k, out, n);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}
int32_t scz_g1_to_affine_dev(scz_ctx *h, const void *a, void *out, size_t n) {
    scz::DeviceGuard dg__(h);
    CHECK_ARGS(h, a && out);
    if (!n) return SCZ_OK;
    Ctx *c = &h->c;
    k_g1_to_affine<<<ceil_div_u32(n, 128), 128, 0, c->stream>>>(a, out, n);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}
int32_t scz_g1_generator_mul_dev(scz_ctx *h, const void *k, void *out, size_t n) {
    scz::DeviceGuard dg__(h);
    CHECK_ARGS(h, k && out);
    if (!n) return SCZ_OK;
    Ctx *c = &h->c;
    k_g1_generator_mul<<<ceil_div_u32(n, 128), 128, 0, c->stream>>>(k, out, n);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}
int32_t scz_g1_apply_inf_mask_dev(scz_ctx *h, void *bases, const uint8_t *mask, size_t n) {
    scz::DeviceGuard dg__(h);
    CHECK_ARGS(h, bases && mask);
    if (!n) return SCZ_OK;
    Ctx *c = &h->c;
    k_g1_apply_inf_mask<<<ceil_div_u32(n, 256), 256, 0, c->stream>>>(bases, mask, n);
    SCZ_LAUNCH_CHECK(c);
    return SCZ_OK;
}

}   // extern "C"
