// Fixed-base tables for the SRS levels.
//
// The bases of every MSM on the prover path are levels of `powers_of_g` (dpoly_comm.rs:242,274,457 and, through
// c_commit, dmsm.rs:23): fixed for the life of the proving key and reused by every proof.  With the window
// multiples 2^(c w) * P_j stored beside the points, all windows of a scalar fall into ONE set of 2^(c-1) buckets
// (digit d of window w adds +-(2^(c w) P_j) to bucket |d|): the bucket reduction shrinks by the number of windows
// and there is no Horner recombination, so the window can be wider (c = 19 / 13 windows at 2^20 points instead of
// 16 / 16) and a 2^20 proof needs 14 % fewer bucket additions and 65 % fewer buckets.  The group element computed
// is the same sum_j s_j P_j; results stay bit-exact.  Cost: (windows - 1) extra copies of each level
// (12.4 GB for the two SRS of a 2^20 proof; HBM is 180 GB), built once at set-up.
#include <algorithm>

#include "ctx.h"
#include "g1.cuh"
#include "msm.h"
#include "msm_digits.cuh"
#include "srs.h"

namespace scz {

// cost in mixed-add equivalents: one addition per (point, window) + the tree reduction of ONE bucket set
uint32_t msm_pick_window_pre(size_t len) {
    if (len == 0) return 1;
    uint32_t best = 1;
    double best_cost = 1e300;
    for (uint32_t c = 1; c <= 22; c++) {
        double cost = (double)msm_num_windows(c) * (double)len + 4.8 * (double)(1u << (c - 1));
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

// pass A: one thread per point; tmp[(w-1) * len + j] = 2^(c w) P_j in XYZZ for w = 1 .. Wd-1; table[0][j] = P_j
__global__ void __launch_bounds__(128) k_srs_multiples(const void *bases, uint32_t len, uint32_t c, uint32_t Wd, void *table,
                                                        void *tmp) {
    uint32_t j = blockIdx.x * 128 + threadIdx.x;
    if (j >= len) return;
    G1Affine p = g1a_load(bases, j);
    g1a_store(table, j, p);
    G1X x = g1x_from_affine(p);
    for (uint32_t w = 1; w < Wd; w++) {
        for (uint32_t i = 0; i < c; i++) x = g1x_double(x);
        g1x_store(tmp, (size_t)(w - 1) * len + j, x);
    }
}
// pass B: XYZZ -> affine with one inversion per point (Montgomery's trick over the point's Wd-1 multiples):
// prefix[(w-1) * len + j] = zzz_1 * ... * zzz_w.  A multiple of a point of prime order is never the identity,
// so only P_j = infinity needs care (all its multiples are the (0, 0) encoding).
__global__ void __launch_bounds__(128) k_srs_normalize(uint32_t len, uint32_t Wd, void *table, const void *tmp, void *prefix) {
    uint32_t j = blockIdx.x * 128 + threadIdx.x;
    if (j >= len || Wd < 2) return;
    if (g1a_load(table, j).is_inf()) {
        G1Affine z;
        z.x = Fq::zero();
        z.y = Fq::zero();
        for (uint32_t w = 1; w < Wd; w++) g1a_store(table, (size_t)w * len + j, z);
        return;
    }
    Fq run = Fq::one();
    for (uint32_t w = 1; w < Wd; w++) {
        run = fp_mul(run, fp_load_rw<FqP>(tmp, ((size_t)(w - 1) * len + j) * 4 + 3));   // zzz is the 4th coordinate
        fp_store<FqP>(prefix, (size_t)(w - 1) * len + j, run);
    }
    Fq inv = fp_inv(run);
    for (uint32_t w = Wd - 1; w >= 1; w--) {
        G1X x = g1x_load(tmp, (size_t)(w - 1) * len + j);
        Fq izzz = w > 1 ? fp_mul(inv, fp_load_rw<FqP>(prefix, (size_t)(w - 2) * len + j)) : inv;   // 1 / zzz_w
        inv = fp_mul(inv, x.zzz);
        Fq iz = fp_mul(x.zz, izzz);   // zz / zzz = 1 / z
        G1Affine a;
        a.x = fp_mul(x.x, fp_sqr(iz));
        a.y = fp_mul(x.y, izzz);
        g1a_store(table, (size_t)w * len + j, a);
    }
}

int32_t srs_precompute(Ctx *ctx, scz_srs *srs) {
    cudaStream_t st = ctx->stream;
    srs->table.assign(srs->level.size(), nullptr);
    srs->table_c.assign(srs->level.size(), 0);
    for (size_t i = 0; i < srs->level.size(); i++) {
        size_t len = srs->len[i];
        if (!len) continue;
        uint32_t c = msm_pick_window_pre(len), Wd = msm_num_windows(c);
        if ((uint64_t)len * Wd >= (1ull << 31)) continue;   // point references are 31 bits: leave this level as it is
        void *table = nullptr, *tmp = nullptr, *prefix = nullptr;
        size_t extra = (size_t)(Wd - 1) * len;
        cudaError_t e = cudaMalloc(&table, (size_t)Wd * len * SCZ_G1_AFFINE_BYTES);
        if (e == cudaSuccess && extra) e = cudaMalloc(&tmp, extra * sizeof(G1X));
        if (e == cudaSuccess && extra) e = cudaMalloc(&prefix, extra * sizeof(Fq));
        if (e != cudaSuccess) {
            cudaGetLastError();
            cudaFree(table), cudaFree(tmp), cudaFree(prefix);
            return ctx->fail(SCZ_ERR_NOMEM, "srs_precompute: level %zu: %s", i, cudaGetErrorString(e));
        }
        uint32_t grid = ceil_div_u32(len, 128);
        k_srs_multiples<<<grid, 128, 0, st>>>(srs->level[i], (uint32_t)len, c, Wd, table, tmp);
        SCZ_LAUNCH_CHECK(ctx);
        k_srs_normalize<<<grid, 128, 0, st>>>((uint32_t)len, Wd, table, tmp, prefix);
        SCZ_LAUNCH_CHECK(ctx);
        SCZ_CUDA(ctx, cudaStreamSynchronize(st));
        cudaFree(tmp), cudaFree(prefix);
        srs->owned.push_back(table);
        srs->table[i] = table;
        srs->table_c[i] = c;
    }
    return SCZ_OK;
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_srs_precompute(scz_ctx *h, scz_srs *srs) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (!srs) return h->c.fail(SCZ_ERR_BAD_ARG, "srs_precompute: null srs");
    return srs_precompute(&h->c, srs);
}

}   // extern "C"

// ---- SRS builders of the reference (set-up, untimed there and here) --------------------------------------------------
namespace scz {

// PolynomialCommitmentCub::new (dpoly_comm.rs:37-67): level i+1 = level i * (1 - s) ++ level i * s with s = s[n-i-1]
__global__ void __launch_bounds__(128) k_srs_extend(const void *prev_jac, uint32_t len, const void *s, void *out_jac) {
    uint32_t j = blockIdx.x * 128 + threadIdx.x;
    if (j >= len) return;
    Fr sv = fp_load<FrP>(s, 0);
    G1X p = g1x_from_jac(g1j_load(prev_jac, j));
    g1j_store(out_jac, j, g1x_to_jac(g1x_mul_fr(p, fp_sub(Fr::one(), sv))));
    g1j_store(out_jac, (size_t)len + j, g1x_to_jac(g1x_mul_fr(p, sv)));
}
// mature() (dpoly_comm.rs:141-150): Jacobian -> packed affine, one inversion per point (set-up code)
__global__ void __launch_bounds__(128) k_jac_to_packed_affine(const void *jac, size_t stride, size_t off, void *out, uint32_t n) {
    uint32_t j = blockIdx.x * 128 + threadIdx.x;
    if (j >= n) return;
    G1Jac p = g1j_load(jac, (size_t)j * stride + off);
    G1Affine a;
    if (p.z.is_zero()) {
        a.x = Fq::zero();
        a.y = Fq::zero();
    } else {
        Fq iz = fp_inv(p.z), iz2 = fp_sqr(iz);
        a.x = fp_mul(p.x, iz2);
        a.y = fp_mul(p.y, fp_mul(iz2, iz));
    }
    g1a_store(out, j, a);
}
// packed affine -> Jacobian, padded with the identity up to `total` entries (to_packed's resize, :179-181)
__global__ void __launch_bounds__(128) k_affine_to_jac_padded(const void *aff, uint32_t n, void *jac, uint32_t total) {
    uint32_t j = blockIdx.x * 128 + threadIdx.x;
    if (j >= total) return;
    G1Jac p;
    p.x = Fq::one(), p.y = Fq::one(), p.z = Fq::zero();
    if (j < n) {
        G1Affine a = g1a_load(aff, j);
        if (!a.is_inf()) p.x = a.x, p.y = a.y, p.z = Fq::one();
    }
    g1j_store(jac, j, p);
}

}   // namespace scz

#include "pss.h"

extern "C" {

int32_t scz_srs_new_dev(scz_ctx *h, const void *d_g_jac, const void *d_s, size_t n, scz_srs **out) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (!d_g_jac || (n && !d_s) || !out || n > 26) return c->fail(SCZ_ERR_BAD_ARG, "srs_new: bad argument");
    scz_srs *srs = new scz_srs();
    srs->device = c->device;
    void *cur = nullptr, *nxt = nullptr;
    auto fail = [&](int32_t rc) {
        cudaFree(cur), cudaFree(nxt);
        for (void *p : srs->owned) cudaFree(p);
        delete srs;
        return rc;
    };
    if (cudaMalloc(&cur, SCZ_G1_JAC_BYTES) != cudaSuccess) return fail(c->fail(SCZ_ERR_NOMEM, "srs_new: out of memory"));
    cudaMemcpyAsync(cur, d_g_jac, SCZ_G1_JAC_BYTES, cudaMemcpyDeviceToDevice, c->stream);
    for (size_t i = 0; i <= n; i++) {
        size_t len = (size_t)1 << i;
        void *aff = nullptr;
        if (cudaMalloc(&aff, len * SCZ_G1_AFFINE_BYTES) != cudaSuccess) return fail(c->fail(SCZ_ERR_NOMEM, "srs_new: out of memory"));
        srs->owned.push_back(aff);
        k_jac_to_packed_affine<<<ceil_div_u32(len, 128), 128, 0, c->stream>>>(cur, 1, 0, aff, (uint32_t)len);
        c->launches++;
        srs->level.push_back(aff);
        srs->len.push_back(len);
        if (i == n) break;
        if (cudaMalloc(&nxt, 2 * len * SCZ_G1_JAC_BYTES) != cudaSuccess) return fail(c->fail(SCZ_ERR_NOMEM, "srs_new: out of memory"));
        k_srs_extend<<<ceil_div_u32(len, 128), 128, 0, c->stream>>>(cur, (uint32_t)len, (const char *)d_s + (n - i - 1) * 32, nxt);
        c->launches++;
        cudaStreamSynchronize(c->stream);
        cudaFree(cur);
        cur = nxt;
        nxt = nullptr;
    }
    cudaStreamSynchronize(c->stream);
    cudaFree(cur);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        cur = nullptr;
        return fail(c->cuda(e, "srs_new"));
    }
    *out = srs;
    return SCZ_OK;
}

int32_t scz_srs_to_packed_dev(scz_ctx *h, const scz_srs *srs, const scz_pp *pp, uint32_t party, scz_srs **out) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (!srs || !pp || !out || party >= pp->n) return c->fail(SCZ_ERR_BAD_ARG, "srs_to_packed: bad argument");
    const size_t l = pp->l, N = pp->n;
    scz_srs *res = new scz_srs();
    res->device = c->device;
    for (size_t i = 0; i < srs->level.size(); i++) {
        size_t len = srs->len[i], padded = len < l ? l : len, chunks = padded / l;
        void *jac = nullptr, *shares = nullptr, *aff = nullptr;
        cudaError_t e = cudaMalloc(&jac, padded * SCZ_G1_JAC_BYTES);
        if (e == cudaSuccess) e = cudaMalloc(&shares, chunks * N * SCZ_G1_JAC_BYTES);
        if (e == cudaSuccess) e = cudaMalloc(&aff, chunks * SCZ_G1_AFFINE_BYTES);
        if (e != cudaSuccess) {
            cudaFree(jac), cudaFree(shares), cudaFree(aff);
            for (void *p : res->owned) cudaFree(p);
            delete res;
            return c->fail(SCZ_ERR_NOMEM, "srs_to_packed: out of memory");
        }
        k_affine_to_jac_padded<<<ceil_div_u32(padded, 128), 128, 0, c->stream>>>(srs->level[i], (uint32_t)len, jac, (uint32_t)padded);
        c->launches++;
        // chunk b = points [b l, (b+1) l) -> pack_from_public -> N shares, stored [chunk][party]
        int32_t rc = pss_apply(c, pp, PSS_PACK, 1, jac, l, l, 1, chunks, shares, N, 1);
        if (rc == SCZ_OK) {
            k_jac_to_packed_affine<<<ceil_div_u32(chunks, 128), 128, 0, c->stream>>>(shares, N, party, aff, (uint32_t)chunks);
            c->launches++;
        }
        cudaStreamSynchronize(c->stream);
        cudaFree(jac), cudaFree(shares);
        res->owned.push_back(aff);
        res->level.push_back(aff);
        res->len.push_back(chunks);
        if (rc != SCZ_OK) {
            for (void *p : res->owned) cudaFree(p);
            delete res;
            return rc;
        }
    }
    *out = res;
    return SCZ_OK;
}

// copy one level to a caller buffer (packed affine)
int32_t scz_srs_level_dev(scz_ctx *h, const scz_srs *srs, size_t level, void *d_out, size_t *len) {
    scz::DeviceGuard dg__(h);
    if (!h) return SCZ_ERR_BAD_ARG;
    if (!srs || level >= srs->level.size()) return h->c.fail(SCZ_ERR_LEVEL_OOB, "srs_level: level %zu", level);
    if (len) *len = srs->len[level];
    if (d_out)
        SCZ_CUDA(&h->c, cudaMemcpyAsync(d_out, srs->level[level], srs->len[level] * SCZ_G1_AFFINE_BYTES, cudaMemcpyDeviceToDevice,
                                        h->c.stream));
    return SCZ_OK;
}

}   // extern "C"
