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
    if (!h) return SCZ_ERR_BAD_ARG;
    if (!srs) return h->c.fail(SCZ_ERR_BAD_ARG, "srs_precompute: null srs");
    return srs_precompute(&h->c, srs);
}

}   // extern "C"
