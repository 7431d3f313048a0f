// Batched Pippenger MSM over BLS12-381 G1 for sm_100a.
//
// Replaces `G::msm` (ark-ec 0.4.2 VariableBaseMSM) at the reference's call sites
// dist-primitive/src/dmsm.rs:23 and dpoly_comm.rs:242,274,457.  One launch sequence
// handles a whole batch of independent MSMs ("segments") -- the shape `d_msm`
// receives from `c_open` (22 MSMs of halving length, dpoly_comm.rs:436) -- so
// the ~800 MSMs of a proof cost ~100 sequences instead of ~800.
//
// Pipeline (all on ctx->stream, nothing returns to the host):
//   1 count     one thread per scalar: Montgomery -> integer, signed c-bit digits,
//               histogram of (segment, window, |digit|) buckets            [atomics]
//   2 scan      exclusive prefix sum of the histogram                      [3 small kernels]
//   3 scatter   digits recomputed, point index (+ sign bit) written to its bucket's slot:
//               a counting sort whose within-bucket order is irrelevant because
//               group addition commutes (the result is bit-exact regardless)
//   4 accumulate one thread per bucket: gathers its bases (96 B, 128-bit loads) and
//               mixed-adds them into an XYZZ accumulator held in registers;
//               oversized buckets (degenerate scalar distributions such as the
//               all-ones test of dmsm.rs:103) go to a block-per-bucket kernel
//   5 reduce    one CTA per window: sum_k k*B_k by chunked running sums + a
//               shared-memory tree with R_AB = R_A + R_B + |A|*S_B
//   6 finish    one thread per segment: Horner over the windows, XYZZ -> Jacobian
// The bucket kernels are bound by the integer multiply pipe (a mixed add is
// ~4.7k IMAD.WIDE for ~100 B of HBM traffic), see DESIGN.md.
#include <algorithm>
#include <vector>

#include "g1.cuh"
#include "msm.h"
#include "msm_digits.cuh"

namespace scz {

constexpr int CNT_THREADS = 256;
constexpr int ACC_THREADS = 128;
constexpr int RED_THREADS = 128;
constexpr int HEAVY_THREADS = 128;
constexpr uint32_t HEAVY_CAP = 1024;   // buckets longer than this are split across a CTA
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;          // per thread
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int seg_by_point(const MsmSeg *segs, int K, uint32_t g) {
    int lo = 0, hi = K - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&segs[mid].point_base) <= g) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ int seg_by_bucket(const MsmSeg *segs, int K, uint32_t b) {
    int lo = 0, hi = K - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&segs[mid].bucket_base) <= b) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ int seg_by_window(const MsmSeg *segs, int K, uint32_t w) {
    int lo = 0, hi = K - 1;
    while (lo < hi) {
        int mid = (lo + hi + 1) >> 1;
        if (__ldg(&segs[mid].window_base) <= w) lo = mid;
        else hi = mid - 1;
    }
    return lo;
}

// ---- 1 + 3: recode, then count (SCATTER = false) or place (SCATTER = true)
template <bool SCATTER>
__global__ void __launch_bounds__(CNT_THREADS) k_msm_recode(const MsmSeg *segs, int K, uint32_t total_points,
                                                             uint32_t *counts, uint32_t *cursor, uint32_t *sorted) {
    uint32_t g = blockIdx.x * CNT_THREADS + threadIdx.x;
    if (g >= total_points) return;
    int s = seg_by_point(segs, K, g);
    const MsmSeg sg = segs[s];
    uint32_t i = g - sg.point_base;
    Fr k = fp_to_canon(fp_load<FrP>(sg.scalars, i));
    uint32_t carry = 0;
    for (uint32_t w = 0; w < sg.W; w++) {
        int32_t d = msm_signed_digit(k.l, sg.c, w, carry);
        if (d == 0) continue;
        uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        uint32_t b = sg.bucket_base + w * sg.nb + (mag - 1);
        if (!SCATTER) {
            atomicAdd(&counts[b], 1u);
        } else {
            uint32_t pos = atomicAdd(&cursor[b], 1u);
            sorted[pos] = i | (d < 0 ? 0x80000000u : 0u);
        }
    }
}

// ---- 2: exclusive scan of counts -> cursor (three passes)
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const uint32_t *in, uint32_t *out, uint32_t *tile_sums,
                                                             uint32_t n) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS], sum = 0;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        v[j] = base + j < n ? in[base + j] : 0;
        sum += v[j];
    }
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0, wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - ws;   // exclusive
        if (lane == SCAN_THREADS / 32 - 1) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    uint32_t run = warp_sums[wid] + incl - sum;
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++) {
        if (base + j < n) out[base + j] = run;
        run += v[j];
    }
}
__global__ void __launch_bounds__(1024) k_scan_tile_sums(uint32_t *tile_sums, uint32_t tiles) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (uint32_t start = 0; start < tiles; start += 1024) {
        uint32_t idx = start + threadIdx.x;
        uint32_t v = idx < tiles ? tile_sums[idx] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warp_sums[wid] = incl;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane], wi = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, wi, o);
                if (lane >= o) wi += t;
            }
            warp_sums[lane] = wi - ws;
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + warp_sums[wid] + incl - v;
        if (idx < tiles) tile_sums[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = excl + v;
        __syncthreads();
    }
}
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t *out, const uint32_t *tile_sums, uint32_t n) {
    uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t add = tile_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < SCAN_ITEMS; j++)
        if (base + j < n) out[base + j] += add;
}

// ---- 4: bucket accumulation, one thread per bucket
// After the scatter pass cursor[b] is the END of bucket b; its entries are sorted[end - count, end).
__global__ void __launch_bounds__(ACC_THREADS) k_msm_accumulate(const MsmSeg *segs, int K, uint32_t total_buckets,
                                                                 const uint32_t *counts, const uint32_t *cursor,
                                                                 const uint32_t *sorted, void *buckets,
                                                                 uint32_t *heavy_list, uint32_t *heavy_count) {
    uint32_t b = blockIdx.x * ACC_THREADS + threadIdx.x;
    if (b >= total_buckets) return;
    uint32_t n = counts[b];
    if (n > HEAVY_CAP) {
        heavy_list[atomicAdd(heavy_count, 1u)] = b;
        return;
    }
    G1X acc = G1X::inf();
    if (n) {
        int s = seg_by_bucket(segs, K, b);
        const void *bases = segs[s].bases;
        const uint32_t *e = sorted + (cursor[b] - n);
        for (uint32_t j = 0; j < n; j++) {
            uint32_t ent = __ldg(e + j);
            G1Affine p = g1a_load(bases, ent & 0x7fffffffu);
            g1x_add_affine(acc, p, (ent >> 31) != 0);
        }
    }
    g1x_store(buckets, b, acc);
}
// tree-sum of one XYZZ value per thread through shared memory; result in thread 0
template <int T>
__device__ __forceinline__ G1X block_sum_g1x(G1X v, G1X *sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int stride = T / 2; stride > 0; stride >>= 1) {
        if ((int)threadIdx.x < stride) {
            v = g1x_add(v, sh[threadIdx.x + stride]);
            sh[threadIdx.x] = v;
        }
        __syncthreads();
    }
    return v;
}
__global__ void __launch_bounds__(HEAVY_THREADS) k_msm_accumulate_heavy(const MsmSeg *segs, int K,
                                                                         const uint32_t *counts, const uint32_t *cursor,
                                                                         const uint32_t *sorted, void *buckets,
                                                                         const uint32_t *heavy_list,
                                                                         const uint32_t *heavy_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    G1X *sh = reinterpret_cast<G1X *>(smem_raw);
    for (uint32_t h = blockIdx.x; h < *heavy_count; h += gridDim.x) {
        uint32_t b = heavy_list[h];
        uint32_t n = counts[b];
        int s = seg_by_bucket(segs, K, b);
        const void *bases = segs[s].bases;
        const uint32_t *e = sorted + (cursor[b] - n);
        G1X acc = G1X::inf();
        for (uint32_t j = threadIdx.x; j < n; j += HEAVY_THREADS) {
            uint32_t ent = __ldg(e + j);
            G1Affine p = g1a_load(bases, ent & 0x7fffffffu);
            g1x_add_affine(acc, p, (ent >> 31) != 0);
        }
        acc = block_sum_g1x<HEAVY_THREADS>(acc, sh);
        if (threadIdx.x == 0) g1x_store(buckets, b, acc);
        __syncthreads();
    }
}

// ---- 5: per-window bucket reduction  sum_{k=1..nb} k * B_k
__device__ __forceinline__ G1X g1x_mul_pow2(G1X p, uint32_t log2k) {
    for (uint32_t i = 0; i < log2k; i++) p = g1x_double(p);
    return p;
}
__global__ void __launch_bounds__(RED_THREADS) k_msm_reduce(const MsmSeg *segs, int K, const void *buckets,
                                                             void *window_sums) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    G1X *shS = reinterpret_cast<G1X *>(smem_raw);
    G1X *shR = shS + RED_THREADS;
    uint32_t gw = blockIdx.x;
    int s = seg_by_window(segs, K, gw);
    const MsmSeg sg = segs[s];
    uint32_t w = gw - sg.window_base;
    uint32_t first = sg.bucket_base + w * sg.nb;
    // chunk length L (a power of two) and number of active threads
    uint32_t L = sg.nb >= RED_THREADS ? sg.nb / RED_THREADS : 1;
    uint32_t active = sg.nb / L;
    uint32_t t = threadIdx.x;
    G1X S = G1X::inf(), R = G1X::inf();
    if (t < active) {
        uint32_t lo = t * L;
        for (uint32_t j = L; j-- > 0;) {   // running sum from the chunk's top bucket down
            S = g1x_add(S, g1x_load(buckets, first + lo + j));
            R = g1x_add(R, S);
        }
    }
    shS[t] = S;
    shR[t] = R;
    __syncthreads();
    // tree: node A = [t, t+stride), node B = [t+stride, t+2*stride);  |A| = stride * L buckets
    uint32_t logL = 31 - __clz(L);
    uint32_t level = 0;
    for (uint32_t stride = 1; stride < active; stride <<= 1, level++) {
        if ((t & (2 * stride - 1)) == 0 && t + stride < active) {
            G1X Sb = shS[t + stride];
            G1X Rb = shR[t + stride];
            R = g1x_add(g1x_add(R, Rb), g1x_mul_pow2(Sb, logL + level));
            S = g1x_add(S, Sb);
            shS[t] = S;
            shR[t] = R;
        }
        __syncthreads();
    }
    if (t == 0) g1x_store(window_sums, gw, R);
}

// ---- 6: Horner over windows, one thread per segment
__global__ void k_msm_finish(const MsmSeg *segs, int K, const void *window_sums, void *out_jac) {
    int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= K) return;
    const MsmSeg sg = segs[s];
    G1X acc = G1X::inf();
    if (sg.len) {
        for (uint32_t w = sg.W; w-- > 0;) {
            acc = g1x_mul_pow2(acc, sg.c);
            acc = g1x_add(acc, g1x_load(window_sums, sg.window_base + w));
        }
    }
    g1j_store(out_jac, s, g1x_to_jac(acc));
}

// ------------------------------------------------------------------ host side
uint32_t msm_pick_window(size_t len) {
    if (len == 0) return 1;
    // cost in mixed-add equivalents: W * (len + 6 * nb): the reduction runs one CTA per
    // window, so a bucket there costs several times a bucket addition
    uint32_t best = 1;
    double best_cost = 1e300;
    for (uint32_t c = 1; c <= 16; c++) {
        double W = (double)msm_num_windows(c), nb = (double)(1u << (c - 1));
        double cost = W * ((double)len + 6.0 * nb);
        if (cost < best_cost) {
            best_cost = cost;
            best = c;
        }
    }
    return best;
}

int32_t msm_g1_batched(Ctx *ctx, const void *const *d_bases, const void *const *d_scalars, const size_t *lens,
                       size_t batch, void *d_out) {
    if (batch == 0) return SCZ_OK;
    if (batch > (1u << 20)) return ctx->fail(SCZ_ERR_BAD_ARG, "msm: batch too large");
    std::vector<MsmSeg> segs(batch);
    uint64_t points = 0, buckets = 0, windows = 0, entries = 0;
    for (size_t k = 0; k < batch; k++) {
        if (lens[k] >= (1ull << 31)) return ctx->fail(SCZ_ERR_BAD_ARG, "msm: segment %zu too long", k);
        if (lens[k] && (!d_bases[k] || !d_scalars[k])) return ctx->fail(SCZ_ERR_BAD_ARG, "msm: null segment %zu", k);
        MsmSeg &s = segs[k];
        s.bases = d_bases[k];
        s.scalars = d_scalars[k];
        s.len = (uint32_t)lens[k];
        s.point_base = (uint32_t)points;
        s.c = ctx->msm_window_override ? ctx->msm_window_override : msm_pick_window(lens[k]);
        s.W = msm_num_windows(s.c);
        s.nb = 1u << (s.c - 1);
        s.bucket_base = (uint32_t)buckets;
        s.window_base = (uint32_t)windows;
        s.pad_ = 0;
        points += s.len;
        buckets += (uint64_t)s.W * s.nb;
        windows += s.W;
        entries += (uint64_t)s.len * s.W;
    }
    if (points >= (1ull << 31) || buckets >= (1ull << 31) || entries >= (1ull << 32))
        return ctx->fail(SCZ_ERR_BAD_ARG, "msm: batch too large (%llu points, %llu buckets)", (unsigned long long)points,
                         (unsigned long long)buckets);
    ctx->msm_bucket_adds = entries;
    ctx->msm_buckets = buckets;
    ctx->msm_windows = windows;

    cudaStream_t st = ctx->stream;
    uint32_t tiles = ceil_div_u32(buckets, SCAN_TILE);
    DevTmp d_segs(ctx), d_counts(ctx), d_cursor(ctx), d_tiles(ctx), d_sorted(ctx), d_buckets(ctx), d_wsums(ctx),
        d_heavy(ctx);
    SCZ_TRY(d_segs.alloc(batch * sizeof(MsmSeg)));
    SCZ_TRY(d_counts.alloc(buckets * 4));
    SCZ_TRY(d_cursor.alloc(buckets * 4));
    SCZ_TRY(d_tiles.alloc((size_t)tiles * 4 + 4));
    SCZ_TRY(d_sorted.alloc((entries ? entries : 1) * 4));
    SCZ_TRY(d_buckets.alloc(buckets * sizeof(G1X)));
    SCZ_TRY(d_wsums.alloc(windows * sizeof(G1X)));
    SCZ_TRY(d_heavy.alloc((buckets + 1) * 4));   // [0] = count, [1..] = list
    // segment table: pageable host -> device; the vector must outlive the copy, so stage through the stream
    SCZ_CUDA(ctx, cudaMemcpyAsync(d_segs.p, segs.data(), batch * sizeof(MsmSeg), cudaMemcpyHostToDevice, st));
    SCZ_CUDA(ctx, cudaStreamSynchronize(st));   // pageable source: make the copy complete before `segs` dies
    SCZ_CUDA(ctx, cudaMemsetAsync(d_counts.p, 0, buckets * 4, st));
    SCZ_CUDA(ctx, cudaMemsetAsync(d_heavy.p, 0, 4, st));
    const MsmSeg *sp = d_segs.as<MsmSeg>();
    int K = (int)batch;
    uint32_t *counts = d_counts.as<uint32_t>(), *cursor = d_cursor.as<uint32_t>(), *sorted = d_sorted.as<uint32_t>();
    uint32_t *heavy = d_heavy.as<uint32_t>();
    if (points) {
        k_msm_recode<false><<<ceil_div_u32(points, CNT_THREADS), CNT_THREADS, 0, st>>>(sp, K, (uint32_t)points, counts,
                                                                                       nullptr, nullptr);
        SCZ_LAUNCH_CHECK(ctx);
    }
    k_scan_tiles<<<tiles, SCAN_THREADS, 0, st>>>(counts, cursor, d_tiles.as<uint32_t>(), (uint32_t)buckets);
    SCZ_LAUNCH_CHECK(ctx);
    k_scan_tile_sums<<<1, 1024, 0, st>>>(d_tiles.as<uint32_t>(), tiles);
    SCZ_LAUNCH_CHECK(ctx);
    k_scan_add<<<tiles, SCAN_THREADS, 0, st>>>(cursor, d_tiles.as<uint32_t>(), (uint32_t)buckets);
    SCZ_LAUNCH_CHECK(ctx);
    if (points) {
        k_msm_recode<true><<<ceil_div_u32(points, CNT_THREADS), CNT_THREADS, 0, st>>>(sp, K, (uint32_t)points, nullptr,
                                                                                      cursor, sorted);
        SCZ_LAUNCH_CHECK(ctx);
    }
    k_msm_accumulate<<<ceil_div_u32(buckets, ACC_THREADS), ACC_THREADS, 0, st>>>(sp, K, (uint32_t)buckets, counts,
                                                                                 cursor, sorted, d_buckets.p, heavy + 1,
                                                                                 heavy);
    SCZ_LAUNCH_CHECK(ctx);
    {
        static bool attr_done = false;
        size_t sh = HEAVY_THREADS * sizeof(G1X);
        if (!attr_done) {
            cudaFuncSetAttribute(k_msm_accumulate_heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
            cudaFuncSetAttribute(k_msm_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)(2 * RED_THREADS * sizeof(G1X)));
            attr_done = true;
        }
        k_msm_accumulate_heavy<<<ctx->sm_count, HEAVY_THREADS, sh, st>>>(sp, K, counts, cursor, sorted, d_buckets.p,
                                                                         heavy + 1, heavy);
        SCZ_LAUNCH_CHECK(ctx);
    }
    k_msm_reduce<<<(uint32_t)windows, RED_THREADS, 2 * RED_THREADS * sizeof(G1X), st>>>(sp, K, d_buckets.p, d_wsums.p);
    SCZ_LAUNCH_CHECK(ctx);
    k_msm_finish<<<ceil_div_u32(batch, 32), 32, 0, st>>>(sp, K, d_wsums.p, d_out);
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_msm_g1_batched_dev(scz_ctx *h, const void *const *d_bases, const void *const *d_scalars, const size_t *lens,
                               size_t batch, void *d_out) {
    if (!h) return SCZ_ERR_BAD_ARG;
    if (batch && (!d_bases || !d_scalars || !lens || !d_out)) return h->c.fail(SCZ_ERR_BAD_ARG, "msm: null argument");
    return msm_g1_batched(&h->c, d_bases, d_scalars, lens, batch, d_out);
}

int32_t scz_msm_g1(scz_ctx *h, const void *bases, const uint8_t *inf_mask, size_t bases_len, const void *scalars,
                   size_t scalars_len, void *out_jac) {
    if (!h) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    if (!out_jac) return c->fail(SCZ_ERR_BAD_ARG, "msm: null output");
    if (bases_len != scalars_len)   // ark-ec returns Err(min len); the reference unwrap()s it (dmsm.rs:23)
        return c->fail(SCZ_ERR_LEN_MISMATCH, "msm: %zu bases vs %zu scalars", bases_len, scalars_len);
    size_t n = bases_len;
    if (n && (!bases || !scalars)) return c->fail(SCZ_ERR_BAD_ARG, "msm: null input");
    DevTmp d_b(c), d_s(c), d_m(c), d_o(c);
    SCZ_TRY(d_b.alloc(n * SCZ_G1_AFFINE_BYTES));
    SCZ_TRY(d_s.alloc(n * SCZ_FR_BYTES));
    SCZ_TRY(d_o.alloc(SCZ_G1_JAC_BYTES));
    if (n) {
        SCZ_CUDA(c, cudaMemcpyAsync(d_b.p, bases, n * SCZ_G1_AFFINE_BYTES, cudaMemcpyHostToDevice, c->stream));
        SCZ_CUDA(c, cudaMemcpyAsync(d_s.p, scalars, n * SCZ_FR_BYTES, cudaMemcpyHostToDevice, c->stream));
        if (inf_mask) {
            SCZ_TRY(d_m.alloc(n));
            SCZ_CUDA(c, cudaMemcpyAsync(d_m.p, inf_mask, n, cudaMemcpyHostToDevice, c->stream));
            SCZ_TRY(scz_g1_apply_inf_mask_dev(h, d_b.p, d_m.as<uint8_t>(), n));
        }
    }
    const void *bp = d_b.p, *sp = d_s.p;
    SCZ_TRY(msm_g1_batched(c, &bp, &sp, &n, 1, d_o.p));
    SCZ_CUDA(c, cudaMemcpyAsync(out_jac, d_o.p, SCZ_G1_JAC_BYTES, cudaMemcpyDeviceToHost, c->stream));
    SCZ_CUDA(c, cudaStreamSynchronize(c->stream));
    return SCZ_OK;
}

int32_t scz_msm_set_window(scz_ctx *h, uint32_t cbits) {
    if (!h || cbits > 20) return SCZ_ERR_BAD_ARG;
    h->c.msm_window_override = cbits;
    return SCZ_OK;
}
int32_t scz_msm_last_stats(const scz_ctx *h, uint64_t *adds, uint64_t *buckets, uint64_t *windows) {
    if (!h) return SCZ_ERR_BAD_ARG;
    if (adds) *adds = h->c.msm_bucket_adds;
    if (buckets) *buckets = h->c.msm_buckets;
    if (windows) *windows = h->c.msm_windows;
    return SCZ_OK;
}

}   // extern "C"
