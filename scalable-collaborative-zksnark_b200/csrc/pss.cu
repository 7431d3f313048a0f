// Packed secret sharing on the device: secret-sharing/src/pss.rs:38-171.
//
// The reference runs two radix-2 (coset) FFTs per call over a vector of n = 8l
// elements (ark-poly Radix2EvaluationDomain).  Every one of the four maps is a
// fixed Fr-linear map, so the FFT pair is folded ONCE, at scz_pp_new, into a small
// matrix (built on the device, no host field code):
//   pack_from_public  shares  = F_share * IF_secret * pad_2l(secrets)            (pss.rs:93-99)
//   pack_single       shares  = PACK * trunc_2l(PACK * [s,0,..])                 (pss.rs:103-113)
//   unpack            secrets = first l of F_secret  * trunc_2l(IF_share * shares) (pss.rs:132-149)
//   unpack2           secrets = even idx < 2l of F_secret2 * trunc_4l(IF_share * shares) (pss.rs:153-171)
// where ark-poly's fft/ifft first RESIZE the vector to the domain size (truncating
// if longer).  All domains are subgroups of <w_n>, so one power table of w_n serves
// every twiddle.  Applying a map is then a batched small mat-vec: one thread per
// output element over Fr, and for G1 operands (the d_msm leader closure) one CTA
// per output point, one scalar multiplication per thread, shared-memory tree sum.
#include "g1_coop.cuh"
#include <string.h>
#include <vector>

#include "pss.h"

namespace scz {

__device__ __forceinline__ Fr fr_from_u32(uint32_t v) {
    Fr x = Fr::zero();
    x.l[0] = v;
    return fp_from_canon(x);
}

// tables layout (Fr each): wn[n] | gpow[4l] | ginvpow[2l] | ninv | inv2l
__global__ void __launch_bounds__(256) k_pss_setup(uint32_t l, void *tables, void *pack, void *pack_single,
                                                    void *unpack, void *unpack2, void *dmsm) {
    const uint32_t n = 8 * l, s1 = 2 * l, s2 = 4 * l;
    Fr *wn = reinterpret_cast<Fr *>(tables);
    Fr *gpow = wn + n, *ginvpow = gpow + s2, *consts = ginvpow + s1;
    if (threadIdx.x == 0) {
        // F::GENERATOR = 7; TWO_ADIC_ROOT_OF_UNITY = 7^((r-1)/2^32)  (2-adicity 32)
        Fr g = fr_from_u32(7);
        uint32_t e[7];
#pragma unroll
        for (int i = 0; i < 7; i++) e[i] = FrP::mod(i + 1);
        Fr w = fp_pow(g, e);
        uint32_t logn = 31 - __clz(n);
        for (uint32_t i = 0; i < 32 - logn; i++) w = fp_sqr(w);   // get_root_of_unity(n)
        Fr acc = Fr::one();
        for (uint32_t k = 0; k < n; k++) {
            wn[k] = acc;
            acc = fp_mul(acc, w);
        }
        acc = Fr::one();
        for (uint32_t i = 0; i < s2; i++) {
            gpow[i] = acc;
            acc = fp_mul(acc, g);
        }
        Fr gi = fp_inv(g);
        acc = Fr::one();
        for (uint32_t i = 0; i < s1; i++) {
            ginvpow[i] = acc;
            acc = fp_mul(acc, gi);
        }
        consts[0] = fp_inv(fr_from_u32(n));
        consts[1] = fp_inv(fr_from_u32(s1));
    }
    __threadfence_block();
    __syncthreads();
    const Fr ninv = consts[0], inv2l = consts[1];
    Fr *P = reinterpret_cast<Fr *>(pack), *PS = reinterpret_cast<Fr *>(pack_single);
    Fr *U = reinterpret_cast<Fr *>(unpack), *U2 = reinterpret_cast<Fr *>(unpack2);
    // PACK[j][k] = sum_{i<2l} w_n^(i j) * (2l)^-1 g^-i w_2l^(-i k),  w_2l = w_n^4
    for (uint32_t idx = threadIdx.x; idx < n * s1; idx += blockDim.x) {
        uint32_t j = idx / s1, k = idx % s1;
        Fr acc = Fr::zero();
        for (uint32_t i = 0; i < s1; i++) {
            uint32_t e1 = (i * j) % n, e2 = (n - (4 * i * k) % n) % n;
            acc = fp_add(acc, fp_mul(fp_mul(wn[e1], wn[e2]), ginvpow[i]));
        }
        P[idx] = fp_mul(acc, inv2l);
    }
    // UNPACK[k][j]  = sum_{i<2l} (g w_2l^k)^i * n^-1 w_n^(-i j)          (rows k < l)
    // UNPACK2[k][j] = sum_{i<4l} (g w_4l^(2k))^i * n^-1 w_n^(-i j)       (rows 2k, k < l; w_4l = w_n^2)
    for (uint32_t idx = threadIdx.x; idx < l * n; idx += blockDim.x) {
        uint32_t k = idx / n, j = idx % n;
        Fr a1 = Fr::zero(), a2 = Fr::zero();
        for (uint32_t i = 0; i < s2; i++) {
            uint32_t ej = (n - (i * j) % n) % n;
            if (i < s1) a1 = fp_add(a1, fp_mul(fp_mul(gpow[i], wn[(4 * k * i) % n]), wn[ej]));
            a2 = fp_add(a2, fp_mul(fp_mul(gpow[i], wn[(2 * (2 * k) * i) % n]), wn[ej]));
        }
        U[idx] = fp_mul(a1, ninv);
        U2[idx] = fp_mul(a2, ninv);
    }
    __threadfence_block();
    __syncthreads();
    // PACK_SINGLE[j] = sum_{i<2l} PACK[j][i] * PACK[i][0]
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
        Fr acc = Fr::zero();
        for (uint32_t i = 0; i < s1; i++) acc = fp_add(acc, fp_mul(P[j * s1 + i], P[i * s1]));
        PS[j] = acc;
    }
    // The d_msm leader closure (dmsm.rs:31-38: unpack2, sum the l secrets, replicate, pack) is the rank-one map
    // out_j = p_j * (sum_i u_i * in_i) with u_i = sum_{b<l} UNPACK2[b][i], p_j = sum_{a<l} PACK[j][a]: stored as u | p
    Fr *D = reinterpret_cast<Fr *>(dmsm);
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        Fr pj = Fr::zero(), ui = Fr::zero();
        for (uint32_t a = 0; a < l; a++) {
            pj = fp_add(pj, P[i * s1 + a]);
            ui = fp_add(ui, U2[a * n + i]);
        }
        D[i] = ui;
        D[n + i] = pj;
    }
}

// out(b, o) = sum_j M[o*mcols + j] * in(b, j)   over Fr
__global__ void __launch_bounds__(128) k_pss_apply_fr(const void *M, uint32_t mcols, uint32_t rows, uint32_t len_in,
                                                      const void *in, size_t in_b, size_t in_j, size_t batch,
                                                      void *out, size_t out_b, size_t out_o) {
    size_t idx = blockIdx.x * (size_t)128 + threadIdx.x;
    if (idx >= batch * rows) return;
    size_t b = idx / rows;
    uint32_t o = (uint32_t)(idx % rows);
    Fr acc = Fr::zero();
    for (uint32_t j = 0; j < len_in; j++)
        acc = fp_add(acc, fp_mul(fp_load<FrP>(M, (size_t)o * mcols + j), fp_load_rw<FrP>(in, b * in_b + j * in_j)));
    fp_store<FrP>(out, b * out_b + o * out_o, acc);
}
// same over G1: one CTA per output point.  Term j = M[o][j] * in(b, j) is a 255-bit scalar multiplication: a
// serial chain, run by a group of 4 cooperating lanes (g1_coop.cuh) with a 4-bit window table in shared
// memory; the groups' terms are then tree-summed.
template <int GROUPS>
__global__ void __launch_bounds__(GROUPS * 4) k_pss_apply_g1(const void *M, uint32_t mcols, uint32_t rows,
                                                              uint32_t len_in, const void *in, size_t in_b, size_t in_j,
                                                              void *out, size_t out_b, size_t out_o) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    G1Jac *tab = reinterpret_cast<G1Jac *>(smem_raw);          // [GROUPS][16]
    G1Jac *res = tab + GROUPS * 16;                             // [GROUPS]
    const Coop g;
    const int gi = threadIdx.x >> 2;
    size_t b = blockIdx.x / rows;
    uint32_t o = blockIdx.x % rows;
    G1Jac acc = g1j_inf();
    for (uint32_t j = gi; j < len_in; j += GROUPS) {
        G1Jac p = g1j_load(in, b * in_b + j * in_j);
        Fr k = fp_to_canon(fp_load<FrP>(M, (size_t)o * mcols + j));
        if (p.z.is_zero() || k.is_zero()) continue;
        G1Jac t = coop_mul_bits(g, p, k.l, tab + gi * 16);
        coop_add(g, acc, t);
    }
    if (g.role == 0) res[gi] = acc;
    __syncthreads();
    for (int stride = GROUPS / 2; stride > 0; stride >>= 1) {
        if (gi < stride) {
            G1Jac o2 = res[gi + stride];
            coop_add(g, acc, o2);
        }
        __syncthreads();
        if (gi < stride && g.role == 0) res[gi] = acc;
        __syncthreads();
    }
    if (threadIdx.x == 0) g1j_store(out, b * out_b + o * out_o, acc);
}

// The d_msm leader closure for a LIST of gathered buffers in one launch: job q maps its party-major input
// [party][k] (n x batch Jacobian points) to S_k = sum_j u_j * in(j, k), then out(o, k) = p_o * S_k: 2n scalar
// multiplications per batch entry instead of the n^2 of the dense n x n map (the reference spends O(n log n) in
// its FFTs over points, pss.rs:124-171 + :93-99).  Two launches; a CTA = 8 groups of 4 cooperating lanes.
struct PssG1Job {
    const void *in;
    void *out;
    uint32_t batch, cta_base;
};
// PHASE 1: CTA (entry e, chunk c) sums u_j * in(j, e) over its 8 parties j = 8c .. 8c+7 into partial[e][c].
// PHASE 2: CTA (e, c) adds the n / 8 partials of its entry (S_e) and writes out(o, e) = p_o * S_e, o = 8c .. 8c+7.
// Every scalar multiplication of a phase runs at the same time: the closure costs two scalar-multiplication
// latencies whatever l is.
template <int PHASE>
__global__ void __launch_bounds__(32) k_pss_dmsm_multi(const void *UP, uint32_t n, const PssG1Job *jobs, uint32_t njobs,
                                                        void *partial) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int GROUPS = 8;
    G1Jac *tab = reinterpret_cast<G1Jac *>(smem_raw);          // [GROUPS][16]
    G1Jac *res = tab + GROUPS * 16;                             // [GROUPS]
    const uint32_t C = n / GROUPS, e = blockIdx.x / C, c = blockIdx.x % C;
    uint32_t lo = 0, hi = njobs - 1;
    while (lo < hi) {
        uint32_t mid = (lo + hi + 1) >> 1;
        if (jobs[mid].cta_base <= e) lo = mid;
        else hi = mid - 1;
    }
    const PssG1Job job = jobs[lo];
    const uint32_t b = e - job.cta_base;                        // batch entry of this job
    const Coop g;
    const int gi = threadIdx.x >> 2;
    if (PHASE == 1) {
        const uint32_t j = c * GROUPS + gi;
        G1Jac acc = g1j_inf();
        {
            G1Jac p = g1j_load(job.in, (size_t)j * job.batch + b);
            Fr k = fp_to_canon(fp_load<FrP>(UP, j));
            if (!p.z.is_zero() && !k.is_zero()) acc = coop_mul_bits(g, p, k.l, tab + gi * 16);
        }
        if (g.role == 0) res[gi] = acc;
        __syncwarp();
        for (int stride = GROUPS / 2; stride > 0; stride >>= 1) {
            if (gi < stride) {
                G1Jac o2 = res[gi + stride];
                coop_add(g, acc, o2);
            }
            __syncwarp();
            if (gi < stride && g.role == 0) res[gi] = acc;
            __syncwarp();
        }
        if (threadIdx.x == 0) g1j_store(partial, (size_t)e * C + c, acc);
    } else {
        G1Jac S = g1j_load(partial, (size_t)e * C);
        for (uint32_t q = 1; q < C; q++) {
            G1Jac t = g1j_load(partial, (size_t)e * C + q);
            coop_add(g, S, t);
        }
        // every lane takes S from ONE shared-memory copy: with S read per lane straight from global memory the
        // products below came out wrong on sm_100a / nvcc 12.9 (memcheck and synccheck clean, the loaded value itself
        // correct) -- the cooperative routines want their replicated operand bit-identical AND identically placed
        if (g.role == 0) res[gi] = S;
        __syncwarp();
        S = res[0];
        const uint32_t o = c * GROUPS + gi;
        Fr k = fp_to_canon(fp_load<FrP>(UP, (size_t)n + o));
        G1Jac t = g1j_inf();
        if (!S.z.is_zero() && !k.is_zero()) t = coop_mul_bits(g, S, k.l, tab + gi * 16);
        if (g.role == 0) g1j_store(job.out, (size_t)o * job.batch + b, t);
    }
}
// jobs: host array of (in, out, batch, -) ; cta_base (first entry of the job) is filled here
int32_t pss_dmsm_multi(Ctx *ctx, const scz_pp *pp, const void *jobs_host, size_t njobs) {
    if (!njobs) return SCZ_OK;
    std::vector<PssG1Job> jobs(njobs);
    memcpy(jobs.data(), jobs_host, njobs * sizeof(PssG1Job));
    uint32_t entries = 0;
    for (auto &j : jobs) {
        j.cta_base = entries;
        entries += j.batch;
    }
    if (!entries) return SCZ_OK;
    const uint32_t C = (uint32_t)pp->n / 8;
    DevTmp d(ctx), partial(ctx);
    SCZ_TRY(d.alloc(njobs * sizeof(PssG1Job)));
    SCZ_TRY(partial.alloc((size_t)entries * C * sizeof(G1Jac)));
    SCZ_TRY(ctx->h2d_staged(d.p, jobs.data(), njobs * sizeof(PssG1Job)));
    constexpr size_t SH8 = 8 * 17 * sizeof(G1Jac);
    k_pss_dmsm_multi<1><<<entries * C, 32, SH8, ctx->stream>>>(pp->d_dmsm, (uint32_t)pp->n, d.as<PssG1Job>(), (uint32_t)njobs, partial.p);
    SCZ_LAUNCH_CHECK(ctx);
    k_pss_dmsm_multi<2><<<entries * C, 32, SH8, ctx->stream>>>(pp->d_dmsm, (uint32_t)pp->n, d.as<PssG1Job>(), (uint32_t)njobs, partial.p);
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}

int32_t pss_apply(Ctx *ctx, const scz_pp *pp, PssMap map, int kind, const void *d_in, size_t len_in, size_t in_b,
                  size_t in_j, size_t batch, void *d_out, size_t out_b, size_t out_o) {
    if (!pp || !d_in || !d_out) return ctx->fail(SCZ_ERR_BAD_ARG, "pss: null argument");
    if (kind != 0 && kind != 1) return ctx->fail(SCZ_ERR_BAD_ARG, "pss: kind must be 0 (Fr) or 1 (G1)");
    if (!batch) return SCZ_OK;
    const void *M;
    uint32_t mcols, rows;
    switch (map) {
        case PSS_PACK:
            if (len_in > 2 * pp->l) len_in = 2 * pp->l;   // ifft_in_place truncates to the secret domain
            M = pp->d_pack, mcols = (uint32_t)(2 * pp->l), rows = (uint32_t)pp->n;
            break;
        case PSS_PACK_SINGLE:
            len_in = 1;
            M = pp->d_pack_single, mcols = 1, rows = (uint32_t)pp->n;
            break;
        case PSS_UNPACK:
            len_in = pp->n;
            M = pp->d_unpack, mcols = (uint32_t)pp->n, rows = (uint32_t)pp->l;
            break;
        case PSS_UNPACK2:
            len_in = pp->n;
            M = pp->d_unpack2, mcols = (uint32_t)pp->n, rows = (uint32_t)pp->l;
            break;
        default:
            return ctx->fail(SCZ_ERR_BAD_ARG, "pss: the d_msm closure runs through pss_dmsm_multi");
    }
    if (kind == 0) {
        k_pss_apply_fr<<<ceil_div_u32(batch * rows, 128), 128, 0, ctx->stream>>>(M, mcols, rows, (uint32_t)len_in, d_in,
                                                                                 in_b, in_j, batch, d_out, out_b, out_o);
    } else {
        constexpr size_t SH8 = 8 * 17 * sizeof(G1Jac), SH32 = 32 * 17 * sizeof(G1Jac);
        if (!ctx->attr_pss) {   // per device, so per ctx (78 KB of dynamic shared memory is above the 48 KB default)
            SCZ_CUDA(ctx, cudaFuncSetAttribute(k_pss_apply_g1<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SH32));
            ctx->attr_pss = true;
        }
        if (len_in <= 8)
            k_pss_apply_g1<8><<<(uint32_t)(batch * rows), 32, SH8, ctx->stream>>>(M, mcols, rows, (uint32_t)len_in, d_in,
                                                                                 in_b, in_j, d_out, out_b, out_o);
        else
            k_pss_apply_g1<32><<<(uint32_t)(batch * rows), 128, SH32, ctx->stream>>>(
                M, mcols, rows, (uint32_t)len_in, d_in, in_b, in_j, d_out, out_b, out_o);
    }
    SCZ_LAUNCH_CHECK(ctx);
    return SCZ_OK;
}

}   // namespace scz

using namespace scz;

extern "C" {

int32_t scz_pp_new(scz_ctx *h, size_t l, scz_pp **out) {
    scz::DeviceGuard dg__(h);
    if (!h || !out) return SCZ_ERR_BAD_ARG;
    Ctx *c = &h->c;
    *out = nullptr;
    if (l == 0 || (l & (l - 1)) || l > 1024)   // Radix2EvaluationDomain sizes; n = 8l must divide 2^32
        return c->fail(SCZ_ERR_NOT_POW2, "PackedSharingParams: l = %zu must be a power of two <= 1024", l);
    scz_pp *pp = new scz_pp();
    pp->l = l;
    pp->n = 8 * l;
    pp->t = l - 1;
    pp->device = c->device;
    size_t n = pp->n;
    char *blk = nullptr;
    size_t total = (n * 2 * l + n + 2 * l * n + 2 * n) * 32;
    cudaError_t e = cudaMalloc(&blk, total);
    if (e != cudaSuccess) {
        delete pp;
        return c->cuda(e, "cudaMalloc(pss matrices)");
    }
    pp->d_pack = blk;
    pp->d_pack_single = blk + n * 2 * l * 32;
    pp->d_unpack = blk + (n * 2 * l + n) * 32;
    pp->d_unpack2 = blk + (n * 2 * l + n + l * n) * 32;
    pp->d_dmsm = blk + (n * 2 * l + n + 2 * l * n) * 32;
    DevTmp tables(c);
    int32_t rc = tables.alloc((n + 4 * l + 2 * l + 2) * 32);
    if (rc == SCZ_OK) {
        k_pss_setup<<<1, 256, 0, c->stream>>>((uint32_t)l, tables.p, pp->d_pack, pp->d_pack_single, pp->d_unpack,
                                              pp->d_unpack2, pp->d_dmsm);
        c->launches++;
        e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) rc = c->cuda(e, "k_pss_setup");
    }
    if (rc != SCZ_OK) {
        cudaFree(blk);
        delete pp;
        return rc;
    }
    *out = pp;
    return SCZ_OK;
}
void scz_pp_free(scz_pp *pp) {
    if (!pp) return;
    cudaSetDevice(pp->device);
    cudaFree(pp->d_pack);
    delete pp;
}
int32_t scz_pp_info(const scz_pp *pp, size_t *t, size_t *l, size_t *n) {
    if (!pp) return SCZ_ERR_BAD_ARG;
    if (t) *t = pp->t;
    if (l) *l = pp->l;
    if (n) *n = pp->n;
    return SCZ_OK;
}
static size_t esz(int kind) { return kind == 0 ? SCZ_FR_BYTES : SCZ_G1_JAC_BYTES; }
int32_t scz_pss_pack_from_public_dev(scz_ctx *h, const scz_pp *pp, int32_t kind, const void *in, size_t len_in,
                                     size_t batch, void *out) {
    scz::DeviceGuard dg__(h);
    if (!h || !pp) return SCZ_ERR_BAD_ARG;
    (void)esz;
    return pss_apply(&h->c, pp, PSS_PACK, kind, in, len_in, len_in, 1, batch, out, pp->n, 1);
}
int32_t scz_pss_pack_single_dev(scz_ctx *h, const scz_pp *pp, int32_t kind, const void *in, size_t batch, void *out) {
    scz::DeviceGuard dg__(h);
    if (!h || !pp) return SCZ_ERR_BAD_ARG;
    return pss_apply(&h->c, pp, PSS_PACK_SINGLE, kind, in, 1, 1, 1, batch, out, pp->n, 1);
}
int32_t scz_pss_unpack_dev(scz_ctx *h, const scz_pp *pp, int32_t kind, const void *in, size_t batch, void *out) {
    scz::DeviceGuard dg__(h);
    if (!h || !pp) return SCZ_ERR_BAD_ARG;
    return pss_apply(&h->c, pp, PSS_UNPACK, kind, in, pp->n, pp->n, 1, batch, out, pp->l, 1);
}
int32_t scz_pss_unpack2_dev(scz_ctx *h, const scz_pp *pp, int32_t kind, const void *in, size_t batch, void *out) {
    scz::DeviceGuard dg__(h);
    if (!h || !pp) return SCZ_ERR_BAD_ARG;
    return pss_apply(&h->c, pp, PSS_UNPACK2, kind, in, pp->n, pp->n, 1, batch, out, pp->l, 1);
}

}   // extern "C"
