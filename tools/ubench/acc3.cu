// dev micro-benchmark: the bucket kernel's inner loop with the gathered points staged in shared memory by cp.async
// (no registers for the current / prefetched point) so that three CTAs of 128 threads fit an SM instead of two.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -o acc3 acc3.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../scalable-collaborative-zksnark_b200/csrc/g1.cuh"
using namespace scz;

__device__ __forceinline__ G1Affine load_pt(const uint4 *pts, uint32_t i) {
    G1Affine p;
    const uint4 *s = pts + (size_t)i * 6;
    uint4 v[6];
#pragma unroll
    for (int k = 0; k < 6; k++) v[k] = __ldcg(s + k);
#pragma unroll
    for (int k = 0; k < 3; k++) {
        p.x.l[4 * k] = v[k].x, p.x.l[4 * k + 1] = v[k].y, p.x.l[4 * k + 2] = v[k].z, p.x.l[4 * k + 3] = v[k].w;
        p.y.l[4 * k] = v[3 + k].x, p.y.l[4 * k + 1] = v[3 + k].y, p.y.l[4 * k + 2] = v[3 + k].z, p.y.l[4 * k + 3] = v[3 + k].w;
    }
    return p;
}

// baseline: points in registers (what k_msm_accumulate does today)
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_reg(const uint4 *pts, uint32_t npts_mask, int T, uint32_t *out) {
    uint32_t t = blockIdx.x * 128 + threadIdx.x;
    uint32_t idx = t * 2654435761u;
    G1X acc = G1X::inf();
    G1Affine p = load_pt(pts, idx & npts_mask);
    for (int j = 0; j < T; j++) {
        uint32_t neg = idx >> 31;
        idx = idx * 1664525u + 1013904223u;
        G1Affine np = load_pt(pts, (idx >> 8) & npts_mask);
        g1x_add_affine(acc, p, neg != 0);
        p = np;
    }
    uint32_t *o = out + (size_t)t * 48;
#pragma unroll
    for (int i = 0; i < 12; i++) o[i] = acc.x.l[i], o[12 + i] = acc.y.l[i], o[24 + i] = acc.zz.l[i], o[36 + i] = acc.zzz.l[i];
}

// staged: slot s of thread t = 6 x uint4 at sm[(s * 6 + k) * 128 + t] (conflict-free 128-bit accesses)
__device__ __forceinline__ void stage_pt(uint4 *sm, int s, const uint4 *pts, uint32_t i) {
    const uint4 *g = pts + (size_t)i * 6;
#pragma unroll
    for (int k = 0; k < 6; k++) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(sm + (s * 6 + k) * 128 + threadIdx.x);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(g + k));
    }
    asm volatile("cp.async.commit_group;");
}
__device__ __forceinline__ Fq lds_fq(const uint4 *sm, int s, int k0) {
    Fq r;
#pragma unroll
    for (int k = 0; k < 3; k++) {
        uint4 v = sm[(s * 6 + k0 + k) * 128 + threadIdx.x];
        r.l[4 * k] = v.x, r.l[4 * k + 1] = v.y, r.l[4 * k + 2] = v.z, r.l[4 * k + 3] = v.w;
    }
    return r;
}
__device__ __forceinline__ void add_affine_staged(G1X &acc, const uint4 *sm, int s, bool neg) {
    if (acc.is_inf()) {
        acc.x = lds_fq(sm, s, 0);
        Fq y = lds_fq(sm, s, 3);
        acc.y = neg ? fp_neg(y) : y;
        acc.zz = Fq::one();
        acc.zzz = Fq::one();
        return;
    }
    Fq u2 = fp_mul(lds_fq(sm, s, 0), acc.zz);
    Fq s2 = fp_mul(lds_fq(sm, s, 3), acc.zzz);
    if (neg) s2 = fp_neg(s2);
    Fq p = fp_sub(u2, acc.x);
    Fq r = fp_sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) {
            Fq y = lds_fq(sm, s, 3);
            acc = g1x_double_affine(lds_fq(sm, s, 0), neg ? fp_neg(y) : y);
        } else acc = G1X::inf();
        return;
    }
    Fq pp = fp_sqr(p);
    Fq ppp = fp_mul(p, pp);
    Fq q = fp_mul(acc.x, pp);
    Fq x3 = fp_sub(fp_sub(fp_sub(fp_sqr(r), ppp), q), q);
    Fq y3 = fp_dot2_sub(r, fp_sub(q, x3), acc.y, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = fp_mul(acc.zz, pp);
    acc.zzz = fp_mul(acc.zzz, ppp);
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_smem(const uint4 *pts, uint32_t npts_mask, int T, uint32_t *out) {
    extern __shared__ uint4 sm[];
    uint32_t t = blockIdx.x * 128 + threadIdx.x;
    uint32_t idx = t * 2654435761u;
    G1X acc = G1X::inf();
    stage_pt(sm, 0, pts, idx & npts_mask);
    for (int j = 0; j < T; j++) {
        uint32_t neg = idx >> 31;
        idx = idx * 1664525u + 1013904223u;
        stage_pt(sm, (j + 1) & 1, pts, (idx >> 8) & npts_mask);
        asm volatile("cp.async.wait_group 1;" ::: "memory");
        add_affine_staged(acc, sm, j & 1, neg != 0);
    }
    uint32_t *o = out + (size_t)t * 48;
#pragma unroll
    for (int i = 0; i < 12; i++) o[i] = acc.x.l[i], o[12 + i] = acc.y.l[i], o[24 + i] = acc.zz.l[i], o[36 + i] = acc.zzz.l[i];
}

static void fill_random(std::vector<uint32_t> &v) {
    uint64_t s = 0x5CA1AB1E12345ull;
    for (size_t i = 0; i < v.size(); i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        v[i] = (uint32_t)(s >> 32);
        if (i % 12 == 11) v[i] &= 0x0fffffffu;   // below p
    }
}

template <class K>
static void run(const char *name, K kern, size_t smem, const uint4 *pts, uint32_t mask, int T, int waves, uint32_t *out,
                std::vector<uint32_t> &ref) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, kern);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, 128, smem);
    int blocks = 148 * 6 * waves;   // the same work for every variant (6 = lcm of the occupancies compared)
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    kern<<<blocks, 128, smem>>>(pts, mask, T, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; r++) kern<<<blocks, 128, smem>>>(pts, mask, T, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 3;
    std::vector<uint32_t> h((size_t)blocks * 128 * 48);
    cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost);
    const char *ok = "ref";
    if (ref.empty()) ref = h;
    else ok = (ref == h) ? "same" : "DIFFERENT";
    double adds = (double)blocks * 128 * T;
    printf("%-28s regs %3d spill %4zu B  CTAs/SM %d: %8.3f ms  %.3f G adds/s  [%s] %s\n", name, fa.numRegs, (size_t)fa.localSizeBytes, occ,
           ms, adds / ms / 1e6, ok, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

int main(int argc, char **argv) {
    int T = argc > 1 ? atoi(argv[1]) : 256;
    int waves = argc > 2 ? atoi(argv[2]) : 2;
    uint32_t npts = argc > 3 ? (1u << atoi(argv[3])) : (1u << 16);
    std::vector<uint32_t> h((size_t)npts * 24);
    fill_random(h);
    uint4 *pts;
    cudaMalloc(&pts, h.size() * 4);
    cudaMemcpy(pts, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    uint32_t *out;
    cudaMalloc(&out, (size_t)148 * 6 * waves * 128 * 48 * 4);
    std::vector<uint32_t> ref;
    size_t smem = 2 * 6 * 128 * 16;
    run("registers, 2 CTAs", k_reg<2>, 0, pts, npts - 1, T, waves, out, ref);
    run("registers, 3 CTAs (spills)", k_reg<3>, 0, pts, npts - 1, T, waves, out, ref);
    run("staged, 2 CTAs", k_smem<2>, smem, pts, npts - 1, T, waves, out, ref);
    run("staged, 3 CTAs", k_smem<3>, smem, pts, npts - 1, T, waves, out, ref);
    run("staged, 4 CTAs", k_smem<4>, smem, pts, npts - 1, T, waves, out, ref);
    return 0;
}
