// dev micro-benchmark: the bucket kernel's inner loop (XYZZ += affine, gathered points) with some of the Fq products
// of the mixed add moved to the FP64 pipe (tools/ubench/fq_f64.cuh).  Prints G1 mixed adds/s per policy and checks that every
// policy produces the same points.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -o hybrid hybrid.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../../scalable-collaborative-zksnark_b200/csrc/g1.cuh"
#include "fq_f64.cuh"
using namespace scz;

// MASK bit k: product k runs on the FP64 pipe.  0 u2, 1 s2, 2 pp, 3 ppp, 4 q, 5 rr, 6 zz', 7 zzz'
template <int MASK, int MCHAIN, int K>
__device__ __forceinline__ Fq pmul(const Fq &a, const Fq &b) {
    if constexpr ((MASK >> K) & 1) return f64::fq_mul_f64<MCHAIN>(a, b);
    else return fp_mul(a, b);
}
template <int MASK, int MCHAIN>
__device__ __forceinline__ void add_affine(G1X &acc, const Fq &x2, const Fq &y2) {
    if (acc.is_inf()) {
        acc.x = x2;
        acc.y = y2;
        acc.zz = Fq::one();
        acc.zzz = Fq::one();
        return;
    }
    Fq u2 = pmul<MASK, MCHAIN, 0>(x2, acc.zz);
    Fq s2 = pmul<MASK, MCHAIN, 1>(y2, acc.zzz);
    Fq p = fp_sub(u2, acc.x);
    Fq r = fp_sub(s2, acc.y);
    if (p.is_zero()) {
        if (r.is_zero()) acc = g1x_double_affine(x2, y2);
        else acc = G1X::inf();
        return;
    }
    Fq pp = pmul<MASK, MCHAIN, 2>(p, p);
    Fq rr = pmul<MASK, MCHAIN, 5>(r, r);
    Fq ppp = pmul<MASK, MCHAIN, 3>(p, pp);
    Fq q = pmul<MASK, MCHAIN, 4>(acc.x, pp);
    Fq zz = pmul<MASK, MCHAIN, 6>(acc.zz, pp);
    Fq x3 = fp_sub(fp_sub(fp_sub(rr, ppp), q), q);
    Fq zzz = pmul<MASK, MCHAIN, 7>(acc.zzz, ppp);
    Fq y3 = fp_dot2_sub(r, fp_sub(q, x3), acc.y, ppp);
    acc.x = x3;
    acc.y = y3;
    acc.zz = zz;
    acc.zzz = zzz;
}

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

template <int MASK, int MCHAIN, int MINB>
__global__ void __launch_bounds__(128, MINB) k_acc(const uint4 *pts, uint32_t npts_mask, int T, uint32_t *out) {
    uint32_t t = blockIdx.x * 128 + threadIdx.x;
    uint32_t idx = t * 2654435761u;
    G1X acc = G1X::inf();
    G1Affine p = load_pt(pts, idx & npts_mask);
    for (int j = 0; j < T; j++) {
        idx = idx * 1664525u + 1013904223u;
        G1Affine np = load_pt(pts, (idx >> 8) & npts_mask);   // prefetch the next point during the add
        add_affine<MASK, MCHAIN>(acc, p.x, p.y);
        p = np;
    }
    uint32_t *o = out + (size_t)t * 48;
#pragma unroll
    for (int i = 0; i < 12; i++) o[i] = acc.x.l[i], o[12 + i] = acc.y.l[i], o[24 + i] = acc.zz.l[i], o[36 + i] = acc.zzz.l[i];
}

// points: k * G for a few k would need a generator; any curve points do.  Build them on the device from one
// point by repeated doubling / adding (affine conversion is not needed: use x = X/ZZ via a host-side trick instead).
// Simpler: the bench does not need valid curve points for timing, but the exceptional branches must not trigger and
// results must be comparable across policies -- random field elements satisfy both (the formulas are polynomial).
static void fill_random(std::vector<uint32_t> &v) {
    uint64_t s = 0x5CA1AB1E12345ull;
    for (size_t i = 0; i < v.size(); i++) {
        s = s * 6364136223846793005ull + 1442695040888963407ull;
        v[i] = (uint32_t)(s >> 32);
        if (i % 12 == 11) v[i] &= 0x0fffffffu;   // below p
    }
}

template <int MASK, int MCHAIN, int MINB>
static void run(const char *name, const uint4 *pts, uint32_t mask, int T, int blocks, uint32_t *out, std::vector<uint32_t> &ref) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_acc<MASK, MCHAIN, MINB>);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_acc<MASK, MCHAIN, MINB>, 128, 0);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_acc<MASK, MCHAIN, MINB><<<blocks, 128>>>(pts, mask, T, out);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; r++) k_acc<MASK, MCHAIN, MINB><<<blocks, 128>>>(pts, mask, T, out);
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
    printf("%-34s regs %3d spill %4zu B  CTAs/SM %d: %8.3f ms  %.3f G adds/s  [%s] %s\n", name, fa.numRegs, (size_t)fa.localSizeBytes, occ,
           ms, adds / ms / 1e6, ok, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

int main(int argc, char **argv) {
    int T = argc > 1 ? atoi(argv[1]) : 256;
    int waves = argc > 2 ? atoi(argv[2]) : 4;
    uint32_t npts = 1u << 16;
    std::vector<uint32_t> h((size_t)npts * 24);
    fill_random(h);
    uint4 *pts;
    cudaMalloc(&pts, h.size() * 4);
    cudaMemcpy(pts, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    int blocks = 148 * 2 * waves;
    uint32_t *out;
    cudaMalloc(&out, (size_t)blocks * 128 * 48 * 4);
    std::vector<uint32_t> ref;
#define RUN(MASK, MCHAIN, MINB) run<MASK, MCHAIN, MINB>("mask " #MASK " mchain " #MCHAIN " minb " #MINB, pts, npts - 1, T, blocks, out, ref)
    RUN(0x00, 0, 1);
    RUN(0x00, 0, 2);
    RUN(0xff, 0, 1);
    RUN(0xff, 1, 1);
    RUN(0x72, 0, 1);   // s2, q, rr, zz'
    RUN(0x72, 1, 1);
    RUN(0x72, 0, 2);
    RUN(0x72, 1, 2);
    RUN(0x52, 0, 1);   // s2, q, zz'
    RUN(0x52, 1, 1);
    RUN(0xf2, 0, 1);   // s2, q, rr, zz', zzz'
    RUN(0xf2, 1, 1);
    RUN(0x32, 1, 1);   // s2, q, rr
    RUN(0x22, 1, 1);   // s2, rr
    RUN(0x02, 1, 1);   // s2
    return 0;
}
