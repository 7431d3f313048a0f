// dev micro-benchmark: chains of Fq products on the integer pipe (fp_mul), on the FP64 pipe (fq_mul_f64) and both in
// one interleaved stream (fq_mul_dual).  Cost per product in SM-cycles per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xptxas -v -o dual dual.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../../scalable-collaborative-zksnark_b200/csrc/field.cuh"
#include "fq_f64.cuh"
using namespace scz;

// MODE 0: I only, 1: F only, 2: I then F (source order), 3: dual (interleaved rows), 4: two I, 5: two F
template <int MODE, int MCHAIN, int MINB>
__global__ void __launch_bounds__(128, MINB) k(const uint32_t *in, uint32_t *out, int T) {
    uint32_t t = blockIdx.x * 128 + threadIdx.x;
    Fq x, y, u, v;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        x.l[i] = in[i] + (i == 0 ? t : 0);
        y.l[i] = in[12 + i] ^ (i == 1 ? t : 0);
        u.l[i] = in[24 + i] + (i == 2 ? t : 0);
        v.l[i] = in[36 + i] ^ (i == 3 ? t : 0);
    }
    for (int j = 0; j < T; j++) {
        if (MODE == 0) {
            x = fp_mul(x, y);
            y = fp_mul(y, x);
        } else if (MODE == 1) {
            u = f64::fq_mul_f64<MCHAIN>(u, v);
            v = f64::fq_mul_f64<MCHAIN>(v, u);
        } else if (MODE == 2) {
            x = fp_mul(x, y);
            u = f64::fq_mul_f64<MCHAIN>(u, v);
            y = fp_mul(y, x);
            v = f64::fq_mul_f64<MCHAIN>(v, u);
        } else if (MODE == 3) {
            f64::fq_mul_dual<MCHAIN>(x, x, y, u, u, v);
            f64::fq_mul_dual<MCHAIN>(y, y, x, v, v, u);
        } else if (MODE == 4) {
            x = fp_mul(x, y);
            u = fp_mul(u, v);
            y = fp_mul(y, x);
            v = fp_mul(v, u);
        } else {
            x = f64::fq_mul_f64<MCHAIN>(x, y);
            u = f64::fq_mul_f64<MCHAIN>(u, v);
            y = f64::fq_mul_f64<MCHAIN>(y, x);
            v = f64::fq_mul_f64<MCHAIN>(v, u);
        }
    }
    uint32_t *o = out + (size_t)t * 48;
#pragma unroll
    for (int i = 0; i < 12; i++) o[i] = x.l[i], o[12 + i] = y.l[i], o[24 + i] = u.l[i], o[36 + i] = v.l[i];
}

template <int MODE, int MCHAIN, int MINB>
static void run(const char *name, const uint32_t *in, uint32_t *out, int T, int products_per_iter) {
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k<MODE, MCHAIN, MINB>);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<MODE, MCHAIN, MINB>, 128, 0);
    int blocks = 148 * occ;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE, MCHAIN, MINB><<<blocks, 128>>>(in, out, T);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; r++) k<MODE, MCHAIN, MINB><<<blocks, 128>>>(in, out, T);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 3;
    // per SM sub-partition: occ warps each run T * products_per_iter products in ms
    double cyc = ms * 1e-3 * 1.965e9 / ((double)T * products_per_iter * occ);   // SMSP cycles per warp-product
    printf("%-44s regs %3d spill %3zu B  warps/SMSP %d: %7.3f ms  %7.0f cycles per warp-product  %s\n", name, fa.numRegs,
           (size_t)fa.localSizeBytes, occ, ms, cyc, cudaGetErrorString(cudaGetLastError()));
    fflush(stdout);
}

int main(int argc, char **argv) {
    int T = argc > 1 ? atoi(argv[1]) : 2000;
    uint32_t h[48];
    for (int i = 0; i < 48; i++) h[i] = 0x9e3779b9u * (i + 1);
    for (int i = 11; i < 48; i += 12) h[i] &= 0x0fffffffu;
    uint32_t *in, *out;
    cudaMalloc(&in, sizeof h);
    cudaMemcpy(in, h, sizeof h, cudaMemcpyHostToDevice);
    cudaMalloc(&out, (size_t)148 * 16 * 128 * 48 * 4);
    unsigned sel = argc > 2 ? strtoul(argv[2], 0, 0) : 0xffffffffu;
    int idx = 0;
#define RUN(MODE, MCHAIN, MINB, PPI, NAME) if ((sel >> idx++) & 1) run<MODE, MCHAIN, MINB>(NAME " mchain " #MCHAIN " minb " #MINB, in, out, T, PPI)
    RUN(0, 0, 2, 2, "I only");
    RUN(0, 0, 3, 2, "I only");
    RUN(0, 0, 4, 2, "I only");
    RUN(4, 0, 2, 4, "two independent I");
    RUN(1, 0, 2, 2, "F only");
    RUN(1, 1, 2, 2, "F only");
    RUN(1, 1, 3, 2, "F only");
    RUN(1, 1, 4, 2, "F only");
    RUN(5, 1, 2, 4, "two independent F");
    RUN(2, 0, 2, 4, "I then F in source order");
    RUN(2, 1, 2, 4, "I then F in source order");
    RUN(3, 0, 2, 4, "dual (interleaved rows)");
    RUN(3, 1, 2, 4, "dual (interleaved rows)");
    RUN(3, 1, 3, 4, "dual (interleaved rows)");
    RUN(3, 1, 4, 4, "dual (interleaved rows)");
    return 0;
}
