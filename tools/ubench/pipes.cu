// dev micro-benchmark: issue rates of IMAD.WIDE and DFMA on sm_100a, alone and interleaved
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define ITERS 4096
template <int MODE>
__global__ void k(uint64_t *out, double *dout, uint32_t a, uint32_t b, double x, double y) {
    uint64_t i0 = threadIdx.x, i1 = threadIdx.x + 1, i2 = threadIdx.x + 2, i3 = threadIdx.x + 3;
    double d0 = threadIdx.x, d1 = d0 + 1, d2 = d0 + 2, d3 = d0 + 3;
    uint32_t aa = a + threadIdx.x, bb = b;
#pragma unroll 16
    for (int i = 0; i < ITERS; i++) {
        if (MODE & 1) {
            i0 = (uint64_t)aa * (uint32_t)i1 + i0;
            i1 = (uint64_t)bb * (uint32_t)i2 + i1;
            i2 = (uint64_t)aa * (uint32_t)i3 + i2;
            i3 = (uint64_t)bb * (uint32_t)i0 + i3;
        }
        if (MODE & 2) {
            d0 = fma(d1, x, d0);
            d1 = fma(d2, y, d1);
            d2 = fma(d3, x, d2);
            d3 = fma(d0, y, d3);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = i0 ^ i1 ^ i2 ^ i3;
    dout[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3;
}
template <int MODE>
void run(const char *name, int warps_per_sm) {
    int sms = 148, threads = 128, blocks = sms * warps_per_sm / 4;
    uint64_t *o;
    double *d;
    cudaMalloc(&o, (size_t)blocks * threads * 8);
    cudaMalloc(&d, (size_t)blocks * threads * 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(o, d, 3, 5, 1.0000001, 0.9999999);
    cudaEventRecord(e0);
    for (int r = 0; r < 5; r++) k<MODE><<<blocks, threads>>>(o, d, 3, 5, 1.0000001, 0.9999999);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= 5;
    double ops = (double)blocks * threads * ITERS * 4;   // per pipe
    printf("%-28s warps/SM %2d: %.3f ms  -> %.2f T ops/s per pipe used (%.1f lanes/clk/SM at 1.965 GHz)\n", name, warps_per_sm, ms,
           ops / ms / 1e9, ops / (ms * 1e-3) / 148 / 1.965e9);
    cudaFree(o);
    cudaFree(d);
}
int main() {
    for (int w : {8, 16, 32}) {
        run<1>("IMAD.WIDE only", w);
        run<2>("DFMA only", w);
        run<3>("IMAD.WIDE + DFMA interleaved", w);
    }
    return 0;
}
