// FP64 tensor-core (mma.sync m8n8k4 f64) issue rate next to the scalar DFMA rate of fp64_peak.cu, and both together
// (do they share a pipe?).  Decides whether the moment contractions of the folded pass belong on DMMA.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
template <int NMMA, int NFMA>
__global__ void mix_kernel(double* out, int iters, double a, double b) {
    double c[NMMA > 0 ? NMMA : 1][2], f[NFMA > 0 ? NFMA : 1];
#pragma unroll
    for (int i = 0; i < (NMMA > 0 ? NMMA : 1); ++i) c[i][0] = c[i][1] = threadIdx.x + i;
#pragma unroll
    for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) f[i] = threadIdx.x + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NMMA; ++i) dmma(c[i][0], c[i][1], a, b);
#pragma unroll
        for (int i = 0; i < NFMA; ++i) f[i] = f[i] * a + b;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < (NMMA > 0 ? NMMA : 1); ++i) s += c[i][0] + c[i][1];
#pragma unroll
    for (int i = 0; i < (NFMA > 0 ? NFMA : 1); ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int NMMA, int NFMA> void run(int threads, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * blocks_per_sm, iters = 20000;
    double* out; cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    mix_kernel<NMMA, NFMA><<<blocks, threads>>>(out, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    mix_kernel<NMMA, NFMA><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warps = (double)blocks * threads / 32;
    double mma_fma = warps * iters * NMMA * 256.0, sc_fma = warps * iters * NFMA * 32.0;
    printf("threads/blk %d blk/SM %d  DMMA x%d + DFMA x%d per iteration: %.3f ms  tensor %.2f TFMA/s  scalar %.2f TFMA/s  "
           "(%.1f + %.1f FMA/clk/SM @1.965GHz)\n", threads, blocks_per_sm, NMMA, NFMA, ms, mma_fma / ms / 1e9, sc_fma / ms / 1e9,
           mma_fma / (ms * 1e-3) / sms / 1.965e9, sc_fma / (ms * 1e-3) / sms / 1.965e9);
    cudaFree(out);
}
int main() {
    run<8, 0>(256, 4);
    run<4, 0>(128, 2);
    run<0, 8>(256, 4);
    run<4, 8>(256, 4);   // both: per-pipe rates unchanged = separate pipes
    run<8, 8>(256, 4);
    run<2, 16>(256, 4);
    return 0;
}
