// FP64 / FP32 FMA throughput microbenchmark (per-SM issue rates for the roofline discussion in DESIGN.md).
#include <cstdio>
#include <cuda_runtime.h>
template <typename T, int ILP>
__global__ void fma_kernel(T* out, int iters, T a, T b) {
    T acc[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = (T)(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) acc[i] = acc[i] * a + b;
    }
    T s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T, int ILP> void run(const char* name, int threads, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sms * blocks_per_sm, iters = 20000;
    T* out; cudaMalloc(&out, sizeof(T) * blocks * threads);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    fma_kernel<T, ILP><<<blocks, threads>>>(out, 100, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e0);
    fma_kernel<T, ILP><<<blocks, threads>>>(out, iters, (T)1.0000001, (T)1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = (double)blocks * threads * iters * ILP;
    printf("%s threads/blk %d blk/SM %d ILP %d: %.2f TFMA/s (%.1f TFLOP/s), %.1f FMA/clk/SM @1.965GHz\n", name, threads,
           blocks_per_sm, ILP, fma / ms / 1e9, 2 * fma / ms / 1e9, fma / (ms * 1e-3) / sms / 1.965e9);
    cudaFree(out);
}
int main() {
    run<double, 8>("fp64", 256, 4);
    run<double, 8>("fp64", 128, 2);   // 8 warps / SM like fold_pass
    run<double, 2>("fp64", 128, 2);
    run<double, 1>("fp64", 128, 2);
    run<float, 8>("fp32", 256, 4);
    run<float, 8>("fp32", 128, 2);
    return 0;
}
