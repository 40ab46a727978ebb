// cudaMemcpy2DAsync rates between pinned host [T x C] arrays and device slabs [T x S] (the host pipeline's copies)
#include <cstdio>
#include <cuda_runtime.h>
int main() {
    const size_t T = 43464, C = 10000;
    float* hy; double* hc; unsigned char* hr;
    cudaMallocHost(&hy, T * C * 4); cudaMallocHost(&hc, T * C * 8); cudaMallocHost(&hr, T * C);
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t S : {512, 1024, 2048, 4096, 10000}) {
        float* dy; double* dc; unsigned char* dr;
        cudaMalloc(&dy, T * S * 4); cudaMalloc(&dc, T * S * 8); cudaMalloc(&dr, T * S);
        float ms;
        cudaMemcpy2DAsync(dy, S * 4, hy, C * 4, S * 4, T, cudaMemcpyHostToDevice, s1); cudaDeviceSynchronize();
        cudaEventRecord(e0, s1);
        cudaMemcpy2DAsync(dy, S * 4, hy, C * 4, S * 4, T, cudaMemcpyHostToDevice, s1);
        cudaEventRecord(e1, s1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("S=%5zu H2D 2D %6.2f GB/s", S, T * S * 4 / ms / 1e6);
        cudaEventRecord(e0, s2);
        cudaMemcpy2DAsync(hc, C * 8, dc, S * 8, S * 8, T, cudaMemcpyDeviceToHost, s2);
        cudaEventRecord(e1, s2); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("  D2H f64 2D %6.2f GB/s", T * S * 8 / ms / 1e6);
        cudaEventRecord(e0, s2);
        cudaMemcpy2DAsync(hr, C, dr, S, S, T, cudaMemcpyDeviceToHost, s2);
        cudaEventRecord(e1, s2); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("  D2H u8 2D %6.2f GB/s", T * S / ms / 1e6);
        // both directions at once
        cudaDeviceSynchronize();
        cudaEventRecord(e0, 0);
        cudaMemcpy2DAsync(dy, S * 4, hy, C * 4, S * 4, T, cudaMemcpyHostToDevice, s1);
        cudaMemcpy2DAsync(hc, C * 8, dc, S * 8, S * 8, T, cudaMemcpyDeviceToHost, s2);
        cudaDeviceSynchronize();
        cudaEventRecord(e1, 0); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("  both %6.2f GB/s (%.2f ms)\n", T * S * 12 / ms / 1e6, ms);
        cudaFree(dy); cudaFree(dc); cudaFree(dr);
    }
    return 0;
}
