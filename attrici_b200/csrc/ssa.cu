// SSA smoothing of the GMT predictor: first elementary component of the singular spectrum analysis.
//
// Replaces calc_gmt_by_ssa (attrici/preprocessing.py:13-40), i.e. row 0 of
// SingularSpectrumAnalysis(window_size, groups=None).transform (attrici/vendored/singularspectrumanalysis.py:161-207):
//     X[i][k] = x[i + k]   (W x K trajectory matrix, K = n - W + 1)               :165-169
//     v = leading eigenvector of G = X X^T                                         :170-172
//     E = v v^T X = v u^T,  u = X^T v                                              :174 (_outer_dot :51-57, slice 0)
//     out[t] = mean of E over the anti-diagonal i + k = t                          :183-192 (_diagonal_averaging :60-76)
// The reference materialises all W elementary matrices - a (W, W, K) float tensor, 4.4 GB and ~17 s of host time
// for the 1901-2023 daily series at subset 10 - although only the first is used.  Here: G in O(W^2 K) FP64 FMAs
// (0.55 GFMA for W = 365), the leading eigenvector by power iteration in one CTA (G stays in L2; the residual
// |G v - lambda v| <= tol lambda ends it), u and the diagonal means in O(W n).  Everything FP64; the reference runs
// float32 when the file holds float32.
#include "common.cuh"

namespace attrici {

namespace {

constexpr int GT = 16;          // G tile edge
constexpr int GK = 128;         // k-tile of the Gram kernel
constexpr int PW_THREADS = 1024;

// G[i][j] = sum_k x[i + k] x[j + k], k < K
__global__ void __launch_bounds__(GT * GT) ssa_gram_kernel(const double* __restrict__ x, int W, int K, double* __restrict__ G) {
    __shared__ double xi[GK + GT], xj[GK + GT];
    const int i0 = blockIdx.y * GT, j0 = blockIdx.x * GT;
    if (j0 > i0) return;   // lower triangle; mirrored below
    const int ti = threadIdx.y, tj = threadIdx.x;
    const int tid = ti * GT + tj;
    double acc = 0.0;
    for (int k0 = 0; k0 < K; k0 += GK) {
        const int kn = min(GK, K - k0);
        __syncthreads();
        for (int q = tid; q < kn + GT; q += GT * GT) {
            // x has n = W + K - 1 entries; rows beyond W are never used in a result that is stored
            const int a = i0 + k0 + q, b = j0 + k0 + q;
            xi[q] = a < W + K - 1 ? x[a] : 0.0;
            xj[q] = b < W + K - 1 ? x[b] : 0.0;
        }
        __syncthreads();
#pragma unroll 4
        for (int k = 0; k < kn; ++k) acc = fma(xi[ti + k], xj[tj + k], acc);
    }
    const int i = i0 + ti, j = j0 + tj;
    if (i < W && j < W) {
        G[(size_t)i * W + j] = acc;
        G[(size_t)j * W + i] = acc;
    }
}

__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < PW_THREADS / 32; ++w) s += red[w];   // same order in every thread
    return s;
}

// power iteration on the symmetric positive semi-definite G (one CTA; warp = row, lanes over columns)
// info[0] = iterations, info[1] = 1 if the residual test was met; lambda[0] = leading eigenvalue
__global__ void __launch_bounds__(PW_THREADS) ssa_power_kernel(const double* __restrict__ G, int W, int max_iter, double tol,
                                                              double* __restrict__ v_out, double* __restrict__ lambda,
                                                              int* __restrict__ info) {
    extern __shared__ double pw_smem[];
    double* v = pw_smem;          // [W]
    double* w = v + W;            // [W]
    __shared__ double red[PW_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // start: a positive, slightly uneven vector (the leading eigenvector of the Gram matrix of a series with a
    // non-zero mean is close to the constant one; the unevenness keeps a zero-mean series from starting orthogonal)
    {
        double pn = 0.0;
        for (int i = tid; i < W; i += PW_THREADS) { const double a = 1.0 + 0.5 * cos(1.7 * i + 0.3); w[i] = a; pn = fma(a, a, pn); }
        const double inv = rsqrt(block_sum(pn, red));
        for (int i = tid; i < W; i += PW_THREADS) v[i] = w[i] * inv;
        __syncthreads();
    }
    int it = 0, ok = 0;
    double lam = 0.0;
    for (; it < max_iter; ++it) {
        for (int i = warp; i < W; i += PW_THREADS / 32) {
            const double* g = G + (size_t)i * W;
            double s = 0.0;
            for (int j = lane; j < W; j += 32) s = fma(g[j], v[j], s);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if (lane == 0) w[i] = s;
        }
        __syncthreads();
        double pl = 0.0, pn = 0.0;
        for (int i = tid; i < W; i += PW_THREADS) { pl = fma(v[i], w[i], pl); pn = fma(w[i], w[i], pn); }
        lam = block_sum(pl, red);               // Rayleigh quotient (|v| = 1)
        const double nrm2 = block_sum(pn, red);
        if (!(nrm2 > 0.0)) break;               // zero matrix
        double pr = 0.0;                        // |G v - lam v|^2, formed directly (|w|^2 - lam^2 would cancel)
        for (int i = tid; i < W; i += PW_THREADS) { const double r = fma(-lam, v[i], w[i]); pr = fma(r, r, pr); }
        const double res2 = block_sum(pr, red);
        const double inv = rsqrt(nrm2);
        for (int i = tid; i < W; i += PW_THREADS) v[i] = w[i] * inv;
        __syncthreads();
        if (res2 <= tol * tol * lam * lam) { ok = 1; ++it; break; }
    }
    for (int i = tid; i < W; i += PW_THREADS) v_out[i] = v[i];
    if (tid == 0) { info[0] = it; info[1] = ok; lambda[0] = lam; }
}

// u[k] = sum_i v[i] x[i + k]
__global__ void ssa_project_kernel(const double* __restrict__ x, const double* __restrict__ v, int W, int K,
                                   double* __restrict__ u) {
    extern __shared__ double pv[];
    for (int i = threadIdx.x; i < W; i += blockDim.x) pv[i] = v[i];
    __syncthreads();
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double s = 0.0;
    for (int i = 0; i < W; ++i) s = fma(pv[i], x[i + k], s);
    u[k] = s;
}

// out[t] = mean over i + k = t of v[i] u[k]
__global__ void ssa_diagonal_kernel(const double* __restrict__ v, const double* __restrict__ u, int W, int K,
                                    double* __restrict__ out) {
    extern __shared__ double pv[];
    for (int i = threadIdx.x; i < W; i += blockDim.x) pv[i] = v[i];
    __syncthreads();
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= W + K - 1) return;
    const int ia = max(0, t - K + 1), ib = min(W - 1, t);
    double s = 0.0;
    for (int i = ia; i <= ib; ++i) s = fma(pv[i], u[t - i], s);
    out[t] = s / (double)(ib - ia + 1);
}

struct SsaLayout { double* G; double* v; double* u; double* lambda; int* info; size_t bytes; };
SsaLayout ssa_layout(int n, int W, char* base) {
    size_t off = 0;
    auto take = [&](size_t b) {
        size_t o = off;
        off = Workspace::align(off + b);
        return base ? (void*)(base + o) : (void*)nullptr;
    };
    SsaLayout L;
    L.G = (double*)take(sizeof(double) * (size_t)W * W);
    L.v = (double*)take(sizeof(double) * (size_t)W);
    L.u = (double*)take(sizeof(double) * (size_t)(n - W + 1));
    L.lambda = (double*)take(256);
    L.info = (int*)take(256);
    L.bytes = off;
    return L;
}

}  // namespace
}  // namespace attrici

using namespace attrici;

extern "C" {

int attrici_ssa_workspace_bytes(int n, int window, size_t* bytes) {
    if (!bytes) { set_error("bytes is null"); return ATTRICI_EINVAL; }
    if (window < 2 || window > n) {
        set_error("window_size must be between 2 and n_timestamps (got %d, n = %d)", window, n);
        return ATTRICI_EINVAL;
    }
    *bytes = ssa_layout(n, window, nullptr).bytes;
    return 0;
}

int attrici_ssa_first_component(const double* x, int n, int window, double* out, void* workspace,
                                size_t workspace_bytes, int max_iter, double tol, int32_t* info, double* eigenvalue,
                                void* stream) {
    if (!x || !out || !workspace) { set_error("x / out / workspace is null"); return ATTRICI_EINVAL; }
    if (window < 2 || window > n) {
        set_error("window_size must be between 2 and n_timestamps (got %d, n = %d)", window, n);
        return ATTRICI_EINVAL;
    }
    if (max_iter < 1 || !(tol > 0.0)) { set_error("max_iter / tol must be positive"); return ATTRICI_EINVAL; }
    if (((uintptr_t)workspace & 255) != 0) { set_error("workspace must be 256-byte aligned"); return ATTRICI_EINVAL; }
    const int W = window, K = n - W + 1;
    SsaLayout L = ssa_layout(n, W, (char*)workspace);
    if (L.bytes > workspace_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, L.bytes);
        return ATTRICI_EWORKSPACE;
    }
    const size_t pw_smem = 2 * sizeof(double) * (size_t)W;
    if (pw_smem > 200 * 1024) { set_error("window_size %d too large for the eigenvector kernel", W); return ATTRICI_EUNSUPPORTED; }
    cudaStream_t s = (cudaStream_t)stream;
    const int nt = (W + GT - 1) / GT;
    ssa_gram_kernel<<<dim3(nt, nt), dim3(GT, GT), 0, s>>>(x, W, K, L.G);
    ATTRICI_LAUNCH_CHECK();
    ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(ssa_power_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pw_smem));
    ssa_power_kernel<<<1, PW_THREADS, pw_smem, s>>>(L.G, W, max_iter, tol, L.v, L.lambda, L.info);
    ATTRICI_LAUNCH_CHECK();
    const size_t vs = sizeof(double) * (size_t)W;
    if (vs > 48 * 1024) {
        ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(ssa_project_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs));
        ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(ssa_diagonal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)vs));
    }
    ssa_project_kernel<<<(K + 127) / 128, 128, vs, s>>>(x, L.v, W, K, L.u);
    ATTRICI_LAUNCH_CHECK();
    ssa_diagonal_kernel<<<(n + 127) / 128, 128, vs, s>>>(L.v, L.u, W, K, out);
    ATTRICI_LAUNCH_CHECK();
    if (info) ATTRICI_CUDA_CHECK(cudaMemcpyAsync(info, L.info, 2 * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    if (eigenvalue) ATTRICI_CUDA_CHECK(cudaMemcpyAsync(eigenvalue, L.lambda, sizeof(double), cudaMemcpyDeviceToDevice, s));
    return 0;
}

}  // extern "C"
