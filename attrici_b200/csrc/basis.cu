// Shared Fourier / GMT basis (replaces attrici/util.py:85-104 and the covariates of
// attrici/estimation/model_scipy.py:180-197, 280-289).
#include "common.cuh"

namespace attrici {

// one thread per day; columns {gmt, cos(k w t) k=1..8, sin(k w t) k=1..8, tau, 0, 0}
__global__ void build_basis_kernel(const int32_t* __restrict__ days, const double* __restrict__ gmt, int T,
                                   int T_pad, float* __restrict__ b32, double* __restrict__ b64) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T_pad) return;
    double row[B_STRIDE];
#pragma unroll
    for (int i = 0; i < B_STRIDE; ++i) row[i] = 0.0;
    if (t < T) {
        // util.py:102: (t - t.min()) / (365 d + 6 h); the phase origin of the shared basis is the
        // first day of the series, per-cell origins are handled by rotating coefficients (step.cu)
        double tau = (double)(days[t] - days[0]) / 365.25;
        row[0] = gmt[t];
        row[B_TAU] = tau;
        const double two_pi = 6.283185307179586;  // 2 * np.pi
#pragma unroll
        for (int k = 1; k <= NH; ++k) {
            // util.py:103: (2 * np.pi * (np.arange(modes) + 1)) * t_scaled
            double x = (two_pi * (double)k) * tau;
            double s, c;
            sincos(x, &s, &c);
            row[k] = c;
            row[NH + k] = s;
        }
    }
#pragma unroll
    for (int i = 0; i < B_STRIDE; ++i) {
        b64[(size_t)t * B_STRIDE + i] = row[i];
        b32[(size_t)t * B_STRIDE + i] = (float)row[i];
    }
}

// counts the days that break days[t] = days[0] + t.  On a gap-free daily axis (count 0) every Fourier column
// repeats after NPHASE = 1461 days, which is what the phase-folded kernels (fold.cu) build on; any other axis
// keeps the day-by-day kernels.
__global__ void fold_check_kernel(const int32_t* __restrict__ days, int T, int* __restrict__ bad) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    bool b = (t < T) && (days[t] - days[0] != t);
    if (__syncthreads_or(b) && threadIdx.x == 0) atomicAdd(bad, 1);
}

int launch_build_basis(Workspace& ws, const int32_t* days, const double* gmt, cudaStream_t s) {
    int T_pad = ws.T + 2 * PASS_TILE;
    build_basis_kernel<<<(T_pad + 127) / 128, 128, 0, s>>>(days, gmt, ws.T, T_pad, ws.basis32, ws.basis64);
    ATTRICI_LAUNCH_CHECK();
    ATTRICI_CUDA_CHECK(cudaMemsetAsync(ws.fold_bad, 0, sizeof(int), s));
    fold_check_kernel<<<(ws.T + 127) / 128, 128, 0, s>>>(days, ws.T, ws.fold_bad);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
