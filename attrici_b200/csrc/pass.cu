// One likelihood pass: for every cell, stream its scaled series once and reduce
//   - the negative log-likelihood                                   (FP64 accumulator),
//   - the gradient moments   sum_t r(t) gmt^p trig_j(t)             (j: 1, cos/sin 1..4),
//   - the Fisher moments     sum_t w(t) gmt^p trig_j(t)             (j: 1, cos/sin 1..8)
// from which step.cu assembles gradient and Fisher information of all 27 coefficients.
//
// This replaces the inner work of the reference objective (DistributionScipy.log_likelihood,
// attrici/estimation/model_scipy.py:311-331: np.dot(covariates, weights) + scipy logpdf + np.sum)
// and supplies what the reference obtains by P+1 finite-difference evaluations per gradient.
//
// Mapping: thread = cell (a warp reads 32 neighbouring cells of one day: one 128-byte line of the
// time-major series), CTA = 128 cells x one time chunk; the basis rows of the chunk are staged through
// shared memory in double-buffered 64-day tiles (cp.async) and read back as warp-uniform 128-bit
// broadcasts.  All cross-day reductions stay in registers; a chunk's partial sums are written once.
//
// Products of two Fourier columns are Fourier columns of the sum / difference harmonic, so the
// 27 x 27 information matrix needs only 17 trig moments per (weight, power of gmt) instead of all
// 378 column products: 102 + 27 accumulators per cell.
#include "common.cuh"
#include "pass_impl.cuh"

namespace attrici {

struct PassArgs {
    const float* ys;
    int64_t ld;
    int C, Cp;
    int i0, i1;
    int chunk_len;
    const void* basis;       // float or double [.][20]
    const double* th_pass;   // [NP][Cp]
    const int* status;       // nullable: only cells with RUNNING status are evaluated
    double* part_nll;        // [n_chunks][Cp]
    void* part_acc;          // [n_chunks][NACC][Cp]
};

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

template <typename R> __device__ __forceinline__ Row<R> load_row(const R* p);
template <> __device__ __forceinline__ Row<float> load_row<float>(const float* p) {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 v0 = q[0], v1 = q[1], v2 = q[2], v3 = q[3], v4 = q[4];
    Row<float> r;
    r.g = v0.x;
    r.c[0] = v0.y; r.c[1] = v0.z; r.c[2] = v0.w;
    r.c[3] = v1.x; r.c[4] = v1.y; r.c[5] = v1.z; r.c[6] = v1.w;
    r.c[7] = v2.x;
    r.s[0] = v2.y; r.s[1] = v2.z; r.s[2] = v2.w;
    r.s[3] = v3.x; r.s[4] = v3.y; r.s[5] = v3.z; r.s[6] = v3.w;
    r.s[7] = v4.x;
    return r;
}
template <> __device__ __forceinline__ Row<double> load_row<double>(const double* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    Row<double> r;
    double2 v = q[0];
    r.g = v.x; r.c[0] = v.y;
#pragma unroll
    for (int i = 1; i < 4; ++i) { v = q[i]; r.c[2 * i - 1] = v.x; r.c[2 * i] = v.y; }
    v = q[4]; r.c[7] = v.x; r.s[0] = v.y;
#pragma unroll
    for (int i = 1; i < 4; ++i) { v = q[4 + i]; r.s[2 * i - 1] = v.x; r.s[2 * i] = v.y; }
    v = q[8]; r.s[7] = v.x;
    return r;
}

template <int KIND, typename R, bool MOMENTS>
__global__ void __launch_bounds__(PASS_CELLS, 2) pass_kernel(PassArgs a) {
    constexpr bool HAS_B = (KIND != K_BERNOULLI);
    constexpr bool HAS_AB = (KIND == K_WEIBULL || KIND == K_BETA);
    __shared__ __align__(16) R tile[2][PASS_TILE * B_STRIDE];

    const int cell = blockIdx.x * PASS_CELLS + threadIdx.x;
    const int chunk = blockIdx.y;
    const int t_begin = a.i0 + chunk * a.chunk_len;
    const int t_end = min(a.i1, t_begin + a.chunk_len);
    bool active = cell < a.C;
    if (active && a.status) active = (a.status[cell] == ATTRICI_STATUS_RUNNING);
    if (__syncthreads_or(active) == 0 || t_begin >= t_end) return;  // uniform per CTA
    const bool warp_active = __any_sync(0xffffffffu, active);

    // coefficients of this cell, shared (series-origin) frame, canonical layout
    R thA[NA], thB[NB];
#pragma unroll
    for (int i = 0; i < NA; ++i) thA[i] = active ? (R)a.th_pass[(size_t)i * a.Cp + cell] : R(0);
#pragma unroll
    for (int i = 0; i < NB; ++i) thB[i] = (active && HAS_B) ? (R)a.th_pass[(size_t)(NA + i) * a.Cp + cell] : R(0);

    PassAcc<R> acc;
    acc.clear();
    double nll = 0.0;

    const R* basis = reinterpret_cast<const R*>(a.basis);
    const int n_tiles = (t_end - t_begin + PASS_TILE - 1) / PASS_TILE;
    constexpr int VEC_PER_TILE = PASS_TILE * B_STRIDE * (int)sizeof(R) / 16;
    auto issue_tile = [&](int it) {
        const char* src = reinterpret_cast<const char*>(basis + (size_t)(t_begin + it * PASS_TILE) * B_STRIDE);
        char* dst = reinterpret_cast<char*>(tile[it & 1]);
        for (int v = threadIdx.x; v < VEC_PER_TILE; v += PASS_CELLS) cp_async16(dst + 16 * v, src + 16 * v);
        cp_async_commit();
    };
    issue_tile(0);

    // cells beyond C read a clamped (valid) column and are masked by `active`
    const int col = min(cell, a.C - 1);
    const float* yp = a.ys + (size_t)t_begin * a.ld + col;
    const float qnan = __int_as_float(0x7fc00000);
    constexpr int G = 4;  // days per prefetch group
    float ycur[G], ynext[G];
#pragma unroll
    for (int j = 0; j < G; ++j) ycur[j] = (active && t_begin + j < t_end) ? yp[(size_t)j * a.ld] : qnan;

    for (int it = 0; it < n_tiles; ++it) {
        if (it + 1 < n_tiles) {
            issue_tile(it + 1);
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (warp_active) {
            const R* tb = tile[it & 1];
            const int t_tile = t_begin + it * PASS_TILE;
            const int n_days = min(PASS_TILE, t_end - t_tile);
            for (int d0 = 0; d0 < n_days; d0 += G) {
                // prefetch the next group of observations while this one is being reduced
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    int tn = t_tile + d0 + G + j;
                    ynext[j] = (active && tn < t_end) ? yp[(size_t)(tn - t_begin) * a.ld] : qnan;
                }
#pragma unroll
                for (int j = 0; j < G; ++j) {
                    if (d0 + j < n_days) {
                        const Row<R> b = load_row<R>(tb + (d0 + j) * B_STRIDE);
                        R d = pass_day<KIND, R, MOMENTS>(ycur[j], active, thA, thB, b, acc);
                        nll += (double)d;
                    }
                }
#pragma unroll
                for (int j = 0; j < G; ++j) ycur[j] = ynext[j];
            }
        }
        __syncthreads();
    }

    if (cell < a.C) {
        a.part_nll[(size_t)chunk * a.Cp + cell] = nll;
        if (MOMENTS) {
            R* out = reinterpret_cast<R*>(a.part_acc) + (size_t)chunk * NACC * a.Cp + cell;
#pragma unroll
            for (int i = 0; i < 3 * NTRIG; ++i) out[(size_t)(ACC_AA + i) * a.Cp] = acc.AA[i];
#pragma unroll
            for (int i = 0; i < 2 * NTRIG; ++i) out[(size_t)(ACC_AB + i) * a.Cp] = HAS_AB ? acc.AB[i] : R(0);
#pragma unroll
            for (int i = 0; i < NTRIG; ++i) out[(size_t)(ACC_BB + i) * a.Cp] = HAS_B ? acc.BB[i] : R(0);
#pragma unroll
            for (int i = 0; i < 2 * NB; ++i) out[(size_t)(ACC_GA + i) * a.Cp] = acc.GA[i];
#pragma unroll
            for (int i = 0; i < NB; ++i) out[(size_t)(ACC_GB + i) * a.Cp] = HAS_B ? acc.GB[i] : R(0);
        }
    }
}

template <int KIND> static int launch_kind(const PassArgs& a, dim3 grid, bool fp64, bool moments, cudaStream_t s) {
    if (fp64) {
        if (moments) pass_kernel<KIND, double, true><<<grid, PASS_CELLS, 0, s>>>(a);
        else pass_kernel<KIND, double, false><<<grid, PASS_CELLS, 0, s>>>(a);
    } else {
        if (moments) pass_kernel<KIND, float, true><<<grid, PASS_CELLS, 0, s>>>(a);
        else pass_kernel<KIND, float, false><<<grid, PASS_CELLS, 0, s>>>(a);
    }
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_pass(Workspace& ws, int term, bool fp64, bool with_moments, const float* y_scaled, int64_t ld,
                int i0, int i1, const int* status, cudaStream_t s) {
    if (i1 <= i0) { set_error("empty fit window [%d, %d)", i0, i1); return ATTRICI_EINVAL; }
    PassArgs a;
    a.ys = y_scaled;
    a.ld = ld;
    a.C = ws.C;
    a.Cp = ws.Cp;
    a.i0 = i0;
    a.i1 = i1;
    // chunks cover the fit window; the chunk count chosen for T bounds the partial buffers
    int len = round_up((i1 - i0 + ws.ck.n_chunks - 1) / ws.ck.n_chunks, PASS_TILE);
    if (len < PASS_TILE) len = PASS_TILE;
    a.chunk_len = len;
    ws.used_chunks = (i1 - i0 + len - 1) / len;
    a.basis = fp64 ? (const void*)ws.basis64 : (const void*)ws.basis32;
    a.th_pass = ws.th_pass;
    a.status = status;
    a.part_nll = ws.part_nll;
    a.part_acc = ws.part_acc;
    dim3 grid((ws.C + PASS_CELLS - 1) / PASS_CELLS, ws.used_chunks);
    switch (term) {
        case T_NORMAL: return launch_kind<K_NORMAL>(a, grid, fp64, with_moments, s);
        case T_GAMMA: return launch_kind<K_GAMMA>(a, grid, fp64, with_moments, s);
        case T_BETA: return launch_kind<K_BETA>(a, grid, fp64, with_moments, s);
        case T_WEIBULL: return launch_kind<K_WEIBULL>(a, grid, fp64, with_moments, s);
        case T_BERNOULLI: return launch_kind<K_BERNOULLI>(a, grid, fp64, with_moments, s);
    }
    set_error("unknown likelihood term %d", term);
    return ATTRICI_EINVAL;
}

}  // namespace attrici
