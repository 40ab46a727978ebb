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
#include <stdlib.h>
#include "common.cuh"
#include "pass_impl.cuh"

namespace attrici {

struct PassArgs {
    const float* ys;
    int64_t ld;
    int Cp;
    const int* list;         // nullable: cells still iterating (cell = list[position]); else cell = position
    int n_active;            // positions
    int Pp;                  // stride of the partials (indexed by position)
    int i0, i1;
    int chunk_len;
    const void* basis;       // float or double [.][20]
    const double* th_pass;   // [NP][Cp]
    const int* status;       // nullable: only cells with RUNNING status are evaluated
    double* part_nll;        // [n_chunks][Pp]
    void* part_acc;          // [n_chunks][NACC][Pp]
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

    const int pos = blockIdx.x * PASS_CELLS + threadIdx.x;
    const int chunk = blockIdx.y;
    const int t_begin = a.i0 + chunk * a.chunk_len;
    const int t_end = min(a.i1, t_begin + a.chunk_len);
    const bool in_range = pos < a.n_active;
    // positions beyond the batch read a clamped (valid) column and are masked by `active`
    const int cell = a.list ? a.list[in_range ? pos : 0] : (in_range ? pos : a.n_active - 1);
    bool active = in_range;
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

    const float* yp = a.ys + (size_t)t_begin * a.ld + cell;
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

    if (in_range) {
        a.part_nll[(size_t)chunk * a.Pp + pos] = nll;
        if (MOMENTS) {
            R* out = reinterpret_cast<R*>(a.part_acc) + (size_t)chunk * NACC * a.Pp + pos;
#pragma unroll
            for (int i = 0; i < 3 * NTRIG; ++i) out[(size_t)(ACC_AA + i) * a.Pp] = acc.AA[i];
#pragma unroll
            for (int i = 0; i < 2 * NTRIG; ++i) out[(size_t)(ACC_AB + i) * a.Pp] = HAS_AB ? acc.AB[i] : R(0);
#pragma unroll
            for (int i = 0; i < NTRIG; ++i) out[(size_t)(ACC_BB + i) * a.Pp] = HAS_B ? acc.BB[i] : R(0);
#pragma unroll
            for (int i = 0; i < 2 * NB; ++i) out[(size_t)(ACC_GA + i) * a.Pp] = acc.GA[i];
#pragma unroll
            for (int i = 0; i < NB; ++i) out[(size_t)(ACC_GB + i) * a.Pp] = HAS_B ? acc.GB[i] : R(0);
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

// ATTRICI_NO_MMA=1 keeps the Normal family on the FMA kernel (A/B measurements only)
static bool use_tensor_cores() {
    static const bool on = [] { const char* e = getenv("ATTRICI_NO_MMA"); return !(e && e[0] == '1'); }();
    return on;
}

// time chunks for the cells of one pass: enough CTAs for ~3 waves of 2 CTAs/SM on 148 SMs, bounded by
// the partial buffers (indexed by position, so fewer cells leave room for more chunks)
PassShape plan_pass(const Workspace& ws, const int* list, int n_active, int i0, int i1, bool fp64) {
    PassShape ps;
    ps.list = list;
    ps.n_active = n_active;
    ps.Pp = round_up(n_active > 0 ? n_active : 1, 32);
    const int window = i1 - i0;
    const int n_cb = (n_active + PASS_CELLS - 1) / PASS_CELLS;
    int want = (148 * 2 * 3 + n_cb - 1) / (n_cb > 0 ? n_cb : 1);
    size_t cap_nll = ws.part_nll_cap / (size_t)ps.Pp;
    size_t cap_acc = (fp64 ? ws.part_acc_cap : 2 * ws.part_acc_cap) / ((size_t)NACC * ps.Pp);
    size_t cap = cap_nll < cap_acc ? cap_nll : cap_acc;
    if (cap > 256) cap = 256;
    if (want > (int)cap) want = (int)cap;
    if (want < 1) want = 1;
    int len = round_up((window + want - 1) / want, PASS_TILE);
    if (len < PASS_TILE) len = PASS_TILE;
    ps.chunk_len = len;
    ps.n_chunks = (window + len - 1) / len;
    return ps;
}

int launch_pass(Workspace& ws, const PassShape& ps, int term, bool fp64, bool with_moments, const float* y_scaled,
                int64_t ld, int i0, int i1, const int* status, cudaStream_t s) {
    if (i1 <= i0) { set_error("empty fit window [%d, %d)", i0, i1); return ATTRICI_EINVAL; }
    if (ps.n_active <= 0) return 0;
    if (term == T_NORMAL && !fp64 && with_moments && use_tensor_cores())
        return launch_pass_mma(ws, ps, y_scaled, ld, i0, i1, status, s);
    PassArgs a;
    a.ys = y_scaled;
    a.ld = ld;
    a.Cp = ws.Cp;
    a.list = ps.list;
    a.n_active = ps.n_active;
    a.Pp = ps.Pp;
    a.i0 = i0;
    a.i1 = i1;
    a.chunk_len = ps.chunk_len;
    a.basis = fp64 ? (const void*)ws.basis64 : (const void*)ws.basis32;
    a.th_pass = ws.th_pass;
    a.status = status;
    a.part_nll = ws.part_nll;
    a.part_acc = ws.part_acc;
    dim3 grid((ps.n_active + PASS_CELLS - 1) / PASS_CELLS, ps.n_chunks);
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
