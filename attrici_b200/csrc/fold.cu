// Phase-folded likelihood pass of the Normal family.
//
// On a gap-free daily axis every Fourier column of the design matrix repeats after NPHASE = 1461 days
// (tau = day / 365.25, attrici/util.py:102), so for the days t = d + 1461 j of one phase d the linear
// predictors are
//     eta_mu(t) = a_d + gmt(t) b_d ,   eta_sigma(t) = c_d        (a_d, b_d, c_d: 9-term trig sums of phase d)
// and the Normal log-likelihood (scipy.stats.norm.logpdf, attrici/estimation/model_scipy.py:468) of those days
// depends on the observations only through  Sy = sum y, Syy = sum y^2, Syg = sum y gmt  (prep.cu gathers them
// once, FP64) and  n, Sg = sum gmt, Sgg = sum gmt^2  (shared by all cells without missing days):
//     R0 = Sy - a n - b Sg            = sum (y - mu)
//     R1 = Syg - a Sg - b Sgg         = sum (y - mu) gmt
//     Q  = Syy - a Sy - b Syg - a R0 - b R1 = sum (y - mu)^2
//     nll_d = n (c + ln sqrt(2 pi)) + e^{-2c} Q / 2
// The gradient and Fisher moments of pass.cu follow with the per-day weights replaced by per-phase sums
//     r_mu: -e2 {R0, R1}    r_sigma: n - e2 Q    w_mumu: e2 {n, Sg, Sgg}    w_sigmasigma: 2 n     (e2 = e^{-2c}).
// One pass therefore touches 1461 x 3 doubles per cell (35 KB) instead of the cell's 43 464-day series, and
// does 1/30 of the arithmetic; the series itself is read once (prep) and once more by the quantile map.
//
// Mapping: thread = cell, CTA = 128 cells x a chunk of phases; phase rows staged through shared memory
// (double-buffered cp.async tiles, warp-uniform broadcasts); partial sums per chunk in the layout of pass.cu,
// so the reduction and the Fisher-scoring step (step.cu) are unchanged.  FP64 throughout except the Fisher
// moments, which only steer the step and may be accumulated in FP32 (RW = float).
//
// Replaces, like pass.cu, DistributionScipy.log_likelihood (model_scipy.py:311-331) and the P + 1
// finite-difference evaluations per gradient of scipy's L-BFGS-B (model_scipy.py:519-523).
#include "common.cuh"
#include "pass_impl.cuh"

namespace attrici {

namespace {

struct FoldArgs {
    int Cp;
    const int* list;
    int n_active, Pp;
    int n_ph, chunk_len;
    const double* pb64;      // [NPHASE + pad][B_STRIDE]
    const double* th_pass;   // [NP][Cp]
    const int* status;
    const int* miss;         // [Cp]
    const double* ystat;     // [NPHASE][3][Cp]
    const double* gstat;     // [NPHASE][3][Cp]
    double* part_nll;        // [chunk][Pp]
    double* part_acc;        // [chunk][NACC][Pp]
};

__device__ __forceinline__ void fcp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

// acc[0] += w, acc[1 + k] += w cos_{k+1}, acc[1 + K + k] += w sin_{k+1}; row = phase row {n, cos 1..8, sin 1..8, ..}
template <int K, typename A> __device__ __forceinline__ void fold_accumulate(A* acc, A w, const double* row) {
    acc[0] += w;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc[1 + k] = fma(w, (A)row[1 + k], acc[1 + k]);
        acc[1 + K + k] = fma(w, (A)row[1 + NH + k], acc[1 + K + k]);
    }
}

template <bool MOMENTS, typename RW>
__global__ void __launch_bounds__(PASS_CELLS, 2) fold_pass_normal_kernel(FoldArgs a) {
    __shared__ __align__(16) double tile[2][PASS_TILE * B_STRIDE];

    const int pos = blockIdx.x * PASS_CELLS + threadIdx.x;
    const int chunk = blockIdx.y;
    const int d_begin = chunk * a.chunk_len;
    const int d_end = min(a.n_ph, d_begin + a.chunk_len);
    const bool in_range = pos < a.n_active;
    const int cell = a.list ? a.list[in_range ? pos : 0] : (in_range ? pos : a.n_active - 1);
    bool active = in_range;
    if (active && a.status) active = (a.status[cell] == ATTRICI_STATUS_RUNNING);
    if (__syncthreads_or(active) == 0 || d_begin >= d_end) return;   // uniform per CTA
    const bool warp_active = __any_sync(0xffffffffu, active);
    const bool own_g = active && a.miss[cell] != 0;

    double thA[NA], thB[NB];
#pragma unroll
    for (int i = 0; i < NA; ++i) thA[i] = active ? a.th_pass[(size_t)i * a.Cp + cell] : 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) thB[i] = active ? a.th_pass[(size_t)(NA + i) * a.Cp + cell] : 0.0;

    RW AA[3 * NTRIG], BB[NTRIG];
    double GA[2 * NB], GB[NB];
    if (MOMENTS) {
#pragma unroll
        for (int i = 0; i < 3 * NTRIG; ++i) AA[i] = RW(0);
#pragma unroll
        for (int i = 0; i < NTRIG; ++i) BB[i] = RW(0);
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) GA[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NB; ++i) GB[i] = 0.0;
    }
    double nll = 0.0;

    const int n_tiles = (d_end - d_begin + PASS_TILE - 1) / PASS_TILE;
    constexpr int VEC_PER_TILE = PASS_TILE * B_STRIDE * (int)sizeof(double) / 16;
    auto issue_tile = [&](int it) {
        const char* src = reinterpret_cast<const char*>(a.pb64 + (size_t)(d_begin + it * PASS_TILE) * B_STRIDE);
        char* dst = reinterpret_cast<char*>(tile[it & 1]);
        for (int v = threadIdx.x; v < VEC_PER_TILE; v += PASS_CELLS) fcp_async16(dst + 16 * v, src + 16 * v);
        asm volatile("cp.async.commit_group;\n" ::);
    };
    issue_tile(0);

    const size_t cp = (size_t)a.Cp;
    const double* ys_p = a.ystat + cell;
    const double* gs_p = a.gstat + cell;

    for (int it = 0; it < n_tiles; ++it) {
        if (it + 1 < n_tiles) {
            issue_tile(it + 1);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        if (warp_active) {
            const double* tb = tile[it & 1];
            const int d_tile = d_begin + it * PASS_TILE;
            const int n_rows = min(PASS_TILE, d_end - d_tile);
            // the three sums of the first phase; the next phase's are requested while this one is reduced
            size_t o = (size_t)d_tile * 3 * cp;
            double sy = active ? ys_p[o] : 0.0, syy = active ? ys_p[o + cp] : 0.0, syg = active ? ys_p[o + 2 * cp] : 0.0;
            for (int r = 0; r < n_rows; ++r) {
                const double* row = tb + r * B_STRIDE;
                const double c_sy = sy, c_syy = syy, c_syg = syg;
                const size_t oc = (size_t)(d_tile + r) * 3 * cp;
                if (r + 1 < n_rows && active) {
                    const size_t on = oc + 3 * cp;
                    sy = ys_p[on]; syy = ys_p[on + cp]; syg = ys_p[on + 2 * cp];
                }
                double n = row[0], sg = row[1 + 2 * NH], sgg = row[2 + 2 * NH];
                if (own_g) { n = gs_p[oc]; sg = gs_p[oc + cp]; sgg = gs_p[oc + 2 * cp]; }
                if (!active) n = sg = sgg = 0.0;
                // a_d, b_d, c_d: two chains each (cos / sin halves) to shorten the dependent FMA depth
                double ac = thA[0], as = 0.0, bc = thA[1], bs = 0.0, cc = thB[0], cs = 0.0;
#pragma unroll
                for (int k = 0; k < MAXM; ++k) {
                    const double ck = row[1 + k], sk = row[1 + NH + k];
                    ac = fma(thA[2 + k], ck, ac);
                    as = fma(thA[2 + MAXM + k], sk, as);
                    bc = fma(thA[2 + 2 * MAXM + k], ck, bc);
                    bs = fma(thA[2 + 3 * MAXM + k], sk, bs);
                    cc = fma(thB[1 + k], ck, cc);
                    cs = fma(thB[1 + MAXM + k], sk, cs);
                }
                const double av = ac + as, bv = bc + bs, cv = cc + cs;
                const double e2 = exp(-2.0 * cv);
                const double R0 = c_sy - av * n - bv * sg;
                const double R1 = c_syg - av * sg - bv * sgg;
                const double Q = c_syy - av * c_sy - bv * c_syg - av * R0 - bv * R1;
                nll += n * (cv + 0.91893853320467274178) + 0.5 * e2 * Q;
                if (MOMENTS) {
                    fold_accumulate<NH, RW>(AA, (RW)(e2 * n), row);
                    fold_accumulate<NH, RW>(AA + NTRIG, (RW)(e2 * sg), row);
                    fold_accumulate<NH, RW>(AA + 2 * NTRIG, (RW)(e2 * sgg), row);
                    fold_accumulate<NH, RW>(BB, (RW)(2.0 * n), row);
                    fold_accumulate<MAXM, double>(GA, -e2 * R0, row);
                    fold_accumulate<MAXM, double>(GA + NB, -e2 * R1, row);
                    fold_accumulate<MAXM, double>(GB, n - e2 * Q, row);
                }
            }
        }
        __syncthreads();
    }

    if (in_range) {
        a.part_nll[(size_t)chunk * a.Pp + pos] = nll;
        if (MOMENTS) {
            double* out = a.part_acc + (size_t)chunk * NACC * a.Pp + pos;
#pragma unroll
            for (int i = 0; i < 3 * NTRIG; ++i) out[(size_t)(ACC_AA + i) * a.Pp] = (double)AA[i];
#pragma unroll
            for (int i = 0; i < 2 * NTRIG; ++i) out[(size_t)(ACC_AB + i) * a.Pp] = 0.0;
#pragma unroll
            for (int i = 0; i < NTRIG; ++i) out[(size_t)(ACC_BB + i) * a.Pp] = (double)BB[i];
#pragma unroll
            for (int i = 0; i < 2 * NB; ++i) out[(size_t)(ACC_GA + i) * a.Pp] = GA[i];
#pragma unroll
            for (int i = 0; i < NB; ++i) out[(size_t)(ACC_GB + i) * a.Pp] = GB[i];
        }
    }
}

}  // namespace

int fold_phases(const Workspace& ws) { return ws.T < NPHASE ? ws.T : NPHASE; }

// exact = true: Fisher moments accumulated in FP64 too (known-answer entry)
int launch_fold_pass(Workspace& ws, const PassShape& ps, bool with_moments, bool exact, const int* status,
                     cudaStream_t s) {
    if (ps.n_active <= 0) return 0;
    if (!ws.ystat) { set_error("workspace was sized for another family"); return ATTRICI_EWORKSPACE; }
    FoldArgs a;
    a.Cp = ws.Cp;
    a.list = ps.list;
    a.n_active = ps.n_active;
    a.Pp = ps.Pp;
    a.n_ph = fold_phases(ws);
    a.chunk_len = ps.chunk_len;
    a.pb64 = ws.pb64;
    a.th_pass = ws.th_pass;
    a.status = status;
    a.miss = ws.miss;
    a.ystat = ws.ystat;
    a.gstat = ws.gstat;
    a.part_nll = ws.part_nll;
    a.part_acc = (double*)ws.part_acc;
    dim3 grid((ps.n_active + PASS_CELLS - 1) / PASS_CELLS, ps.n_chunks);
    if (!with_moments) fold_pass_normal_kernel<false, float><<<grid, PASS_CELLS, 0, s>>>(a);
    else if (exact) fold_pass_normal_kernel<true, double><<<grid, PASS_CELLS, 0, s>>>(a);
    else fold_pass_normal_kernel<true, float><<<grid, PASS_CELLS, 0, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
