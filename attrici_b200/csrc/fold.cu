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
#include <stdlib.h>
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

__device__ __forceinline__ void fcp_async8(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(sa), "l"(gmem));
}

constexpr int FOLD_RING = 8;     // phases whose sums are in flight per thread (cp.async ring)
constexpr int FOLD_TILE = 64;    // phase rows per shared-memory tile

// acc[0] += w, acc[1 + k] += w cos_{k+1}, acc[1 + K + k] += w sin_{k+1}; trig = {cos 1..8, sin 1..8} of the phase
template <int K, typename A> __device__ __forceinline__ void fold_accumulate(A* acc, A w, const A* trig) {
    acc[0] += w;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc[1 + k] = fma(w, trig[k], acc[1 + k]);
        acc[1 + K + k] = fma(w, trig[NH + k], acc[1 + K + k]);
    }
}

// MODE 0: nll only; 1: nll + gradient moments (the Fisher moments of an earlier pass are reused); 2: everything;
// 3: everything at a flat point (only the two intercepts are non-zero, the state fit_init leaves): mu and sigma do
// not depend on the phase, so there is no predictor chain and no exp per phase, and for cells without missing
// days the Fisher moments are e^{-2c} times sums of the phase rows alone, formed once per CTA
template <int MODE, typename RW>
__global__ void __launch_bounds__(PASS_CELLS, MODE >= 2 ? 2 : 3) fold_pass_normal_kernel(FoldArgs a) {
    constexpr bool FLAT = MODE == 3;
    constexpr bool MOMENTS = MODE >= 1;   // gradient moments
    constexpr bool FISHER = MODE >= 2;    // Fisher moments
    __shared__ double flat_sums[FLAT ? 4 * NTRIG : 1];   // sum over the chunk's phases of {n, Sg, Sgg, 2 n} x trig17
    // phase rows of the current tile: FP64 {n, cos 1..8, sin 1..8, sum gmt, sum gmt^2, 0} and an RW copy of the
    // 16 trig columns for the Fisher moments; per-thread ring of the three sums of the next FOLD_RING phases
    __shared__ __align__(16) double tile[FOLD_TILE * B_STRIDE];
    __shared__ __align__(16) RW trig_w[FOLD_TILE * 2 * NH];
    __shared__ double ring[FOLD_RING][3][PASS_CELLS];

    const int tid = threadIdx.x;
    const int pos = blockIdx.x * PASS_CELLS + tid;
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
    for (int i = 0; i < NA; ++i) thA[i] = (active && (!FLAT || i == 0)) ? a.th_pass[(size_t)i * a.Cp + cell] : 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) thB[i] = (active && (!FLAT || i == 0)) ? a.th_pass[(size_t)(NA + i) * a.Cp + cell] : 0.0;
    const double e2_flat = FLAT ? exp(-2.0 * thB[0]) : 0.0;
    double flat_acc = 0.0;   // threads 0 .. 67: one of the 4 x 17 sums each, gathered tile by tile

    RW AA[3 * NTRIG], BB[NTRIG];
    double GA[2 * NB], GB[NB];
    if (FISHER) {
#pragma unroll
        for (int i = 0; i < 3 * NTRIG; ++i) AA[i] = RW(0);
#pragma unroll
        for (int i = 0; i < NTRIG; ++i) BB[i] = RW(0);
    }
    if (MOMENTS) {
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) GA[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NB; ++i) GB[i] = 0.0;
    }
    double nll = 0.0;

    const size_t cp = (size_t)a.Cp;
    const double* ys_p = a.ystat + cell;
    const double* gs_p = a.gstat + cell;
    const int n_ph = d_end - d_begin;
    // one cp.async group per phase (possibly empty) keeps the wait count uniform
    auto issue = [&](int q) {
        if (q < n_ph && active) {
            const double* src = ys_p + (size_t)(d_begin + q) * 3 * cp;
            double* dst = &ring[q % FOLD_RING][0][tid];
            fcp_async8(dst, src);
            fcp_async8(dst + PASS_CELLS, src + cp);
            fcp_async8(dst + 2 * PASS_CELLS, src + 2 * cp);
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
#pragma unroll
    for (int q = 0; q < FOLD_RING - 1; ++q) issue(q);

    for (int q0 = 0; q0 < n_ph; q0 += FOLD_TILE) {
        const int n_rows = min(FOLD_TILE, n_ph - q0);
        __syncthreads();   // the previous tile is no longer read
        for (int i = tid; i < n_rows * B_STRIDE; i += PASS_CELLS) {
            const double v = a.pb64[(size_t)(d_begin + q0) * B_STRIDE + i];
            tile[i] = v;
            const int r = i / B_STRIDE, c = i - r * B_STRIDE;
            if (FISHER && c >= 1 && c <= 2 * NH) trig_w[r * 2 * NH + (c - 1)] = (RW)v;
        }
        __syncthreads();
        if (FLAT && tid < 4 * NTRIG) {
            const int p = tid / NTRIG, j = tid % NTRIG;
            const int wcol = p == 0 ? 0 : (p == 1 ? 1 + 2 * NH : (p == 2 ? 2 + 2 * NH : 0));
            const double wscale = p == 3 ? 2.0 : 1.0;
            for (int r = 0; r < n_rows; ++r) {
                const double* row = tile + r * B_STRIDE;
                flat_acc = fma(wscale * row[wcol], j == 0 ? 1.0 : row[j], flat_acc);
            }
        }
        if (!warp_active) {   // keep the group accounting of this warp's (empty) ring consistent
            for (int r = 0; r < n_rows; ++r) issue(q0 + r + FOLD_RING - 1);
            continue;
        }
        for (int r = 0; r < n_rows; ++r) {
            const int q = q0 + r;
            issue(q + FOLD_RING - 1);
            asm volatile("cp.async.wait_group %0;\n" ::"n"(FOLD_RING - 1));
            const double* row = tile + r * B_STRIDE;
            const double* slot = &ring[q % FOLD_RING][0][tid];
            const double c_sy = active ? slot[0] : 0.0, c_syy = active ? slot[PASS_CELLS] : 0.0,
                         c_syg = active ? slot[2 * PASS_CELLS] : 0.0;
            double n = row[0], sg = row[1 + 2 * NH], sgg = row[2 + 2 * NH];
            if (own_g) {
                const size_t oc = (size_t)(d_begin + q) * 3 * cp;
                n = gs_p[oc]; sg = gs_p[oc + cp]; sgg = gs_p[oc + 2 * cp];
            }
            if (!active) n = sg = sgg = 0.0;
            double av, bv, cv, e2;
            if (FLAT) {
                av = thA[0]; bv = 0.0; cv = thB[0]; e2 = e2_flat;
            } else {
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
                av = ac + as; bv = bc + bs; cv = cc + cs;
                e2 = exp(-2.0 * cv);
            }
            const double R0 = c_sy - av * n - bv * sg;
            const double R1 = c_syg - av * sg - bv * sgg;
            const double Q = c_syy - av * c_sy - bv * c_syg - av * R0 - bv * R1;
            nll += n * (cv + 0.91893853320467274178) + 0.5 * e2 * Q;
            if (FISHER && (!FLAT || own_g)) {   // flat point, no missing day: formed from flat_sums after the loop
                const RW* tw = trig_w + r * 2 * NH;
                fold_accumulate<NH, RW>(AA, (RW)(e2 * n), tw);
                fold_accumulate<NH, RW>(AA + NTRIG, (RW)(e2 * sg), tw);
                fold_accumulate<NH, RW>(AA + 2 * NTRIG, (RW)(e2 * sgg), tw);
                fold_accumulate<NH, RW>(BB, (RW)(2.0 * n), tw);
            }
            if (MOMENTS) {
                fold_accumulate<MAXM, double>(GA, -e2 * R0, row + 1);
                fold_accumulate<MAXM, double>(GA + NB, -e2 * R1, row + 1);
                fold_accumulate<MAXM, double>(GB, n - e2 * Q, row + 1);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    if (FLAT) {
        if (tid < 4 * NTRIG) flat_sums[tid] = flat_acc;
        __syncthreads();
    }

    if (in_range) {
        a.part_nll[(size_t)chunk * a.Pp + pos] = nll;
        if (MOMENTS) {
            double* out = a.part_acc + (size_t)chunk * NACC * a.Pp + pos;
            if (FISHER) {
                const bool from_sums = FLAT && !own_g;
                const double fs = active ? 1.0 : 0.0;
#pragma unroll
                for (int i = 0; i < 3 * NTRIG; ++i)
                    out[(size_t)(ACC_AA + i) * a.Pp] = from_sums ? fs * e2_flat * flat_sums[i] : (double)AA[i];
#pragma unroll
                for (int i = 0; i < 2 * NTRIG; ++i) out[(size_t)(ACC_AB + i) * a.Pp] = 0.0;
#pragma unroll
                for (int i = 0; i < NTRIG; ++i)
                    out[(size_t)(ACC_BB + i) * a.Pp] = from_sums ? fs * flat_sums[3 * NTRIG + i] : (double)BB[i];
            }
#pragma unroll
            for (int i = 0; i < 2 * NB; ++i) out[(size_t)(ACC_GA + i) * a.Pp] = GA[i];
#pragma unroll
            for (int i = 0; i < NB; ++i) out[(size_t)(ACC_GB + i) * a.Pp] = GB[i];
        }
    }
}

// ---- paired phases ------------------------------------------------------------------------------------------
// 1461 / 365.25 = 4 exactly, so phase 1461 - d has the cosines of phase d and the negated sines: with
//     C_x = x_0 + sum_k x_{c,k} cos_k(d) ,  S_x = sum_k x_{s,k} sin_k(d)        (x = a, b, c)
// the predictors of the two phases are C_x + S_x and C_x - S_x, and their moment contributions combine into one
// accumulation with the weights (w + w') on the constant and cosine moments and (w - w') on the sine moments.
// A pair of phases therefore costs one 27-term predictor sum and one set of 95 accumulations instead of two
// (the exp, the residual sums and the loads stay per phase): ~1.6x fewer instructions per phase.  Items are
// d = 0 (alone) and the pairs (d, 1461 - d), d = 1 .. 730; used when the series covers a whole phase cycle.
constexpr int NPAIR = NPHASE / 2 + 1;   // 731
constexpr int PAIR_RING = 4;            // items (two phases each) in flight per thread

template <int K, typename A> __device__ __forceinline__ void fold_accumulate2(A* acc, A wc, A ws, const A* trig) {
    acc[0] += wc;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc[1 + k] = fma(wc, trig[k], acc[1 + k]);
        acc[1 + K + k] = fma(ws, trig[NH + k], acc[1 + K + k]);
    }
}

template <int MODE, typename RW>
__global__ void __launch_bounds__(PASS_CELLS, MODE >= 2 ? 2 : 3) fold_pass_pair_kernel(FoldArgs a) {
    constexpr bool FLAT = MODE == 3;      // the flat start point: only the two intercepts are non-zero
    constexpr bool MOMENTS = MODE >= 1;   // gradient moments
    constexpr bool FISHER = MODE >= 2;    // Fisher moments
    __shared__ double flat_sums[FLAT ? 4 * NTRIG : 1];   // sum over the chunk's phases of {n, Sg, Sgg, 2 n} x trig17
    __shared__ __align__(16) double tile[FOLD_TILE * B_STRIDE];   // rows of the phases d of the tile
    __shared__ double mirror[FOLD_TILE][4];                       // {n, Sg, Sgg} of the phases 1461 - d
    __shared__ __align__(16) RW trig_w[FOLD_TILE * 2 * NH];
    __shared__ double ring[PAIR_RING][6][PASS_CELLS];

    const int tid = threadIdx.x;
    const int pos = blockIdx.x * PASS_CELLS + tid;
    const int chunk = blockIdx.y;
    const int d_begin = chunk * a.chunk_len;
    const int d_end = min(NPAIR, d_begin + a.chunk_len);
    const bool in_range = pos < a.n_active;
    const int cell = a.list ? a.list[in_range ? pos : 0] : (in_range ? pos : a.n_active - 1);
    bool active = in_range;
    if (active && a.status) active = (a.status[cell] == ATTRICI_STATUS_RUNNING);
    if (__syncthreads_or(active) == 0 || d_begin >= d_end) return;   // uniform per CTA
    const bool warp_active = __any_sync(0xffffffffu, active);
    const bool own_g = active && a.miss[cell] != 0;

    double thA[NA], thB[NB];
#pragma unroll
    for (int i = 0; i < NA; ++i) thA[i] = (active && (!FLAT || i == 0)) ? a.th_pass[(size_t)i * a.Cp + cell] : 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) thB[i] = (active && (!FLAT || i == 0)) ? a.th_pass[(size_t)(NA + i) * a.Cp + cell] : 0.0;
    const double e2_flat = FLAT ? exp(-2.0 * thB[0]) : 0.0;
    double flat_acc = 0.0;   // threads 0 .. 67: one of the 4 x 17 sums each, gathered tile by tile

    RW AA[3 * NTRIG], BB[NTRIG];
    double GA[2 * NB], GB[NB];
    if (FISHER) {
#pragma unroll
        for (int i = 0; i < 3 * NTRIG; ++i) AA[i] = RW(0);
#pragma unroll
        for (int i = 0; i < NTRIG; ++i) BB[i] = RW(0);
    }
    if (MOMENTS) {
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) GA[i] = 0.0;
#pragma unroll
        for (int i = 0; i < NB; ++i) GB[i] = 0.0;
    }
    double nll = 0.0;

    const size_t cp = (size_t)a.Cp;
    const double* ys_p = a.ystat + cell;
    const double* gs_p = a.gstat + cell;
    const int n_items = d_end - d_begin;
    // one cp.async group per item (possibly empty) keeps the wait count uniform
    auto issue = [&](int q) {
        if (q < n_items && active) {
            const int d = d_begin + q;
            double* dst = &ring[q % PAIR_RING][0][tid];
            const double* src = ys_p + (size_t)d * 3 * cp;
            fcp_async8(dst, src);
            fcp_async8(dst + PASS_CELLS, src + cp);
            fcp_async8(dst + 2 * PASS_CELLS, src + 2 * cp);
            if (d > 0) {
                const double* srm = ys_p + (size_t)(NPHASE - d) * 3 * cp;
                fcp_async8(dst + 3 * PASS_CELLS, srm);
                fcp_async8(dst + 4 * PASS_CELLS, srm + cp);
                fcp_async8(dst + 5 * PASS_CELLS, srm + 2 * cp);
            }
        }
        asm volatile("cp.async.commit_group;\n" ::);
    };
#pragma unroll
    for (int q = 0; q < PAIR_RING - 1; ++q) issue(q);

    for (int q0 = 0; q0 < n_items; q0 += FOLD_TILE) {
        const int n_rows = min(FOLD_TILE, n_items - q0);
        __syncthreads();   // the previous tile is no longer read
        for (int i = tid; i < n_rows * B_STRIDE; i += PASS_CELLS) {
            const double v = a.pb64[(size_t)(d_begin + q0) * B_STRIDE + i];
            tile[i] = v;
            const int r = i / B_STRIDE, c = i - r * B_STRIDE;
            if (FISHER && c >= 1 && c <= 2 * NH) trig_w[r * 2 * NH + (c - 1)] = (RW)v;
        }
        for (int i = tid; i < n_rows * 3; i += PASS_CELLS) {
            const int r = i / 3, c = i - r * 3;
            const int d = d_begin + q0 + r;
            mirror[r][c] = d > 0 ? a.pb64[(size_t)(NPHASE - d) * B_STRIDE + (c == 0 ? 0 : c + 2 * NH)] : 0.0;
        }
        __syncthreads();
        if (FLAT && tid < 4 * NTRIG) {
            // both phases of a pair: the mirrored phase has the same cosines and the negated sines
            const int p = tid / NTRIG, j = tid % NTRIG;
            const int wcol = p == 0 ? 0 : (p == 1 ? 1 + 2 * NH : (p == 2 ? 2 + 2 * NH : 0));
            const int mcol = p == 0 ? 0 : (p == 1 ? 1 : (p == 2 ? 2 : 0));
            const double wscale = p == 3 ? 2.0 : 1.0;
            const double msign = j > NH ? -1.0 : 1.0;
            for (int r = 0; r < n_rows; ++r) {
                const double* row = tile + r * B_STRIDE;
                const double w = wscale * fma(msign, mirror[r][mcol], row[wcol]);
                flat_acc = fma(w, j == 0 ? 1.0 : row[j], flat_acc);
            }
        }
        if (!warp_active) {   // keep the group accounting of this warp's (empty) ring consistent
            for (int r = 0; r < n_rows; ++r) issue(q0 + r + PAIR_RING - 1);
            continue;
        }
        for (int r = 0; r < n_rows; ++r) {
            const int q = q0 + r;
            const int d = d_begin + q;
            issue(q + PAIR_RING - 1);
            asm volatile("cp.async.wait_group %0;\n" ::"n"(PAIR_RING - 1));
            const double* row = tile + r * B_STRIDE;
            const double* slot = &ring[q % PAIR_RING][0][tid];
            const bool has_m = d > 0;
            const double sy = active ? slot[0] : 0.0, syy = active ? slot[PASS_CELLS] : 0.0,
                         syg = active ? slot[2 * PASS_CELLS] : 0.0;
            const double sym = (active && has_m) ? slot[3 * PASS_CELLS] : 0.0,
                         syym = (active && has_m) ? slot[4 * PASS_CELLS] : 0.0,
                         sygm = (active && has_m) ? slot[5 * PASS_CELLS] : 0.0;
            double n = row[0], sg = row[1 + 2 * NH], sgg = row[2 + 2 * NH];
            double nm = mirror[r][0], sgm = mirror[r][1], sggm = mirror[r][2];
            if (own_g) {
                const size_t oc = (size_t)d * 3 * cp;
                n = gs_p[oc]; sg = gs_p[oc + cp]; sgg = gs_p[oc + 2 * cp];
                if (has_m) {
                    const size_t om = (size_t)(NPHASE - d) * 3 * cp;
                    nm = gs_p[om]; sgm = gs_p[om + cp]; sggm = gs_p[om + 2 * cp];
                }
            }
            if (!active) n = sg = sgg = nm = sgm = sggm = 0.0;
            double av, bv, cv, am, bm, cm, e2, e2m;
            if (FLAT) {
                av = am = thA[0]; bv = bm = 0.0; cv = cm = thB[0]; e2 = e2m = e2_flat;
            } else {
                // cosine and sine halves of the three predictors
                double ca = thA[0], sa = 0.0, cb = thA[1], sb = 0.0, cc = thB[0], sc = 0.0;
#pragma unroll
                for (int k = 0; k < MAXM; ++k) {
                    const double ck = row[1 + k], sk = row[1 + NH + k];
                    ca = fma(thA[2 + k], ck, ca);
                    sa = fma(thA[2 + MAXM + k], sk, sa);
                    cb = fma(thA[2 + 2 * MAXM + k], ck, cb);
                    sb = fma(thA[2 + 3 * MAXM + k], sk, sb);
                    cc = fma(thB[1 + k], ck, cc);
                    sc = fma(thB[1 + MAXM + k], sk, sc);
                }
                av = ca + sa; bv = cb + sb; cv = cc + sc;      // phase d
                am = ca - sa; bm = cb - sb; cm = cc - sc;      // phase 1461 - d
                e2 = exp(-2.0 * cv); e2m = exp(-2.0 * cm);
            }
            const double R0 = sy - av * n - bv * sg, R0m = sym - am * nm - bm * sgm;
            const double R1 = syg - av * sg - bv * sgg, R1m = sygm - am * sgm - bm * sggm;
            const double Q = syy - av * sy - bv * syg - av * R0 - bv * R1;
            const double Qm = syym - am * sym - bm * sygm - am * R0m - bm * R1m;
            nll += n * (cv + 0.91893853320467274178) + 0.5 * e2 * Q + nm * (cm + 0.91893853320467274178) + 0.5 * e2m * Qm;
            if (FISHER && (!FLAT || own_g)) {   // flat point, no missing day: formed from flat_sums after the loop
                const RW* tw = trig_w + r * 2 * NH;
                const double w0 = e2 * n, w0m = e2m * nm, w1 = e2 * sg, w1m = e2m * sgm, w2 = e2 * sgg, w2m = e2m * sggm;
                fold_accumulate2<NH, RW>(AA, (RW)(w0 + w0m), (RW)(w0 - w0m), tw);
                fold_accumulate2<NH, RW>(AA + NTRIG, (RW)(w1 + w1m), (RW)(w1 - w1m), tw);
                fold_accumulate2<NH, RW>(AA + 2 * NTRIG, (RW)(w2 + w2m), (RW)(w2 - w2m), tw);
                fold_accumulate2<NH, RW>(BB, (RW)(2.0 * (n + nm)), (RW)(2.0 * (n - nm)), tw);
            }
            if (MOMENTS) {
                const double g0 = -e2 * R0, g0m = -e2m * R0m, g1 = -e2 * R1, g1m = -e2m * R1m;
                const double gb = n - e2 * Q, gbm = nm - e2m * Qm;
                fold_accumulate2<MAXM, double>(GA, g0 + g0m, g0 - g0m, row + 1);
                fold_accumulate2<MAXM, double>(GA + NB, g1 + g1m, g1 - g1m, row + 1);
                fold_accumulate2<MAXM, double>(GB, gb + gbm, gb - gbm, row + 1);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::);
    if (FLAT) {
        if (tid < 4 * NTRIG) flat_sums[tid] = flat_acc;
        __syncthreads();
    }

    if (in_range) {
        a.part_nll[(size_t)chunk * a.Pp + pos] = nll;
        if (MOMENTS) {
            double* out = a.part_acc + (size_t)chunk * NACC * a.Pp + pos;
            if (FISHER) {
                const bool from_sums = FLAT && !own_g;
                const double fs = active ? 1.0 : 0.0;
#pragma unroll
                for (int i = 0; i < 3 * NTRIG; ++i)
                    out[(size_t)(ACC_AA + i) * a.Pp] = from_sums ? fs * e2_flat * flat_sums[i] : (double)AA[i];
#pragma unroll
                for (int i = 0; i < 2 * NTRIG; ++i) out[(size_t)(ACC_AB + i) * a.Pp] = 0.0;
#pragma unroll
                for (int i = 0; i < NTRIG; ++i)
                    out[(size_t)(ACC_BB + i) * a.Pp] = from_sums ? fs * flat_sums[3 * NTRIG + i] : (double)BB[i];
            }
#pragma unroll
            for (int i = 0; i < 2 * NB; ++i) out[(size_t)(ACC_GA + i) * a.Pp] = GA[i];
#pragma unroll
            for (int i = 0; i < NB; ++i) out[(size_t)(ACC_GB + i) * a.Pp] = GB[i];
        }
    }
}

}  // namespace

int fold_phases(const Workspace& ws) { return ws.T < NPHASE ? ws.T : NPHASE; }

// number of items a folded Normal pass is planned over: phases, or phase pairs when the series covers a whole cycle
int fold_items(const Workspace& ws, int mode, bool exact) {
    static const bool pair_on = [] { const char* e = getenv("ATTRICI_NO_PAIR"); return !(e && e[0] == '1'); }();
    const bool pair = pair_on && ws.T >= NPHASE && !exact;
    return pair ? NPAIR : fold_phases(ws);
}

// mode: PASS_NLL / PASS_GRAD / PASS_FULL; exact = true: Fisher moments accumulated in FP64 too (known-answer entry)
int launch_fold_pass(Workspace& ws, const PassShape& ps, int mode, bool exact, const int* status, cudaStream_t s) {
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
    if (fold_items(ws, mode, exact) == NPAIR && ws.T >= NPHASE) {
        if (mode == PASS_FLAT) fold_pass_pair_kernel<3, float><<<grid, PASS_CELLS, 0, s>>>(a);
        else if (mode == PASS_NLL) fold_pass_pair_kernel<0, float><<<grid, PASS_CELLS, 0, s>>>(a);
        else if (mode == PASS_GRAD) fold_pass_pair_kernel<1, float><<<grid, PASS_CELLS, 0, s>>>(a);
        else fold_pass_pair_kernel<2, float><<<grid, PASS_CELLS, 0, s>>>(a);
        ATTRICI_LAUNCH_CHECK();
        return 0;
    }
    if (mode == PASS_FLAT && !exact) fold_pass_normal_kernel<3, float><<<grid, PASS_CELLS, 0, s>>>(a);
    else if (mode == PASS_NLL) fold_pass_normal_kernel<0, float><<<grid, PASS_CELLS, 0, s>>>(a);
    else if (mode == PASS_GRAD) fold_pass_normal_kernel<1, float><<<grid, PASS_CELLS, 0, s>>>(a);
    else if (exact) fold_pass_normal_kernel<2, double><<<grid, PASS_CELLS, 0, s>>>(a);
    else fold_pass_normal_kernel<2, float><<<grid, PASS_CELLS, 0, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}


// ---------------------------------------------------------------------------------------------------------
// Phase-folded pass of the other families (Gamma, Weibull, Beta, Bernoulli), in two stages.
//
// Their likelihoods are not functions of a few sums, so the series is still streamed once per pass, but in phase
// order: for the days t = d + 1461 j of one phase the 27-term predictor sums collapse to
//     eta_A(t) = a_d + gmt(t) b_d ,  eta_B(t) = c_d ,
// everything that depends on eta_B only (e^c, lgamma / digamma / trigamma of the Gamma shape) is evaluated once
// per phase, and a day contributes to eleven per-phase sums
//     sum w_AA {1, g, g^2}, sum w_AB {1, g}, sum w_BB, sum r_A {1, g}, sum r_B, sum nll, (count)
// instead of 129 moment accumulators.
//   stage 1  fold_stream_kernel<KIND, float, false>: thread = cell, CTA = 128 cells x <= 16 phases; one load, one FMA, the
//            family's day terms and ten accumulations per day; ~60 registers, so enough warps stay resident to hide
//            the latency of the per-day special functions (the day-ordered pass.cu kernel carries the 129
//            accumulators in registers and runs 8 warps per SM).  Writes psums[phase][11][cell] (FP32).
//   stage 2  fold_sums_moments_kernel: folds the per-phase sums into the 129 moments of pass.cu with the trig
//            columns of each phase (once per phase instead of once per day) and writes the chunk partials, so the
//            reduction and the Fisher-scoring step are unchanged.
// FP32 day arithmetic with FP64 accumulation of the nll across phases, as pass_kernel<KIND, float>.
// ---------------------------------------------------------------------------------------------------------
namespace {

constexpr int NSUM = 11;        // per-phase sums: wAA {1,g,g2}, wAB {1,g}, wBB, rA {1,g}, rB, nll, spare
constexpr int ST_MAXPH = 16;    // phases per CTA of stage 1
constexpr int ST_U = 10;        // days per batch of loads

struct StreamArgs {
    const float* ys;
    int64_t ld;
    int T, Cp;
    const int* list;
    int n_active, Pp;
    int i0, i1;
    int n_ph, ph_len, njp;   // ph_len: phases per CTA; njp: days per phase rounded up to whole batches
    const double* pb64;      // phase rows (trig columns)
    const float* basis32;    // gmt(t) = basis32[t * B_STRIDE]
    const double* basis64;
    const double* th_pass;
    const int* status;
    float* psums;            // [n_ph][NSUM][Pp]           (sums mode)
    double* part_nll;        // [chunk][Pp]                 (nll-only mode: one chunk per CTA row)
};

// terms of the Gamma family that depend on eta_B only (shape a = nu^2 = e^{2 eta_B})
template <typename R> struct GammaPhase { R a, lna, lgam_a, psi_a, wbb; };

template <int KIND, typename R>
__device__ __forceinline__ DayTerms<R> gen_day(R y, R ea, R eb, const GammaPhase<R>& gp) {
    if (KIND == K_GAMMA) {
        DayTerms<R> d;
        const R ly = log(y);
        const R lnu = ly - ea;
        const R u = exp(lnu);
        const R dd = lnu + gp.lna + R(1) - u - gp.psi_a;
        d.nll = -((gp.a - R(1)) * ly - gp.a * u - gp.lgam_a - gp.a * ea + gp.a * gp.lna);
        d.rA = -gp.a * (u - R(1));
        d.rB = R(-2) * gp.a * dd;
        d.wAA = gp.a;
        d.wAB = R(0);
        d.wBB = gp.wbb;
        return d;
    }
    return family_day_t<KIND, R>(y, ea, eb);
}

// R = float, NLL_ONLY = false: per-phase sums of all day terms (stage 1 of a likelihood pass)
// R = double, NLL_ONLY = true: the FP64 log posterior of the returned point (model_scipy.py:525), summed per CTA row
template <int KIND, typename R, bool NLL_ONLY>
__global__ void __launch_bounds__(PASS_CELLS) fold_stream_kernel(StreamArgs a) {
    constexpr bool HAS_B = (KIND != K_BERNOULLI);
    constexpr bool HAS_AB = (KIND == K_WEIBULL || KIND == K_BETA);
    extern __shared__ __align__(16) unsigned char st_smem[];
    R* coef = reinterpret_cast<R*>(st_smem);               // [NP][128]
    R* trig = coef + NP * PASS_CELLS;                      // [ST_MAXPH][8]: cos 1..4, sin 1..4
    R* g_s = trig + ST_MAXPH * 2 * MAXM;                   // [ST_MAXPH][njp] gmt of day d + 1461 j

    const int tid = threadIdx.x;
    const int pos = blockIdx.x * PASS_CELLS + tid;
    const int d_begin = blockIdx.y * a.ph_len, d_end = min(a.n_ph, d_begin + a.ph_len);
    const bool in_range = pos < a.n_active;
    const int cell = a.list ? a.list[in_range ? pos : 0] : (in_range ? pos : a.n_active - 1);
    bool active = in_range;
    if (active && a.status) active = (a.status[cell] == ATTRICI_STATUS_RUNNING);
    if (__syncthreads_or(active) == 0 || d_begin >= d_end) return;   // uniform per CTA

#pragma unroll
    for (int i = 0; i < NP; ++i)
        coef[i * PASS_CELLS + tid] = (active && (HAS_B || i < NA)) ? (R)a.th_pass[(size_t)i * a.Cp + cell] : R(0);
    const R* cf = coef + tid;
    const float* yp0 = a.ys + cell;
    const size_t jstride = (size_t)NPHASE * a.ld;
    double nll_total = 0.0;

    for (int d0 = d_begin; d0 < d_end; d0 += ST_MAXPH) {
        const int nph = min(ST_MAXPH, d_end - d0);
        __syncthreads();   // the previous tile (and, the first time, the coefficient columns) are settled
        for (int i = tid; i < nph * 2 * MAXM; i += PASS_CELLS) {
            const int r = i / (2 * MAXM), k = i - r * 2 * MAXM;
            trig[i] = (R)a.pb64[(size_t)(d0 + r) * B_STRIDE + (k < MAXM ? 1 + k : 1 + NH + (k - MAXM))];
        }
        for (int i = tid; i < nph * a.njp; i += PASS_CELLS) {
            const int r = i / a.njp, j = i - r * a.njp;
            const int t = d0 + r + j * NPHASE;
            R g = R(0);
            if (t < a.T) g = (sizeof(R) == 8) ? (R)a.basis64[(size_t)t * B_STRIDE] : (R)a.basis32[(size_t)t * B_STRIDE];
            g_s[i] = g;
        }
        __syncthreads();
        if (!active) continue;
        for (int r = 0; r < nph; ++r) {
            const int d = d0 + r;
            // days of this phase inside the fit window: j in [ja, jb)
            const int nj = (a.T - d + NPHASE - 1) / NPHASE;
            const int ja = a.i0 > d ? (a.i0 - d + NPHASE - 1) / NPHASE : 0;
            const int jb = a.i1 > d ? min(nj, (a.i1 - d + NPHASE - 1) / NPHASE) : 0;
            R w0 = 0, w1 = 0, w2 = 0, ab0 = 0, ab1 = 0, bb = 0, ra0 = 0, ra1 = 0, rb = 0, nl = 0;
            if (ja < jb) {
                const R* tr = trig + r * 2 * MAXM;
                R av = cf[0], bv = cf[1 * PASS_CELLS], cv = HAS_B ? cf[NA * PASS_CELLS] : R(0);
#pragma unroll
                for (int k = 0; k < MAXM; ++k) {
                    const R ck = tr[k], sk = tr[MAXM + k];
                    av = fma(cf[(2 + k) * PASS_CELLS], ck, av);
                    av = fma(cf[(2 + MAXM + k) * PASS_CELLS], sk, av);
                    bv = fma(cf[(2 + 2 * MAXM + k) * PASS_CELLS], ck, bv);
                    bv = fma(cf[(2 + 3 * MAXM + k) * PASS_CELLS], sk, bv);
                    if (HAS_B) {
                        cv = fma(cf[(NA + 1 + k) * PASS_CELLS], ck, cv);
                        cv = fma(cf[(NA + 1 + MAXM + k) * PASS_CELLS], sk, cv);
                    }
                }
                GammaPhase<R> gp = {R(0), R(0), R(0), R(0), R(0)};
                if (KIND == K_GAMMA) {
                    gp.a = exp(R(2) * cv);
                    gp.lna = R(2) * cv;
                    gp.lgam_a = lgam<R>(gp.a);
                    gp.psi_a = digamma<R>(gp.a);
                    gp.wbb = NLL_ONLY ? R(0) : R(4) * gp.a * (gp.a * trigamma<R>(gp.a) - R(1));
                }
                const R* gr = g_s + r * a.njp;
                const float* yp = yp0 + (size_t)d * a.ld + (size_t)ja * jstride;
                for (int j0 = ja; j0 < jb; j0 += ST_U, yp += ST_U * jstride) {
                    float yv[ST_U];
                    const int last = jb - 1 - j0;
#pragma unroll
                    for (int u = 0; u < ST_U; ++u) yv[u] = (u <= last) ? __ldg(yp + (size_t)u * jstride) : 0.f;
#pragma unroll
                    for (int u = 0; u < ST_U; ++u) {
                        if (u <= last) {
                            const float yf = yv[u];
                            bool valid;
                            R y;
                            if (KIND == K_BERNOULLI) {
                                valid = true;                     // every day of the window counts; NaN (dry) is the event
                                y = (yf != yf) ? R(1) : R(0);
                            } else {
                                valid = (yf == yf);
                                y = valid ? (R)yf : R(0.5);
                            }
                            const R g = gr[j0 + u];
                            const DayTerms<R> dt = gen_day<KIND, R>(y, fma(g, bv, av), cv, gp);
                            if (valid) {
                                nl += dt.nll;
                                if (!NLL_ONLY) {
                                    w0 += dt.wAA; w1 = fma(dt.wAA, g, w1); w2 = fma(dt.wAA * g, g, w2);
                                    ra0 += dt.rA; ra1 = fma(dt.rA, g, ra1);
                                    if (HAS_AB) { ab0 += dt.wAB; ab1 = fma(dt.wAB, g, ab1); }
                                    if (HAS_B) { bb += dt.wBB; rb += dt.rB; }
                                }
                            }
                        }
                    }
                }
            }
            if (NLL_ONLY) {
                nll_total += (double)nl;
            } else {
                float* out = a.psums + (size_t)d * NSUM * a.Pp + pos;
                out[0] = (float)w0; out[(size_t)1 * a.Pp] = (float)w1; out[(size_t)2 * a.Pp] = (float)w2;
                out[(size_t)3 * a.Pp] = (float)ab0; out[(size_t)4 * a.Pp] = (float)ab1; out[(size_t)5 * a.Pp] = (float)bb;
                out[(size_t)6 * a.Pp] = (float)ra0; out[(size_t)7 * a.Pp] = (float)ra1; out[(size_t)8 * a.Pp] = (float)rb;
                out[(size_t)9 * a.Pp] = (float)nl;
            }
        }
    }
    if (NLL_ONLY && in_range) a.part_nll[(size_t)blockIdx.y * a.Pp + pos] = nll_total;
}

struct SumsArgs {
    int Cp;
    const int* list;
    int n_active, Pp;
    int n_ph, chunk_len;
    const double* pb64;
    const int* status;
    const float* psums;      // [n_ph][NSUM][Pp]
    double* part_nll;        // [chunk][Pp]
    float* part_acc;         // [chunk][NACC][Pp]
};

template <bool HAS_B, bool HAS_AB>
__global__ void __launch_bounds__(PASS_CELLS, 2) fold_sums_moments_kernel(SumsArgs a) {
    __shared__ float trig[FOLD_TILE * 2 * NH];   // cos 1..8, sin 1..8 per phase
    const int tid = threadIdx.x;
    const int pos = blockIdx.x * PASS_CELLS + tid;
    const int chunk = blockIdx.y;
    const int d_begin = chunk * a.chunk_len;
    const int d_end = min(a.n_ph, d_begin + a.chunk_len);
    const bool in_range = pos < a.n_active;
    const int cell = a.list ? a.list[in_range ? pos : 0] : (in_range ? pos : a.n_active - 1);
    bool active = in_range;
    if (active && a.status) active = (a.status[cell] == ATTRICI_STATUS_RUNNING);
    if (__syncthreads_or(active) == 0 || d_begin >= d_end) return;   // uniform per CTA

    PassAcc<float> acc;
    acc.clear();
    double nll = 0.0;
    const float* ps = a.psums + pos;
    const size_t pp = (size_t)a.Pp;
    for (int q0 = d_begin; q0 < d_end; q0 += FOLD_TILE) {
        const int n_rows = min(FOLD_TILE, d_end - q0);
        __syncthreads();
        for (int i = tid; i < n_rows * 2 * NH; i += PASS_CELLS) {
            const int r = i / (2 * NH), c = i - r * 2 * NH;
            trig[i] = (float)a.pb64[(size_t)(q0 + r) * B_STRIDE + 1 + c];
        }
        __syncthreads();
        if (!active) continue;
        // the sums of the next phase are requested while this one is folded
        float cur[10], nxt[10];
#pragma unroll
        for (int k = 0; k < 10; ++k) cur[k] = ps[((size_t)q0 * NSUM + k) * pp];
        for (int r = 0; r < n_rows; ++r) {
            if (r + 1 < n_rows) {
#pragma unroll
                for (int k = 0; k < 10; ++k) nxt[k] = ps[((size_t)(q0 + r + 1) * NSUM + k) * pp];
            }
            const float* tr = trig + r * 2 * NH;
            fold_accumulate<NH, float>(acc.AA, cur[0], tr);
            fold_accumulate<NH, float>(acc.AA + NTRIG, cur[1], tr);
            fold_accumulate<NH, float>(acc.AA + 2 * NTRIG, cur[2], tr);
            if (HAS_AB) {
                fold_accumulate<NH, float>(acc.AB, cur[3], tr);
                fold_accumulate<NH, float>(acc.AB + NTRIG, cur[4], tr);
            }
            if (HAS_B) {
                fold_accumulate<NH, float>(acc.BB, cur[5], tr);
                fold_accumulate<MAXM, float>(acc.GB, cur[8], tr);
            }
            fold_accumulate<MAXM, float>(acc.GA, cur[6], tr);
            fold_accumulate<MAXM, float>(acc.GA + NB, cur[7], tr);
            nll += (double)cur[9];
#pragma unroll
            for (int k = 0; k < 10; ++k) cur[k] = nxt[k];
        }
    }

    if (in_range) {
        a.part_nll[(size_t)chunk * a.Pp + pos] = nll;
        float* out = a.part_acc + (size_t)chunk * NACC * a.Pp + pos;
#pragma unroll
        for (int i = 0; i < 3 * NTRIG; ++i) out[(size_t)(ACC_AA + i) * a.Pp] = acc.AA[i];
#pragma unroll
        for (int i = 0; i < 2 * NTRIG; ++i) out[(size_t)(ACC_AB + i) * a.Pp] = HAS_AB ? acc.AB[i] : 0.f;
#pragma unroll
        for (int i = 0; i < NTRIG; ++i) out[(size_t)(ACC_BB + i) * a.Pp] = HAS_B ? acc.BB[i] : 0.f;
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) out[(size_t)(ACC_GA + i) * a.Pp] = acc.GA[i];
#pragma unroll
        for (int i = 0; i < NB; ++i) out[(size_t)(ACC_GB + i) * a.Pp] = HAS_B ? acc.GB[i] : 0.f;
    }
}

}  // namespace

// Pass of the non-Normal terms on a gap-free daily axis; `ps` planned over the phases.
//   nll_only = false: FP32 pass with all moments (stage 1 sums, stage 2 moments; partials in float)
//   nll_only = true : FP64 day arithmetic, nll only (part_nll per chunk of `ps`)
int launch_fold_pass_generic(Workspace& ws, const PassShape& ps, int term, bool nll_only, const float* y_scaled, int64_t ld,
                             int i0, int i1, const int* status, cudaStream_t s) {
    if (ps.n_active <= 0) return 0;
    if (!ws.psums) { set_error("workspace was sized for another family"); return ATTRICI_EWORKSPACE; }
    StreamArgs a;
    a.ys = y_scaled; a.ld = ld; a.T = ws.T; a.Cp = ws.Cp;
    a.list = ps.list; a.n_active = ps.n_active; a.Pp = ps.Pp;
    a.i0 = i0; a.i1 = i1;
    a.n_ph = fold_phases(ws);
    a.njp = ((ws.T + NPHASE - 1) / NPHASE + ST_U - 1) / ST_U * ST_U;
    a.pb64 = ws.pb64; a.basis32 = ws.basis32; a.basis64 = ws.basis64; a.th_pass = ws.th_pass; a.status = status;
    a.psums = ws.psums; a.part_nll = ws.part_nll;
    const int n_cb = (ps.n_active + PASS_CELLS - 1) / PASS_CELLS;
    const size_t elem = nll_only ? sizeof(double) : sizeof(float);
    const size_t smem = elem * ((size_t)NP * PASS_CELLS + ST_MAXPH * 2 * MAXM + (size_t)ST_MAXPH * a.njp);
    if (smem > 200 * 1024) { set_error("series too long for the phase-ordered pass (T=%d)", ws.T); return ATTRICI_EUNSUPPORTED; }
    auto launch = [&](auto kernel, dim3 grid) -> int {
        ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, PASS_CELLS, smem, s>>>(a);
        return 0;
    };
    if (nll_only) {
        a.ph_len = ps.chunk_len;   // one CTA row per chunk of the plan: part_nll[chunk][Pp]
        dim3 grid(n_cb, ps.n_chunks);
        switch (term) {
            case T_GAMMA: ATTRICI_TRY(launch(fold_stream_kernel<K_GAMMA, double, true>, grid)); break;
            case T_BETA: ATTRICI_TRY(launch(fold_stream_kernel<K_BETA, double, true>, grid)); break;
            case T_WEIBULL: ATTRICI_TRY(launch(fold_stream_kernel<K_WEIBULL, double, true>, grid)); break;
            case T_BERNOULLI: ATTRICI_TRY(launch(fold_stream_kernel<K_BERNOULLI, double, true>, grid)); break;
            default: set_error("no phase-ordered generic pass for term %d", term); return ATTRICI_EINVAL;
        }
        ATTRICI_LAUNCH_CHECK();
        return 0;
    }
    int want = (148 * 16 + n_cb - 1) / n_cb;
    if (want > a.n_ph) want = a.n_ph;
    if (want < 1) want = 1;
    a.ph_len = (a.n_ph + want - 1) / want;
    if (a.ph_len > ST_MAXPH) a.ph_len = ST_MAXPH;
    dim3 grid1(n_cb, (a.n_ph + a.ph_len - 1) / a.ph_len);
    bool has_b = true, has_ab = false;
    switch (term) {
        case T_GAMMA: ATTRICI_TRY(launch(fold_stream_kernel<K_GAMMA, float, false>, grid1)); break;
        case T_BETA: ATTRICI_TRY(launch(fold_stream_kernel<K_BETA, float, false>, grid1)); has_ab = true; break;
        case T_WEIBULL: ATTRICI_TRY(launch(fold_stream_kernel<K_WEIBULL, float, false>, grid1)); has_ab = true; break;
        case T_BERNOULLI: ATTRICI_TRY(launch(fold_stream_kernel<K_BERNOULLI, float, false>, grid1)); has_b = false; break;
        default: set_error("no phase-ordered generic pass for term %d", term); return ATTRICI_EINVAL;
    }
    ATTRICI_LAUNCH_CHECK();
    SumsArgs b;
    b.Cp = ws.Cp; b.list = ps.list; b.n_active = ps.n_active; b.Pp = ps.Pp;
    b.n_ph = a.n_ph; b.chunk_len = ps.chunk_len; b.pb64 = ws.pb64; b.status = status; b.psums = ws.psums;
    b.part_nll = ws.part_nll; b.part_acc = (float*)ws.part_acc;
    dim3 grid2(n_cb, ps.n_chunks);
    if (has_ab) fold_sums_moments_kernel<true, true><<<grid2, PASS_CELLS, 0, s>>>(b);
    else if (has_b) fold_sums_moments_kernel<true, false><<<grid2, PASS_CELLS, 0, s>>>(b);
    else fold_sums_moments_kernel<false, false><<<grid2, PASS_CELLS, 0, s>>>(b);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
