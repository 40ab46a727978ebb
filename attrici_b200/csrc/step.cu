// Kernels around the per-cell optimiser mathematics of step_impl.cuh (one warp = one cell, the cell's
// 27 x 27 system in shared memory), the cross-chunk reduction of the pass partials, and the conversions
// between the reference's theta layout and the canonical per-term layout.
#include "common.cuh"
#include "step_impl.cuh"
#include "step_warp.cuh"

namespace attrici {

namespace {

constexpr int STEP_WARPS = 4;
constexpr int STEP_THREADS = 32 * STEP_WARPS;
constexpr int RED_THREADS = 128;
constexpr int RED_GROUP = 2;    // accumulators per thread of the reduction kernel

// table of the information-matrix entries for this launch's (term, modes), built once per CTA
__device__ __forceinline__ void build_h_table(const StepArgs& a, HEntry* tab) {
    for (int e = threadIdx.x; e < NTRI; e += blockDim.x) {
        // e = i (i + 1) / 2 + j, j <= i
        int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        tab[e] = h_entry(a.term, a.modes, i, e - i * (i + 1) / 2);
    }
    __syncthreads();
}

// cell handled by this warp: position in the active list (or the identity), -1 if none
__device__ __forceinline__ int warp_cell(const int* list, int n) {
    int pos = blockIdx.x * STEP_WARPS + (threadIdx.x >> 5);
    if (pos >= n) return -1;
    return list ? list[pos] : pos;
}

__global__ void __launch_bounds__(STEP_THREADS) fit_init_kernel(StepArgs a, int* list_out, int* counter) {
    __shared__ CellScratch S[STEP_WARPS];
    int cell = warp_cell(nullptr, a.C);
    if (cell < 0) return;
    WarpCtx ctx;
    bool run = cell_init(a, cell, S[threadIdx.x >> 5], ctx);
    if (run && ctx.lane() == 0) list_out[atomicAdd(counter, 1)] = cell;
}

// the table once per term (launch_fit_init), so that the step kernel only copies it
__global__ void h_table_kernel(StepArgs a, HEntry* tab) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < NTRI; e += gridDim.x * blockDim.x) {
        int i = (int)((sqrtf(8.f * (float)e + 1.f) - 1.f) * 0.5f);
        while (i * (i + 1) / 2 > e) --i;
        while ((i + 1) * (i + 2) / 2 <= e) ++i;
        tab[e] = h_entry(a.term, a.modes, i, e - i * (i + 1) / 2);
    }
}

__global__ void __launch_bounds__(STEP_THREADS, 5) fit_step_kernel(StepArgs a, const int* list_in, int n_in,
                                                                int* list_out, int* counter, int fresh_fisher,
                                                                const HEntry* __restrict__ tab_g) {
    __shared__ CellScratch S[STEP_WARPS];
    __shared__ HEntry tab[NTRI];
    static_assert(sizeof(HEntry) == 12, "HEntry is copied as three words");
    for (int i = threadIdx.x; i < NTRI * 3; i += STEP_THREADS)
        reinterpret_cast<int*>(tab)[i] = __ldg(reinterpret_cast<const int*>(tab_g) + i);
    __syncthreads();
    int cell = warp_cell(list_in, n_in);
    if (cell < 0) return;
    bool run;
    if (!term_has_ab(a.term)) {   // no cross block: both diagonal blocks are factorised side by side (step_warp.cuh)
        WarpCtxT<true> ctx;
        run = cell_step(a, cell, S[threadIdx.x >> 5], ctx, tab, fresh_fisher != 0);
    } else {
        WarpCtxT<false> ctx;
        run = cell_step(a, cell, S[threadIdx.x >> 5], ctx, tab, fresh_fisher != 0);
    }
    if (run && (threadIdx.x & 31) == 0) list_out[atomicAdd(counter, 1)] = cell;
}

__global__ void __launch_bounds__(STEP_THREADS) fit_final_point_kernel(StepArgs a) {
    __shared__ CellScratch S[STEP_WARPS];
    int cell = warp_cell(nullptr, a.C);
    if (cell < 0) return;
    WarpCtx ctx;
    cell_final_point(a, cell, S[threadIdx.x >> 5], ctx);
}

__global__ void __launch_bounds__(STEP_THREADS) fit_store_kernel(StepArgs a) {
    __shared__ CellScratch S[STEP_WARPS];
    int cell = warp_cell(nullptr, a.C);
    if (cell < 0) return;
    WarpCtx ctx;
    cell_store(a, cell, S[threadIdx.x >> 5], ctx);
}

// known-answer evaluation: objective, gradient, Fisher information at th_trial
__global__ void __launch_bounds__(STEP_THREADS) assemble_kernel(StepArgs a, double* nll, double* grad /*[Cp][NP]*/,
                                                                double* fisher /*[Cp][NTRI], nullable*/) {
    __shared__ CellScratch Ss[STEP_WARPS];
    __shared__ HEntry tab[NTRI];
    build_h_table(a, tab);
    int cell = warp_cell(nullptr, a.C);
    if (cell < 0) return;
    WarpCtx ctx;
    CellScratch& S = Ss[threadIdx.x >> 5];
    for (int i = ctx.lane(); i < NP; i += 32) S.th[i] = a.th_trial[(size_t)cell * NP + i];
    load_moments(a, cell, S, ctx);
    double f = assemble(a, S, fisher != nullptr, ctx, tab);
    if (ctx.lane() == 0) nll[cell] = f;
    if (grad) for (int i = ctx.lane(); i < NP; i += 32) grad[(size_t)cell * NP + i] = S.g[i];
    if (fisher) for (int i = ctx.lane(); i < NTRI; i += 32) fisher[(size_t)cell * NTRI + i] = S.H[i];
}

__global__ void __launch_bounds__(STEP_THREADS) phase_kernel(StepArgs a) {
    int cell = warp_cell(nullptr, a.C);
    if (cell < 0) return;
    WarpCtx ctx;
    cell_phase(a, cell, ctx);
}

// th_trial (cell frame) -> th_pass (shared frame)
__global__ void __launch_bounds__(STEP_THREADS) rotate_kernel(StepArgs a) {
    __shared__ double th[STEP_WARPS][NP];
    int cell = warp_cell(nullptr, a.C);
    if (cell < 0) return;
    WarpCtx ctx;
    double* t = th[threadIdx.x >> 5];
    for (int i = ctx.lane(); i < NP; i += 32) t[i] = a.th_trial[(size_t)cell * NP + i];
    ctx.sync();
    rotate_to_shared(a, cell, t, ctx);
}

// cross-chunk reduction of the pass partials: thread = active cell (coalesced), blockIdx.y = group of
// accumulators.  partials are indexed by list position when a list is given (stride Pp), results by cell.
template <typename R>
__global__ void __launch_bounds__(RED_THREADS) reduce_partials_kernel(const int* __restrict__ list, int n, int Pp, int Cp,
                                                                      int n_chunks, int row0,
                                                                      const double* __restrict__ part_nll,
                                                                      const R* __restrict__ part_acc,
                                                                      double* __restrict__ momd,
                                                                      const int* __restrict__ status) {
    int pos = blockIdx.x * RED_THREADS + threadIdx.x;
    if (pos >= n) return;
    int cell = list ? list[pos] : pos;
    if (status && status[cell] != ATTRICI_STATUS_RUNNING) return;
    int i0 = row0 + blockIdx.y * RED_GROUP;
    for (int i = i0; i < i0 + RED_GROUP && i <= NACC; ++i) {
        double s = 0.0;
        if (i == NACC) {
            for (int ch = 0; ch < n_chunks; ++ch) s += part_nll[(size_t)ch * Pp + pos];
        } else {
            const R* p = part_acc + (size_t)i * Pp + pos;
            for (int ch = 0; ch < n_chunks; ++ch) s += (double)p[(size_t)ch * NACC * Pp];
        }
        momd[(size_t)i * Cp + cell] = s;
    }
}

// reference layout theta[C x P] -> th_trial [Cp][NP] of one term slot (inactive entries zero)
__global__ void load_theta_kernel(int C, int family, int modes, int slot, const double* __restrict__ theta,
                                  double* __restrict__ th_trial) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= C) return;
    const int P = family_num_params(family, modes);
    for (int i = 0; i < NP; ++i) th_trial[(size_t)cell * NP + i] = 0.0;
    for (int j = 0; j < P; ++j) {
        int sl, cn;
        ref_to_canon(family, modes, j, &sl, &cn);
        if (sl == slot) th_trial[(size_t)cell * NP + cn] = theta[(size_t)cell * P + j];
    }
}

// canonical per-term results -> reference layout params[C x P], logp, status, n_pass
__global__ void store_params_kernel(int C, int Cp, int family, int modes, const double* __restrict__ th_out,
                                    const double* __restrict__ logp_part, const int* __restrict__ status_part,
                                    const int* __restrict__ npass_part, double* __restrict__ params,
                                    double* __restrict__ logp, int32_t* __restrict__ status,
                                    int32_t* __restrict__ n_pass) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= C) return;
    const int P = family_num_params(family, modes);
    const int nt = family_terms_count(family);
    for (int j = 0; j < P; ++j) {
        int sl, cn;
        ref_to_canon(family, modes, j, &sl, &cn);
        params[(size_t)cell * P + j] = th_out[((size_t)sl * NP + cn) * Cp + cell];
    }
    double lp = 0.0;
    int st = 0, np = 0;
    for (int s = 0; s < nt; ++s) {
        lp += logp_part[(size_t)s * Cp + cell];
        int ss = status_part[(size_t)s * Cp + cell];
        st = ss > st ? ss : st;
        np += npass_part[(size_t)s * Cp + cell];
    }
    if (st == ATTRICI_STATUS_NO_DATA || st == ATTRICI_STATUS_NONFINITE) {
        // a cell without data in any of its terms has no model at all (detrend.py:315-317 skips it)
        for (int j = 0; j < P; ++j) params[(size_t)cell * P + j] = NAN;
        lp = NAN;
    }
    if (logp) logp[cell] = lp;
    if (status) status[cell] = st;
    if (n_pass) n_pass[cell] = np;
}

// canonical nll / grad [Cp][NP] / packed fisher [Cp][NTRI] of one term slot -> reference layout
// (accumulating over slots)
__global__ void scatter_kat_kernel(int C, int family, int modes, int slot, const double* __restrict__ nll_c,
                                   const double* __restrict__ grad_c, const double* __restrict__ fisher_c,
                                   double* __restrict__ nll, double* __restrict__ grad, double* __restrict__ fisher,
                                   int accumulate) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= C) return;
    const int P = family_num_params(family, modes);
    if (nll) nll[cell] = (accumulate ? nll[cell] : 0.0) + nll_c[cell];
    for (int j = 0; j < P; ++j) {
        int sl, cn;
        ref_to_canon(family, modes, j, &sl, &cn);
        if (sl != slot) {
            if (!accumulate && fisher)
                for (int k = 0; k < P; ++k) fisher[((size_t)cell * P + j) * P + k] = 0.0;
            continue;
        }
        if (grad) grad[(size_t)cell * P + j] = grad_c[(size_t)cell * NP + cn];
        if (fisher) {
            for (int k = 0; k < P; ++k) {
                int sl2, cn2;
                ref_to_canon(family, modes, k, &sl2, &cn2);
                fisher[((size_t)cell * P + j) * P + k] =
                    (sl2 == slot) ? fisher_c[(size_t)cell * NTRI + tri(cn, cn2)] : 0.0;
            }
        }
    }
}

inline int warp_blocks(int n) { return (n + STEP_WARPS - 1) / STEP_WARPS; }
inline int thread_blocks(int n) { return (n + 127) / 128; }

}  // namespace

StepArgs make_step_args(Workspace& ws, const StepParams& sp) {
    StepArgs a;
    a.C = ws.C; a.Cp = ws.Cp;
    a.term = sp.term; a.modes = sp.modes; a.slot = sp.slot;
    a.max_pass = sp.max_pass; a.tol_pred = sp.tol_pred; a.lambda0 = sp.lambda0; a.zero_init = sp.zero_init;
    a.basis64 = ws.basis64;
    a.st_cnt = ws.st_cnt; a.st_sum = ws.st_sum; a.st_sq = ws.st_sq; a.st_ln = ws.st_ln; a.st_first = ws.st_first;
    a.momd = ws.momd;
    a.th_acc = ws.th_acc; a.th_trial = ws.th_trial; a.th_pass = ws.th_pass; a.g_acc = ws.g_acc;
    a.h_acc = ws.h_acc; a.f_acc = ws.f_acc; a.pred = ws.pred; a.lam = ws.lam; a.rot = ws.rot;
    a.status = ws.status; a.npass = ws.npass;
    a.th_out = ws.th_out; a.logp_part = ws.logp_part; a.status_part = ws.status_part; a.npass_part = ws.npass_part;
    return a;
}

int launch_fit_init(Workspace& ws, const StepParams& sp, int* list_out, int* counter, cudaStream_t s) {
    ATTRICI_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), s));
    fit_init_kernel<<<warp_blocks(ws.C), STEP_THREADS, 0, s>>>(make_step_args(ws, sp), list_out, counter);
    ATTRICI_LAUNCH_CHECK();
    h_table_kernel<<<3, 128, 0, s>>>(make_step_args(ws, sp), reinterpret_cast<HEntry*>(ws.h_tab));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_fit_step(Workspace& ws, const StepParams& sp, const int* list_in, int n_in, int* list_out, int* counter,
                    cudaStream_t s, bool fresh_fisher) {
    ATTRICI_CUDA_CHECK(cudaMemsetAsync(counter, 0, sizeof(int), s));
    if (n_in <= 0) return 0;
    fit_step_kernel<<<warp_blocks(n_in), STEP_THREADS, 0, s>>>(make_step_args(ws, sp), list_in, n_in, list_out, counter,
                                                               fresh_fisher ? 1 : 0,
                                                               reinterpret_cast<const HEntry*>(ws.h_tab));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_fit_final_point(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    fit_final_point_kernel<<<warp_blocks(ws.C), STEP_THREADS, 0, s>>>(make_step_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_fit_store(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    fit_store_kernel<<<warp_blocks(ws.C), STEP_THREADS, 0, s>>>(make_step_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_reduce(Workspace& ws, const PassShape& ps, bool fp64, int mode, cudaStream_t s, const int* status) {
    if (ps.n_active <= 0) return 0;
    // rows reduced: everything, the gradient moments + nll, or only row NACC (the sum of the day terms of the nll)
    const int row0 = (mode == PASS_FULL || mode == PASS_FLAT) ? 0 : (mode == PASS_GRAD ? ACC_GA : NACC);
    dim3 grid((ps.n_active + RED_THREADS - 1) / RED_THREADS, (NACC + 1 - row0 + RED_GROUP - 1) / RED_GROUP);
    if (fp64)
        reduce_partials_kernel<double><<<grid, RED_THREADS, 0, s>>>(ps.list, ps.n_active, ps.Pp, ws.Cp, ps.n_chunks,
                                                                   row0, ws.part_nll,
                                                                   (const double*)ws.part_acc, ws.momd, status);
    else
        reduce_partials_kernel<float><<<grid, RED_THREADS, 0, s>>>(ps.list, ps.n_active, ps.Pp, ws.Cp, ps.n_chunks,
                                                                  row0, ws.part_nll,
                                                                  (const float*)ws.part_acc, ws.momd, status);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_assemble(Workspace& ws, const StepParams& sp, double* nll, double* grad, double* fisher, cudaStream_t s) {
    assemble_kernel<<<warp_blocks(ws.C), STEP_THREADS, 0, s>>>(make_step_args(ws, sp), nll, grad, fisher);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_phase_only(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    phase_kernel<<<warp_blocks(ws.C), STEP_THREADS, 0, s>>>(make_step_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_rotate(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    rotate_kernel<<<warp_blocks(ws.C), STEP_THREADS, 0, s>>>(make_step_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_load_theta(Workspace& ws, int family, int modes, int slot, const double* theta, cudaStream_t s) {
    load_theta_kernel<<<thread_blocks(ws.C), 128, 0, s>>>(ws.C, family, modes, slot, theta, ws.th_trial);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_store_params(Workspace& ws, int family, int modes, double* params, double* logp, int32_t* status,
                        int32_t* n_pass, cudaStream_t s) {
    store_params_kernel<<<thread_blocks(ws.C), 128, 0, s>>>(ws.C, ws.Cp, family, modes, ws.th_out, ws.logp_part,
                                                          ws.status_part, ws.npass_part, params, logp, status, n_pass);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_scatter_kat(Workspace& ws, int family, int modes, int slot, const double* nll_c, const double* grad_c,
                       const double* fisher_c, double* nll, double* grad, double* fisher, int accumulate,
                       cudaStream_t s) {
    scatter_kat_kernel<<<thread_blocks(ws.C), 128, 0, s>>>(ws.C, family, modes, slot, nll_c, grad_c, fisher_c, nll, grad,
                                                         fisher, accumulate);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
