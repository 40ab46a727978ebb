// Kernels around the per-cell optimiser mathematics of step_impl.cuh (one thread = one cell) and the
// conversions between the reference's theta layout and the canonical per-term layout.
#include "common.cuh"
#include "step_impl.cuh"

namespace attrici {

namespace {

constexpr int STEP_THREADS = 64;

__global__ void fit_init_kernel(StepArgs a) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.C) return;
    if (cell_init(a, cell)) atomicAdd(a.n_running, 1);
}

__global__ void fit_step_kernel(StepArgs a) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.C) return;
    if (cell_step(a, cell)) atomicAdd(a.n_running, 1);
}

__global__ void fit_final_point_kernel(StepArgs a) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.C) return;
    cell_final_point(a, cell);
}

__global__ void fit_store_kernel(StepArgs a) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.C) return;
    cell_store(a, cell);
}

// known-answer evaluation: objective, gradient, Fisher information at th_trial
__global__ void assemble_kernel(StepArgs a, double* nll, double* grad /*[NP][Cp] canonical*/,
                                double* fisher /*[NTRI][Cp] canonical packed, nullable*/) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.C) return;
    const size_t cp = a.Cp;
    double th[NP], g[NP];
    for (int i = 0; i < NP; ++i) th[i] = a.th_trial[(size_t)i * cp + cell];
    double f = assemble(a, cell, th, g, fisher ? fisher + cell : nullptr);
    nll[cell] = f;
    if (grad) for (int i = 0; i < NP; ++i) grad[(size_t)i * cp + cell] = g[i];
}

__global__ void phase_kernel(StepArgs a) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.C) return;
    cell_phase(a, cell);
}

// th_trial (cell frame) -> th_pass (shared frame)
__global__ void rotate_kernel(StepArgs a) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= a.C) return;
    double th[NP];
    for (int i = 0; i < NP; ++i) th[i] = a.th_trial[(size_t)i * a.Cp + cell];
    rotate_to_shared(a, cell, th);
}

// reference layout theta[C x P] -> th_trial of one term slot (inactive entries zero)
__global__ void load_theta_kernel(int C, int Cp, int family, int modes, int slot, const double* __restrict__ theta,
                                  double* __restrict__ th_trial) {
    int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= C) return;
    const int P = family_num_params(family, modes);
    for (int i = 0; i < NP; ++i) th_trial[(size_t)i * Cp + cell] = 0.0;
    for (int j = 0; j < P; ++j) {
        int sl, cn;
        ref_to_canon(family, modes, j, &sl, &cn);
        if (sl == slot) th_trial[(size_t)cn * Cp + cell] = theta[(size_t)cell * P + j];
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

// canonical nll / grad / packed fisher of one term slot -> reference layout (accumulating over slots)
__global__ void scatter_kat_kernel(int C, int Cp, int family, int modes, int slot, const double* __restrict__ nll_c,
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
        if (grad) grad[(size_t)cell * P + j] = grad_c[(size_t)cn * Cp + cell];
        if (fisher) {
            for (int k = 0; k < P; ++k) {
                int sl2, cn2;
                ref_to_canon(family, modes, k, &sl2, &cn2);
                fisher[((size_t)cell * P + j) * P + k] =
                    (sl2 == slot) ? fisher_c[(size_t)tri(cn, cn2) * Cp + cell] : 0.0;
            }
        }
    }
}

StepArgs make_args(Workspace& ws, const StepParams& sp) {
    StepArgs a;
    a.C = ws.C; a.Cp = ws.Cp;
    a.n_chunks = ws.used_chunks;
    a.fp64 = sp.fp64;
    a.term = sp.term; a.modes = sp.modes; a.slot = sp.slot;
    a.max_pass = sp.max_pass; a.tol_pred = sp.tol_pred; a.lambda0 = sp.lambda0; a.zero_init = sp.zero_init;
    a.basis64 = ws.basis64;
    a.st_cnt = ws.st_cnt; a.st_sum = ws.st_sum; a.st_sq = ws.st_sq; a.st_ln = ws.st_ln; a.st_first = ws.st_first;
    a.th_acc = ws.th_acc; a.th_trial = ws.th_trial; a.th_pass = ws.th_pass; a.g_acc = ws.g_acc;
    a.h_acc = ws.h_acc; a.h_wrk = ws.h_wrk; a.f_acc = ws.f_acc; a.pred = ws.pred; a.lam = ws.lam; a.rot = ws.rot;
    a.status = ws.status; a.npass = ws.npass; a.n_running = ws.n_running;
    a.part_nll = ws.part_nll; a.part_acc = ws.part_acc;
    a.th_out = ws.th_out; a.logp_part = ws.logp_part; a.status_part = ws.status_part; a.npass_part = ws.npass_part;
    return a;
}

inline int blocks(int C) { return (C + STEP_THREADS - 1) / STEP_THREADS; }

}  // namespace

int launch_fit_init(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    ATTRICI_CUDA_CHECK(cudaMemsetAsync(ws.n_running, 0, sizeof(int), s));
    fit_init_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(make_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_fit_step(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    ATTRICI_CUDA_CHECK(cudaMemsetAsync(ws.n_running, 0, sizeof(int), s));
    fit_step_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(make_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_fit_final_point(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    fit_final_point_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(make_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_fit_store(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    fit_store_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(make_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_assemble(Workspace& ws, const StepParams& sp, double* nll, double* grad, double* fisher, cudaStream_t s) {
    assemble_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(make_args(ws, sp), nll, grad, fisher);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_phase_only(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    phase_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(make_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_rotate(Workspace& ws, const StepParams& sp, cudaStream_t s) {
    rotate_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(make_args(ws, sp));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_load_theta(Workspace& ws, int family, int modes, int slot, const double* theta, cudaStream_t s) {
    load_theta_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(ws.C, ws.Cp, family, modes, slot, theta, ws.th_trial);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_store_params(Workspace& ws, int family, int modes, double* params, double* logp, int32_t* status,
                        int32_t* n_pass, cudaStream_t s) {
    store_params_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(ws.C, ws.Cp, family, modes, ws.th_out, ws.logp_part,
                                                            ws.status_part, ws.npass_part, params, logp, status,
                                                            n_pass);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_scatter_kat(Workspace& ws, int family, int modes, int slot, const double* nll_c, const double* grad_c,
                       const double* fisher_c, double* nll, double* grad, double* fisher, int accumulate,
                       cudaStream_t s) {
    scatter_kat_kernel<<<blocks(ws.C), STEP_THREADS, 0, s>>>(ws.C, ws.Cp, family, modes, slot, nll_c, grad_c, fisher_c,
                                                           nll, grad, fisher, accumulate);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
