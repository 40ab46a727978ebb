// Shared definitions of the attrici_b200 CUDA library (sm_100a).
//
// Layout vocabulary
//   series   : [T x ld] time-major, cell-minor (a warp reads 32 neighbouring cells of one day
//              = one 128-byte line)
//   basis32  : [T][20] float  {gmt, cos(k w t) k=1..8, sin(k w t) k=1..8, tau, pad x2}  - moments need
//              harmonics up to 2*modes because products of two Fourier columns are Fourier columns
//   basis64  : [T][20] double, same columns                                  - parameter evaluation, FP64 passes
//   per-cell state arrays are [field][Cp] so that thread = cell accesses coalesce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/attrici_b200.h"
#include "layout.cuh"

namespace attrici {

struct Chunking {
    int n_chunks;
    int chunk_len;
};

inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

// time chunks per cell block: enough CTAs for ~3 waves of 2 CTAs/SM on 148 SMs
inline Chunking choose_chunks(int T, int C) {
    int n_cb = (C + PASS_CELLS - 1) / PASS_CELLS;
    int want = (148 * 2 * 3 + n_cb - 1) / n_cb;
    if (want < 1) want = 1;
    if (want > MAX_CHUNKS) want = MAX_CHUNKS;
    int len = round_up((T + want - 1) / want, PASS_TILE);
    if (len < PASS_TILE) len = PASS_TILE;
    Chunking c;
    c.chunk_len = len;
    c.n_chunks = (T + len - 1) / len;
    if (c.n_chunks < 1) c.n_chunks = 1;
    return c;
}

// carve-up of the caller's scratch buffer
struct Workspace {
    int T, C, Cp;
    Chunking ck;
    char* base;
    size_t bytes;
    float* basis32;
    double* basis64;
    float* zmat;       // [T + pad][ZM_STRIDE] TF32 hi / lo parts of the moment matrix (tensor-core pass)
    // prep results
    float* cmin;       // [Cp]
    float* cscale;     // [Cp]
    // start values / phase origins of the fit window, per term slot (0: main or bernoulli, 1: gamma of pr)
    double* st_cnt;    // [2][Cp]
    double* st_sum;    // [2][Cp]
    double* st_sq;     // [2][Cp]
    double* st_ln;     // [2][Cp]  sum of ln y (gamma / weibull start)
    int* st_first;     // [2][Cp]  first used day index of the fit window (-1 = none)
    // fit state ("cell-major" = [Cp][n], read by one warp per cell)
    double* th_acc;    // [Cp][NP] accepted point, cell frame
    double* th_trial;  // [Cp][NP] trial point, cell frame
    double* th_pass;   // [NP][Cp] trial point, shared frame (what the pass kernel reads, thread = cell)
    double* g_acc;     // [Cp][NP]
    double* h_acc;     // [Cp][NTRI] (also the canonical Fisher output of the known-answer entry)
    double* f_acc;     // [Cp]
    double* pred;      // [Cp]
    double* lam;       // [Cp]
    double* rot;       // [Cp][2*NH] cos/sin of the per-cell phase offset, harmonics 1..8
    int* status;       // [Cp]
    int* npass;        // [Cp]
    int* list[2];      // [Cp] each: cells still iterating (ping-pong)
    int* counters;     // [2] lengths of the two lists (+pad)
    double* momd;      // [NACC + 1][Cp] chunk-reduced moments, row NACC = sum of the day terms of the nll
    double* part_nll;  // [chunk][Pp]
    size_t part_nll_cap;   // doubles
    void* part_acc;    // [chunk][NACC][Pp] float or double
    size_t part_acc_cap;   // 8-byte units
    // results of the (up to) two likelihood terms, canonical layout, cell frame
    double* th_out;    // [2][NP][Cp]
    double* logp_part; // [2][Cp]
    int* status_part;  // [2][Cp]
    int* npass_part;   // [2][Cp]
    double* uniforms;  // [T][Cp] MT19937 doubles (pr only, else nullptr)
    // phase-folded path (fold.cu): on a gap-free daily axis every Fourier column has period NPHASE days
    int* fold_bad;     // [0] number of days that break days[t] = days[0] + t (0: the folded kernels apply); [1], [2] fit window of the per-phase sums
    double* pb64;      // [NPHASE + pad][B_STRIDE] phase rows {n, cos 1..8, sin 1..8, sum gmt, sum gmt^2, 0} of the fit window
    int* miss;         // [Cp] 1: the cell has missing days inside the fit window
    double* ystat;     // [NPHASE][3][Cp] per-phase sums {y, y^2, y gmt} over the cell's valid window days (Normal only)
    double* gstat;     // [NPHASE][3][Cp] per-phase {n, sum gmt, sum gmt^2} of cells with missing days (Normal only)
    float* psums;      // [NPHASE][11][Cp] per-phase sums of the day terms (two-stage pass of the other families)
    // fused fit (fit_fused.cu): cell-major per-phase sums, per-cell affine map of the sums, pair tables
    double* ysumT;     // [Cp][3][FUSED_SP] {u, u^2, u gmt} per phase, u = y_scaled or y - pivot (Normal only)
    double* gsumT;     // [Cp][3][FUSED_SP] {n, sum gmt, sum gmt^2} per phase of cells with missing days (Normal only)
    uint8_t* cntT;     // [Cp][FUSED_SP] valid window days per phase (read for cells with missing days)
    float* pivot;      // [Cp] u = y - pivot (one-pass prep)
    double* aff;       // [Cp][2] {m, s}: y_scaled = (u - m) / s
    double* ftab_d;    // [FUSED_ND][FUSED_TP]
    float* ftab_f;     // [16][FUSED_FP]
    void* h_tab;       // [NTRI] moment indices of the information-matrix entries of the term being fitted (step.cu)

    static size_t align(size_t x) { return (x + 255) & ~size_t(255); }

    // returns total bytes; if base_ != nullptr also sets the pointers
    size_t layout(int T_, int C_, int family, void* base_) {
        T = T_;
        C = C_;
        Cp = round_up(C_, 32);
        ck = choose_chunks(T_, C_);
        base = (char*)base_;
        size_t off = 0;
        auto take = [&](size_t n) {
            size_t o = off;
            off = align(off + n);
            return base ? (void*)(base + o) : (void*)nullptr;
        };
        size_t c = (size_t)Cp;
        basis32 = (float*)take(sizeof(float) * B_STRIDE * (size_t)(T + 2 * PASS_TILE));
        basis64 = (double*)take(sizeof(double) * B_STRIDE * (size_t)(T + 2 * PASS_TILE));
        zmat = (float*)take(sizeof(float) * ZM_STRIDE * (size_t)(T + 2 * PASS_TILE));
        cmin = (float*)take(4 * c);
        cscale = (float*)take(4 * c);
        st_cnt = (double*)take(8 * 2 * c);
        st_sum = (double*)take(8 * 2 * c);
        st_sq = (double*)take(8 * 2 * c);
        st_ln = (double*)take(8 * 2 * c);
        st_first = (int*)take(4 * 2 * c);
        th_acc = (double*)take(8 * NP * c);
        th_trial = (double*)take(8 * NP * c);
        th_pass = (double*)take(8 * NP * c);
        g_acc = (double*)take(8 * NP * c);
        h_acc = (double*)take(8 * NTRI * c);
        f_acc = (double*)take(8 * c);
        pred = (double*)take(8 * c);
        lam = (double*)take(8 * c);
        rot = (double*)take(8 * 2 * NH * c);
        status = (int*)take(4 * c);
        npass = (int*)take(4 * c);
        list[0] = (int*)take(4 * c);
        list[1] = (int*)take(4 * c);
        counters = (int*)take(256);
        momd = (double*)take(8 * (size_t)(NACC + 1) * c);
        part_nll_cap = (size_t)MAX_CHUNKS * c;
        part_nll = (double*)take(8 * part_nll_cap);
        part_acc_cap = (size_t)ck.n_chunks * NACC * c;
        part_acc = take(8 * part_acc_cap);
        th_out = (double*)take(8 * 2 * NP * c);
        logp_part = (double*)take(8 * 2 * c);
        status_part = (int*)take(4 * 2 * c);
        npass_part = (int*)take(4 * 2 * c);
        uniforms = (family == ATTRICI_BERNOULLI_GAMMA) ? (double*)take(8 * (size_t)T * c) : nullptr;
        fold_bad = (int*)take(256);
        pb64 = (double*)take(sizeof(double) * B_STRIDE * (size_t)(NPHASE + 2 * PASS_TILE));
        miss = (int*)take(4 * c);
        ystat = (family == ATTRICI_NORMAL) ? (double*)take(8 * 3 * (size_t)NPHASE * c) : nullptr;
        gstat = (family == ATTRICI_NORMAL) ? (double*)take(8 * 3 * (size_t)NPHASE * c) : nullptr;
        psums = (family != ATTRICI_NORMAL) ? (float*)take(4 * 11 * (size_t)NPHASE * c) : nullptr;
        ysumT = (family == ATTRICI_NORMAL) ? (double*)take(8 * 3 * (size_t)FUSED_SP * c) : nullptr;
        gsumT = (family == ATTRICI_NORMAL) ? (double*)take(8 * 3 * (size_t)FUSED_SP * c) : nullptr;
        cntT = (family == ATTRICI_NORMAL) ? (uint8_t*)take((size_t)FUSED_SP * c) : nullptr;
        pivot = (float*)take(4 * c);
        aff = (double*)take(8 * 2 * c);
        ftab_d = (double*)take(8 * (size_t)FUSED_ND * FUSED_TP);
        ftab_f = (float*)take(4 * 16 * (size_t)FUSED_FP);
        h_tab = take(16 * (size_t)NTRI);
        bytes = off;
        return off;
    }
};

// shape of one likelihood pass: which cells (list of still-iterating cells or all), how the fit window is
// cut into time chunks, and the stride of the partial buffers (indexed by list position)
struct PassShape {
    const int* list;   // nullptr: cell = position
    int n_active;
    int Pp;            // stride of the partials = round_up(n_active, 32)
    int n_chunks, chunk_len;
};
PassShape plan_pass(const Workspace& ws, const int* list, int n_active, int i0, int i1, bool fp64);

void set_error(const char* fmt, ...);
void count_launch();

#define ATTRICI_CUDA_CHECK(expr)                                                             \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess) {                                                             \
            ::attrici::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),     \
                                 __FILE__, __LINE__);                                        \
            return (int)_e;                                                                  \
        }                                                                                    \
    } while (0)

#define ATTRICI_LAUNCH_CHECK()                                                               \
    do {                                                                                     \
        ::attrici::count_launch();                                                           \
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            ::attrici::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                        \
            return (int)_e;                                                                  \
        }                                                                                    \
    } while (0)

#define ATTRICI_TRY(expr)             \
    do {                              \
        int _rc = (expr);             \
        if (_rc != 0) return _rc;     \
    } while (0)

// ---- kernel launchers (one per .cu file) -------------------------------------------------
int launch_build_basis(Workspace& ws, const int32_t* days, const double* gmt, cudaStream_t s);

int launch_prep(Workspace& ws, const attrici_variable_t& var, const float* y, float* y_scaled, int64_t ld,
                int i0, int i1, double* scaling, int flags, cudaStream_t s);

// one likelihood pass over the fit window [i0, i1) at ws.th_pass for the cells of `ps`; fills ws.part_nll /
// ws.part_acc.  status != nullptr (only without a list): only cells whose status is RUNNING are evaluated.
int launch_pass(Workspace& ws, const PassShape& ps, int term, bool fp64, bool with_moments, const float* y_scaled,
                int64_t ld, int i0, int i1, const int* status, cudaStream_t s);
// tensor-core variant for the Normal family, FP32 path with moments (pass_mma.cu)
int launch_build_zmat(Workspace& ws, cudaStream_t s);
int launch_pass_mma(Workspace& ws, const PassShape& ps, const float* y_scaled, int64_t ld, int i0, int i1,
                    const int* status, cudaStream_t s);
// phase-folded pass of the Normal family over the per-phase sums of prep (fold.cu); `ps` is planned over
// [0, fold_phases(ws)); partials are FP64
// what a pass produces: the nll only, nll + gradient moments, or nll + gradient + Fisher moments
// PASS_FLAT = PASS_FULL at a point whose only non-zero coefficients are the two intercepts (folded Normal pass)
enum { PASS_NLL = 0, PASS_GRAD = 1, PASS_FULL = 2, PASS_FLAT = 3 };
int fold_phases(const Workspace& ws);
// items a folded Normal pass of this mode is planned over (phases, or phase pairs: fold.cu)
int fold_items(const Workspace& ws, int mode, bool exact);
int launch_fold_pass(Workspace& ws, const PassShape& ps, int mode, bool exact, const int* status, cudaStream_t s);
// phase-ordered pass of the Gamma / Weibull / Beta / Bernoulli terms over the scaled series: FP32 with all moments,
// or (nll_only) the FP64 nll of the returned point
int launch_fold_pass_generic(Workspace& ws, const PassShape& ps, int term, bool nll_only, const float* y_scaled, int64_t ld,
                             int i0, int i1, const int* status, cudaStream_t s);
// cross-chunk reduction of the partials of the last pass into ws.momd
// status != nullptr (only without a list): cells whose status is not RUNNING are skipped
int launch_reduce(Workspace& ws, const PassShape& ps, bool fp64, int mode, cudaStream_t s,
                  const int* status = nullptr);

// fused fit of the folded Normal term (fit_fused.cu): pair tables of the fit window, the cell-major copy of sums
// gathered in the phase-major layout, and the whole iteration of a term in one persistent kernel
int launch_fused_tables(Workspace& ws, cudaStream_t s);
int launch_fused_from_phase_major(Workspace& ws, cudaStream_t s);
// one pass over the series: min / max and the per-phase sums of y - pivot, cell-major (prep_onepass.cu); Normal family
// on a gap-free daily axis when the caller does not ask for y_scaled.  flags: ATTRICI_PREP_* of the C ABI
bool prep_onepass_applies(const Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled, int64_t ld,
                          int flags);
int launch_prep_onepass(Workspace& ws, const attrici_variable_t& var, const float* y, int64_t ld, int i0, int i1,
                        double* scaling, cudaStream_t s);

struct StepParams {
    int term;
    int modes;
    int slot;        // which start-value / phase-origin slot of the workspace
    int max_pass;
    double tol_pred;
    double lambda0;
    int zero_init;
};
// start values and phase origins; list_out / counter receive the cells that iterate
int launch_fit_init(Workspace& ws, const StepParams& sp, int* list_out, int* counter, cudaStream_t s);
// accept / reject / propose for the cells of list_in; list_out / counter receive those that go on
// fresh_fisher = false: the pass before only produced gradient moments, the information matrix of the accepted
// point is reused
int launch_fit_step(Workspace& ws, const StepParams& sp, const int* list_in, int n_in, int* list_out, int* counter,
                    cudaStream_t s, bool fresh_fisher = true);
int launch_fit_fused(Workspace& ws, const StepParams& sp, int fisher_passes, int chord_passes, cudaStream_t s);
// th_trial := th_acc, rotated into th_pass (before the final FP64 evaluation)
int launch_fit_final_point(Workspace& ws, const StepParams& sp, cudaStream_t s);
// after an FP64 pass without moments at th_acc (+ reduce): store coefficients / logp / status of this term slot
int launch_fit_store(Workspace& ws, const StepParams& sp, cudaStream_t s);
// after a pass at th_trial (+ reduce): nll [Cp], grad [Cp][NP], fisher [Cp][NTRI] in the cell frame (KAT entry)
int launch_assemble(Workspace& ws, const StepParams& sp, double* nll, double* grad, double* fisher, cudaStream_t s);
// phase offsets only (KAT entry: no fit_init)
int launch_phase_only(Workspace& ws, const StepParams& sp, cudaStream_t s);
int launch_rotate(Workspace& ws, const StepParams& sp, cudaStream_t s);

// reference theta layout <-> canonical per-term layout
int launch_load_theta(Workspace& ws, int family, int modes, int slot, const double* theta, cudaStream_t s);
int launch_store_params(Workspace& ws, int family, int modes, double* params, double* logp, int32_t* status,
                        int32_t* n_pass, cudaStream_t s);
// scatter canonical grad / packed fisher of one term slot into the reference layout
int launch_scatter_kat(Workspace& ws, int family, int modes, int slot, const double* nll_c, const double* grad_c,
                       const double* fisher_c, double* nll, double* grad, double* fisher, int accumulate,
                       cudaStream_t s);

int launch_qmap(Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled,
                int64_t ld, const double* params, const double* scaling, const uint32_t* seeds, double* cfact,
                uint8_t* replaced, double* cfact_scaled, int64_t ld_out, cudaStream_t s, float* cfact32 = nullptr);
// does the lean Normal kernel (the one that can store a float32 counterfactual) serve this request?
bool qmap_fold_lean_applies(const attrici_variable_t& var, const double* cfact_scaled, const uint8_t* replaced);
// Normal family in phase order (qmap_fold.cu); runs only if ws.fold_bad[0] == 0 (checked on the device)
int launch_qmap_fold(Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled, int64_t ld,
                     const double* params, const double* scaling, double* cfact, uint8_t* replaced,
                     double* cfact_scaled, int64_t ld_out, cudaStream_t s, float* cfact32 = nullptr);
// Gamma / Weibull families in phase order (scale shortcut), same device-side dispatch; needs y_scaled
int launch_qmap_fold_pr(Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled, int64_t ld,
                        const double* params, const double* scaling, const double* uniforms, double* cfact,
                        uint8_t* replaced, double* cfact_scaled, int64_t ld_out, cudaStream_t s);
int launch_qmap_fold_scale(Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled, int64_t ld,
                           const double* params, const double* scaling, double* cfact, uint8_t* replaced,
                           double* cfact_scaled, int64_t ld_out, cudaStream_t s);
int launch_eval_parameter(Workspace& ws, const attrici_variable_t& var, const double* params, int which,
                          int counterfactual, double* out, int64_t ld_out, cudaStream_t s);
// rolling-window model of the Normal family (window.cu)
size_t window_workspace_bytes(int C);
int launch_window_fit(void* workspace, const float* ys, int64_t ld, int T, int C, int i0, int i1, const double* gmt,
                      const int* run_start, const int* run_first, const int* run_len, int n_runs, int window_size,
                      double* params, double* logp, int32_t* status, int32_t* n_iter, cudaStream_t s);
int launch_window_qmap(const attrici_variable_t& var, const float* y, const float* ys, int64_t ld, int T, int C,
                       const double* gmt, const int16_t* doy, const double* params, const double* scaling, double* cfact,
                       uint8_t* replaced, double* cfact_scaled, int64_t ld_out, cudaStream_t s);
int launch_mt19937(const uint32_t* seeds, int C, int n, double* out, int64_t ld, cudaStream_t s);

}  // namespace attrici
