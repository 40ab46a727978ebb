// Shared definitions of the attrici_b200 CUDA library (sm_100a).
//
// Layout vocabulary
//   series   : [T x ld] time-major, cell-minor (a warp reads 32 neighbouring cells of one day
//              = one 128-byte line)
//   basis32  : [T][20] float  {gmt, cos(k w t) k=1..8, sin(k w t) k=1..8, pad x3}  - moments need
//              harmonics up to 2*modes because products of two Fourier columns are Fourier columns
//   basis64  : [T][20] double, same columns                                        - parameter evaluation, FP64 passes
//   per-cell state arrays are [field][C] so that thread = cell accesses coalesce.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/attrici_b200.h"

namespace attrici {

constexpr int MAXM = ATTRICI_MAX_MODES;      // 4 Fourier modes
constexpr int NA = 2 + 4 * MAXM;             // 18 coefficients of a GMT-dependent parameter
constexpr int NB = 1 + 2 * MAXM;             // 9 coefficients of an independent parameter
constexpr int NP = NA + NB;                  // 27 coefficients of one likelihood term (canonical layout)
constexpr int NTRI = NP * (NP + 1) / 2;      // 378 packed lower-triangular entries
constexpr int NH = 2 * MAXM;                 // 8 harmonics carried for the moments
constexpr int NTRIG = 1 + 2 * NH;            // 17 trig columns {1, cos 1..8, sin 1..8}
constexpr int B32_STRIDE = 20;               // floats per day in basis32
constexpr int B64_STRIDE = 20;               // doubles per day in basis64 (same layout as basis32)

// accumulator layout of one likelihood pass (per cell)
constexpr int ACC_AA = 0;                    // 3 powers of gmt x 17 trig
constexpr int ACC_AB = ACC_AA + 3 * NTRIG;   // 2 x 17
constexpr int ACC_BB = ACC_AB + 2 * NTRIG;   // 17
constexpr int ACC_GA = ACC_BB + NTRIG;       // 2 powers x {1, cos 1..4, sin 1..4}
constexpr int ACC_GB = ACC_GA + 2 * NB;      // 9
constexpr int NACC = ACC_GB + NB;            // 129

constexpr int PASS_WARPS = 4;                // warps per CTA of the pass kernel
constexpr int PASS_CELLS = 32 * PASS_WARPS;  // cells per CTA
constexpr int PASS_TILE = 64;                // days per shared-memory basis tile
constexpr int MAX_CHUNKS = 32;

// likelihood term kinds (a Bernoulli-Gamma fit is a Bernoulli term plus a Gamma term that share
// no coefficient: model_scipy.py:433-446)
enum Term : int { T_NORMAL = 0, T_GAMMA = 1, T_BETA = 2, T_WEIBULL = 3, T_BERNOULLI = 4 };

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
    // offsets
    float* basis32;
    double* basis64;
    // prep results
    float* cmin;       // [C]
    float* cmax;       // [C]
    double* psum;      // [2][C] pr: sum x, sum ln x over wet days
    int* pcnt;         // [C]    pr: wet days
    // start values / phase origins of the fit window, per term slot (0: main or bernoulli, 1: gamma of pr)
    double* st_cnt;    // [2][C]
    double* st_sum;    // [2][C]
    double* st_sq;     // [2][C]
    double* st_ln;     // [2][C]  sum of ln y (gamma / weibull start)
    int* st_first;     // [2][C]  first used day index of the fit window (-1 = none)
    // fit state
    double* th_acc;    // [NP][C]
    double* th_trial;  // [NP][C] cell frame
    double* th_pass;   // [NP][C] shared frame (what the pass kernel reads)
    double* g_acc;     // [NP][C]
    double* h_acc;     // [NTRI][C]
    double* h_wrk;     // [NTRI][C]
    double* f_acc;     // [C]
    double* pred;      // [C]
    double* lam;       // [C]
    double* rot;       // [2*NH][C] cos/sin of the per-cell phase offset, harmonics 1..8
    int* status;       // [C]
    int* npass;        // [C]
    int* n_running;    // [1] (+pad)
    double* part_nll;  // [MAX_CHUNKS][C]
    void* part_acc;    // [n_chunks][NACC][C] float or double
    // params staging for two-term fits: canonical per-term coefficients
    double* th_out;    // [2][NP][C]
    double* logp_part; // [2][C]

    static size_t align(size_t x) { return (x + 255) & ~size_t(255); }

    // returns total bytes; if base != nullptr also sets the pointers
    size_t layout(int T_, int C_, void* base_) {
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
        basis32 = (float*)take(sizeof(float) * B32_STRIDE * (size_t)(T + 2 * PASS_TILE));
        basis64 = (double*)take(sizeof(double) * B64_STRIDE * (size_t)(T + 2 * PASS_TILE));
        cmin = (float*)take(4 * c);
        cmax = (float*)take(4 * c);
        psum = (double*)take(8 * 2 * c);
        pcnt = (int*)take(4 * c);
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
        h_wrk = (double*)take(8 * NTRI * c);
        f_acc = (double*)take(8 * c);
        pred = (double*)take(8 * c);
        lam = (double*)take(8 * c);
        rot = (double*)take(8 * 2 * NH * c);
        status = (int*)take(4 * c);
        npass = (int*)take(4 * c);
        n_running = (int*)take(256);
        part_nll = (double*)take(8 * (size_t)MAX_CHUNKS * c);
        part_acc = take(8 * (size_t)ck.n_chunks * NACC * c);
        th_out = (double*)take(8 * 2 * NP * c);
        logp_part = (double*)take(8 * 2 * c);
        bytes = off;
        return off;
    }
};

void set_error(const char* fmt, ...);

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
        cudaError_t _e = cudaGetLastError();                                                 \
        if (_e != cudaSuccess) {                                                             \
            ::attrici::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), \
                                 __FILE__, __LINE__);                                        \
            return (int)_e;                                                                  \
        }                                                                                    \
    } while (0)

// ---- kernel launchers (one per .cu file) -------------------------------------------------
int launch_build_basis(Workspace& ws, const int32_t* days, const double* gmt, cudaStream_t s);

int launch_prep(Workspace& ws, const attrici_variable_t& var, const float* y, float* y_scaled, int64_t ld,
                int i0, int i1, double* scaling, cudaStream_t s);

// one likelihood pass over the fit window at ws.th_pass; fills ws.part_nll / ws.part_acc
int launch_pass(Workspace& ws, int term, bool fp64, bool with_moments, const float* y_scaled, int64_t ld,
                int i0, int i1, cudaStream_t s);

struct StepParams {
    int term;
    int modes;
    int slot;        // which start-value / phase-origin slot of the workspace
    int fp64;
    int max_pass;
    double tol_pred;
    double lambda0;
    int zero_init;
};
int launch_fit_init(Workspace& ws, const StepParams& sp, cudaStream_t s);
int launch_fit_step(Workspace& ws, const StepParams& sp, cudaStream_t s);
// reduce partials of a pass at th_pass into nll/grad/fisher in the cell frame (KAT entry, final logp)
int launch_assemble(Workspace& ws, const StepParams& sp, double* nll, double* grad, double* fisher,
                    int p_stride, cudaStream_t s);
int launch_load_theta(Workspace& ws, const StepParams& sp, const double* theta_canon, cudaStream_t s);

int launch_qmap(Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled,
                int64_t ld, const double* params, const double* scaling, const uint32_t* seeds, double* cfact,
                uint8_t* replaced, double* cfact_scaled, int64_t ld_out, cudaStream_t s);
int launch_eval_parameter(Workspace& ws, const attrici_variable_t& var, const double* params, int which,
                          int counterfactual, double* out, int64_t ld_out, cudaStream_t s);
int launch_mt19937(const uint32_t* seeds, int C, int n, double* out, int64_t ld, cudaStream_t s);

}  // namespace attrici
