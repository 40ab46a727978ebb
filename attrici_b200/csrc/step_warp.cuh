// One warp = one cell: the cooperative group the per-cell optimiser of step_impl.cuh runs on (device side), with the
// register-resident Cholesky solve.  Shared by step.cu (one launch per step) and fit_fused.cu (the whole iteration
// of a cell inside one persistent kernel).
#pragma once
#include "common.cuh"
#include "step_impl.cuh"

namespace attrici {

// pointers of the fit state in the workspace, as the per-cell optimiser takes them (step.cu)
StepArgs make_step_args(Workspace& ws, const StepParams& sp);

// BLOCK: the information matrix is block-diagonal, diag(A [NA x NA], B [NB x NB]) - terms without a cross block
// (Normal, Gamma, Bernoulli): both blocks are factorised at the same time by different lanes
template <bool BLOCK> struct WarpCtxT {
    static constexpr int W = 32;
    static constexpr bool kRegisterSolve = true;
    __device__ __forceinline__ int lane() const { return threadIdx.x & 31; }
    template <class Scr> __device__ bool solve(Scr& S, double lam) const {
        if constexpr (BLOCK) return solve_block(S, lam);
        else return solve_full(S, lam);
    }
    template <class Scr> __device__ bool solve_full(Scr& S, double lam) const;
    template <class Scr> __device__ bool solve_block(Scr& S, double lam) const;
    __device__ __forceinline__ void sync() const { __syncwarp(); }
    __device__ __forceinline__ double sum(double v) const {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        return v;
    }
    __device__ __forceinline__ bool all(bool p) const { return __all_sync(0xffffffffu, p); }
};
using WarpCtx = WarpCtxT<false>;

// (H + lam D) delta = -g with lane = row: the lane's row of H (then of the Cholesky factor L) lives in registers,
// the factorisation is right-looking: column j is scaled and published through shared memory (S.y), then every
// later column k takes its rank-1 update with L[k][j] read back as a warp-wide broadcast.  The forward
// substitution is column-oriented in registers; for the backward one the rows of L are parked in S.L so that
// lane k can read L[i][k] while delta[i] is broadcast.  Same mathematics as lm_solve() (step_impl.cuh) in about a
// quarter of the instructions; returns false if a pivot is not positive.  Every loop is fully unrolled so that
// the row stays in registers.
template <bool BLOCK> template <class Scr> __device__ inline bool WarpCtxT<BLOCK>::solve_full(Scr& S, double lam) const {
    constexpr unsigned FULL = 0xffffffffu;
    const int l = lane();
    const bool row = l < NP;
    double a[NP];
#pragma unroll
    for (int k = 0; k < NP; ++k) {
        double v = (row && k <= l) ? S.H[l * (l + 1) / 2 + k] : 0.0;
        if (k == l) { S.hd[l] = v; v += lam * fmax(fabs(v), 1e-12); }
        a[k] = v;
    }
    __syncwarp();   // every row of H is in registers before rows of L (which may share its storage) are written
    double inv_diag = 1.0;   // 1 / L[l][l]
    bool ok = true;
    double* col = S.y;       // column j of L, indexed by row (free until the substitutions)
#pragma unroll
    for (int j = 0; j < NP; ++j) {
        const double d = __shfl_sync(FULL, a[j], j);
        if (!(d > 0.0) || !isfinite(d)) { ok = false; break; }   // uniform: d is the same in every lane
        const double inv = rsqrt(d), ljj = d * inv;   // 1 ulp-level: the factor only shapes a damped step
        if (l == j) { a[j] = ljj; inv_diag = inv; }
        else a[j] *= inv;                                  // L[l][j] for l > j (rows l < j hold zeros)
        if (j + 1 < NP) {
            if (row) col[l] = a[j];
            __syncwarp();
#pragma unroll
            for (int k = j + 1; k < NP; ++k) a[k] = fma(-a[j], col[k], a[k]);   // only entries k <= l are ever read
            __syncwarp();
        }
    }
    if (!ok) return false;
    // rows of L for the backward substitution
    if (row) {
#pragma unroll
        for (int k = 0; k < NP; ++k)
            if (k <= l) S.L()[l * (l + 1) / 2 + k] = a[k];
    }
    // L y = -g, column-oriented: once y[i] is known every later row subtracts L[.][i] y[i]
    double r = row ? -S.g[l] : 0.0, y = 0.0;
#pragma unroll
    for (int i = 0; i < NP; ++i) {
        const double yi = __shfl_sync(FULL, r * inv_diag, i);
        if (l == i) y = yi;
        if (l > i) r = fma(-a[i], yi, r);
    }
    __syncwarp();
    // L^T delta = y, column-oriented from the last row: delta[i] = r[i] / L[i][i], then r[k] -= L[i][k] delta[i], k < i
    double rb = y, delta = 0.0;
#pragma unroll
    for (int i = NP - 1; i >= 0; --i) {
        const double di = __shfl_sync(FULL, rb * inv_diag, i);
        if (l == i) delta = di;
        if (l < i) rb = fma(-S.L()[i * (i + 1) / 2 + l], di, rb);
    }
    if (row) S.delta[l] = delta;
    __syncwarp();
    return true;
}

// Block-diagonal variant.  Lanes 0 .. NA-1 own the rows of block A, lanes NA .. NP-1 the rows of block B; a lane
// keeps its row of its own block (at most NA entries) in registers and both factorisations advance together: NA
// pivot steps with rows of <= NA entries instead of NP steps with rows of <= NP (153 + 36 instead of 351 rank-1
// updates per lane set, a third of the code).  Same mathematics as solve_full on a matrix whose cross block is zero.
template <bool BLOCK> template <class Scr> __device__ inline bool WarpCtxT<BLOCK>::solve_block(Scr& S, double lam) const {
    constexpr unsigned FULL = 0xffffffffu;
    const int l = lane();
    const bool row = l < NP;
    const int base = l < NA ? 0 : NA;     // first row / column of the lane's block
    const int nb = l < NA ? NA : NB;      // size of the lane's block
    const int r = l - base;               // row within the block
    double a[NA];
#pragma unroll
    for (int k = 0; k < NA; ++k) {
        double v = (row && k <= r) ? S.H[l * (l + 1) / 2 + base + k] : 0.0;
        if (row && k == r) { S.hd[l] = v; v += lam * fmax(fabs(v), 1e-12); }
        a[k] = v;
    }
    __syncwarp();   // every row of H is in registers before rows of L (which may share its storage) are written
    double inv_diag = 1.0;   // 1 / L[l][l]
    bool ok = true;
    double* col = S.y;       // column j of both blocks, indexed by row (free until the substitutions)
#pragma unroll
    for (int j = 0; j < NA; ++j) {
        const bool in = row && j < nb;
        const double d = __shfl_sync(FULL, a[j], in ? base + j : l);
        const bool good = !in || ((d > 0.0) && isfinite(d));
        if (!__all_sync(FULL, good)) { ok = false; break; }
        const double inv = rsqrt(in ? d : 1.0), ljj = d * inv;
        if (in) {
            if (r == j) { a[j] = ljj; inv_diag = inv; }
            else a[j] *= inv;               // L[r][j] for r > j (rows r < j hold zeros)
        }
        if (j + 1 < NA) {
            if (row) col[l] = a[j];
            __syncwarp();
#pragma unroll
            for (int k = j + 1; k < NA; ++k)
                a[k] = fma(-a[j], (k < nb) ? col[base + k] : 0.0, a[k]);   // only entries k <= r are ever read
            __syncwarp();
        }
    }
    if (!ok) return false;
    double* Ls = S.L();
    if (row) {
#pragma unroll
        for (int k = 0; k < NA; ++k)
            if (k <= r) Ls[l * (l + 1) / 2 + base + k] = a[k];
    }
    // L y = -g per block, column-oriented
    double rr = row ? -S.g[l] : 0.0, y = 0.0;
#pragma unroll
    for (int i = 0; i < NA; ++i) {
        const bool in = row && i < nb;
        const double yi = __shfl_sync(FULL, rr * inv_diag, in ? base + i : l);
        if (in) {
            if (r == i) y = yi;
            if (r > i) rr = fma(-a[i], yi, rr);
        }
    }
    __syncwarp();
    // L^T delta = y per block, from the last row
    double rb = y, delta = 0.0;
#pragma unroll
    for (int i = NA - 1; i >= 0; --i) {
        const bool in = row && i < nb;
        const double di = __shfl_sync(FULL, rb * inv_diag, in ? base + i : l);
        if (in) {
            if (r == i) delta = di;
            if (r < i) rb = fma(-Ls[(base + i) * (base + i + 1) / 2 + l], di, rb);
        }
    }
    if (row) S.delta[l] = delta;
    __syncwarp();
    return true;
}

}  // namespace attrici
