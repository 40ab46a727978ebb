// Per-cell optimiser mathematics: Fisher-scoring Levenberg-Marquardt with per-cell convergence.
// Replaces scipy.optimize.minimize(method="L-BFGS-B") of ModelScipy.fit
// (attrici/estimation/model_scipy.py:519-523; finite differences, ~900-2400 objective calls per cell)
// by ~5-15 likelihood passes per cell.
//
// Execution model: one cooperative group of W lanes per cell (a warp on the device, W = 1 in the CPU
// unit tests), the cell's 27 x 27 system in shared memory, Cholesky / substitutions parallel over rows.
//
// Frames: the reference builds each cell's Fourier columns with phase origin = first day the
// cell's likelihood term uses (attrici/util.py:102 via model_scipy.py:189, 289 and the wet-day
// selection :424-431), while all cells here share ONE basis with origin = first day of the series.
// Coefficients are therefore optimised in the cell's own frame (where the priors of
// model_scipy.py:142-171, 256-270 apply) and rotated harmonic by harmonic into the shared frame
// for the likelihood pass; moments come back through the transposed rotation.
//
// Everything here is __host__ __device__ so that tests/hostmath can run the identical code for one
// cell on the CPU against the oracle (unit test of the mathematics; the product never does).
#pragma once
#include "layout.cuh"

namespace attrici {

// Layouts: "cell-major" arrays are [Cp][n] (a group's lanes read consecutive entries of one cell);
// th_pass and momd are [n][Cp] (the pass / reduce kernels run thread = cell).
struct StepArgs {
    int C, Cp;
    int term, modes, slot;
    int max_pass;
    double tol_pred, lambda0;
    int zero_init;
    const double* basis64;
    const double* st_cnt; const double* st_sum; const double* st_sq; const double* st_ln; const int* st_first;
    const double* momd;      // [NACC + 1][Cp] chunk-reduced moments; row NACC = sum of the day terms of the nll
    double* th_acc;          // [Cp][NP]
    double* th_trial;        // [Cp][NP]
    double* th_pass;         // [NP][Cp]
    double* g_acc;           // [Cp][NP]
    double* h_acc;           // [Cp][NTRI]
    double* f_acc; double* pred; double* lam;
    double* rot;             // [Cp][2 NH]
    int* status; int* npass;
    double* th_out; double* logp_part; int* status_part; int* npass_part;   // [2][NP][Cp], [2][Cp] ...
};

constexpr int H_ZERO = NACC + 1;   // CellScratch::mom slot that holds 0

// per-cell scratch (shared memory on the device).  SHARE_L: the Cholesky factor overwrites the packed information
// matrix (the register-resident solves of step_warp.cuh read their rows of H before they store rows of L; the
// diagonal of H the predicted decrease needs is kept in hd) - 5.2 KB per cell instead of 8.2 KB.
template <bool SHARE_L> struct CellScratchT {
    double mom[NACC + 2];   // [NACC] = sum of the day terms of the nll, [NACC + 1] = 0 (the "no moment" slot of HEntry)
    double H[NTRI];
    double Lbuf[SHARE_L ? 1 : NTRI];
    double g[NP], th[NP], delta[NP], y[NP];
    double hd[NP];          // diagonal of H as assembled (before damping), filled by the solve
    ATTRICI_HD double* L() { return SHARE_L ? H : Lbuf; }
};
using CellScratch = CellScratchT<false>;

// serial "group" of one lane (CPU unit tests)
struct SerialCtx {
    static constexpr int W = 1;
    static constexpr bool kRegisterSolve = false;
    ATTRICI_HD int lane() const { return 0; }
    ATTRICI_HD void sync() const {}
    ATTRICI_HD double sum(double v) const { return v; }
    ATTRICI_HD bool all(bool p) const { return p; }
};

// prior precision 1/s^2 in the cell frame (model_scipy.py:142-171, 256-270): only the cos
// coefficients of each Fourier block carry a prior
ATTRICI_HD double prior_prec(const Col& c) {
    if (c.k == 0) return c.pw == 0 ? 1.0 : 100.0;
    if (c.sn) return 0.0;
    if (c.pw) return 100.0;
    double s = 2.0 * (c.k - 1) + 1.0;
    return s * s;
}

// sum over prior terms of (-ln s - ln sqrt(2 pi)): part of the reference's logp
ATTRICI_HD double prior_const(int term, int modes) {
    const double lsq = 0.91893853320467274178, ln10 = 2.30258509299404568402;
    double sl = 0.0;
    for (int i = 0; i < modes; ++i) sl += log(2.0 * i + 1.0);
    double a = ln10 * (1 + modes) + sl - (2 + 2 * modes) * lsq;
    double b = sl - (1 + modes) * lsq;
    return term_has_b(term) ? a + b : a;
}

// cell frame -> shared frame: cos/sin pair of harmonic k rotated by the cell's phase offset.
// With x = w (t - t_cell) = X - phi (X the shared-frame phase): a cos kx + b sin kx =
// (a ck - b sk) cos kX + (a sk + b ck) sin kX, ck = cos k phi, sk = sin k phi.
// th_cell: [NP] readable by every lane; writes th_pass[.][cell]
template <class Ctx>
ATTRICI_HD void rotate_to_shared(const StepArgs& a, int cell, const double* th_cell, const Ctx& ctx) {
    const size_t cp = a.Cp;
    const double* rot = a.rot + (size_t)cell * 2 * NH;
    for (int i = ctx.lane(); i < NP; i += Ctx::W) {
        Col c = col_of(i);
        double v = th_cell[i];
        if (c.k > 0 && c.k <= MAXM) {
            // partner coefficient of the same (block, power) pair
            int partner = c.sn ? i - MAXM : i + MAXM;
            double ck = rot[c.k - 1], sk = rot[NH + c.k - 1];
            double other = th_cell[partner];
            v = c.sn ? (other * sk + v * ck) : (v * ck - other * sk);
        }
        a.th_pass[(size_t)i * cp + cell] = v;
    }
}

// chunk-reduced moments of one cell -> S.mom, every trig moment rotated into the cell frame
// (cos kx = ck cos kX + sk sin kX ; sin kx = ck sin kX - sk cos kX).  mom layout = pass.cu's.
// in_scratch: the pass of this group left the moments in S.mom already (fit_fused.cu)
template <class Ctx, class Scr>
ATTRICI_HD void load_moments(const StepArgs& a, int cell, Scr& S, const Ctx& ctx, bool in_scratch = false) {
    const size_t cp = a.Cp;
    if (!in_scratch)
        for (int i = ctx.lane(); i <= NACC; i += Ctx::W) S.mom[i] = a.momd[(size_t)i * cp + cell];
    if (ctx.lane() == 0) S.mom[H_ZERO] = 0.0;
    ctx.sync();
    const double* rot = a.rot + (size_t)cell * 2 * NH;
    // 9 (block, harmonic) families: AA p0..2, AB p0..1, BB carry 8 harmonics; GA p0, GA p1, GB carry 4
    for (int w = ctx.lane(); w < 6 * NH + 3 * MAXM; w += Ctx::W) {
        double* m;
        int k, half;
        if (w < 6 * NH) { m = S.mom + (w / NH) * NTRIG; k = w % NH + 1; half = NH; }
        else { int q = w - 6 * NH; m = S.mom + ACC_GA + (q / MAXM) * NB; k = q % MAXM + 1; half = MAXM; }
        double c = rot[k - 1], s = rot[NH + k - 1];
        double C_ = m[k], S_ = m[half + k];
        m[k] = c * C_ + s * S_;
        m[half + k] = c * S_ - s * C_;
    }
    ctx.sync();
}

// moment of (column i) x (column j) under weight stream `m` (17 trig moments at the right power)
ATTRICI_HD double trig_product(const double* m, const Col& a, const Col& b) {
    const double* Cm = m;           // Cm[0] const, Cm[k] cos k
    const double* Sm = m + NH;      // Sm[k] sin k (Sm[0] aliases cos 8 and is never read)
    int sum = a.k + b.k, dif = a.k - b.k;
    int ad = dif < 0 ? -dif : dif;
    if (!a.sn && !b.sn) return 0.5 * (Cm[ad] + Cm[sum]);
    if (a.sn && b.sn) return 0.5 * (Cm[ad] - Cm[sum]);
    // sin(x) cos(y) = 0.5 (sin(x + y) + sin(x - y))
    int sgn = a.sn ? dif : -dif;    // (sin harmonic) - (cos harmonic)
    double lo = ad == 0 ? 0.0 : (sgn > 0 ? Sm[ad] : -Sm[ad]);
    return 0.5 * (Sm[sum] + lo);
}

// Packed entry e = tri(i, j) of the information matrix as a table row: H[e] = 0.5 (mom[i1] + s2 mom[i2]) + add
// (the same case analysis as trig_product / assemble, resolved once per launch instead of once per cell)
struct HEntry {
    short i1, i2;
    float s2, add;
};

ATTRICI_HD HEntry h_entry(int term, int modes, int i, int j) {
    HEntry e;
    e.i1 = H_ZERO; e.i2 = H_ZERO; e.s2 = 0.f; e.add = 0.f;
    const Col ci = col_of(i), cj = col_of(j);
    if (!col_active(ci, term, modes) || !col_active(cj, term, modes)) {
        e.add = (i == j) ? 1.f : 0.f;
        return e;
    }
    const int pw = ci.pw + cj.pw;
    int base;
    if (ci.blk == 0 && cj.blk == 0) base = ACC_AA + pw * NTRIG;
    else if (ci.blk == 1 && cj.blk == 1) base = ACC_BB;
    else base = ACC_AB + pw * NTRIG;
    const int cm = base, sm = base + NH;   // cm + k: cos k moment (k = 0: constant), sm + k: sin k moment
    const int sum = ci.k + cj.k, dif = ci.k - cj.k;
    const int ad = dif < 0 ? -dif : dif;
    if (!ci.sn && !cj.sn) { e.i1 = (short)(cm + ad); e.i2 = (short)(cm + sum); e.s2 = 1.f; }
    else if (ci.sn && cj.sn) { e.i1 = (short)(cm + ad); e.i2 = (short)(cm + sum); e.s2 = -1.f; }
    else {
        const int sgn = ci.sn ? dif : -dif;
        e.i1 = (short)(sm + sum);
        if (ad != 0) { e.i2 = (short)(sm + ad); e.s2 = sgn > 0 ? 1.f : -1.f; }
    }
    if (i == j) e.add = (float)prior_prec(ci);
    return e;
}

// objective, gradient (S.g) and Fisher information + prior (S.H, packed lower triangle) of one cell in
// the cell frame, canonical layout, at S.th.  Needs S.mom (load_moments).  Returns the objective.
template <class Ctx, class Scr>
ATTRICI_HD double assemble(const StepArgs& a, Scr& S, bool want_H, const Ctx& ctx, const HEntry* tab = nullptr) {
    double fp = 0.0;
    for (int i = ctx.lane(); i < NP; i += Ctx::W) {
        Col ci = col_of(i);
        double gi = 0.0;
        if (col_active(ci, a.term, a.modes)) {
            const double* gm = (ci.blk == 0) ? S.mom + ACC_GA + ci.pw * NB : S.mom + ACC_GB;
            gi = ci.k == 0 ? gm[0] : (ci.sn ? gm[MAXM + ci.k] : gm[ci.k]);
            double pr = prior_prec(ci);
            gi += pr * S.th[i];
            fp += 0.5 * pr * S.th[i] * S.th[i];
        }
        S.g[i] = gi;
    }
    if (want_H && tab) {
        for (int e = ctx.lane(); e < NTRI; e += Ctx::W) {
            const HEntry t = tab[e];
            S.H[e] = 0.5 * (S.mom[t.i1] + (double)t.s2 * S.mom[t.i2]) + (double)t.add;
        }
    } else if (want_H) {
        for (int i = 0; i < NP; ++i) {
            Col ci = col_of(i);
            bool act_i = col_active(ci, a.term, a.modes);
            for (int j = ctx.lane(); j <= i; j += Ctx::W) {
                Col cj = col_of(j);
                bool act_j = col_active(cj, a.term, a.modes);
                double h;
                if (!act_i || !act_j) {
                    h = (i == j) ? 1.0 : 0.0;
                } else {
                    int pw = ci.pw + cj.pw;
                    const double* m;
                    if (ci.blk == 0 && cj.blk == 0) m = S.mom + ACC_AA + pw * NTRIG;
                    else if (ci.blk == 1 && cj.blk == 1) m = S.mom + ACC_BB;
                    else m = S.mom + ACC_AB + pw * NTRIG;
                    h = trig_product(m, ci, cj);
                    if (i == j) h += prior_prec(ci);
                }
                S.H[tri(i, j)] = h;
            }
        }
    }
    double f = S.mom[NACC] + ctx.sum(fp) - prior_const(a.term, a.modes);
    ctx.sync();
    return f;
}

// solve (H + lam D) delta = -g by Cholesky (S.L), rows in parallel; returns false if not SPD
template <class Ctx, class Scr> ATTRICI_HD bool lm_solve(Scr& S, double lam, const Ctx& ctx) {
    for (int i = ctx.lane(); i < NP; i += Ctx::W) S.hd[i] = S.H[tri(i, i)];
    ctx.sync();
    for (int j = 0; j < NP; ++j) {
        double part = 0.0;
        for (int k = ctx.lane(); k < j; k += Ctx::W) { double l = S.L()[tri(j, k)]; part += l * l; }
        double hjj = S.H[tri(j, j)];
        double s = hjj + lam * fmax(fabs(hjj), 1e-12) - ctx.sum(part);
        if (!(s > 0.0) || !isfinite(s)) return false;   // uniform: s is the same in every lane
        double ljj = sqrt(s);
        double inv = 1.0 / ljj;
        for (int i = j + 1 + ctx.lane(); i < NP; i += Ctx::W) {
            double t = S.H[tri(i, j)];
            for (int k = 0; k < j; ++k) t -= S.L()[tri(i, k)] * S.L()[tri(j, k)];
            S.L()[tri(i, j)] = t * inv;
        }
        if (ctx.lane() == 0) S.L()[tri(j, j)] = ljj;
        ctx.sync();
    }
    for (int i = 0; i < NP; ++i) {   // L y = -g
        double part = 0.0;
        for (int k = ctx.lane(); k < i; k += Ctx::W) part += S.L()[tri(i, k)] * S.y[k];
        double t = (-S.g[i] - ctx.sum(part)) / S.L()[tri(i, i)];
        if (ctx.lane() == 0) S.y[i] = t;
        ctx.sync();
    }
    for (int i = NP - 1; i >= 0; --i) {   // L^T delta = y
        double part = 0.0;
        for (int k = i + 1 + ctx.lane(); k < NP; k += Ctx::W) part += S.L()[tri(k, i)] * S.delta[k];
        double t = (S.y[i] - ctx.sum(part)) / S.L()[tri(i, i)];
        if (ctx.lane() == 0) S.delta[i] = t;
        ctx.sync();
    }
    return true;
}

// Relative size of the predicted decrease below which a step is applied WITHOUT being evaluated: the
// iteration is in its quadratic phase there (the next predicted decrease is orders of magnitude below
// tol_pred), and the float32 noise of the objective (~1e-9 relative) could not confirm the step anyway.
constexpr double TOL_FINAL_STEP = 1e-8;
// relative noise level of the objective evaluated with float32 / split-TF32 day arithmetic
constexpr double F_NOISE = 2e-7;

// Normal family, first step from the moment start: with sigma constant the objective is exactly quadratic in
// the coefficients of mu, so the mu block is solved alone (ordinary least squares) and the intercept of
// log sigma is set to the residual standard deviation that solution will have.  A joint scoring step
// would fit log sigma to residuals that still contain the whole annual cycle and overshoot.
struct FirstStep {
    bool on;
    double sum_z2;   // sum of squared standardised residuals at the start
    double n;        // valid days
};

// from the accepted state in S (th, g, H) and lam: next trial point, or termination.
// Every lane returns the same status.
template <class Ctx, class Scr>
ATTRICI_HD int propose(const StepArgs& a, int cell, Scr& S, const Ctx& ctx, FirstStep first = FirstStep{false, 0.0, 0.0}) {
    double lam = a.lam[cell];
    bool ok = false;
    for (int tries = 0; tries < 40 && !ok; ++tries) {
        if constexpr (Ctx::kRegisterSolve) ok = ctx.solve(S, lam);
        else ok = lm_solve(S, lam, ctx);
        if (ok) {
            bool fin = true;
            for (int i = ctx.lane(); i < NP; i += Ctx::W) fin = fin && isfinite(S.delta[i]);
            ok = ctx.all(fin);
        }
        if (!ok) lam = fmax(lam * 10.0, 1e-6);
    }
    int status = ATTRICI_STATUS_RUNNING;
    double pred = 0.0;
    if (!ok) {
        status = ATTRICI_STATUS_STALLED;
    } else {
        // predicted decrease of the quadratic model, -g.d - 0.5 d.H.d, with H d = -g - lam D d from the solve:
        // -0.5 g.d + 0.5 lam d.D.d
        double part = 0.0;
        for (int i = ctx.lane(); i < NP; i += Ctx::W) {
            const double di = S.delta[i], hii = S.hd[i];
            part += -0.5 * S.g[i] * di + 0.5 * lam * fmax(fabs(hii), 1e-12) * di * di;
        }
        pred = ctx.sum(part);
        double f = a.f_acc[cell];
        double fs = fmax(fabs(f), 1.0);
        if (pred <= a.tol_pred * fs) {
            status = ATTRICI_STATUS_CONVERGED;
        } else if (!first.on && pred <= fmax(TOL_FINAL_STEP, a.tol_pred) * fs) {
            // final step: applied to the accepted point, not evaluated by another moment pass
            ctx.sync();
            for (int i = ctx.lane(); i < NP; i += Ctx::W) a.th_acc[(size_t)cell * NP + i] = S.th[i] + S.delta[i];
            status = ATTRICI_STATUS_CONVERGED;
        }
        if (first.on && status == ATTRICI_STATUS_RUNNING) {
            double z2 = first.sum_z2 - 2.0 * pred;
            if (z2 > 0.0 && first.n > 0.0) {
                ctx.sync();
                if (ctx.lane() == 0) S.delta[NA] = 0.5 * log(z2 / first.n);
            }
        }
    }
    const bool go_on = (status == ATTRICI_STATUS_RUNNING) && (a.npass[cell] < a.max_pass);
    ctx.sync();
    if (go_on) {
        for (int i = ctx.lane(); i < NP; i += Ctx::W) {
            double v = S.th[i] + S.delta[i];
            S.y[i] = v;   // y is free after the solve: holds the trial point for the rotation
            a.th_trial[(size_t)cell * NP + i] = v;
        }
        ctx.sync();
        rotate_to_shared(a, cell, S.y, ctx);
    }
    if (ctx.lane() == 0) {
        a.lam[cell] = lam;
        a.pred[cell] = pred;
        a.status[cell] = status;   // stays RUNNING if max_pass was hit: reported as not converged
    }
    ctx.sync();
    return go_on ? ATTRICI_STATUS_RUNNING : (status == ATTRICI_STATUS_RUNNING ? -1 : status);
}

// phase offset of the cell frame: tau of the first day this term uses (basis64 column B_TAU)
template <class Ctx> ATTRICI_HD void cell_phase(const StepArgs& a, int cell, const Ctx& ctx) {
    int first = a.st_first[(size_t)a.slot * a.Cp + cell];
    double tau0 = (first >= 0) ? a.basis64[(size_t)first * B_STRIDE + B_TAU] : 0.0;
    double* rot = a.rot + (size_t)cell * 2 * NH;
    for (int k = 1 + ctx.lane(); k <= NH; k += Ctx::W) {
        double x = (6.283185307179586 * (double)k) * tau0;
        rot[k - 1] = cos(x);
        rot[NH + k - 1] = sin(x);
    }
    ctx.sync();
}

// returns true if the cell takes part in the iteration
template <class Ctx, class Scr> ATTRICI_HD bool cell_init(const StepArgs& a, int cell, Scr& S, const Ctx& ctx) {
    const size_t so = (size_t)a.slot * a.Cp + cell;
    double n = a.st_cnt[so];
    int first = a.st_first[so];
    int status = ATTRICI_STATUS_RUNNING;
    if (!(n > 0.0) || first < 0) status = ATTRICI_STATUS_NO_DATA;
    cell_phase(a, cell, ctx);
    double v0 = 0.0, vb = 0.0;   // start values of the two intercepts
    if (status == ATTRICI_STATUS_RUNNING && !a.zero_init) {
        // moment-matched intercepts (the reference starts from zero, model_scipy.py:113,234; the MAP
        // estimate does not depend on the start, the number of passes does)
        double mean = a.st_sum[so] / n;
        double var = fmax(a.st_sq[so] / n - mean * mean, 1e-12);
        switch (a.term) {
            case T_NORMAL: v0 = mean; vb = 0.5 * log(var); break;
            case T_GAMMA:
                if (mean > 0.0) { v0 = log(mean); vb = 0.5 * log(fmax(mean * mean / var, 1e-3)); }
                break;
            case T_WEIBULL:
                if (mean > 0.0) {
                    double c = pow(sqrt(var) / mean, -1.086);
                    c = fmin(fmax(c, 0.05), 50.0);
                    vb = log(c);
                    v0 = log(mean) - lgamma(1.0 + 1.0 / c);
                }
                break;
            case T_BETA:
                if (mean > 0.0 && mean < 1.0) {
                    double phi = mean * (1.0 - mean) / var - 1.0;
                    if (!(phi > 1e-3)) phi = 1.0;
                    v0 = log(mean / (1.0 - mean));
                    vb = log(phi);
                }
                break;
            case T_BERNOULLI: {
                double p = fmin(fmax(mean, 1e-6), 1.0 - 1e-6);
                v0 = log(p / (1.0 - p));
            } break;
        }
        // stay inside the bulk of the N(0,1) priors of the intercepts
        if (!isfinite(v0)) v0 = 0.0;
        if (!isfinite(vb)) vb = 0.0;
        v0 = fmin(fmax(v0, -6.0), 6.0);
        vb = fmin(fmax(vb, -6.0), 6.0);
    }
    for (int i = ctx.lane(); i < NP; i += Ctx::W) {
        double v = (i == 0) ? v0 : (i == NA ? vb : 0.0);
        S.th[i] = v;
        a.th_trial[(size_t)cell * NP + i] = v;
        a.th_acc[(size_t)cell * NP + i] = v;
    }
    ctx.sync();
    rotate_to_shared(a, cell, S.th, ctx);
    if (ctx.lane() == 0) {
        a.lam[cell] = a.lambda0;
        a.f_acc[cell] = INFINITY;
        a.pred[cell] = 0.0;
        a.status[cell] = status;
        a.npass[cell] = 0;
    }
    ctx.sync();
    return status == ATTRICI_STATUS_RUNNING;
}

// after a pass at th_trial (moments reduced into momd): accept / reject, update damping, propose the
// next trial.  Returns true if the cell needs another pass.
template <class Ctx, class Scr>
ATTRICI_HD bool cell_step(const StepArgs& a, int cell, Scr& S, const Ctx& ctx, const HEntry* tab = nullptr,
                          bool fresh_fisher = true, bool moments_in_scratch = false) {
    if (a.status[cell] != ATTRICI_STATUS_RUNNING) return false;
    for (int i = ctx.lane(); i < NP; i += Ctx::W) S.th[i] = a.th_trial[(size_t)cell * NP + i];
    load_moments(a, cell, S, ctx, moments_in_scratch);
    // without fresh Fisher moments the information matrix of the accepted point stays in use (the matrix only
    // steers the step; near the optimum it changes by less than the damping does)
    double f_t = assemble(a, S, fresh_fisher, ctx, tab);
    if (!fresh_fisher) {
        for (int i = ctx.lane(); i < NTRI; i += Ctx::W) S.H[i] = a.h_acc[(size_t)cell * NTRI + i];
        ctx.sync();
    }
    bool fin = isfinite(f_t);
    for (int i = ctx.lane(); i < NP; i += Ctx::W) fin = fin && isfinite(S.g[i]);
    const bool finite = ctx.all(fin);
    const int np = a.npass[cell];
    const double f = a.f_acc[cell];
    double lam = a.lam[cell];
    const double pred = a.pred[cell];
    ctx.sync();
    bool accept;
    int status = ATTRICI_STATUS_RUNNING;
    if (np == 0) {
        accept = finite;
        if (!finite) status = ATTRICI_STATUS_NONFINITE;
    } else {
        // The float32 / split-TF32 day arithmetic leaves up to ~5e-8 relative noise in f.  Differences of f
        // below that level carry no information: such steps are accepted on the strength of the
        // positive-definite Fisher model alone, the damping is left as it is, and the predicted-decrease
        // tests in propose() end the iteration.
        const double noise = F_NOISE * fabs(f);
        accept = finite && (f_t <= f + noise) && (f - f_t >= 1e-4 * pred - noise);
        if (accept) {
            if (pred > 10.0 * noise) {
                double rho = (f - f_t) / pred;
                if (rho > 0.75) lam = fmax(lam * 0.2, 1e-9);
                else if (rho < 0.25) lam = lam * 2.0;
            }
        } else {
            lam = fmax(lam * 8.0, 1e-4);
            if (lam > 1e12) status = ATTRICI_STATUS_STALLED;
        }
    }
    if (ctx.lane() == 0) {
        a.npass[cell] = np + 1;
        a.lam[cell] = lam;
        if (accept) a.f_acc[cell] = f_t;
        if (status != ATTRICI_STATUS_RUNNING) a.status[cell] = status;
    }
    if (accept) {
        for (int i = ctx.lane(); i < NP; i += Ctx::W) {
            a.th_acc[(size_t)cell * NP + i] = S.th[i];
            a.g_acc[(size_t)cell * NP + i] = S.g[i];
        }
        if (fresh_fisher)
            for (int i = ctx.lane(); i < NTRI; i += Ctx::W) a.h_acc[(size_t)cell * NTRI + i] = S.H[i];
    } else {
        for (int i = ctx.lane(); i < NP; i += Ctx::W) {
            S.th[i] = a.th_acc[(size_t)cell * NP + i];
            S.g[i] = a.g_acc[(size_t)cell * NP + i];
        }
        for (int i = ctx.lane(); i < NTRI; i += Ctx::W) S.H[i] = a.h_acc[(size_t)cell * NTRI + i];
    }
    ctx.sync();
    if (status != ATTRICI_STATUS_RUNNING) return false;
    FirstStep first{false, 0.0, 0.0};
    if (np == 0 && a.term == T_NORMAL && !a.zero_init) {
        // start point has a constant sigma = exp(th[NA]): nll day terms = 0.5 z^2 + th[NA] + ln sqrt(2 pi)
        first.on = true;
        first.n = 0.5 * S.mom[ACC_BB];   // w_BB = 2 on every valid day
        first.sum_z2 = 2.0 * (S.mom[NACC] - first.n * (S.th[NA] + 0.91893853320467274178));
        ctx.sync();
        for (int i = NA + ctx.lane(); i < NP; i += Ctx::W) S.g[i] = 0.0;   // mu block alone
        ctx.sync();
    }
    return propose(a, cell, S, ctx, first) == ATTRICI_STATUS_RUNNING;
}

// th_trial := th_acc and its shared-frame image (before the final FP64 evaluation)
template <class Ctx, class Scr> ATTRICI_HD void cell_final_point(const StepArgs& a, int cell, Scr& S, const Ctx& ctx) {
    for (int i = ctx.lane(); i < NP; i += Ctx::W) {
        double v = a.th_acc[(size_t)cell * NP + i];
        S.th[i] = v;
        a.th_trial[(size_t)cell * NP + i] = v;
    }
    ctx.sync();
    rotate_to_shared(a, cell, S.th, ctx);
    ctx.sync();
}

// after the FP64 pass at th_acc (day terms reduced into momd row NACC): log posterior
// (model_scipy.py:525 "logp"), coefficients, status
// use_f_acc (converged cells): no pass at the returned point.  Either it is the accepted point the last pass evaluated
// in FP64 (objective a.f_acc), or that point plus the final step propose() applied without evaluation because its
// predicted decrease was <= TOL_FINAL_STEP |f|: the objective there is f_acc - pred up to the error of the quadratic
// model on a step of that size (observed < 1e-2 pred, i.e. < 1e-10 |f|).
template <class Ctx, class Scr> ATTRICI_HD void cell_store(const StepArgs& a, int cell, Scr& S, const Ctx& ctx, bool use_f_acc = false) {
    const size_t cp = a.Cp;
    double fp = 0.0;
    for (int i = ctx.lane(); i < NP; i += Ctx::W) {
        double v = a.th_acc[(size_t)cell * NP + i];
        S.th[i] = v;
        Col c = col_of(i);
        if (col_active(c, a.term, a.modes)) fp += 0.5 * prior_prec(c) * v * v;
    }
    double f = a.momd[(size_t)NACC * cp + cell] + ctx.sum(fp) - prior_const(a.term, a.modes);
    if (use_f_acc) {
        const double fa = a.f_acc[cell], pred = a.pred[cell];
        f = (pred <= a.tol_pred * fmax(fabs(fa), 1.0)) ? fa : fa - pred;
    }
    int st = a.status[cell];
    bool bad = (st == ATTRICI_STATUS_NO_DATA || st == ATTRICI_STATUS_NONFINITE);
    ctx.sync();
    for (int i = ctx.lane(); i < NP; i += Ctx::W) a.th_out[((size_t)a.slot * NP + i) * cp + cell] = bad ? NAN : S.th[i];
    if (ctx.lane() == 0) {
        const size_t so = (size_t)a.slot * cp + cell;
        a.logp_part[so] = bad ? NAN : -f;
        a.status_part[so] = st;
        a.npass_part[so] = a.npass[cell];
    }
    ctx.sync();
}

}  // namespace attrici
