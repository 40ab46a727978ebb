// Per-cell optimiser mathematics: Fisher-scoring Levenberg-Marquardt with per-cell convergence.
// Replaces scipy.optimize.minimize(method="L-BFGS-B") of ModelScipy.fit
// (attrici/estimation/model_scipy.py:519-523; finite differences, ~900-2400 objective calls per cell)
// by ~5-15 likelihood passes per cell.  One thread = one cell; all per-cell state is laid out
// [field][cell] so that the 27 x 27 factorisations of 32 neighbouring cells coalesce.
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

struct StepArgs {
    int C, Cp;
    int n_chunks;        // chunks written by the last pass
    int fp64;            // partial moments are double
    int term, modes, slot;
    int max_pass;
    double tol_pred, lambda0;
    int zero_init;
    const double* basis64;
    const double* st_cnt; const double* st_sum; const double* st_sq; const double* st_ln; const int* st_first;
    double* th_acc; double* th_trial; double* th_pass; double* g_acc; double* h_acc; double* h_wrk;
    double* f_acc; double* pred; double* lam; double* rot;
    int* status; int* npass; int* n_running;
    const double* part_nll; const void* part_acc;
    double* th_out; double* logp_part; int* status_part; int* npass_part;
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

// 0.5 sum prec theta^2 over the active coefficients
ATTRICI_HD_NOINLINE double prior_energy(const double* th, int term, int modes) {
    double e = 0.0;
    for (int i = 0; i < NP; ++i) {
        Col c = col_of(i);
        if (col_active(c, term, modes)) e += 0.5 * prior_prec(c) * th[i] * th[i];
    }
    return e;
}

// cell frame -> shared frame: cos/sin pair of harmonic k rotated by the cell's phase offset.
// With x = w (t - t_cell) = X - phi (X the shared-frame phase): a cos kx + b sin kx =
// (a ck - b sk) cos kX + (a sk + b ck) sin kX, ck = cos k phi, sk = sin k phi.
ATTRICI_HD_NOINLINE void rotate_to_shared(const StepArgs& a, int cell, const double* th_cell /*local [NP]*/) {
    const size_t cp = a.Cp;
    double out[NP];
    for (int i = 0; i < NP; ++i) out[i] = th_cell[i];
    for (int k = 1; k <= MAXM; ++k) {
        double c = a.rot[(size_t)(k - 1) * cp + cell], s = a.rot[(size_t)(NH + k - 1) * cp + cell];
        // A block: plain and gmt-multiplied pairs; B block
        const int ci[3] = {2 + (k - 1), 2 + 2 * MAXM + (k - 1), NA + 1 + (k - 1)};
        const int si[3] = {2 + MAXM + (k - 1), 2 + 3 * MAXM + (k - 1), NA + 1 + MAXM + (k - 1)};
        for (int q = 0; q < 3; ++q) {
            double cc = th_cell[ci[q]], ss = th_cell[si[q]];
            out[ci[q]] = cc * c - ss * s;
            out[si[q]] = cc * s + ss * c;
        }
    }
    for (int i = 0; i < NP; ++i) a.th_pass[(size_t)i * cp + cell] = out[i];
}

// reduce the chunk partials of one cell and rotate every trig moment into the cell frame
// (cos kx = ck cos kX + sk sin kX ; sin kx = ck sin kX - sk cos kX).  mom layout = pass.cu's.
ATTRICI_HD_NOINLINE double reduce_moments(const StepArgs& a, int cell, double* mom /*[NACC]*/) {
    const size_t cp = a.Cp;
    double nll = 0.0;
    for (int ch = 0; ch < a.n_chunks; ++ch) nll += a.part_nll[(size_t)ch * cp + cell];
    for (int i = 0; i < NACC; ++i) {
        double s = 0.0;
        if (a.fp64) {
            const double* p = reinterpret_cast<const double*>(a.part_acc) + (size_t)i * cp + cell;
            for (int ch = 0; ch < a.n_chunks; ++ch) s += p[(size_t)ch * NACC * cp];
        } else {
            const float* p = reinterpret_cast<const float*>(a.part_acc) + (size_t)i * cp + cell;
            for (int ch = 0; ch < a.n_chunks; ++ch) s += (double)p[(size_t)ch * NACC * cp];
        }
        mom[i] = s;
    }
    for (int k = 1; k <= NH; ++k) {
        double c = a.rot[(size_t)(k - 1) * cp + cell], s = a.rot[(size_t)(NH + k - 1) * cp + cell];
        for (int blk = 0; blk < 6; ++blk) {  // AA p0..2, AB p0..1, BB
            double* m = mom + blk * NTRIG;
            double C_ = m[k], S_ = m[NH + k];
            m[k] = c * C_ + s * S_;
            m[NH + k] = c * S_ - s * C_;
        }
        if (k <= MAXM) {
            for (int blk = 0; blk < 3; ++blk) {  // GA p0, GA p1, GB
                double* m = mom + ACC_GA + blk * NB;
                double C_ = m[k], S_ = m[MAXM + k];
                m[k] = c * C_ + s * S_;
                m[MAXM + k] = c * S_ - s * C_;
            }
        }
    }
    return nll;
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

// objective, gradient and Fisher information (+ prior) of one cell, cell frame, canonical layout.
// g: local [NP]; Hglob: packed [NTRI] column of this cell with stride Cp (nullable)
ATTRICI_HD_NOINLINE double assemble(const StepArgs& a, int cell, const double* th /*local [NP], cell frame*/,
                                    double* g, double* Hglob) {
    const size_t cp = a.Cp;
    double mom[NACC];
    double f = reduce_moments(a, cell, mom);
    // cos(0) moment doubles as "cos k with k = 0": trig_product reads Cm[0] for equal harmonics
    for (int i = 0; i < NP; ++i) {
        Col ci = col_of(i);
        bool act_i = col_active(ci, a.term, a.modes);
        double gi = 0.0;
        if (act_i) {
            const double* gm = (ci.blk == 0) ? mom + ACC_GA + ci.pw * NB : mom + ACC_GB;
            gi = ci.k == 0 ? gm[0] : (ci.sn ? gm[MAXM + ci.k] : gm[ci.k]);
            double pr = prior_prec(ci);
            gi += pr * th[i];
            f += 0.5 * pr * th[i] * th[i];
        }
        g[i] = gi;
        if (!Hglob) continue;
        for (int j = 0; j <= i; ++j) {
            Col cj = col_of(j);
            bool act_j = col_active(cj, a.term, a.modes);
            double h;
            if (!act_i || !act_j) {
                h = (i == j) ? 1.0 : 0.0;
            } else {
                int pw = ci.pw + cj.pw;
                const double* m;
                if (ci.blk == 0 && cj.blk == 0) m = mom + ACC_AA + pw * NTRIG;
                else if (ci.blk == 1 && cj.blk == 1) m = mom + ACC_BB;
                else m = mom + ACC_AB + pw * NTRIG;
                h = trig_product(m, ci, cj);
                if (i == j) h += prior_prec(ci);
            }
            Hglob[(size_t)tri(i, j) * cp] = h;
        }
    }
    return f - prior_const(a.term, a.modes);
}

// solve (H + lam D) delta = -g by Cholesky in the global work column; returns false if not SPD
ATTRICI_HD_NOINLINE bool lm_solve(const StepArgs& a, int cell, const double* g, double lam, double* delta) {
    const size_t cp = a.Cp;
    const double* H = a.h_acc + cell;
    double* L = a.h_wrk + cell;
    for (int j = 0; j < NP; ++j) {
        double hjj = H[(size_t)tri(j, j) * cp];
        double dj = fmax(fabs(hjj), 1e-12);
        double s = hjj + lam * dj;
        for (int k = 0; k < j; ++k) {
            double l = L[(size_t)tri(j, k) * cp];
            s -= l * l;
        }
        if (!(s > 0.0) || !isfinite(s)) return false;
        double ljj = sqrt(s);
        L[(size_t)tri(j, j) * cp] = ljj;
        double inv = 1.0 / ljj;
        for (int i = j + 1; i < NP; ++i) {
            double t = H[(size_t)tri(i, j) * cp];
            for (int k = 0; k < j; ++k) t -= L[(size_t)tri(i, k) * cp] * L[(size_t)tri(j, k) * cp];
            L[(size_t)tri(i, j) * cp] = t * inv;
        }
    }
    double y[NP];
    for (int i = 0; i < NP; ++i) {
        double t = -g[i];
        for (int k = 0; k < i; ++k) t -= L[(size_t)tri(i, k) * cp] * y[k];
        y[i] = t / L[(size_t)tri(i, i) * cp];
    }
    for (int i = NP - 1; i >= 0; --i) {
        double t = y[i];
        for (int k = i + 1; k < NP; ++k) t -= L[(size_t)tri(k, i) * cp] * delta[k];
        delta[i] = t / L[(size_t)tri(i, i) * cp];
    }
    return true;
}

// from the accepted state (th_acc, f_acc, g_acc, h_acc, lam): next trial point, or termination
ATTRICI_HD_NOINLINE void propose(const StepArgs& a, int cell) {
    const size_t cp = a.Cp;
    double g[NP], th[NP], delta[NP];
    for (int i = 0; i < NP; ++i) {
        g[i] = a.g_acc[(size_t)i * cp + cell];
        th[i] = a.th_acc[(size_t)i * cp + cell];
    }
    double lam = a.lam[cell];
    bool ok = false;
    for (int tries = 0; tries < 40 && !ok; ++tries) {
        ok = lm_solve(a, cell, g, lam, delta);
        if (ok) for (int i = 0; i < NP; ++i) ok = ok && isfinite(delta[i]);
        if (!ok) lam = fmax(lam * 10.0, 1e-6);
    }
    a.lam[cell] = lam;
    if (!ok) { a.status[cell] = ATTRICI_STATUS_STALLED; return; }
    // predicted decrease of the quadratic model: -g.d - 0.5 d.H.d
    double gd = 0.0, dhd = 0.0;
    const double* H = a.h_acc + cell;
    for (int i = 0; i < NP; ++i) {
        gd += g[i] * delta[i];
        double hi = 0.0;
        for (int j = 0; j < NP; ++j) hi += H[(size_t)tri(i, j) * cp] * delta[j];
        dhd += delta[i] * hi;
    }
    double pred = -gd - 0.5 * dhd;
    a.pred[cell] = pred;
    double f = a.f_acc[cell];
    if (pred <= a.tol_pred * fmax(fabs(f), 1.0)) { a.status[cell] = ATTRICI_STATUS_CONVERGED; return; }
    if (a.npass[cell] >= a.max_pass) return;  // stays RUNNING: reported as not converged
    for (int i = 0; i < NP; ++i) {
        th[i] += delta[i];
        a.th_trial[(size_t)i * cp + cell] = th[i];
    }
    rotate_to_shared(a, cell, th);
}

// phase offset of the cell frame: tau of the first day this term uses (basis64 column B_TAU)
ATTRICI_HD_NOINLINE void cell_phase(const StepArgs& a, int cell) {
    const size_t cp = a.Cp;
    int first = a.st_first[(size_t)a.slot * cp + cell];
    double tau0 = (first >= 0) ? a.basis64[(size_t)first * B_STRIDE + B_TAU] : 0.0;
    for (int k = 1; k <= NH; ++k) {
        double x = (6.283185307179586 * (double)k) * tau0;
        a.rot[(size_t)(k - 1) * cp + cell] = cos(x);
        a.rot[(size_t)(NH + k - 1) * cp + cell] = sin(x);
    }
}

// returns true if the cell takes part in the iteration
ATTRICI_HD_NOINLINE bool cell_init(const StepArgs& a, int cell) {
    const size_t cp = a.Cp;
    const size_t so = (size_t)a.slot * cp + cell;
    double th[NP];
    for (int i = 0; i < NP; ++i) th[i] = 0.0;
    double n = a.st_cnt[so];
    int first = a.st_first[so];
    int status = ATTRICI_STATUS_RUNNING;
    if (!(n > 0.0) || first < 0) status = ATTRICI_STATUS_NO_DATA;
    cell_phase(a, cell);
    if (status == ATTRICI_STATUS_RUNNING && !a.zero_init) {
        // moment-matched intercepts (the reference starts from zero, model_scipy.py:113,234; the MAP
        // estimate does not depend on the start, the number of passes does)
        double mean = a.st_sum[so] / n;
        double var = fmax(a.st_sq[so] / n - mean * mean, 1e-12);
        switch (a.term) {
            case T_NORMAL: th[0] = mean; th[NA] = 0.5 * log(var); break;
            case T_GAMMA:
                if (mean > 0.0) { th[0] = log(mean); th[NA] = 0.5 * log(fmax(mean * mean / var, 1e-3)); }
                break;
            case T_WEIBULL:
                if (mean > 0.0) {
                    double c = pow(sqrt(var) / mean, -1.086);
                    c = fmin(fmax(c, 0.05), 50.0);
                    th[NA] = log(c);
                    th[0] = log(mean) - lgamma(1.0 + 1.0 / c);
                }
                break;
            case T_BETA:
                if (mean > 0.0 && mean < 1.0) {
                    double phi = mean * (1.0 - mean) / var - 1.0;
                    if (!(phi > 1e-3)) phi = 1.0;
                    th[0] = log(mean / (1.0 - mean));
                    th[NA] = log(phi);
                }
                break;
            case T_BERNOULLI: {
                double p = fmin(fmax(mean, 1e-6), 1.0 - 1e-6);
                th[0] = log(p / (1.0 - p));
            } break;
        }
        // stay inside the bulk of the N(0,1) priors of the intercepts
        for (int i = 0; i < NP; ++i) {
            if (!isfinite(th[i])) th[i] = 0.0;
            th[i] = fmin(fmax(th[i], -6.0), 6.0);
        }
    }
    for (int i = 0; i < NP; ++i) {
        a.th_trial[(size_t)i * cp + cell] = th[i];
        a.th_acc[(size_t)i * cp + cell] = th[i];
    }
    rotate_to_shared(a, cell, th);
    a.lam[cell] = a.lambda0;
    a.f_acc[cell] = INFINITY;
    a.pred[cell] = 0.0;
    a.status[cell] = status;
    a.npass[cell] = 0;
    return status == ATTRICI_STATUS_RUNNING;
}

// after a pass at th_trial: accept / reject, update damping, propose the next trial.
// returns true if the cell needs another pass
ATTRICI_HD_NOINLINE bool cell_step(const StepArgs& a, int cell) {
    if (a.status[cell] != ATTRICI_STATUS_RUNNING) return false;
    const size_t cp = a.Cp;
    double th[NP], g[NP];
    for (int i = 0; i < NP; ++i) th[i] = a.th_trial[(size_t)i * cp + cell];
    // Fisher information of the trial goes to the work column first; promoted on acceptance
    double f_t = assemble(a, cell, th, g, a.h_wrk + cell);
    bool finite = isfinite(f_t);
    for (int i = 0; i < NP; ++i) finite = finite && isfinite(g[i]);
    int np = a.npass[cell];
    a.npass[cell] = np + 1;
    double f = a.f_acc[cell];
    double lam = a.lam[cell];
    bool accept;
    if (np == 0) {
        if (!finite) { a.status[cell] = ATTRICI_STATUS_NONFINITE; return false; }
        accept = true;
    } else {
        // float32 day arithmetic leaves ~1e-9 relative noise in f: tolerate it, the predicted-decrease
        // test in propose() is what ends the iteration
        double slack = 4e-9 * fabs(f);
        double pred = a.pred[cell];
        accept = finite && (f_t <= f + slack) && (f - f_t >= 1e-4 * pred - slack);
        if (accept) {
            double rho = (f - f_t) / pred;
            if (rho > 0.75) lam = fmax(lam * 0.2, 1e-9);
            else if (rho < 0.25) lam = lam * 2.0;
        } else {
            lam = fmax(lam * 8.0, 1e-4);
            if (lam > 1e12) { a.status[cell] = ATTRICI_STATUS_STALLED; return false; }
        }
    }
    a.lam[cell] = lam;
    if (accept) {
        a.f_acc[cell] = f_t;
        for (int i = 0; i < NP; ++i) {
            a.th_acc[(size_t)i * cp + cell] = th[i];
            a.g_acc[(size_t)i * cp + cell] = g[i];
        }
        for (int i = 0; i < NTRI; ++i) a.h_acc[(size_t)i * cp + cell] = a.h_wrk[(size_t)i * cp + cell];
    }
    propose(a, cell);
    return a.status[cell] == ATTRICI_STATUS_RUNNING && a.npass[cell] < a.max_pass;
}

// th_trial := th_acc and its shared-frame image (before the final FP64 evaluation)
ATTRICI_HD_NOINLINE void cell_final_point(const StepArgs& a, int cell) {
    const size_t cp = a.Cp;
    double th[NP];
    for (int i = 0; i < NP; ++i) {
        th[i] = a.th_acc[(size_t)i * cp + cell];
        a.th_trial[(size_t)i * cp + cell] = th[i];
    }
    rotate_to_shared(a, cell, th);
}

// after the FP64 pass at th_acc: log posterior (model_scipy.py:525 "logp"), coefficients, status
ATTRICI_HD_NOINLINE void cell_store(const StepArgs& a, int cell) {
    const size_t cp = a.Cp;
    double th[NP];
    for (int i = 0; i < NP; ++i) th[i] = a.th_acc[(size_t)i * cp + cell];
    double nll = 0.0;
    for (int ch = 0; ch < a.n_chunks; ++ch) nll += a.part_nll[(size_t)ch * cp + cell];
    double f = nll + prior_energy(th, a.term, a.modes) - prior_const(a.term, a.modes);
    int st = a.status[cell];
    bool bad = (st == ATTRICI_STATUS_NO_DATA || st == ATTRICI_STATUS_NONFINITE);
    const size_t so = (size_t)a.slot * cp + cell;
    for (int i = 0; i < NP; ++i) a.th_out[((size_t)a.slot * NP + i) * cp + cell] = bad ? NAN : th[i];
    a.logp_part[so] = bad ? NAN : -f;
    a.status_part[so] = st;
    a.npass_part[so] = a.npass[cell];
}

}  // namespace attrici
