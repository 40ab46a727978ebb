// CPU build of the __host__ __device__ mathematics of attrici_b200/csrc (special.cuh, families.cuh,
// pass_impl.cuh, step_impl.cuh, qmap_impl.cuh) for UNIT TESTS against scipy and the oracle
// (tests/test_hostmath.py).  Test scaffolding only: nothing in the product loads this library, and
// the product has no CPU path.  The drivers below emulate "one thread = one cell" for a single cell.
#include <cstdint>
#include <cstring>
#include <vector>
#include <cmath>
#include "../../attrici_b200/csrc/layout.cuh"
#include "../../attrici_b200/csrc/families.cuh"
#include "../../attrici_b200/csrc/pass_impl.cuh"
#include "../../attrici_b200/csrc/step_impl.cuh"
#include "../../attrici_b200/csrc/qmap_impl.cuh"
using namespace attrici;

namespace {

// same arithmetic as build_basis_kernel (basis.cu)
void host_basis(const int32_t* days, const double* gmt, int T, std::vector<double>& b64, std::vector<float>& b32) {
    b64.assign((size_t)(T + 2 * PASS_TILE) * B_STRIDE, 0.0);
    b32.assign((size_t)(T + 2 * PASS_TILE) * B_STRIDE, 0.0f);
    for (int t = 0; t < T; ++t) {
        double* row = &b64[(size_t)t * B_STRIDE];
        double tau = (double)(days[t] - days[0]) / 365.25;
        row[0] = gmt[t];
        row[B_TAU] = tau;
        for (int k = 1; k <= NH; ++k) {
            double x = (6.283185307179586 * (double)k) * tau;
            row[k] = cos(x);
            row[NH + k] = sin(x);
        }
        for (int i = 0; i < B_STRIDE; ++i) b32[(size_t)t * B_STRIDE + i] = (float)row[i];
    }
}

template <int KIND, typename R, bool MOMENTS>
void host_pass_t(const float* ys, int i0, int i1, int chunk_len, int n_chunks, const R* basis, const double* th_pass,
                 double* part_nll, R* part_acc) {
    R thA[NA], thB[NB];
    for (int i = 0; i < NA; ++i) thA[i] = (R)th_pass[i];
    for (int i = 0; i < NB; ++i) thB[i] = (KIND != K_BERNOULLI) ? (R)th_pass[NA + i] : R(0);
    for (int ch = 0; ch < n_chunks; ++ch) {
        int t0 = i0 + ch * chunk_len, t1 = std::min(i1, t0 + chunk_len);
        PassAcc<R> acc;
        acc.clear();
        double nll = 0.0;
        for (int t = t0; t < t1; ++t) {
            Row<R> b = load_row_scalar<R>(basis + (size_t)t * B_STRIDE);
            nll += (double)pass_day<KIND, R, MOMENTS>(ys[t], true, thA, thB, b, acc);
        }
        part_nll[ch] = nll;
        if (MOMENTS) {
            R* out = part_acc + (size_t)ch * NACC;
            for (int i = 0; i < 3 * NTRIG; ++i) out[ACC_AA + i] = acc.AA[i];
            for (int i = 0; i < 2 * NTRIG; ++i) out[ACC_AB + i] = (KIND == K_WEIBULL || KIND == K_BETA) ? acc.AB[i] : R(0);
            for (int i = 0; i < NTRIG; ++i) out[ACC_BB + i] = (KIND != K_BERNOULLI) ? acc.BB[i] : R(0);
            for (int i = 0; i < 2 * NB; ++i) out[ACC_GA + i] = acc.GA[i];
            for (int i = 0; i < NB; ++i) out[ACC_GB + i] = (KIND != K_BERNOULLI) ? acc.GB[i] : R(0);
        }
    }
}

template <typename R, bool MOMENTS>
void host_pass(int term, const float* ys, int i0, int i1, int chunk_len, int n_chunks, const R* basis,
               const double* th_pass, double* part_nll, R* part_acc) {
    switch (term) {
        case T_NORMAL: host_pass_t<K_NORMAL, R, MOMENTS>(ys, i0, i1, chunk_len, n_chunks, basis, th_pass, part_nll, part_acc); break;
        case T_GAMMA: host_pass_t<K_GAMMA, R, MOMENTS>(ys, i0, i1, chunk_len, n_chunks, basis, th_pass, part_nll, part_acc); break;
        case T_BETA: host_pass_t<K_BETA, R, MOMENTS>(ys, i0, i1, chunk_len, n_chunks, basis, th_pass, part_nll, part_acc); break;
        case T_WEIBULL: host_pass_t<K_WEIBULL, R, MOMENTS>(ys, i0, i1, chunk_len, n_chunks, basis, th_pass, part_nll, part_acc); break;
        default: host_pass_t<K_BERNOULLI, R, MOMENTS>(ys, i0, i1, chunk_len, n_chunks, basis, th_pass, part_nll, part_acc); break;
    }
}

// single-cell workspace (Cp = 1)
struct HostCell {
    std::vector<double> b64;
    std::vector<float> b32;
    double st_cnt[2], st_sum[2], st_sq[2], st_ln[2];
    int st_first[2];
    double th_acc[NP], th_trial[NP], th_pass[NP], g_acc[NP], h_acc[NTRI], momd[NACC + 1];
    double f_acc, pred, lam, rot[2 * NH];
    int status, npass;
    double part_nll[MAX_CHUNKS];
    std::vector<double> part_acc;   // holds float or double partials
    double th_out[2 * NP], logp_part[2];
    int status_part[2], npass_part[2];
    CellScratch S;

    StepArgs args(int term, int modes, int slot, int max_pass, double tol_pred, double lambda0, int zero_init) {
        StepArgs a;
        a.C = 1; a.Cp = 1;
        a.term = term; a.modes = modes; a.slot = slot;
        a.max_pass = max_pass; a.tol_pred = tol_pred; a.lambda0 = lambda0; a.zero_init = zero_init;
        a.basis64 = b64.data();
        a.st_cnt = st_cnt; a.st_sum = st_sum; a.st_sq = st_sq; a.st_ln = st_ln; a.st_first = st_first;
        a.momd = momd;
        a.th_acc = th_acc; a.th_trial = th_trial; a.th_pass = th_pass; a.g_acc = g_acc; a.h_acc = h_acc;
        a.f_acc = &f_acc; a.pred = &pred; a.lam = &lam; a.rot = rot;
        a.status = &status; a.npass = &npass;
        a.th_out = th_out; a.logp_part = logp_part; a.status_part = status_part; a.npass_part = npass_part;
        return a;
    }

    // reduce_partials_kernel
    void reduce(bool fp64, bool moments, int n_chunks) {
        double s = 0.0;
        for (int ch = 0; ch < n_chunks; ++ch) s += part_nll[ch];
        momd[NACC] = s;
        if (!moments) return;
        for (int i = 0; i < NACC; ++i) {
            double t = 0.0;
            for (int ch = 0; ch < n_chunks; ++ch)
                t += fp64 ? part_acc[(size_t)ch * NACC + i]
                          : (double)reinterpret_cast<const float*>(part_acc.data())[(size_t)ch * NACC + i];
            momd[i] = t;
        }
    }
};

// start statistics of the fit window (prep_scale_kernel / prep_stats_kernel)
void host_stats(HostCell& hc, int term, int slot, const float* ys, int i0, int i1) {
    double n = 0, s = 0, q = 0, l = 0;
    int first = -1;
    for (int t = i0; t < i1; ++t) {
        bool valid = !(ys[t] != ys[t]);
        if (term == T_BERNOULLI) {
            n += 1.0; s += valid ? 0.0 : 1.0; q += valid ? 0.0 : 1.0;
            if (first < 0) first = t;
        } else if (valid) {
            double d = (double)ys[t];
            n += 1.0; s += d; q += d * d;
            if (d > 0.0) l += log(d);
            if (first < 0) first = t;
        }
    }
    hc.st_cnt[slot] = n; hc.st_sum[slot] = s; hc.st_sq[slot] = q; hc.st_ln[slot] = l; hc.st_first[slot] = first;
}

void run_pass(HostCell& hc, int term, bool fp64, bool moments, const float* ys, int i0, int i1, int chunk_len, int n_chunks) {
    if (fp64) {
        double* pa = hc.part_acc.data();
        if (moments) host_pass<double, true>(term, ys, i0, i1, chunk_len, n_chunks, hc.b64.data(), hc.th_pass, hc.part_nll, pa);
        else host_pass<double, false>(term, ys, i0, i1, chunk_len, n_chunks, hc.b64.data(), hc.th_pass, hc.part_nll, pa);
    } else {
        float* pa = reinterpret_cast<float*>(hc.part_acc.data());
        if (moments) host_pass<float, true>(term, ys, i0, i1, chunk_len, n_chunks, hc.b32.data(), hc.th_pass, hc.part_nll, pa);
        else host_pass<float, false>(term, ys, i0, i1, chunk_len, n_chunks, hc.b32.data(), hc.th_pass, hc.part_nll, pa);
    }
}

}  // namespace

extern "C" {

void hm_unary(int which, const double* x, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        switch (which) {
            case 0: out[i] = digamma<double>(x[i]); break;
            case 1: out[i] = trigamma<double>(x[i]); break;
            case 2: out[i] = ndtr(x[i]); break;
            case 3: out[i] = ndtri(x[i]); break;
            case 4: out[i] = (double)digamma<float>((float)x[i]); break;
            case 5: out[i] = (double)trigamma<float>((float)x[i]); break;
        }
    }
}

void hm_binary(int which, const double* a, const double* x, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        switch (which) {
            case 0: out[i] = gamma_p(a[i], x[i]); break;
            case 1: out[i] = gamma_q(a[i], x[i]); break;
            case 2: out[i] = gamma_p_inv(a[i], x[i]); break;
        }
    }
}

// the inverse as the precipitation quantile map uses it (qmap_pr.cu): P, Q and the prefactor at the starting point x0
// from one evaluation, then the Halley iteration from x0 to the quantile of p
void hm_gamma_inv_from(const double* a, const double* x0, const double* p, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        const double gln = lgamma(a[i]);
        double P, Q, pref;
        gamma_pq_pref_l(a[i], x0[i], gln, &P, &Q, &pref);
        out[i] = gamma_p_inv_from(a[i], p[i], x0[i], P, Q, pref, gln);
    }
}

void hm_ternary(int which, const double* a, const double* b, const double* x, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        switch (which) {
            case 0: out[i] = beta_i(a[i], b[i], x[i]); break;
            case 1: out[i] = beta_i_inv(a[i], b[i], x[i]); break;
        }
    }
}

// kind, y, etaA, etaB -> out[6] = nll, rA, rB, wAA, wAB, wBB (Fisher weights)
void hm_family_f64(int kind, const double* y, const double* ea, const double* eb, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        DayTerms<double> d = family_day<double>(kind, y[i], ea[i], eb[i]);
        out[6 * i + 0] = d.nll; out[6 * i + 1] = d.rA; out[6 * i + 2] = d.rB;
        out[6 * i + 3] = d.wAA; out[6 * i + 4] = d.wAB; out[6 * i + 5] = d.wBB;
    }
}
void hm_family_f32(int kind, const double* y, const double* ea, const double* eb, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        DayTerms<float> d = family_day<float>(kind, (float)y[i], (float)ea[i], (float)eb[i]);
        out[6 * i + 0] = d.nll; out[6 * i + 1] = d.rA; out[6 * i + 2] = d.rB;
        out[6 * i + 3] = d.wAA; out[6 * i + 4] = d.wAB; out[6 * i + 5] = d.wBB;
    }
}

int hm_num_params(int family, int modes) { return family_num_params(family, modes); }

// reference-layout index -> (slot, canonical index)
void hm_ref_to_canon(int family, int modes, int j, int* slot, int* canon) { ref_to_canon(family, modes, j, slot, canon); }

// objective / gradient / Fisher information of one likelihood term at theta (canonical layout, cell frame)
// through the pass + assemble code of the device path.  fisher: [NP*NP] full symmetric (nullable)
int hm_nll_grad_cell(int term, int modes, int T, int i0, int i1, const float* ys, const int32_t* days, const double* gmt,
                     int fp64, int n_chunks, const double* theta, double* nll, double* grad, double* fisher) {
    HostCell hc;
    host_basis(days, gmt, T, hc.b64, hc.b32);
    hc.part_acc.assign((size_t)MAX_CHUNKS * NACC, 0.0);
    host_stats(hc, term, 0, ys, i0, i1);
    if (n_chunks > MAX_CHUNKS) n_chunks = MAX_CHUNKS;
    int chunk_len = (i1 - i0 + n_chunks - 1) / n_chunks;
    n_chunks = (i1 - i0 + chunk_len - 1) / chunk_len;
    StepArgs a = hc.args(term, modes, 0, 60, 1e-11, 1e-3, 1);
    SerialCtx ctx;
    cell_phase(a, 0, ctx);
    for (int i = 0; i < NP; ++i) { hc.th_trial[i] = theta[i]; hc.S.th[i] = theta[i]; }
    rotate_to_shared(a, 0, hc.th_trial, ctx);
    run_pass(hc, term, fp64 != 0, true, ys, i0, i1, chunk_len, n_chunks);
    hc.reduce(fp64 != 0, true, n_chunks);
    load_moments(a, 0, hc.S, ctx);
    *nll = assemble(a, hc.S, true, ctx);
    for (int i = 0; i < NP; ++i) grad[i] = hc.S.g[i];
    if (fisher)
        for (int i = 0; i < NP; ++i)
            for (int j = 0; j < NP; ++j) fisher[i * NP + j] = hc.S.H[tri(i, j)];
    return 0;
}

// the Fisher-scoring fit of one likelihood term for one cell, same control flow as api.cu fit_term
int hm_fit_cell(int term, int modes, int T, int i0, int i1, const float* ys, const int32_t* days, const double* gmt,
                int fp64, int zero_init, int max_pass, double tol_pred, double lambda0, int n_chunks,
                double* theta_out, double* logp, int* status, int* npass, double* trace /*[max_pass][4], nullable*/) {
    HostCell hc;
    host_basis(days, gmt, T, hc.b64, hc.b32);
    hc.part_acc.assign((size_t)MAX_CHUNKS * NACC, 0.0);
    host_stats(hc, term, 0, ys, i0, i1);
    if (n_chunks > MAX_CHUNKS) n_chunks = MAX_CHUNKS;
    int chunk_len = (i1 - i0 + n_chunks - 1) / n_chunks;
    n_chunks = (i1 - i0 + chunk_len - 1) / chunk_len;
    StepArgs a = hc.args(term, modes, 0, max_pass, tol_pred, lambda0, zero_init);
    SerialCtx ctx;
    bool running = cell_init(a, 0, hc.S, ctx);
    for (int it = 0; running && it < max_pass; ++it) {
        run_pass(hc, term, fp64 != 0, true, ys, i0, i1, chunk_len, n_chunks);
        hc.reduce(fp64 != 0, true, n_chunks);
        running = cell_step(a, 0, hc.S, ctx);
        if (trace) { trace[4 * it] = hc.momd[NACC]; trace[4 * it + 1] = hc.f_acc; trace[4 * it + 2] = hc.pred; trace[4 * it + 3] = hc.lam; }
    }
    cell_final_point(a, 0, hc.S, ctx);
    run_pass(hc, term, true, false, ys, i0, i1, chunk_len, n_chunks);
    hc.reduce(true, false, n_chunks);
    cell_store(a, 0, hc.S, ctx);
    for (int i = 0; i < NP; ++i) theta_out[i] = hc.th_out[i];
    *logp = hc.logp_part[0];
    *status = hc.status_part[0];
    *npass = hc.npass_part[0];
    return 0;
}

// quantile map of one cell: params in the REFERENCE layout; uniforms may be null unless family = pr
int hm_qmap_cell(int family, int scaling, int post_rule, int modes, int T, const float* y, const float* ys,
                 const int32_t* days, const double* gmt, const double* params, double datamin, double scale,
                 const double* uniforms, double* cfact, double* cfact_scaled, int* replaced, double* par_ref,
                 double* par_cf) {
    std::vector<double> b64;
    std::vector<float> b32;
    host_basis(days, gmt, T, b64, b32);
    CellCoef c;
    memset(&c, 0, sizeof(c));
    const int P = family_num_params(family, modes);
    for (int j = 0; j < P; ++j) {
        int sl, cn;
        ref_to_canon(family, modes, j, &sl, &cn);
        if (cn < NA) c.a[sl][cn] = params[j];
        else c.b[cn - NA] = params[j];
    }
    const int n_dep = family == ATTRICI_BERNOULLI_GAMMA ? 2 : 1;
    const bool mask_in_place = (scaling == ATTRICI_SCALE_PERCENT || scaling == ATTRICI_SCALE_RANGE);
    float af[NA], bf[NB];
    for (int i = 0; i < NA; ++i) af[i] = (float)c.a[0][i];
    for (int i = 0; i < NB; ++i) bf[i] = (float)c.b[i];
    for (int t = 0; t < T; ++t) {
        DayParams d;
        day_predictors(c, &b64[(size_t)t * B_STRIDE], n_dep, d);
        DistPar ref, cf;
        family_params(family, d, ref, cf);
        float yv = y[t];
        if (std::isinf(yv) || (mask_in_place && (ys[t] != ys[t]))) yv = NAN;
        QmapOut o;
        double cs_fast;
        // same dispatch as qmap_kernel: shortcut for the bulk of the Normal family, full map otherwise
        if (family == ATTRICI_NORMAL && normal_fast_path(c, af, bf, &b32[(size_t)t * B_STRIDE], &b64[(size_t)t * B_STRIDE], ys[t], &cs_fast))
            o = qmap_finish(scaling, post_rule, cs_fast, yv, datamin, scale);
        else
            o = qmap_day(family, scaling, post_rule, yv, ys[t], ref, cf, datamin, scale, uniforms ? uniforms[t] : 0.0);
        cfact[t] = o.cfact;
        if (cfact_scaled) cfact_scaled[t] = o.cfact_scaled;
        if (replaced) replaced[t] = o.replaced;
        if (par_ref) { par_ref[3 * t] = ref.p0; par_ref[3 * t + 1] = ref.p1; par_ref[3 * t + 2] = ref.p2; }
        if (par_cf) { par_cf[3 * t] = cf.p0; par_cf[3 * t + 1] = cf.p1; par_cf[3 * t + 2] = cf.p2; }
    }
    return 0;
}

// NumPy legacy MT19937 doubles
struct HostState {
    uint32_t* p;
    uint32_t& operator()(int i) const { return p[i]; }
};
void hm_mt19937(uint32_t seed, int n, double* out) {
    uint32_t mt[Mt19937::N];
    HostState st{mt};
    Mt19937::seed(st, seed);
    int pos = Mt19937::N;
    for (int i = 0; i < n; ++i) {
        uint32_t ab[2];
        for (int j = 0; j < 2; ++j) {
            if (pos >= Mt19937::N) { Mt19937::twist(st); pos = 0; }
            ab[j] = Mt19937::temper(mt[pos++]);
        }
        out[i] = Mt19937::to_double(ab[0], ab[1]);
    }
}

}  // extern "C"
