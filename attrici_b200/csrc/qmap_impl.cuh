// Per-day quantile map: factual / counterfactual distribution parameters, cdf -> inverse cdf,
// rescale and NaN/Inf replacement, all in FP64 like the reference.
// Replaces ModelScipy.estimate_distribution x2 (attrici/estimation/model_scipy.py:547-570 with the real
// predictor and with predictor * 0, attrici/detrend.py:362-370), attrici/distributions.py:81-216,
// Variable.quantile_mapping and its overrides (attrici/variables.py:287-303, 422-480, 645-653, 691-698),
// rescale (variables.py:71-89, 114-131, 483-486) and the replacement of attrici/detrend.py:397-411.
// __host__ __device__: tests/hostmath runs the identical code on the CPU against scipy.
#pragma once
#include "layout.cuh"

namespace attrici {

// coefficients of one cell in the canonical layout, per term slot (cell frame = what the reference
// stores in trace["params"]; evaluation uses the series origin)
struct CellCoef {
    double a[2][NA];
    double b[NB];      // the independent parameter (sigma / nu / phi / alpha)
};

struct DayParams {
    double a_ref[2], a_cf[2];   // linear predictors of the dependent parameters, factual / counterfactual
    double b;                   // linear predictor of the independent parameter
};

// row: basis64 row {gmt, cos 1..8, sin 1..8, ...}
ATTRICI_HD void day_predictors(const CellCoef& c, const double* row, int n_dep, DayParams& d) {
    const double g = row[0];
    for (int s = 0; s < n_dep; ++s) {
        double base = c.a[s][0], trend = c.a[s][1];
#pragma unroll
        for (int k = 0; k < MAXM; ++k) {
            double ck = row[1 + k], sk = row[1 + NH + k];
            base = fma(c.a[s][2 + k], ck, base);
            base = fma(c.a[s][2 + MAXM + k], sk, base);
            trend = fma(c.a[s][2 + 2 * MAXM + k], ck, trend);
            trend = fma(c.a[s][2 + 3 * MAXM + k], sk, trend);
        }
        d.a_cf[s] = base;                 // predictor == 0 (detrend.py:366-370)
        d.a_ref[s] = fma(g, trend, base);
    }
    double eb = c.b[0];
#pragma unroll
    for (int k = 0; k < MAXM; ++k) {
        eb = fma(c.b[1 + k], row[1 + k], eb);
        eb = fma(c.b[1 + MAXM + k], row[1 + NH + k], eb);
    }
    d.b = eb;
}

ATTRICI_HD double invlogit(double x) { return 1.0 / (1.0 + exp(-x)); }

// scipy.stats.gamma.cdf(y, a, scale) / ppf(q, a, scale)
ATTRICI_HD double gamma_cdf(double y, double a, double scale) {
    if (isnan(y) || isnan(a) || isnan(scale) || !(a > 0.0) || !(scale > 0.0)) return NAN;
    double x = y / scale;
    if (x <= 0.0) return 0.0;
    return gamma_p(a, x);
}
// x0: optional starting value for the standardised quantile (<= 0: none)
// p0 / q0: P(a, x0) and Q(a, x0) when the caller has them (see gamma_p_inv)
ATTRICI_HD double gamma_ppf(double q, double a, double scale, double x0 = -1.0, double p0 = -1.0, double q0 = -1.0) {
    if (isnan(q) || isnan(a) || isnan(scale) || !(a > 0.0) || !(scale > 0.0) || q < 0.0 || q > 1.0) return NAN;
    return gamma_p_inv(a, q, x0, p0, q0) * scale;
}

// distribution parameters of the family (the values `report_variables` would store)
struct DistPar {
    double p0, p1, p2;        // in the family's own order (distributions.py dataclass fields)
};

// family parameters from the linear predictors: links of attrici/variables.py create_model tables
ATTRICI_HD void family_params(int family, const DayParams& d, DistPar& ref, DistPar& cf) {
    switch (family) {
        case ATTRICI_NORMAL: {  // mu (identity), sigma (exp)
            double sg = exp(d.b);
            ref.p0 = d.a_ref[0]; cf.p0 = d.a_cf[0]; ref.p1 = cf.p1 = sg; ref.p2 = cf.p2 = 0.0;
        } break;
        case ATTRICI_GAMMA: {  // mu (exp), nu (exp)
            double nu = exp(d.b);
            ref.p0 = exp(d.a_ref[0]); cf.p0 = exp(d.a_cf[0]); ref.p1 = cf.p1 = nu; ref.p2 = cf.p2 = 0.0;
        } break;
        case ATTRICI_BETA: {  // mu (invlogit), phi (exp)
            double phi = exp(d.b);
            ref.p0 = invlogit(d.a_ref[0]); cf.p0 = invlogit(d.a_cf[0]); ref.p1 = cf.p1 = phi; ref.p2 = cf.p2 = 0.0;
        } break;
        case ATTRICI_WEIBULL: {  // alpha (exp, independent), beta (exp, dependent)
            double al = exp(d.b);
            ref.p0 = cf.p0 = al; ref.p1 = exp(d.a_ref[0]); cf.p1 = exp(d.a_cf[0]); ref.p2 = cf.p2 = 0.0;
        } break;
        default: {  // bernoulli-gamma: p (invlogit), mu (exp), nu (exp)
            double nu = exp(d.b);
            ref.p0 = invlogit(d.a_ref[0]); cf.p0 = invlogit(d.a_cf[0]);
            ref.p1 = exp(d.a_ref[1]); cf.p1 = exp(d.a_cf[1]);
            ref.p2 = cf.p2 = nu;
        } break;
    }
}

// cdf under the factual parameters, inverse cdf under the counterfactual ones (distributions.py:81-216)
ATTRICI_HD double dist_cdf(int family, const DistPar& p, double y) {
    switch (family) {
        case ATTRICI_NORMAL: return ndtr((y - p.p0) / p.p1);
        case ATTRICI_GAMMA: { double a = p.p1 * p.p1; return gamma_cdf(y, a, p.p0 / a); }
        case ATTRICI_BETA: {
            if (isnan(y)) return NAN;
            double A = p.p0 * p.p1, B = (1.0 - p.p0) * p.p1;
            if (y <= 0.0) return 0.0;
            if (y >= 1.0) return 1.0;
            return beta_i(A, B, y);
        }
        case ATTRICI_WEIBULL: {
            if (isnan(y)) return NAN;
            if (y <= 0.0) return 0.0;
            return -expm1(-pow(y / p.p1, p.p0));
        }
        default: { double a = p.p2 * p.p2; return p.p0 + (1.0 - p.p0) * gamma_cdf(y, a, p.p1 / a); }
    }
}

// y_hint: a value whose quantile under these parameters is close to q (<= 0: none); it only seeds the
// iterations of the gamma / beta inverses, the result is the same root
ATTRICI_HD double dist_invcdf(int family, const DistPar& p, double q, double y_hint = -1.0) {
    switch (family) {
        case ATTRICI_NORMAL: return ndtri(q) * p.p1 + p.p0;
        case ATTRICI_GAMMA: { double a = p.p1 * p.p1; return gamma_ppf(q, a, p.p0 / a, y_hint * a / p.p0); }
        case ATTRICI_BETA: {
            double A = p.p0 * p.p1, B = (1.0 - p.p0) * p.p1;
            return beta_i_inv(A, B, q, y_hint);
        }
        case ATTRICI_WEIBULL: {
            if (isnan(q) || q < 0.0 || q > 1.0) return NAN;
            return pow(-log1p(-q), 1.0 / p.p0) * p.p1;
        }
        default: {
            double a = p.p2 * p.p2;
            return gamma_ppf((q - p.p0) / (1.0 - p.p0), a, p.p1 / a, y_hint * a / p.p1);
        }
    }
}

// Gamma family shortcut.  The shape a = nu^2 does not depend on the predictor, so factual and counterfactual
// distributions differ by their scale only and F_cf^-1(F_ref(y)) = y mu_cf / mu_ref exactly; the reference's
// gammainc -> gammaincinv round trip only adds its own rounding, which stays below ~1e-9 relative while the upper
// tail probability Q(a, x) is above ~1e-7 (1.1e-16 / Q).  Days further out in the upper tail keep the full
// evaluation, which reproduces the reference's saturation (cdf rounds to 1 -> ppf = Inf -> `replaced`).
// Leading term of ln Q for x > a + 1:  (a - 1) ln x - x - lgamma(a).
ATTRICI_HD bool gamma_fast_path(const DistPar& ref, const DistPar& cf, double ys, double* cs) {
    if (!(ys > 0.0) || !(ref.p0 > 0.0) || !(cf.p0 > 0.0) || !(ref.p1 > 0.0) || !isfinite(ref.p0) || !isfinite(cf.p0)) return false;
    const double a = ref.p1 * ref.p1;
    const double x = ys * a / ref.p0;
    if (!isfinite(x)) return false;
    if (x > a + 1.0) {
        const double ln_q = (a - 1.0) * log(x) - x - lgamma(a);
        if (!(ln_q > -16.0)) return false;
    }
    *cs = ys * (cf.p0 / ref.p0);
    return true;
}

// Weibull family shortcut.  The shape alpha does not depend on the predictor, so factual and counterfactual
// distributions differ by their scale only and F_cf^-1(F_ref(y)) = y beta_cf / beta_ref exactly (with
// r = (y / beta_ref)^alpha: q = -expm1(-r), invcdf = beta_cf (-log1p(-q))^(1/alpha) = beta_cf r^(1/alpha)).
// The reference's round trip only adds its own rounding while r stays below ~16 (1 - q = e^-r > 1e-7); beyond
// that it loses digits of 1 - q and finally saturates (q rounds to 1 -> Inf -> `replaced`), so those days keep
// the full evaluation.  r is only needed as a classifier: float32 is ample.
ATTRICI_HD bool weibull_fast_path(const DistPar& ref, const DistPar& cf, double ys, double* cs) {
    if (!(ys > 0.0) || !(ref.p0 > 0.0) || !(ref.p1 > 0.0) || !(cf.p1 > 0.0) || !(cf.p0 == ref.p0) || !isfinite(ref.p1) ||
        !isfinite(cf.p1) || !isfinite(ref.p0))
        return false;
    const float r = powf((float)(ys / ref.p1), (float)ref.p0);
    if (!(r < 16.0f)) return false;
    *cs = ys * (cf.p1 / ref.p1);
    return true;
}

// Normal family shortcut.  sigma does not depend on the predictor, so F_cf^-1(F_ref(y)) = mu_cf + sigma z
// with z = (y - mu_ref) / sigma, i.e. y + (mu_cf - mu_ref): the reference's ndtri(ndtr(z)) round trip only
// adds its own rounding, at most 1.1e-16 / pdf(z) in z (8e-13 at |z| = 4).  Days with |z| < 4 (float32
// estimate) therefore take y - gmt * trend(t) in FP64; the tails keep the full cdf -> ppf evaluation, which
// reproduces the reference's saturation to +-Inf (and hence its `replaced` flags) exactly.
// af / bf: float copies of the mu / log sigma coefficients; row32 / row64: the day's basis rows.
ATTRICI_HD bool normal_fast_path(const CellCoef& c, const float* af, const float* bf, const float* row32,
                                 const double* row64, float ys, double* cs) {
    if (ys != ys) return false;
    const float g = row32[0];
    float mu = fmaf(af[1], g, af[0]);
    float es = bf[0];
#pragma unroll
    for (int k = 0; k < MAXM; ++k) {
        const float ck = row32[1 + k], sk = row32[1 + NH + k];
        mu = fmaf(fmaf(g, af[2 + 2 * MAXM + k], af[2 + k]), ck, mu);
        mu = fmaf(fmaf(g, af[2 + 3 * MAXM + k], af[2 + MAXM + k]), sk, mu);
        es = fmaf(bf[1 + k], ck, es);
        es = fmaf(bf[1 + MAXM + k], sk, es);
    }
    const float z = (ys - mu) * expf(-es);
    if (!(fabsf(z) < 4.0f)) return false;
    double trend = c.a[0][1];
#pragma unroll
    for (int k = 0; k < MAXM; ++k) {
        trend = fma(c.a[0][2 + 2 * MAXM + k], row64[1 + k], trend);
        trend = fma(c.a[0][2 + 3 * MAXM + k], row64[1 + NH + k], trend);
    }
    *cs = (double)ys - row64[0] * trend;
    return true;
}

struct QmapOut {
    double cfact_scaled, cfact;
    int replaced;
};

// rescale (variables.py:71-89, 114-131, 483-486, 657) and replacement (detrend.py:397-411) of a mapped value
ATTRICI_HD QmapOut qmap_finish(int scaling, int post_rule, double cs, float y, double datamin, double scale) {
    QmapOut o;
    if (post_rule == ATTRICI_POST_LE0_NAN) { if (cs <= 0.0) cs = NAN; }               // Rsds, :697
    else if (post_rule == ATTRICI_POST_OUTSIDE01_NAN) { if (cs >= 1.0 || cs <= 0.0) cs = NAN; }  // Tasskew
    o.cfact_scaled = cs;
    double c;
    switch (scaling) {
        case ATTRICI_SCALE_UNITY: c = cs * scale + datamin; break;      // variables.py:131
        case ATTRICI_SCALE_NONE: c = cs * 1.0; break;                    // :657
        case ATTRICI_SCALE_PR:                                           // :483-486
            c = cs * scale + 0.0000011574;
            if (cs <= 0.0) c = 0.0;
            break;
        default: c = cs * scale; break;                                  // :89
    }
    o.replaced = ATTRICI_REPLACED_NONE;
    if (c != c) { c = (double)y; o.replaced = ATTRICI_REPLACED_NAN; }
    else if (c == INFINITY) { c = (double)y; o.replaced = ATTRICI_REPLACED_PINF; }
    else if (c == -INFINITY) { c = (double)y; o.replaced = ATTRICI_REPLACED_NINF; }
    o.cfact = c;
    return o;
}

// one day of one cell.  y: original observation (after the in-place masks of hurs / tasrange / sfcWind),
// ys: scaled observation, u: the day's MT19937 uniform (pr only)
ATTRICI_HD QmapOut qmap_day(int family, int scaling, int post_rule, float y, float ys, const DistPar& ref,
                            const DistPar& cf, double datamin, double scale, double u) {
    double cs;
    if (family != ATTRICI_BERNOULLI_GAMMA) {
        // variables.py:303: invcdf_cf(cdf_ref(y_scaled)); NaN in -> NaN out
        if (!(family == ATTRICI_GAMMA && gamma_fast_path(ref, cf, (double)ys, &cs)) &&
            !(family == ATTRICI_WEIBULL && weibull_fast_path(ref, cf, (double)ys, &cs))) {
            double q = dist_cdf(family, ref, (double)ys);
            // the counterfactual value is close to the observation: start the gamma / beta inverses there
            cs = dist_invcdf(family, cf, q, (double)ys);
        }
    } else {
        // Pr.quantile_mapping, variables.py:422-480
        const bool dry = (ys != ys);
        const double y0 = dry ? 0.0 : (double)ys;
        // factual quantile as dist_cdf forms it, keeping P and Q of the gamma part: the shape nu^2 does not depend
        // on the predictor, so x = y0 / scale_ref is also where the counterfactual inverse starts (for equal dry-day
        // probabilities it is the root), with P(a, x) already known
        const double a = ref.p2 * ref.p2, sc_ref = ref.p1 / a;
        double x = -1.0, P = -1.0, Q = -1.0, q;
        if (isnan(a) || isnan(sc_ref) || !(a > 0.0) || !(sc_ref > 0.0)) {
            q = NAN;
        } else {
            x = y0 / sc_ref;
            if (x <= 0.0) { q = ref.p0; x = -1.0; }
            else { gamma_pq_l(a, x, lgamma(a), &P, &Q); q = ref.p0 + (1.0 - ref.p0) * P; }
        }
        const bool drier_cf = cf.p0 > ref.p0;
        const bool qm0 = drier_cf && (q > cf.p0);
        const bool qm1 = !drier_cf && !dry;
        const bool to_wet = !qm1 && (u * ref.p0 > cf.p0);
        cs = 0.0;
        if (qm0 || qm1 || to_wet) {
            const double acf = cf.p2 * cf.p2;
            const bool known = !dry && acf == a && isfinite(x) && x > 0.0;
            cs = gamma_ppf((q - cf.p0) / (1.0 - cf.p0), acf, cf.p1 / acf, known ? x : (dry ? -1.0 : y0 * acf / cf.p1),
                           known ? P : -1.0, known ? Q : -1.0);
        }
    }
    return qmap_finish(scaling, post_rule, cs, y, datamin, scale);
}

// NumPy legacy MT19937 (np.random.seed(int) = init_genrand; np.random.rand -> random_sample):
// state mt[624] with stride `st` between consecutive words
struct Mt19937 {
    static constexpr int N = 624, M = 397;
    template <typename Acc> ATTRICI_HD static void seed(Acc mt, uint32_t s) {
        mt(0) = s;
        for (int i = 1; i < N; ++i) {
            uint32_t p = mt(i - 1);
            mt(i) = 1812433253u * (p ^ (p >> 30)) + (uint32_t)i;
        }
    }
    template <typename Acc> ATTRICI_HD static void twist(Acc mt) {
        int kk;
        for (kk = 0; kk < N - M; ++kk) {
            uint32_t y = (mt(kk) & 0x80000000u) | (mt(kk + 1) & 0x7fffffffu);
            mt(kk) = mt(kk + M) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        for (; kk < N - 1; ++kk) {
            uint32_t y = (mt(kk) & 0x80000000u) | (mt(kk + 1) & 0x7fffffffu);
            mt(kk) = mt(kk + (M - N)) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        uint32_t y = (mt(N - 1) & 0x80000000u) | (mt(0) & 0x7fffffffu);
        mt(N - 1) = mt(M - 1) ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
    }
    ATTRICI_HD static uint32_t temper(uint32_t y) {
        y ^= (y >> 11);
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= (y >> 18);
        return y;
    }
    // random_sample: (a >> 5, b >> 6) -> (a * 2^26 + b) / 2^53
    ATTRICI_HD static double to_double(uint32_t a, uint32_t b) {
        return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) / 9007199254740992.0;
    }
};

}  // namespace attrici
