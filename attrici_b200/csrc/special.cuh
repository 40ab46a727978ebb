// Special functions needed by the likelihood and the quantile map, written for the device
// (CUDA has no regularised incomplete gamma / beta, their inverses, digamma or trigamma).
// Every function is __host__ __device__ so that tests/hostmath can compile this header with g++
// and compare it against scipy.special on the CPU (a unit test of the math, not a product path).
//
// What the reference calls for these (through scipy.stats, attrici/distributions.py:81-216 and
// attrici/estimation/model_scipy.py:352,373,468,494):
//   norm.cdf / ppf        -> ndtr, ndtri          -> ndtr(), ndtri()
//   gamma.cdf / ppf       -> gammainc, gammaincinv-> gamma_p(), gamma_p_inv()
//   beta.cdf / ppf        -> betainc, betaincinv  -> beta_i(), beta_i_inv()
//   gamma/beta.logpdf     -> gammaln              -> lgam()   (+ digamma/trigamma for gradients)
#pragma once
#include <math.h>
#include <float.h>

#if defined(__CUDACC__)
#define ATTRICI_HD __host__ __device__ __forceinline__
#define ATTRICI_HD_NOINLINE inline __host__ __device__
#else
#define ATTRICI_HD inline
#define ATTRICI_HD_NOINLINE inline
#endif

namespace attrici {

// correctly rounded reciprocal without the special-case handling of an IEEE division (one rounding more than
// x / y when used as x * rcp(y); the series and continued fractions below run to 1e-16 regardless)
ATTRICI_HD double rcp(double x) {
#if defined(__CUDA_ARCH__)
    return __drcp_rn(x);
#else
    return 1.0 / x;
#endif
}

template <typename R> ATTRICI_HD R lgam(R x);
template <> ATTRICI_HD double lgam<double>(double x) { return lgamma(x); }
template <> ATTRICI_HD float lgam<float>(float x) { return lgammaf(x); }

// digamma for x > 0: upward recurrence to x >= X0, then the asymptotic series
template <typename R> ATTRICI_HD R digamma(R x) {
    const R X0 = (sizeof(R) == 8) ? R(10) : R(6);
    R acc = 0;
    while (x < X0) {
        acc -= R(1) / x;
        x += R(1);
    }
    R inv = R(1) / x, inv2 = inv * inv;
    R s = inv2 * (R(1.0 / 12) - inv2 * (R(1.0 / 120) - inv2 * (R(1.0 / 252) - inv2 * (R(1.0 / 240) - inv2 * (R(1.0 / 132) - inv2 * R(691.0 / 32760))))));
    return acc + log(x) - R(0.5) * inv - s;
}

template <typename R> ATTRICI_HD R trigamma(R x) {
    const R X0 = (sizeof(R) == 8) ? R(10) : R(6);
    R acc = 0;
    while (x < X0) {
        R i = R(1) / x;
        acc += i * i;
        x += R(1);
    }
    R inv = R(1) / x, inv2 = inv * inv;
    R s = inv * (R(1) + inv * (R(0.5) + inv * (R(1.0 / 6) - inv2 * (R(1.0 / 30) - inv2 * (R(1.0 / 42) - inv2 * (R(1.0 / 30) - inv2 * (R(5.0 / 66) - inv2 * R(691.0 / 2730))))))));
    return acc + s;
}

// digamma and trigamma of one argument x > 0 together: one upward recurrence to X >= X0 shared by both (sum of
// 1 / (x + k) for digamma, of its square for trigamma), then their asymptotic series on the shared powers of 1 / X.
// The Beta likelihood needs both at A, B and phi every day.  (lgamma stays with the library function: a float32
// Stirling form loses too much of the nll to cancellation.)
template <typename R> ATTRICI_HD void psi_tri(R x, R& psi, R& tri) {
    const R X0 = (sizeof(R) == 8) ? R(10) : R(6);
    R s1 = R(0), s2 = R(0);
    while (x < X0) {
        const R i = R(1) / x;
        s1 += i;
        s2 = fma(i, i, s2);
        x += R(1);
    }
    const R inv = R(1) / x, inv2 = inv * inv;
    const R ps = inv2 * (R(1.0 / 12) - inv2 * (R(1.0 / 120) - inv2 * (R(1.0 / 252) - inv2 * (R(1.0 / 240) - inv2 * (R(1.0 / 132) - inv2 * R(691.0 / 32760))))));
    psi = log(x) - R(0.5) * inv - ps - s1;
    const R ts = inv * (R(1) + inv * (R(0.5) + inv * (R(1.0 / 6) - inv2 * (R(1.0 / 30) - inv2 * (R(1.0 / 42) - inv2 * (R(1.0 / 30) - inv2 * (R(5.0 / 66) - inv2 * R(691.0 / 2730))))))));
    tri = ts + s2;
}

// ---------------------------------------------------------------------------------------
// normal distribution
// ---------------------------------------------------------------------------------------
// same branch structure as cephes ndtr (what scipy.special.ndtr evaluates)
ATTRICI_HD double ndtr(double a) {
    if (isnan(a)) return a;
    double x = a * 0.70710678118654752440;
    double z = fabs(x);
    if (z < 0.70710678118654752440) return 0.5 + 0.5 * erf(x);
    double y = 0.5 * erfc(z);
    return x > 0 ? 1.0 - y : y;
}

// inverse normal cdf: Wichura's AS 241 (PPND16), relative error ~1e-16
ATTRICI_HD_NOINLINE double ndtri(double p) {
    if (isnan(p) || p < 0.0 || p > 1.0) return NAN;
    if (p == 0.0) return -INFINITY;
    if (p == 1.0) return INFINITY;
    double q = p - 0.5, r, val;
    if (fabs(q) <= 0.425) {
        r = 0.180625 - q * q;
        val = q * (((((((2.5090809287301226727e3 * r + 3.3430575583588128105e4) * r + 6.7265770927008700853e4) * r + 4.5921953931549871457e4) * r + 1.3731693765509461125e4) * r + 1.9715909503065514427e3) * r + 1.3314166789178437745e2) * r + 3.3871328727963666080e0) /
              (((((((5.2264952788528545610e3 * r + 2.8729085735721942674e4) * r + 3.9307895800092710610e4) * r + 2.1213794301586595867e4) * r + 5.3941960214247511077e3) * r + 6.8718700749205790830e2) * r + 4.2313330701600911252e1) * r + 1.0);
        return val;
    }
    r = q < 0 ? p : 1.0 - p;
    r = sqrt(-log(r));
    if (r <= 5.0) {
        r -= 1.6;
        val = (((((((7.74545014278341407640e-4 * r + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r + 1.27045825245236838258e0) * r + 3.64784832476320460504e0) * r + 5.76949722146069140550e0) * r + 4.63033784615654529590e0) * r + 1.42343711074968357734e0) /
              (((((((1.05075007164441684324e-9 * r + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r + 1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r + 1.67638483018380384940e0) * r + 2.05319162663775882187e0) * r + 1.0);
    } else {
        r -= 5.0;
        val = (((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r + 2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r + 1.78482653991729133580e0) * r + 5.46378491116411436990e0) * r + 6.65790464350110377720e0) /
              (((((((2.04426310338993978564e-15 * r + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r + 7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r + 1.36929880922735805310e-1) * r + 5.99832206555887937690e-1) * r + 1.0);
    }
    return q < 0 ? -val : val;
}

// ---------------------------------------------------------------------------------------
// regularised incomplete gamma P(a, x), Q(a, x) and the inverse of P
// ---------------------------------------------------------------------------------------
// x^a e^-x / Gamma(a)
// gln = lgamma(a), hoisted by callers that evaluate several times at the same shape
ATTRICI_HD double gamma_prefactor(double a, double x, double gln) { return exp(a * log(x) - x - gln); }

// series for P without its prefactor: sum_n x^n / (a (a + 1) .. (a + n)), converges quickly for x < a + 1.
// Four terms per reciprocal - t_{n+k} = t_n x^k / ((a + n + 1) .. (a + n + k)), all four over the common
// R = 1 / ((a+n+1)(a+n+2)(a+n+3)(a+n+4)) - and the stopping rule checked every fourth term (it asks for terms below
// 1e-17 of the sum, beyond the precision of the sum, so the extra terms change nothing): ~8 instructions per term
// instead of ~20 with a reciprocal per term.
ATTRICI_HD double gamma_p_series_sum(double a, double x) {
    double ap = a, del = 1.0 / a, sum = del;
    const double x2 = x * x, x3 = x2 * x, x4 = x2 * x2;
    for (int n = 0; n < 500; ++n) {
        const double d1 = ap + 1.0, d2 = ap + 2.0, d3 = ap + 3.0, d4 = ap + 4.0;
        const double p12 = d1 * d2, p34 = d3 * d4;
        const double dr = del * rcp(p12 * p34);
        const double t1 = dr * x * (d2 * p34), t2 = dr * x2 * p34, t3 = dr * x3 * d4, t4 = dr * x4;
        sum += (t1 + t2) + (t3 + t4);
        del = t4;
        ap = d4;
        if (fabs(t4) < fabs(sum) * 1e-17) break;
    }
    return sum;
}

ATTRICI_HD_NOINLINE double gamma_p_series(double a, double x, double gln) {
    return gamma_p_series_sum(a, x) * gamma_prefactor(a, x, gln);
}

// modified Lentz continued fraction for Q, converges quickly for x > a + 1
ATTRICI_HD_NOINLINE double gamma_q_cf(double a, double x, double gln) {
    const double tiny = 1e-300;
    double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
    for (int i = 1; i < 2000; ++i) {
        double an = -i * (i - a);
        b += 2.0;
        d = an * d + b;
        if (fabs(d) < tiny) d = tiny;
        c = b + an * rcp(c);
        if (fabs(c) < tiny) c = tiny;
        d = rcp(d);
        double del = d * c;
        h *= del;
        if (fabs(del - 1.0) < 1e-16) break;
    }
    return h * gamma_prefactor(a, x, gln);
}

// Switch between the series for P and the continued fraction for Q.  The textbook switch is x = a + 1; the series
// still converges quickly (about x + 6 sqrt(x) + 20 positive terms, no cancellation) far beyond it, and a warp
// whose lanes fall on both sides of the switch pays for both branches, so the series is kept up to x = 30: for the
// shapes of pr / tasrange (a ~ 0.3 .. 3) practically every lane then takes the same branch.
ATTRICI_HD bool gamma_use_series(double a, double x) { return x < a + 1.0 || x < 30.0; }

// P and Q with lgamma(a) supplied
ATTRICI_HD double gamma_p_l(double a, double x, double gln) {
    if (isnan(a) || isnan(x) || a <= 0.0 || x < 0.0) return NAN;
    if (x == 0.0) return 0.0;
    if (isinf(x)) return 1.0;
    if (gamma_use_series(a, x)) return gamma_p_series(a, x, gln);
    return 1.0 - gamma_q_cf(a, x, gln);
}

ATTRICI_HD double gamma_q_l(double a, double x, double gln) {
    if (isnan(a) || isnan(x) || a <= 0.0 || x < 0.0) return NAN;
    if (x == 0.0) return 1.0;
    if (isinf(x)) return 0.0;
    if (gamma_use_series(a, x)) {
        // 1 - P keeps full relative accuracy only while Q is not small; the far upper tail (rare) takes the
        // continued fraction, which computes Q itself
        const double p = gamma_p_series(a, x, gln);
        if (p < 0.99 || x < a + 1.0) return 1.0 - p;
    }
    return gamma_q_cf(a, x, gln);
}

// P and Q from one series evaluation (one call site: the lanes of a warp do not diverge on which of the two is
// wanted); the far upper tail takes the continued fraction for Q
ATTRICI_HD void gamma_pq_l(double a, double x, double gln, double* p, double* q) {
    if (isnan(a) || isnan(x) || a <= 0.0 || x < 0.0) { *p = NAN; *q = NAN; return; }
    if (x == 0.0) { *p = 0.0; *q = 1.0; return; }
    if (isinf(x)) { *p = 1.0; *q = 0.0; return; }
    if (gamma_use_series(a, x)) {
        const double pp = gamma_p_series(a, x, gln);
        if (pp < 0.99 || x < a + 1.0) { *p = pp; *q = 1.0 - pp; return; }
    }
    *q = gamma_q_cf(a, x, gln);
    *p = 1.0 - *q;
}

ATTRICI_HD double gamma_p(double a, double x) { return gamma_p_l(a, x, lgamma(a)); }
ATTRICI_HD double gamma_q(double a, double x) { return gamma_q_l(a, x, lgamma(a)); }

// x with P(a, x) = p   (scipy.special.gammaincinv); safeguarded Halley iteration
// x0 > 0: starting value supplied by the caller (the quantile map knows a point whose quantile is close)
// p0 >= 0: P(a, x0) and q0 = Q(a, x0), already known to the caller - the first iteration then costs no evaluation of
// the incomplete gamma function (pr: the factual quantile was just computed at that very point)
ATTRICI_HD_NOINLINE double gamma_p_inv(double a, double p, double x0 = -1.0, double p0 = -1.0, double q0 = -1.0) {
    if (isnan(a) || isnan(p) || a <= 0.0 || p < 0.0 || p > 1.0) return NAN;
    if (p == 0.0) return 0.0;
    if (p == 1.0) return INFINITY;
    const double gln = lgamma(a);
    const double a1 = a - 1.0;
    double x;
    // starting value (Abramowitz & Stegun 26.2.22 / 26.4.17, as in Numerical Recipes invgammp)
    if (x0 > 0.0 && isfinite(x0)) {
        x = x0;
    } else if (a > 1.0) {
        double pp = p < 0.5 ? p : 1.0 - p;
        double t = sqrt(-2.0 * log(pp));
        x = (2.30753 + t * 0.27061) / (1.0 + t * (0.99229 + t * 0.04481)) - t;
        if (p < 0.5) x = -x;
        double w = 1.0 - 1.0 / (9.0 * a) - x / (3.0 * sqrt(a));
        x = a * w * w * w;
        if (x < 1e-3) x = 1e-3;
    } else {
        double t = 1.0 - a * (0.253 + a * 0.12);
        if (p < t)
            x = pow(p / t, 1.0 / a);
        else
            x = 1.0 - log(1.0 - (p - t) / (1.0 - t));
    }
    double lo = 0.0, hi = INFINITY;
    const double lna1 = a > 1.0 ? log(a1) : 0.0;
    const double afac = a > 1.0 ? exp(a1 * (lna1 - 1.0) - gln) : 0.0;
    for (int it = 0; it < 60; ++it) {
        if (x <= 0.0) x = 0.5 * (lo + (isinf(hi) ? lo + 1e-300 : hi));
        // error in the better-conditioned tail
        double pv, qv;
        if (it == 0 && p0 >= 0.0 && x == x0) { pv = p0; qv = q0; }
        else gamma_pq_l(a, x, gln, &pv, &qv);
        double err = (p < 0.5) ? pv - p : (1.0 - p) - qv;
        if (err > 0.0) hi = x; else lo = x;
        double dens;  // dP/dx
        if (a > 1.0)
            dens = afac * exp(-(x - a1) + a1 * (log(x) - lna1));
        else
            dens = exp(-x + a1 * log(x) - gln);
        if (!(dens > 0.0)) {
            x = isinf(hi) ? 2.0 * x + 1.0 : 0.5 * (lo + hi);
            continue;
        }
        double u = err / dens;
        // Halley correction, limited as in Numerical Recipes
        double corr = u * (a1 / x - 1.0);
        if (corr > 1.0) corr = 1.0;
        double dx = u / (1.0 - 0.5 * corr);
        double xn = x - dx;
        // Halley converges cubically (error after the step ~ (a1 / x - 1)^2 / 12 ... times the cube of the step, in
        // relative terms a small multiple of (dx / x)^3): a step below 2e-5 x leaves an error below 1e-13 x, the
        // level of the series itself, so the point after the step is taken without evaluating there again.  Only a
        // plain Halley step from inside the bracket qualifies.
        if (fabs(xn - x) <= 2e-5 * x && xn >= lo && xn <= hi && corr <= 0.5 && corr >= -0.5) { x = xn; break; }
        if (!(xn >= lo) || !(xn <= hi)) xn = isinf(hi) ? 2.0 * x : 0.5 * (lo + hi);
        x = xn;
    }
    return x;
}

// gamma_pq_l with the series inlined (the cost of an item of the precipitation quantile map is that loop); also returns the prefactor x^a e^-x / Gamma(a), from which the
// inverse takes the density (x^(a-1) e^-x / Gamma(a) = prefactor / x) instead of a second exp / log pair
ATTRICI_HD void gamma_pq_pref_l(double a, double x, double gln, double* p, double* q, double* pref) {
    *pref = 0.0;
    if (isnan(a) || isnan(x) || a <= 0.0 || x < 0.0) { *p = NAN; *q = NAN; return; }
    if (x == 0.0) { *p = 0.0; *q = 1.0; return; }
    if (isinf(x)) { *p = 1.0; *q = 0.0; return; }
    *pref = gamma_prefactor(a, x, gln);
    if (gamma_use_series(a, x)) {
        const double pp = gamma_p_series_sum(a, x) * *pref;
        if (pp < 0.99 || x < a + 1.0) { *p = pp; *q = 1.0 - pp; return; }
    }
    *q = gamma_q_cf(a, x, gln);   // far upper tail: the continued fraction computes Q itself (rare)
    *p = 1.0 - *q;
}

// gamma_p_inv for the precipitation quantile map (qmap_pr.cu), with lgamma(a) and - when the caller has them - the
// values at the starting point supplied.  Two deviations, both inside the tolerance of the
// path (1e-4 relative on the counterfactual; the tests ask 1e-5 of the oracle):
//  * the density is prefactor / x (one rounding away from the closed form the scalar routine evaluates),
//  * a Halley step below 3e-4 x ends the iteration (cubic convergence: the error after it is a small multiple of
//    (3e-4)^3 = 2.7e-11 relative), so the typical day costs ONE evaluation after the factual quantile and a day whose
//    quantile moves a lot (drizzle: P ~ x^a) two; the scalar routine's 2e-5 asked for one more each.
ATTRICI_HD double gamma_p_inv_from(double a, double p, double x0, double p0, double q0, double pref0, double gln) {
    if (isnan(a) || isnan(p) || a <= 0.0 || p < 0.0 || p > 1.0) return NAN;
    if (p == 0.0) return 0.0;
    if (p == 1.0) return INFINITY;
    const double a1 = a - 1.0;
    double x;
    if (x0 > 0.0 && isfinite(x0)) {
        x = x0;
    } else if (a > 1.0) {
        double pp = p < 0.5 ? p : 1.0 - p;
        double t = sqrt(-2.0 * log(pp));
        x = (2.30753 + t * 0.27061) / (1.0 + t * (0.99229 + t * 0.04481)) - t;
        if (p < 0.5) x = -x;
        double w = 1.0 - 1.0 / (9.0 * a) - x / (3.0 * sqrt(a));
        x = a * w * w * w;
        if (x < 1e-3) x = 1e-3;
    } else {
        double t = 1.0 - a * (0.253 + a * 0.12);
        if (p < t) x = pow(p / t, 1.0 / a);
        else x = 1.0 - log(1.0 - (p - t) / (1.0 - t));
    }
    double lo = 0.0, hi = INFINITY;
    for (int it = 0; it < 60; ++it) {
        if (x <= 0.0) x = 0.5 * (lo + (isinf(hi) ? lo + 1e-300 : hi));
        double pv, qv, pref;
        if (it == 0 && p0 >= 0.0 && x == x0) { pv = p0; qv = q0; pref = pref0; }
        else gamma_pq_pref_l(a, x, gln, &pv, &qv, &pref);
        const double err = (p < 0.5) ? pv - p : (1.0 - p) - qv;
        if (err > 0.0) hi = x; else lo = x;
        const double dens = pref / x;   // dP/dx
        if (!(dens > 0.0)) {
            x = isinf(hi) ? 2.0 * x + 1.0 : 0.5 * (lo + hi);
            continue;
        }
        const double u = err / dens;
        double corr = u * (a1 / x - 1.0);
        if (corr > 1.0) corr = 1.0;
        const double dx = u / (1.0 - 0.5 * corr);
        double xn = x - dx;
        if (fabs(xn - x) <= 3e-4 * x && xn >= lo && xn <= hi && corr <= 0.5 && corr >= -0.5) { x = xn; break; }
        if (!(xn >= lo) || !(xn <= hi)) xn = isinf(hi) ? 2.0 * x : 0.5 * (lo + hi);
        x = xn;
    }
    return x;
}

// ---------------------------------------------------------------------------------------
// regularised incomplete beta I_x(a, b) and its inverse
// ---------------------------------------------------------------------------------------
// Continued fraction of the incomplete beta function (the h of Numerical Recipes' betacf), evaluated by the forward
// recurrence of its convergents A_n / B_n (A_n = A_{n-1} + d_n A_{n-2}, the same for B) instead of the modified Lentz
// form: one reciprocal per double step - both partial numerators share 1 / ((a + 2m - 1)(a + 2m)(a + 2m + 1)) - where
// Lentz needs five, and the convergence test |A_n B_{n-2} - A_{n-2} B_n| < eps |A_n B_{n-2}| needs none.  The
// convergents are rescaled when they leave [2^-200, 2^200].  The quantile map of hurs spends its time here.
ATTRICI_HD_NOINLINE double beta_cf(double a, double b, double x) {
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    // convergents after the first partial numerator d_1 = -(a + b) x / (a + 1):  1 / (1 + d_1)
    double am = 1.0, bm = 1.0, az = 1.0, bz = 1.0 - qab * x / qap;
    for (int m = 1; m < 3000; ++m) {
        const double em = (double)m, t2 = em + em;
        const double p1 = qam + t2, p2 = a + t2, p3 = qap + t2;
        const double r = rcp(p1 * p2 * p3);
        const double de = em * (b - em) * x * (p3 * r);               // m (b - m) x / ((a + 2m - 1)(a + 2m))
        const double dd = -(a + em) * (qab + em) * x * (p1 * r);      // -(a + m)(a + b + m) x / ((a + 2m)(a + 2m + 1))
        const double ap = az + de * am, bp = bz + de * bm;
        const double app = ap + dd * az, bpp = bp + dd * bz;
        const double cross = app * bz;
        const bool done = fabs(az * bpp - cross) < 1e-16 * fabs(cross);
        am = ap; bm = bp; az = app; bz = bpp;
        if (done) break;
        const double mag = fabs(bz);
        if (mag > 1.6e60 || mag < 6.2e-61) {   // 2^+-200
            const double sc = mag > 1.0 ? 6.223015277861142e-61 : 1.6069380442589903e60;
            am *= sc; bm *= sc; az *= sc; bz *= sc;
        }
    }
    return az / bz;
}

// returns I_x(a,b); if comp != nullptr also 1 - I_x(a,b) computed without cancellation.
// lnb = lgamma(a + b) - lgamma(a) - lgamma(b), hoisted by callers that evaluate several times at the same (a, b)
// pref (optional): x^a (1 - x)^b / B(a, b), from which the inverse takes the density (pref / (x (1 - x)))
ATTRICI_HD_NOINLINE double beta_i2_l(double a, double b, double x, double lnb, double* comp, double* pref = nullptr) {
    if (pref) *pref = 0.0;
    if (isnan(a) || isnan(b) || isnan(x) || a <= 0.0 || b <= 0.0 || x < 0.0 || x > 1.0) {
        if (comp) *comp = NAN;
        return NAN;
    }
    if (x == 0.0) { if (comp) *comp = 1.0; return 0.0; }
    if (x == 1.0) { if (comp) *comp = 0.0; return 1.0; }
    double bt = exp(lnb + a * log(x) + b * log1p(-x));
    if (pref) *pref = bt;
    // one continued-fraction call with the arguments swapped where needed: lanes on either side of the switch
    // then run the same code instead of two divergent calls
    const bool flip = !(x < (a + 1.0) / (a + b + 2.0));
    const double fa = flip ? b : a, fb = flip ? a : b, fx = flip ? 1.0 - x : x;
    const double v = bt * beta_cf(fa, fb, fx) / fa;
    const double p = flip ? 1.0 - v : v, q = flip ? v : 1.0 - v;
    if (comp) *comp = q;
    return p;
}

ATTRICI_HD double beta_i2(double a, double b, double x, double* comp) {
    return beta_i2_l(a, b, x, lgamma(a + b) - lgamma(a) - lgamma(b), comp);
}

ATTRICI_HD double beta_i(double a, double b, double x) { return beta_i2(a, b, x, nullptr); }

// x with I_x(a, b) = p   (scipy.special.betaincinv); safeguarded Halley iteration
// x0 in (0, 1): starting value supplied by the caller
ATTRICI_HD_NOINLINE double beta_i_inv(double a, double b, double p, double x0 = -1.0) {
    if (isnan(a) || isnan(b) || isnan(p) || a <= 0.0 || b <= 0.0 || p < 0.0 || p > 1.0) return NAN;
    if (p == 0.0) return 0.0;
    if (p == 1.0) return 1.0;
    const double a1 = a - 1.0, b1 = b - 1.0;
    double x;
    if (x0 > 0.0 && x0 < 1.0) {
        x = x0;
    } else if (a >= 1.0 && b >= 1.0) {
        double pp = p < 0.5 ? p : 1.0 - p;
        double t = sqrt(-2.0 * log(pp));
        x = (2.30753 + t * 0.27061) / (1.0 + t * (0.99229 + t * 0.04481)) - t;
        if (p < 0.5) x = -x;
        double al = (x * x - 3.0) / 6.0;
        double h = 2.0 / (1.0 / (2.0 * a - 1.0) + 1.0 / (2.0 * b - 1.0));
        double w = x * sqrt(al + h) / h - (1.0 / (2.0 * b - 1.0) - 1.0 / (2.0 * a - 1.0)) * (al + 5.0 / 6.0 - 2.0 / (3.0 * h));
        x = a / (a + b * exp(2.0 * w));
    } else {
        double lna = log(a / (a + b)), lnb = log(b / (a + b));
        double t = exp(a * lna) / a, u = exp(b * lnb) / b, w = t + u;
        if (p < t / w)
            x = pow(a * w * p, 1.0 / a);
        else
            x = 1.0 - pow(b * w * (1.0 - p), 1.0 / b);
    }
    const double afac = -lgamma(a) - lgamma(b) + lgamma(a + b);
    double lo = 0.0, hi = 1.0;
    for (int it = 0; it < 80; ++it) {
        if (!(x > 0.0) || !(x < 1.0)) x = 0.5 * (lo + hi);
        double q, pref;
        double pv = beta_i2_l(a, b, x, afac, &q, &pref);
        double err = (p < 0.5) ? pv - p : (1.0 - p) - q;
        if (err > 0.0) hi = x; else lo = x;
        // dI/dx = x^(a-1) (1 - x)^(b-1) / B(a, b): the prefactor of the evaluation just made, over x (1 - x)
        double dens = pref / (x * (1.0 - x));
        double xn;
        if (!(dens > 0.0) || isinf(dens)) {
            xn = 0.5 * (lo + hi);
        } else {
            double u = err / dens;
            double corr = u * (a1 / x - b1 / (1.0 - x));
            if (corr > 1.0) corr = 1.0;
            xn = x - u / (1.0 - 0.5 * corr);
            // Halley converges cubically, in units of the distance to the nearer end of (0, 1): a plain step from inside
            // the bracket below 1e-6 of that distance leaves an error far below one ulp, so the point after the step is
            // taken without evaluating there again (waiting for a step below 1e-10 asked for one evaluation more)
            const double edge = fmin(x, 1.0 - x);
            if (fabs(xn - x) <= 1e-6 * edge && xn >= lo && xn <= hi && corr <= 0.5 && corr >= -0.5) { x = xn; break; }
            if (fabs(xn - x) <= 1e-10 * x) { x = xn; break; }
            if (!(xn >= lo) || !(xn <= hi)) xn = 0.5 * (lo + hi);
        }
        x = xn;
    }
    return x;
}

}  // namespace attrici
