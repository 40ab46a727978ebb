// Per-day negative log-likelihood of each distribution, its derivatives with respect to the two
// linear predictors and the Fisher-information weights used by the scoring iteration.
//
// eta_a: linear predictor of the GMT-dependent parameter, eta_b: of the independent one.
// The likelihoods are scipy's, as called by the reference at
// attrici/estimation/model_scipy.py:352 (beta), :373 (gamma), :442 (bernoulli), :468 (norm),
// :494 (weibull_min); links per attrici/variables.py (identity / exp / invlogit), applied to BOTH
// kinds of parameter as attrici/estimation/model_pymc5.py:253-256 does (DESIGN.md "link correction").
//
//   r*  = d nll / d eta        w** = E[d2 nll / d eta d eta]  (expected, hence positive semi-definite)
#pragma once
#include "special.cuh"

namespace attrici {

enum : int { K_NORMAL = 0, K_GAMMA = 1, K_BETA = 2, K_WEIBULL = 3, K_BERNOULLI = 4 };

template <typename R> struct DayTerms {
    R nll, rA, rB, wAA, wAB, wBB;
};

template <typename R> ATTRICI_HD R softplus(R x) {
    // log(1 + e^x) without overflow
    R ax = fabs(x);
    return (x > 0 ? x : R(0)) + log1p(exp(-ax));
}

template <typename R> ATTRICI_HD R sigmoid(R x) { return R(1) / (R(1) + exp(-x)); }

template <int KIND, typename R> ATTRICI_HD DayTerms<R> family_day_t(R y, R ea, R eb) {
    DayTerms<R> d;
    if (KIND == K_NORMAL) {
        // mu = ea, sigma = exp(eb); scipy norm._logpdf(z) - ln sigma
        R e = exp(-eb);
        R z = (y - ea) * e;
        d.nll = R(0.5) * z * z + eb + R(0.91893853320467274178);
        d.rA = -z * e;
        d.rB = R(1) - z * z;
        d.wAA = e * e;
        d.wAB = R(0);
        d.wBB = R(2);
    } else if (KIND == K_GAMMA) {
        // mu = exp(ea), a = nu^2 = exp(2 eb); scipy gamma(a, scale = mu / a)
        R a = exp(R(2) * eb);
        R ly = log(y);
        R lnu = ly - ea;
        R u = exp(lnu);
        R lna = R(2) * eb;
        R dd = lnu + lna + R(1) - u - digamma<R>(a);
        d.nll = -((a - R(1)) * ly - a * u - lgam<R>(a) - a * ea + a * lna);
        d.rA = -a * (u - R(1));
        d.rB = R(-2) * a * dd;
        d.wAA = a;
        d.wAB = R(0);
        d.wBB = R(4) * a * (a * trigamma<R>(a) - R(1));
    } else if (KIND == K_WEIBULL) {
        // c = alpha = exp(eb) (independent), scale beta = exp(ea) (dependent); scipy weibull_min(c, scale)
        R c = exp(eb);
        R L = log(y) - ea;
        R r = exp(c * L);
        d.nll = -(eb + (c - R(1)) * L - r - ea);
        d.rA = -c * (r - R(1));
        d.rB = -(R(1) + c * L * (R(1) - r));
        d.wAA = c * c;
        d.wAB = -c * R(0.42278433509846713939);   // -(1 - EulerGamma) c
        d.wBB = R(1.82368066085287938958);        // pi^2/6 + (1 - EulerGamma)^2
    } else if (KIND == K_BETA) {
        // mu = sigmoid(ea), phi = exp(eb); scipy beta(mu phi, (1 - mu) phi)
        R mu = sigmoid<R>(ea);
        R phi = exp(eb);
        R A = mu * phi, B = (R(1) - mu) * phi;
        R ly = log(y), l1y = log1p(-y);
        // digamma and trigamma of each argument from one shared recurrence
        R psiA, tA, psiB, tB, psiP, tP;
        psi_tri<R>(A, psiA, tA);
        psi_tri<R>(B, psiB, tB);
        psi_tri<R>(phi, psiP, tP);
        const R lgA = lgam<R>(A), lgB = lgam<R>(B), lgP = lgam<R>(phi);
        R m1 = mu * (R(1) - mu);
        R delta = ly - l1y - psiA + psiB;
        d.nll = -((A - R(1)) * ly + (B - R(1)) * l1y - lgA - lgB + lgP);
        d.rA = -phi * m1 * delta;
        d.rB = -(A * (ly - psiA) + B * (l1y - psiB) + phi * psiP);
        R pm = phi * m1;
        d.wAA = pm * pm * (tA + tB);
        d.wAB = pm * (A * tA - B * tB);
        d.wBB = A * A * tA + B * B * tB - phi * phi * tP;
    } else {
        // bernoulli: y = 1 on a dry day, p = sigmoid(ea) = P(dry); scipy bernoulli.logpmf(k, p)
        R p = sigmoid<R>(ea);
        d.nll = y * softplus<R>(-ea) + (R(1) - y) * softplus<R>(ea);
        d.rA = p - y;
        d.rB = R(0);
        d.wAA = p * (R(1) - p);
        d.wAB = R(0);
        d.wBB = R(0);
    }
    return d;
}

template <typename R> ATTRICI_HD_NOINLINE DayTerms<R> family_day(int kind, R y, R ea, R eb) {
    switch (kind) {
        case K_NORMAL: return family_day_t<K_NORMAL, R>(y, ea, eb);
        case K_GAMMA: return family_day_t<K_GAMMA, R>(y, ea, eb);
        case K_BETA: return family_day_t<K_BETA, R>(y, ea, eb);
        case K_WEIBULL: return family_day_t<K_WEIBULL, R>(y, ea, eb);
        default: return family_day_t<K_BERNOULLI, R>(y, ea, eb);
    }
}

}  // namespace attrici
