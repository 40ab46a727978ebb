// Per-day body of the likelihood pass (see pass.cu).  __host__ __device__ so that tests/hostmath
// can run the identical arithmetic for one cell on the CPU against the oracle.
#pragma once
#include "layout.cuh"
#include "families.cuh"

namespace attrici {

template <typename R> struct Row {
    R g;        // gmt
    R c[NH];    // cos k
    R s[NH];    // sin k
};

template <typename R> ATTRICI_HD Row<R> load_row_scalar(const R* p) {
    Row<R> r;
    r.g = p[0];
#pragma unroll
    for (int k = 0; k < NH; ++k) { r.c[k] = p[1 + k]; r.s[k] = p[1 + NH + k]; }
    return r;
}

template <typename R> struct PassAcc {
    R AA[3 * NTRIG], AB[2 * NTRIG], BB[NTRIG], GA[2 * NB], GB[NB];
    ATTRICI_HD void clear() {
#pragma unroll
        for (int i = 0; i < 3 * NTRIG; ++i) AA[i] = R(0);
#pragma unroll
        for (int i = 0; i < 2 * NTRIG; ++i) AB[i] = R(0);
#pragma unroll
        for (int i = 0; i < NTRIG; ++i) BB[i] = R(0);
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) GA[i] = R(0);
#pragma unroll
        for (int i = 0; i < NB; ++i) GB[i] = R(0);
    }
};

// acc[0] += w, acc[1 + k] += w cos_{k+1}, acc[1 + K + k] += w sin_{k+1}  for k = 0..K-1
template <int K, typename R> ATTRICI_HD void accumulate(R* acc, R w, const Row<R>& b) {
    acc[0] += w;
#pragma unroll
    for (int k = 0; k < K; ++k) {
        acc[1 + k] = fma(w, b.c[k], acc[1 + k]);
        acc[1 + K + k] = fma(w, b.s[k], acc[1 + K + k]);
    }
}

// linear predictors of the two parameters at one day (shared frame, canonical coefficient layout)
template <bool HAS_B, typename R>
ATTRICI_HD void linear_predictors(const R* thA, const R* thB, const Row<R>& b, R& etaA, R& etaB) {
    etaA = fma(thA[1], b.g, thA[0]);
    etaB = HAS_B ? thB[0] : R(0);
#pragma unroll
    for (int k = 0; k < MAXM; ++k) {
        R ck = fma(b.g, thA[2 + 2 * MAXM + k], thA[2 + k]);
        R sk = fma(b.g, thA[2 + 3 * MAXM + k], thA[2 + MAXM + k]);
        etaA = fma(ck, b.c[k], etaA);
        etaA = fma(sk, b.s[k], etaA);
        if (HAS_B) {
            etaB = fma(thB[1 + k], b.c[k], etaB);
            etaB = fma(thB[1 + MAXM + k], b.s[k], etaB);
        }
    }
}

// one day of one cell: returns the day's nll contribution (0 on unused days)
template <int KIND, typename R, bool MOMENTS>
ATTRICI_HD R pass_day(float yf, bool cell_active, const R* thA, const R* thB, const Row<R>& b, PassAcc<R>& acc) {
    constexpr bool HAS_B = (KIND != K_BERNOULLI);
    constexpr bool HAS_AB = (KIND == K_WEIBULL || KIND == K_BETA);
    bool valid;
    R y;
    if (KIND == K_BERNOULLI) {
        // every day of the window counts; NaN (dry) is the event (model_scipy.py:444)
        valid = cell_active;
        y = (yf != yf) ? R(1) : R(0);
    } else {
        valid = cell_active && !(yf != yf);
        y = valid ? (R)yf : R(0.5);
    }
    R etaA, etaB;
    linear_predictors<HAS_B, R>(thA, thB, b, etaA, etaB);
    DayTerms<R> dt = family_day_t<KIND, R>(y, etaA, etaB);
    if (!valid) { dt.nll = R(0); dt.rA = R(0); dt.rB = R(0); dt.wAA = R(0); dt.wAB = R(0); dt.wBB = R(0); }
    if (MOMENTS) {
        R w = dt.wAA;
        accumulate<NH, R>(acc.AA, w, b);
        w *= b.g;
        accumulate<NH, R>(acc.AA + NTRIG, w, b);
        w *= b.g;
        accumulate<NH, R>(acc.AA + 2 * NTRIG, w, b);
        if (HAS_AB) {
            w = dt.wAB;
            accumulate<NH, R>(acc.AB, w, b);
            accumulate<NH, R>(acc.AB + NTRIG, w * b.g, b);
        }
        if (HAS_B) {
            accumulate<NH, R>(acc.BB, dt.wBB, b);
            accumulate<MAXM, R>(acc.GB, dt.rB, b);
        }
        accumulate<MAXM, R>(acc.GA, dt.rA, b);
        accumulate<MAXM, R>(acc.GA + NB, dt.rA * b.g, b);
    }
    return dt.nll;
}

}  // namespace attrici
