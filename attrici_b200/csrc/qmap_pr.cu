// Quantile map of precipitation (Bernoulli-Gamma) in phase order, with the wet days compacted.
//
// The day-ordered kernel (qmap.cu, thread = cell) spends its time in the incomplete-gamma evaluations of the wet
// days: cdf_ref(y), then the inverse under the counterfactual parameters.  A warp holds 32 cells of one day, about
// a third of them wet, so two thirds of the lanes idle through every series term (measured: 9.8 of 32 lanes active,
// profiles/r2_qmap_pr_hurs_ncu.txt), and the per-day predictors (45 FMAs, five exp, lgamma) are evaluated by all.
// Here, on a gap-free daily axis (fold.cu):
//
//  * pass 1, thread = (cell, half of a batch of 10 days of one phase): the 45-term predictor sums, nu, a = nu^2,
//    lgamma(a), p_cf, mu_cf once per (cell, phase); per day two FMAs and two exp (p_ref, mu_ref).  A dry day that
//    stays dry is finished on the spot (cfact = 0, coalesced store).  A wet day - or a dry day the MT19937 draw turns
//    wet - becomes an item {x = y / scale_ref, p_ref, y, (day, cell)} on one of two shared-memory stacks (short / long
//    series: the number of terms grows with x, and a warp runs as long as its longest lane).
//  * pass 2, thread = item: whenever a stack holds at least 256 items, its top 256 are popped, one per thread:
//    P(a, x) and Q(a, x) from one series, the factual quantile, the decision of Pr.quantile_mapping, and the
//    counterfactual inverse started AT x with P, Q known (the shape does not depend on the predictor, so for equal
//    dry-day probabilities x is the root; the first Halley step costs no evaluation).  All 32 lanes of a warp work,
//    whatever the wet-day pattern of the cells.  The stack is emptied at the end of each phase (the per-phase values
//    of a cell live in one slot per cell).
//
// The decisions and formulas per day are those of qmap_impl.cuh::qmap_day; the incomplete gamma function is evaluated
// by this file's own copy of the Halley iteration of special.cuh (the density from the prefactor, a cubic stopping rule
// at 3e-4 x - rounding-level and 3e-11-level differences, inside
// the 1e-5 the tests ask of the counterfactual); the trig columns are those of the phase's first day, as in every
// phase-ordered kernel.  Measured (10 000 cells x 43 464 days): 253 ms for the day-ordered kernel and its MT19937
// stream at the start of round 2, 41 ms now.
//
// Replaces Pr.quantile_mapping (attrici/variables.py:422-480), BernoulliGamma.cdf / invcdf
// (attrici/distributions.py:178-216), ModelScipy.estimate_distribution x2 (model_scipy.py:547-570), Pr.rescale
// (variables.py:483-486) and the replacement of detrend.py:392-411, like qmap.cu.
#include <stdio.h>
#include <stdlib.h>
#include "common.cuh"
#include "qmap_impl.cuh"

namespace attrici {

namespace {

constexpr int PR_CELLS = 128;
constexpr int PR_THREADS = 256;                   // two threads per cell in pass 1 (days 0..4 / 5..9 of a batch)
constexpr int PR_U = 10;                          // days per batch
constexpr int PR_MAXPH = 16;                      // phases per CTA
constexpr int PR_NCOEF = 2 * NA + NB;             // p (18) | mu (18) | nu (9), canonical layout
constexpr int PR_CAP = PR_CELLS * PR_U + 2 * PR_THREADS;   // one batch on top of two remainders below one round each

struct PrArgs {
    const float* y; const float* ys;
    int64_t ld, ld_out;
    int T, C, Cp, modes, n_ph, ph_len, nj_max;
    const double* basis64;
    const double* params;      // [C x P] reference layout
    const double* scaling_v;   // [C x 2]
    const double* uniforms;    // [T][Cp]
    double* cfact; uint8_t* replaced; double* cfact_scaled;
    const int* fold_bad;
    int scaling, post_rule;
    double xl_a, xl_b;         // an item goes to the long-series stack if x > xl_a * shape + xl_b
};

struct PrPhase { double a, gln, p_cf, sc_cf; };   // per (cell, phase): shape nu^2, lgamma(a), counterfactual p and scale

// one item: the quantile map of qmap_day for a wet day (x > 0) or a dry day the uniform draw turned wet (x = -1;
// NaN: the same with degenerate parameters)
__device__ __forceinline__ double pr_item(double x, double p_ref, const PrPhase& ph) {
    double q, P = -1.0, Q = -1.0, pref = 0.0;
    const bool dry = !(x > 0.0);
    if (isnan(x)) q = NAN;
    else if (dry) q = p_ref;
    else { gamma_pq_pref_l(ph.a, x, ph.gln, &P, &Q, &pref); q = p_ref + (1.0 - p_ref) * P; }
    const bool drier_cf = ph.p_cf > p_ref;
    const bool qm0 = drier_cf && (q > ph.p_cf);
    const bool qm1 = !drier_cf && !dry;
    // (a dry day is only pushed when its uniform draw turns it wet: to_wet of qmap_day)
    double cs = 0.0;
    if (qm0 || qm1 || dry) {
        const double target = (q - ph.p_cf) / (1.0 - ph.p_cf);
        // gamma_ppf (qmap_impl.cuh)
        if (isnan(target) || isnan(ph.a) || isnan(ph.sc_cf) || !(ph.a > 0.0) || !(ph.sc_cf > 0.0) || target < 0.0 || target > 1.0) {
            cs = NAN;
        } else {
            const bool known = !dry && isfinite(x);
            cs = gamma_p_inv_from(ph.a, target, known ? x : -1.0, known ? P : -1.0, known ? Q : -1.0, pref, ph.gln) * ph.sc_cf;
        }
    }
    return cs;
}

// a wet day whose standardised value is not positive (or not a number: degenerate parameters): the generic day
__device__ __noinline__ QmapOut pr_rare_day(int scaling, int post_rule, float y, float s, double p_ref, double mu_ref, double nu,
                                            double p_cf, double mu_cf, double datamin, double scale, double u) {
    DistPar ref, cf;
    ref.p0 = p_ref; ref.p1 = mu_ref; ref.p2 = nu;
    cf.p0 = p_cf; cf.p1 = mu_cf; cf.p2 = nu;
    return qmap_day(ATTRICI_BERNOULLI_GAMMA, scaling, post_rule, y, s, ref, cf, datamin, scale, u);
}

__global__ void __launch_bounds__(PR_THREADS, 2) qmap_fold_pr_kernel(PrArgs a) {
    if (a.fold_bad[0] != 0) return;   // not a gap-free daily axis: the day-ordered kernel runs instead
    extern __shared__ __align__(16) unsigned char pr_smem[];
    double* coef_s = reinterpret_cast<double*>(pr_smem);                        // [45][128]
    double* g_s = coef_s + PR_NCOEF * PR_CELLS;                                  // [ph][nj_max]
    double* row_s = g_s + PR_MAXPH * a.nj_max;                                   // [ph][8]: cos 1..4, sin 1..4
    PrPhase* ph_s = reinterpret_cast<PrPhase*>(row_s + PR_MAXPH * 2 * MAXM);     // [128]
    double* sc_s = reinterpret_cast<double*>(ph_s + PR_CELLS);                   // [128][2] datamin, scale
    double* it_x = sc_s + 2 * PR_CELLS;                                          // [PR_CAP]
    double* it_p = it_x + PR_CAP;                                                // [PR_CAP]
    float* it_y = reinterpret_cast<float*>(it_p + PR_CAP);                       // [PR_CAP]
    unsigned short* it_i = reinterpret_cast<unsigned short*>(it_y + PR_CAP);     // [PR_CAP] (j << 7) | cell
    int* count_s = reinterpret_cast<int*>(it_i + PR_CAP + (PR_CAP & 1));        // [2][2], by batch parity: short-series stack (grows up from 0), long (down from PR_CAP)

    const int tid = threadIdx.x;
    const int c = tid & (PR_CELLS - 1), half = tid >> 7;
    const int cell = blockIdx.x * PR_CELLS + c;
    const bool live = cell < a.C;
    const int ccol = live ? cell : a.C - 1;
    const int d0 = blockIdx.y * a.ph_len, d1 = min(a.n_ph, d0 + a.ph_len);
    if (d0 >= d1) return;
    const int nph = d1 - d0;
    {
        const int P = family_num_params(ATTRICI_BERNOULLI_GAMMA, a.modes);
        if (half == 0) {
            const double* pp = a.params + (size_t)ccol * P;
#pragma unroll 1
            for (int i = 0; i < PR_NCOEF; ++i) coef_s[i * PR_CELLS + c] = 0.0;
            for (int j = 0; j < P; ++j) {
                int slot, canon;
                ref_to_canon(ATTRICI_BERNOULLI_GAMMA, a.modes, j, &slot, &canon);
                coef_s[(canon < NA ? slot * NA + canon : 2 * NA + canon - NA) * PR_CELLS + c] = pp[j];
            }
            sc_s[2 * c] = a.scaling_v[2 * (size_t)ccol];
            sc_s[2 * c + 1] = a.scaling_v[2 * (size_t)ccol + 1];
        }
        for (int i = tid; i < nph * a.nj_max; i += PR_THREADS) {
            const int p = i / a.nj_max, j = i - p * a.nj_max;
            const int t = d0 + p + j * NPHASE;
            g_s[i] = (t < a.T) ? a.basis64[(size_t)t * B_STRIDE] : 0.0;
        }
        for (int i = tid; i < nph * 2 * MAXM; i += PR_THREADS) {
            const int p = i / (2 * MAXM), k = i - p * (2 * MAXM);
            row_s[i] = a.basis64[(size_t)(d0 + p) * B_STRIDE + (k < MAXM ? 1 + k : 1 + NH + (k - MAXM))];
        }
        if (tid < 4) count_s[tid] = 0;
    }
    __syncthreads();
    const bool want_rep = a.replaced != nullptr, want_cs = a.cfact_scaled != nullptr;
    const size_t jstride = (size_t)NPHASE * a.ld, jstride_o = (size_t)NPHASE * a.ld_out, jstride_u = (size_t)NPHASE * a.Cp;
    const double* cf = coef_s + c;
    const int lane = tid & 31;

    // pops `n` items from the top of the stack (n <= PR_THREADS), one per thread, and writes their results
    // pops n <= PR_THREADS items, one per thread, at first + step * tid, and writes their results
    auto pop_round = [&](int first, int step, int n, int d) {
        if (tid < n) {
            const int i = first + step * tid;
            const unsigned idx = it_i[i];
            const int cc = idx & (PR_CELLS - 1), j = idx >> 7;
            const float y = it_y[i];
            const double cs = pr_item(it_x[i], it_p[i], ph_s[cc]);
            const QmapOut o = qmap_finish(a.scaling, a.post_rule, cs, y, sc_s[2 * cc], sc_s[2 * cc + 1]);
            const size_t oi = (size_t)d * a.ld_out + (size_t)j * jstride_o + (size_t)(blockIdx.x * PR_CELLS + cc);
            a.cfact[oi] = o.cfact;
            if (want_rep) a.replaced[oi] = (uint8_t)o.replaced;
            if (want_cs) a.cfact_scaled[oi] = o.cfact_scaled;
        }
    };

    int batch = 0;
    for (int p = 0; p < nph; ++p) {
        const int d = d0 + p;
        const int nj = (a.T - d + NPHASE - 1) / NPHASE;
        const double* row = row_s + p * 2 * MAXM;
        // predictor sums of the phase in the accumulation order of day_predictors (qmap_impl.cuh)
        double base[2], trend[2], eb = cf[2 * NA * PR_CELLS];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const double* cs_ = cf + s * NA * PR_CELLS;
            double b = cs_[0], tr = cs_[1 * PR_CELLS];
#pragma unroll
            for (int k = 0; k < MAXM; ++k) {
                const double ck = row[k], sk = row[MAXM + k];
                b = fma(cs_[(2 + k) * PR_CELLS], ck, b);
                b = fma(cs_[(2 + MAXM + k) * PR_CELLS], sk, b);
                tr = fma(cs_[(2 + 2 * MAXM + k) * PR_CELLS], ck, tr);
                tr = fma(cs_[(2 + 3 * MAXM + k) * PR_CELLS], sk, tr);
            }
            base[s] = b; trend[s] = tr;
        }
#pragma unroll
        for (int k = 0; k < MAXM; ++k) {
            eb = fma(cf[(2 * NA + 1 + k) * PR_CELLS], row[k], eb);
            eb = fma(cf[(2 * NA + 1 + MAXM + k) * PR_CELLS], row[MAXM + k], eb);
        }
        // family_params (qmap_impl.cuh): p = invlogit, mu = exp, nu = exp; scale of the gamma part = mu / nu^2
        const double nu = exp(eb), sh = nu * nu;
        const double p_cf = invlogit(base[0]), mu_cf = exp(base[1]);
        if (half == 0) {
            PrPhase ph;
            ph.a = sh; ph.gln = lgamma(sh); ph.p_cf = p_cf; ph.sc_cf = mu_cf / sh;
            ph_s[c] = ph;
        }
        __syncthreads();   // the slots of the phase are set (the stack is empty here)
        const double* gp = g_s + p * a.nj_max;
        const float* yp = a.y + (size_t)d * a.ld + ccol;
        const float* sp = a.ys + (size_t)d * a.ld + ccol;
        const double* up = a.uniforms + (size_t)d * a.Cp + ccol;
        const size_t out0 = (size_t)d * a.ld_out + cell;
        for (int j0 = 0; j0 < nj; j0 += PR_U) {
            // ---- pass 1: this thread's five days of the batch
            constexpr int H = (PR_U + 1) / 2;
            const int jb = j0 + half * H, je = min(nj, j0 + (half ? PR_U : H));   // days jb .. je - 1
            int* cnt = count_s + 2 * (batch & 1);
            float yv[H], sv[H];
#pragma unroll
            for (int u = 0; u < H; ++u) {
                const size_t off = (size_t)min(jb + u, nj - 1) * jstride;
                yv[u] = __ldg(yp + off);
                sv[u] = __ldg(sp + off);
            }
#pragma unroll
            for (int u = 0; u < H; ++u) {
                const int j = jb + u;
                const bool in = j < je && live;
                float y = yv[u];
                const float s = sv[u];
                if (isinf(y)) y = __int_as_float(0x7fc00000);   // detrend.py:311
                const double g = gp[min(j, nj - 1)];
                const double p_ref = invlogit(fma(g, trend[0], base[0]));
                const bool dry = (s != s);
                bool push = in && !dry;
                double x = -1.0;
                if (!dry) {
                    // x = y / scale_ref = y nu^2 / mu_ref, without a division: y (nu^2 e^{-eta})
                    const double eta = fma(g, trend[1], base[1]);
                    x = (double)s * (sh * exp(-eta));
                    if (!(sh > 0.0) || !(x > 0.0) || !isfinite(x)) {
                        // not a positive standardised value (never for a series Pr.scale produced): the generic day
                        push = false;
                        if (in) {
                            const double mu_ref = exp(eta);
                            const QmapOut o = pr_rare_day(a.scaling, a.post_rule, y, s, p_ref, mu_ref, nu, p_cf, mu_cf, sc_s[2 * c],
                                                          sc_s[2 * c + 1], __ldg(up + (size_t)j * jstride_u));
                            const size_t oi = out0 + (size_t)j * jstride_o;
                            a.cfact[oi] = o.cfact;
                            if (want_rep) a.replaced[oi] = (uint8_t)o.replaced;
                            if (want_cs) a.cfact_scaled[oi] = o.cfact_scaled;
                        }
                    }
                } else if (in) {
                    // dry day: wet in the counterfactual if the uniform draw says so (variables.py:458-466)
                    const double uu = __ldg(up + (size_t)j * jstride_u);
                    push = uu * p_ref > p_cf;   // (qm0 of qmap_day cannot hold on a dry day: it needs p_ref > p_cf > p_ref)
                    if (push) {
                        // the quantile of a dry day is p_ref - or NaN for degenerate parameters, as dist_cdf gives it
                        const double sc_ref = exp(fma(g, trend[1], base[1])) / sh;
                        x = (!(sh > 0.0) || !(sc_ref > 0.0)) ? NAN : -1.0;
                    } else {
                        const QmapOut o = qmap_finish(a.scaling, a.post_rule, 0.0, y, sc_s[2 * c], sc_s[2 * c + 1]);
                        const size_t oi = out0 + (size_t)j * jstride_o;
                        a.cfact[oi] = o.cfact;
                        if (want_rep) a.replaced[oi] = (uint8_t)o.replaced;
                        if (want_cs) a.cfact_scaled[oi] = o.cfact_scaled;
                    }
                }
                // warp-aggregated push.  Two stacks: the length of the series grows with x (about x + 6 sqrt(x) + 20
                // terms) and a warp runs as long as its longest lane, so items with a long series are kept together
                const bool longer = !(x <= fma(a.xl_a, sh, a.xl_b));   // also the dry days turned wet (several evaluations)
                const unsigned m_lo = __ballot_sync(0xffffffffu, push && !longer), m_hi = __ballot_sync(0xffffffffu, push && longer);
                if (m_lo | m_hi) {
                    int b_lo = 0, b_hi = 0;
                    if (lane == 0) {
                        if (m_lo) b_lo = atomicAdd(cnt, __popc(m_lo));
                        if (m_hi) b_hi = atomicAdd(cnt + 1, __popc(m_hi));
                    }
                    b_lo = __shfl_sync(0xffffffffu, b_lo, 0);
                    b_hi = __shfl_sync(0xffffffffu, b_hi, 0);
                    if (push) {
                        const unsigned below = (1u << lane) - 1u;
                        const int i = longer ? PR_CAP - 1 - (b_hi + __popc(m_hi & below)) : b_lo + __popc(m_lo & below);
                        it_x[i] = x; it_p[i] = p_ref; it_y[i] = y; it_i[i] = (unsigned short)((j << 7) | c);
                    }
                }
            }
            __syncthreads();
            // ---- pass 2: full rounds from the tops of the stacks; everything at the end of the phase
            const int n_lo = cnt[0], n_hi = cnt[1];
            const bool last_batch = j0 + PR_U >= nj;
            const int t_lo = last_batch ? n_lo : (n_lo / PR_THREADS) * PR_THREADS;
            const int t_hi = last_batch ? n_hi : (n_hi / PR_THREADS) * PR_THREADS;
            // the next batch counts on from the remainders in the other pair of counters (nobody reads that pair any
            // more: its last readers passed the barrier above), so one barrier after the pops is enough
            if (tid == 0) { count_s[2 * ((batch + 1) & 1)] = n_lo - t_lo; count_s[2 * ((batch + 1) & 1) + 1] = n_hi - t_hi; }
            for (int done = 0; done < t_hi; done += PR_THREADS) pop_round(PR_CAP - n_hi + done, 1, min(PR_THREADS, t_hi - done), d);
            for (int done = 0; done < t_lo; done += PR_THREADS) pop_round(n_lo - 1 - done, -1, min(PR_THREADS, t_lo - done), d);
            __syncthreads();
            ++batch;
        }
    }
}

}  // namespace

// Bernoulli-Gamma family in phase order; runs only if ws.fold_bad[0] == 0 (checked on the device)
int launch_qmap_fold_pr(Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled, int64_t ld,
                        const double* params, const double* scaling, const double* uniforms, double* cfact,
                        uint8_t* replaced, double* cfact_scaled, int64_t ld_out, cudaStream_t s) {
    PrArgs a;
    a.y = y; a.ys = y_scaled; a.ld = ld; a.ld_out = ld_out;
    a.T = ws.T; a.C = ws.C; a.Cp = ws.Cp; a.modes = var.modes;
    a.n_ph = ws.T < NPHASE ? ws.T : NPHASE;
    a.nj_max = (ws.T + NPHASE - 1) / NPHASE;
    a.basis64 = ws.basis64; a.params = params; a.scaling_v = scaling; a.uniforms = uniforms;
    a.cfact = cfact; a.replaced = replaced; a.cfact_scaled = cfact_scaled;
    a.fold_bad = ws.fold_bad;
    a.scaling = var.scaling; a.post_rule = var.post_rule;
    a.xl_a = 0.0; a.xl_b = 2.5;
    if (const char* e = getenv("ATTRICI_PR_XL")) sscanf(e, "%lf,%lf", &a.xl_a, &a.xl_b);   // tuning only
    const int n_cb = (ws.C + PR_CELLS - 1) / PR_CELLS;
    int want = (148 * 8 + n_cb - 1) / n_cb;
    if (want > a.n_ph) want = a.n_ph;
    if (want < 1) want = 1;
    a.ph_len = (a.n_ph + want - 1) / want;
    if (a.ph_len > PR_MAXPH) a.ph_len = PR_MAXPH;
    const int n_chunks = (a.n_ph + a.ph_len - 1) / a.ph_len;
    if (a.nj_max > 511) { set_error("series too long for the phase-ordered precipitation quantile map (T=%d)", ws.T); return ATTRICI_EUNSUPPORTED; }
    const size_t smem = sizeof(double) * ((size_t)PR_NCOEF * PR_CELLS + (size_t)PR_MAXPH * a.nj_max + PR_MAXPH * 2 * MAXM + 2 * PR_CELLS +
                                          2 * (size_t)PR_CAP) +
                        sizeof(PrPhase) * PR_CELLS + sizeof(float) * PR_CAP + sizeof(unsigned short) * (PR_CAP + (PR_CAP & 1)) + 16;
    if (smem > 200 * 1024) { set_error("series too long for the phase-ordered precipitation quantile map (T=%d)", ws.T); return ATTRICI_EUNSUPPORTED; }
    ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(qmap_fold_pr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qmap_fold_pr_kernel<<<dim3(n_cb, n_chunks), PR_THREADS, smem, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
