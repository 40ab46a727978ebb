// Rolling-window model of the Normal family (the reference's second model family, `--window-size`):
// every day of the year k = 1..366 has its own coefficients
//     mu(t) = a_k + b_k gmt(t) ,  sigma(t) = exp(c_k) ,   k = dayofyear(t)
// (attrici/estimation/model_pymc5.py:366-547: weights_<par>_longterm_intercept / _trend [366], all with N(0, 1) priors)
// fitted on the windows collect_windows gathers (attrici/util.py:107-172): for every day t of the fit series with day
// of year k the `window_size` days around it (series mirrored at both ends without repeating the edge day); day 366
// takes the windows around the day AFTER each day 365 (:126-133).  An observation enters ~window_size of the 366 sets.
//
// The log posterior separates over k, and for the Normal family the set of day k enters through six sums
//     n, sum g, sum g^2, sum y, sum y^2, sum y g            (over the multiset of windowed days)
// so a cell is 366 independent three-parameter MAP problems on sufficient statistics.
//
//   window_sums_kernel   thread = cell, CTA = 128 cells x 16 days of the year; per calendar-year run the window sum of
//                        the first centre is formed directly and then slides (one day enters, one leaves per step),
//                        accumulators in shared memory (thread-private columns): ~4 loads per (cell, year, k) instead
//                        of window_size.  Missing days are skipped (the reference drops them before it windows by
//                        index - a documented deviation for cells with gaps).
//   window_map_kernel    thread = (cell, k): damped Newton on (a, b, c) with the exact 3 x 3 Hessian, FP64
//   window_logp_kernel   thread = cell: log posterior = sum over k, status
//   window_qmap_kernel   thread = cell over a chunk of days: distribution parameters by day of year, quantile map /
//                        rescale / replacement through the same qmap_day as every other kernel (qmap_impl.cuh)
//
// The reference has no SciPy implementation of this model (model_scipy.py:58-59 raises), no test and no fixture for
// it, and PyMC is not in this image: parity is against oracle/window_oracle.py only ("parity unpinned").
#include "common.cuh"
#include "qmap_impl.cuh"

namespace attrici {

namespace {

constexpr int NDOY = 366;
constexpr int WK = 16;           // days of the year per CTA of the sums kernel
constexpr int W_THREADS = 128;

struct WindowArgs {
    const float* ys; int64_t ld;
    int T0, Tn, C, Cp;            // fit series = days [T0, T0 + Tn) of the axis
    const double* gmt;            // [T] scaled predictor on the axis
    const int* run_start;         // [n_runs] first day of a calendar-year run, relative to T0
    const int* run_first;         // [n_runs] its first day of the year (1-based)
    const int* run_len;           // [n_runs]
    int n_runs, half;
    double* sums;                 // [NDOY][6][Cp] {n, sum g, sum g^2, sum y, sum y^2, sum y g}
};

// day index of the mirrored series (util.py:136-150): reflection about the first and the last day
__device__ __forceinline__ int w_reflect(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * n - 2 - i;
    return min(max(i, 0), n - 1);
}

struct WSum { double n, g, gg, y, yy, yg; };

__device__ __forceinline__ void w_add(WSum& s, const WindowArgs& a, const float* col, int i, double sign) {
    const int t = a.T0 + w_reflect(i, a.Tn);
    const float v = __ldg(col + (size_t)t * a.ld);
    if (v == v) {
        const double y = (double)v, g = __ldg(a.gmt + t);
        s.n += sign; s.g += sign * g; s.gg = fma(sign * g, g, s.gg);
        s.y += sign * y; s.yy = fma(sign * y, y, s.yy); s.yg = fma(sign * y, g, s.yg);
    }
}

__global__ void __launch_bounds__(W_THREADS) window_sums_kernel(WindowArgs a) {
    extern __shared__ double w_acc[];   // [WK][6][W_THREADS]
    const int tid = threadIdx.x;
    const int cell = blockIdx.x * W_THREADS + tid;
    const int k0 = 1 + blockIdx.y * WK;                // first day of the year of this CTA (1-based)
    for (int i = 0; i < WK * 6; ++i) w_acc[i * W_THREADS + tid] = 0.0;
    if (cell >= a.C) return;
    const float* col = a.ys + cell;
    auto deposit = [&](int k, const WSum& s) {
        double* p = w_acc + (size_t)(k - k0) * 6 * W_THREADS + tid;
        p[0] += s.n; p[W_THREADS] += s.g; p[2 * W_THREADS] += s.gg;
        p[3 * W_THREADS] += s.y; p[4 * W_THREADS] += s.yy; p[5 * W_THREADS] += s.yg;
    };
    for (int r = 0; r < a.n_runs; ++r) {
        const int s0 = a.run_start[r], kf = a.run_first[r], len = a.run_len[r];
        const int klast = min(kf + len - 1, 365);     // the run's own day 366 is not a centre (util.py:126-133)
        const int ka = max(k0, kf), kb = min(k0 + WK - 1, klast);
        if (ka <= kb) {
            int c = s0 + (ka - kf);
            WSum s = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int o = -a.half; o <= a.half; ++o) w_add(s, a, col, c + o, 1.0);
            deposit(ka, s);
            for (int k = ka + 1; k <= kb; ++k) {
                // window(c + 1) = window(c) - {c - half} + {c + 1 + half} as multisets of mirrored days
                w_add(s, a, col, c - a.half, -1.0);
                ++c;
                w_add(s, a, col, c + a.half, 1.0);
                deposit(k, s);
            }
        }
        if (k0 <= NDOY && NDOY <= k0 + WK - 1 && kf <= 365 && 365 <= kf + len - 1) {
            // day 366: the window around the day after this run's day 365
            const int c = s0 + (365 - kf) + 1;
            WSum s = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            for (int o = -a.half; o <= a.half; ++o) w_add(s, a, col, c + o, 1.0);
            deposit(NDOY, s);
        }
    }
    for (int k = k0; k < k0 + WK && k <= NDOY; ++k)
        for (int j = 0; j < 6; ++j)
            a.sums[((size_t)(k - 1) * 6 + j) * a.Cp + cell] = w_acc[((size_t)(k - k0) * 6 + j) * W_THREADS + tid];
}

// MAP of one (cell, day of year): minimise  f = n c + e^{-2c} Q(a, b) / 2 + (a^2 + b^2 + c^2) / 2
// (Normal likelihood of the windowed days + the N(0, 1) priors of the three weights; constants added in logp)
__global__ void __launch_bounds__(W_THREADS) window_map_kernel(const double* __restrict__ sums, int C, int Cp,
                                                               double* __restrict__ params, double* __restrict__ obj,
                                                               int* __restrict__ iters) {
    const int cell = blockIdx.x * W_THREADS + threadIdx.x;
    const int k = blockIdx.y;
    if (cell >= C) return;
    const double* s = sums + (size_t)k * 6 * Cp + cell;
    const double n = s[0], sg = s[(size_t)Cp], sgg = s[2 * (size_t)Cp], sy = s[3 * (size_t)Cp], syy = s[4 * (size_t)Cp],
                 syg = s[5 * (size_t)Cp];
    double a = 0.0, b = 0.0, c = 0.0, f = 0.0;
    int it = 0, ok = 1;
    if (n > 0.0) {
        // start: ordinary least squares of y on (1, g), log of the residual standard deviation
        const double det = n * sgg - sg * sg;
        if (det > 1e-12 * fmax(n * sgg, 1e-300)) { a = (sy * sgg - syg * sg) / det; b = (n * syg - sg * sy) / det; }
        else a = sy / n;
        auto resid = [&](double a_, double b_, double& r0, double& r1) {
            r0 = sy - a_ * n - b_ * sg;
            r1 = syg - a_ * sg - b_ * sgg;
            return syy - a_ * sy - b_ * syg - a_ * r0 - b_ * r1;
        };
        double r0, r1;
        double Q = resid(a, b, r0, r1);
        c = 0.5 * log(fmax(Q / n, 1e-30));
        c = fmin(fmax(c, -20.0), 20.0);
        auto value = [&](double a_, double b_, double c_) {
            double q0, q1;
            const double Q_ = resid(a_, b_, q0, q1);
            return n * c_ + 0.5 * exp(-2.0 * c_) * Q_ + 0.5 * (a_ * a_ + b_ * b_ + c_ * c_);
        };
        f = value(a, b, c);
        ok = 0;
        double lam = 0.0;
        for (it = 0; it < 200; ++it) {
            Q = resid(a, b, r0, r1);
            const double e2 = exp(-2.0 * c);
            const double ga = -e2 * r0 + a, gb = -e2 * r1 + b, gc = n - e2 * Q + c;
            double haa = e2 * n + 1.0, hab = e2 * sg, hbb = e2 * sgg + 1.0, hac = 2.0 * e2 * r0, hbc = 2.0 * e2 * r1,
                   hcc = 2.0 * e2 * Q + 1.0;
            double da = 0.0, db = 0.0, dc = 0.0, dec = 0.0;
            bool solved = false;
            for (int tr = 0; tr < 60 && !solved; ++tr) {
                // Cholesky of H + lam diag(H)
                const double l11s = haa * (1.0 + lam);
                if (l11s > 0.0) {
                    const double l11 = sqrt(l11s), l21 = hab / l11, l31 = hac / l11;
                    const double l22s = hbb * (1.0 + lam) - l21 * l21;
                    if (l22s > 0.0) {
                        const double l22 = sqrt(l22s), l32 = (hbc - l31 * l21) / l22;
                        const double l33s = hcc * (1.0 + lam) - l31 * l31 - l32 * l32;
                        if (l33s > 0.0) {
                            const double l33 = sqrt(l33s);
                            const double y1 = -ga / l11, y2 = (-gb - l21 * y1) / l22, y3 = (-gc - l31 * y1 - l32 * y2) / l33;
                            dc = y3 / l33; db = (y2 - l32 * dc) / l22; da = (y1 - l21 * db - l31 * dc) / l11;
                            dec = -(ga * da + gb * db + gc * dc);     // Newton decrement (> 0)
                            const double ft = value(a + da, b + db, c + dc);
                            // sufficient decrease, or a step so small that the objective cannot resolve it
                            if (isfinite(ft) && (ft <= f - 1e-4 * dec || dec <= 1e-12 * fmax(fabs(f), 1.0))) {
                                solved = true;
                                f = fmin(ft, f);
                                lam = lam * 0.25;
                                if (lam < 1e-12) lam = 0.0;
                                break;
                            }
                        }
                    }
                }
                lam = fmax(lam * 8.0, 1e-6);
            }
            if (!solved) break;
            a += da; b += db; c += dc;
            if (dec <= 1e-13 * fmax(fabs(f), 1.0)) { ok = 1; ++it; break; }
        }
        f = value(a, b, c);
    }
    double* p = params + (size_t)cell * 3 * NDOY + k;
    p[0] = a; p[NDOY] = b; p[2 * NDOY] = c;
    obj[(size_t)k * Cp + cell] = f + (n + 3.0) * 0.91893853320467274178;   // ln sqrt(2 pi) of the n days and of the three priors
    iters[(size_t)k * Cp + cell] = ok ? it : -1;
}

__global__ void window_logp_kernel(const double* __restrict__ obj, const int* __restrict__ iters, const double* __restrict__ sums,
                                   int C, int Cp, double* __restrict__ logp, int32_t* __restrict__ status,
                                   int32_t* __restrict__ n_iter) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= C) return;
    double f = 0.0, n_tot = 0.0;
    int worst = 0, bad = 0;
    for (int k = 0; k < NDOY; ++k) {
        f += obj[(size_t)k * Cp + cell];
        n_tot += sums[(size_t)k * 6 * Cp + cell];
        const int it = iters[(size_t)k * Cp + cell];
        if (it < 0) bad = 1; else worst = max(worst, it);
    }
    const int st = !(n_tot > 0.0) ? ATTRICI_STATUS_NO_DATA : (!isfinite(f) ? ATTRICI_STATUS_NONFINITE
                   : (bad ? ATTRICI_STATUS_STALLED : ATTRICI_STATUS_CONVERGED));
    logp[cell] = (st == ATTRICI_STATUS_NO_DATA || st == ATTRICI_STATUS_NONFINITE) ? NAN : -f;
    if (status) status[cell] = st;
    if (n_iter) n_iter[cell] = worst;
}

struct WQmapArgs {
    const float* y; const float* ys; int64_t ld, ld_out;
    int T, C, chunk_len;
    const double* gmt; const int16_t* doy; const double* params; const double* scaling_v;
    double* cfact; uint8_t* replaced; double* cfact_scaled;
    int family, scaling, post_rule;
};

__global__ void __launch_bounds__(W_THREADS) window_qmap_kernel(WQmapArgs a) {
    const int cell = blockIdx.x * W_THREADS + threadIdx.x;
    if (cell >= a.C) return;
    const int t0 = blockIdx.y * a.chunk_len, t1 = min(a.T, t0 + a.chunk_len);
    const double* p = a.params + (size_t)cell * 3 * NDOY;
    const double datamin = a.scaling_v[2 * (size_t)cell], scale = a.scaling_v[2 * (size_t)cell + 1];
    const bool mask_in_place = (a.scaling == ATTRICI_SCALE_PERCENT || a.scaling == ATTRICI_SCALE_RANGE);
    for (int t = t0; t < t1; ++t) {
        const size_t i = (size_t)t * a.ld + cell;
        float y = __ldg(a.y + i);
        const float s = __ldg(a.ys + i);
        if (fabsf(y) == INFINITY) y = __int_as_float(0x7fc00000);            // detrend.py:311
        if (mask_in_place && s != s) y = __int_as_float(0x7fc00000);
        const int k = (int)a.doy[t] - 1;
        const double g = a.gmt[t];
        DistPar ref, cf;
        const double sg = exp(p[2 * NDOY + k]);
        ref.p0 = fma(p[NDOY + k], g, p[k]); cf.p0 = p[k];
        ref.p1 = cf.p1 = sg; ref.p2 = cf.p2 = 0.0;
        const QmapOut o = qmap_day(a.family, a.scaling, a.post_rule, y, s, ref, cf, datamin, scale, 0.0);
        const size_t oi = (size_t)t * a.ld_out + cell;
        a.cfact[oi] = o.cfact;
        if (a.replaced) a.replaced[oi] = (uint8_t)o.replaced;
        if (a.cfact_scaled) a.cfact_scaled[oi] = o.cfact_scaled;
    }
}

}  // namespace

size_t window_workspace_bytes(int C) {
    const size_t cp = (size_t)round_up(C, 32);
    return Workspace::align(8 * 6 * (size_t)NDOY * cp) + Workspace::align(8 * (size_t)NDOY * cp) + Workspace::align(4 * (size_t)NDOY * cp);
}

int launch_window_fit(void* workspace, const float* ys, int64_t ld, int T, int C, int i0, int i1, const double* gmt,
                      const int* run_start, const int* run_first, const int* run_len, int n_runs, int window_size,
                      double* params, double* logp, int32_t* status, int32_t* n_iter, cudaStream_t s) {
    const int cp = round_up(C, 32);
    char* base = (char*)workspace;
    double* sums = (double*)base;
    double* obj = (double*)(base + Workspace::align(8 * 6 * (size_t)NDOY * cp));
    int* iters = (int*)(base + Workspace::align(8 * 6 * (size_t)NDOY * cp) + Workspace::align(8 * (size_t)NDOY * cp));
    WindowArgs a;
    a.ys = ys; a.ld = ld; a.T0 = i0; a.Tn = i1 - i0; a.C = C; a.Cp = cp; a.gmt = gmt;
    a.run_start = run_start; a.run_first = run_first; a.run_len = run_len; a.n_runs = n_runs; a.half = window_size / 2;
    a.sums = sums;
    (void)T;
    const size_t smem = sizeof(double) * WK * 6 * W_THREADS;
    ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(window_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((C + W_THREADS - 1) / W_THREADS, (NDOY + WK - 1) / WK);
    window_sums_kernel<<<grid, W_THREADS, smem, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    window_map_kernel<<<dim3((C + W_THREADS - 1) / W_THREADS, NDOY), W_THREADS, 0, s>>>(sums, C, cp, params, obj, iters);
    ATTRICI_LAUNCH_CHECK();
    window_logp_kernel<<<(C + 127) / 128, 128, 0, s>>>(obj, iters, sums, C, cp, logp, status, n_iter);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_window_qmap(const attrici_variable_t& var, const float* y, const float* ys, int64_t ld, int T, int C,
                       const double* gmt, const int16_t* doy, const double* params, const double* scaling, double* cfact,
                       uint8_t* replaced, double* cfact_scaled, int64_t ld_out, cudaStream_t s) {
    WQmapArgs a;
    a.y = y; a.ys = ys; a.ld = ld; a.ld_out = ld_out; a.T = T; a.C = C;
    a.gmt = gmt; a.doy = doy; a.params = params; a.scaling_v = scaling;
    a.cfact = cfact; a.replaced = replaced; a.cfact_scaled = cfact_scaled;
    a.family = var.family; a.scaling = var.scaling; a.post_rule = var.post_rule;
    const int n_cb = (C + W_THREADS - 1) / W_THREADS;
    int want = (148 * 16 + n_cb - 1) / n_cb;
    if (want > 512) want = 512;
    a.chunk_len = (T + want - 1) / want;
    if (a.chunk_len < 8) a.chunk_len = 8;
    dim3 grid(n_cb, (T + a.chunk_len - 1) / a.chunk_len);
    window_qmap_kernel<<<grid, W_THREADS, 0, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
