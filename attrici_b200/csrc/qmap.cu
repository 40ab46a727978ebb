// Quantile-map kernels: thread = cell (coalesced 128-byte lines of the time-major series), CTA = 128
// cells x one time chunk; FP64 throughout like the reference (scipy.stats cdf / ppf).
// Mathematics in qmap_impl.cuh; reference lines cited there.
#include <stdlib.h>
#include "common.cuh"
#include "qmap_impl.cuh"

namespace attrici {

namespace {

constexpr int QMAP_THREADS = 128;

struct QmapArgs {
    const float* y; const float* ys;
    int64_t ld;
    int T, C, Cp;
    int chunk_len;
    int family, scaling, post_rule, modes;
    const double* basis64;
    const float* basis32;
    const double* params;    // [C x P] reference layout
    const double* scaling_v; // [C x 2]
    const double* uniforms;  // [T][Cp] (pr only)
    double* cfact; uint8_t* replaced; double* cfact_scaled;
    int64_t ld_out;
    const int* fold_bad;     // device flag of attrici_build_basis: 0 = gap-free daily axis
    int want_fold;           // this launch runs only if (fold_bad[0] == 0) == want_fold (-1: always)
    int ph_len;              // phases per CTA of the folded kernel
    float lo, hi; int has_lo, has_hi;   // thresholds, to recompute y_scaled when it is not given
};

__device__ void load_coef(int family, int modes, const double* __restrict__ params, int cell, CellCoef& c) {
    const int P = family_num_params(family, modes);
#pragma unroll
    for (int i = 0; i < NA; ++i) { c.a[0][i] = 0.0; c.a[1][i] = 0.0; }
#pragma unroll
    for (int i = 0; i < NB; ++i) c.b[i] = 0.0;
    for (int j = 0; j < P; ++j) {
        int sl, cn;
        ref_to_canon(family, modes, j, &sl, &cn);
        double v = params[(size_t)cell * P + j];
        if (cn < NA) c.a[sl][cn] = v;
        else c.b[cn - NA] = v;
    }
}

// y_scaled of one observation, recomputed with the float32 arithmetic of prep.cu (bit-identical), for callers
// that do not keep the scaled series (Normal family: unity / none scaling)
__device__ __forceinline__ float rescale_input(const QmapArgs& a, float v, float datamin, float scale) {
    if (isinf(v)) v = __int_as_float(0x7fc00000);
    if (a.scaling == ATTRICI_SCALE_UNITY) return __fdiv_rn(__fsub_rn(v, datamin), scale);
    if (a.has_lo && v <= a.lo) v = __int_as_float(0x7fc00000);
    if (a.has_hi && v >= a.hi) v = __int_as_float(0x7fc00000);
    if (a.scaling == ATTRICI_SCALE_RANGE) return __fdiv_rn(v, scale);
    if (a.scaling == ATTRICI_SCALE_PERCENT) return __fdiv_rn(v, 100.0f);
    return v;
}

constexpr int QT = 64;   // days per shared-memory tile of the basis rows

__device__ __forceinline__ void qcp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}

// thread = cell, CTA = 128 cells x one time chunk.  The basis rows of the chunk (float row for the z
// estimate of the Normal shortcut, double row for everything that reaches the output) are staged through
// shared memory in double-buffered 64-day tiles and read back as warp-uniform broadcasts.
__global__ void __launch_bounds__(QMAP_THREADS) qmap_kernel(QmapArgs a) {
    __shared__ __align__(16) float tile32[2][QT * B_STRIDE];
    __shared__ __align__(16) double tile64[2][QT * B_STRIDE];
    if (a.want_fold >= 0 && (a.fold_bad[0] == 0) != (a.want_fold != 0)) return;   // uniform: the folded kernel runs instead
    const int cell = blockIdx.x * QMAP_THREADS + threadIdx.x;
    const bool live = cell < a.C;
    const int ccol = live ? cell : a.C - 1;
    const int t0 = blockIdx.y * a.chunk_len, t1 = min(a.T, t0 + a.chunk_len);
    if (t0 >= t1) return;
    CellCoef coef;
    load_coef(a.family, a.modes, a.params, ccol, coef);
    const int n_dep = (a.family == ATTRICI_BERNOULLI_GAMMA) ? 2 : 1;
    const double datamin = a.scaling_v[2 * (size_t)ccol], scale = a.scaling_v[2 * (size_t)ccol + 1];
    const bool mask_in_place = (a.scaling == ATTRICI_SCALE_PERCENT || a.scaling == ATTRICI_SCALE_RANGE);
    const bool normal = (a.family == ATTRICI_NORMAL);
    float af[NA], bf[NB];
#pragma unroll
    for (int i = 0; i < NA; ++i) af[i] = (float)coef.a[0][i];
#pragma unroll
    for (int i = 0; i < NB; ++i) bf[i] = (float)coef.b[i];

    const int n_tiles = (t1 - t0 + QT - 1) / QT;
    auto issue_tile = [&](int it) {
        // rows beyond the series exist (the basis arrays are padded by two pass tiles of zeros)
        const size_t r0 = (size_t)(t0 + it * QT) * B_STRIDE;
        const float4* s32 = reinterpret_cast<const float4*>(a.basis32 + r0);
        const double2* s64 = reinterpret_cast<const double2*>(a.basis64 + r0);
        for (int v = threadIdx.x; v < QT * B_STRIDE / 4; v += QMAP_THREADS)
            qcp_async16(reinterpret_cast<float4*>(tile32[it & 1]) + v, s32 + v);
        for (int v = threadIdx.x; v < QT * B_STRIDE / 2; v += QMAP_THREADS)
            qcp_async16(reinterpret_cast<double2*>(tile64[it & 1]) + v, s64 + v);
        asm volatile("cp.async.commit_group;\n" ::);
    };
    issue_tile(0);
    for (int it = 0; it < n_tiles; ++it) {
        if (it + 1 < n_tiles) {
            issue_tile(it + 1);
            asm volatile("cp.async.wait_group 1;\n" ::);
        } else {
            asm volatile("cp.async.wait_group 0;\n" ::);
        }
        __syncthreads();
        if (live) {
            const int tt0 = t0 + it * QT;
            const int nd = min(QT, t1 - tt0);
            const float* r32 = tile32[it & 1];
            const double* r64 = tile64[it & 1];
            const float* yp = a.y + (size_t)tt0 * a.ld + cell;
            const float* sp = a.ys ? a.ys + (size_t)tt0 * a.ld + cell : nullptr;
            float ynx = __ldg(yp), snx = sp ? __ldg(sp) : 0.f;
            for (int d = 0; d < nd; ++d) {
                float yv = ynx;
                const float sv = sp ? snx : rescale_input(a, yv, (float)datamin, (float)scale);
                if (d + 1 < nd) {
                    ynx = __ldg(yp + (size_t)(d + 1) * a.ld);
                    if (sp) snx = __ldg(sp + (size_t)(d + 1) * a.ld);
                }
                const double* row = r64 + d * B_STRIDE;
                // the caller-visible series after Inf -> NaN (detrend.py:311) and the in-place masks of
                // hurs / tasrange / sfcWind (variables.py:583, 723, 771)
                if (isinf(yv) || (mask_in_place && isnan(sv))) yv = __int_as_float(0x7fc00000);
                QmapOut o;
                double cs_fast;
                if (normal && normal_fast_path(coef, af, bf, r32 + d * B_STRIDE, row, sv, &cs_fast)) {
                    o = qmap_finish(a.scaling, a.post_rule, cs_fast, yv, datamin, scale);
                } else {
                    DayParams dp;
                    day_predictors(coef, row, n_dep, dp);
                    DistPar ref, cf;
                    family_params(a.family, dp, ref, cf);
                    const double u = a.uniforms ? a.uniforms[(size_t)(tt0 + d) * a.Cp + cell] : 0.0;
                    o = qmap_day(a.family, a.scaling, a.post_rule, yv, sv, ref, cf, datamin, scale, u);
                }
                const size_t oi = (size_t)(tt0 + d) * a.ld_out + cell;
                a.cfact[oi] = o.cfact;
                if (a.replaced) a.replaced[oi] = (uint8_t)o.replaced;
                if (a.cfact_scaled) a.cfact_scaled[oi] = o.cfact_scaled;
            }
        }
        __syncthreads();
    }
}

struct EvalArgs {
    int T, C, chunk_len, family, modes, which, counterfactual;
    const double* basis64;
    const double* params;
    double* out;
    int64_t ld_out;
    const int* fold_bad;   // [0] = days off the gap-free daily axis (Workspace::fold_bad)
    int skip_if_fold;      // day-ordered kernel: return at once if the phase-ordered one applies
    int n_ph, ph_len;      // phase-ordered kernel: phases, phases per CTA
};

// value of the requested field of the family's distribution (or its expectation) for one day
__device__ __forceinline__ double eval_field(const EvalArgs& a, const DayParams& d) {
    DistPar ref, cf;
    family_params(a.family, d, ref, cf);
    const DistPar& p = a.counterfactual ? cf : ref;
    if (a.which < 0) {
        // Distribution.expectation() (attrici/distributions.py:89-90, 128-129, 157-158, 186-187, 215-216)
        switch (a.family) {
            case ATTRICI_WEIBULL: return p.p1 * tgamma(1.0 + 1.0 / p.p0);   // weibull_min.mean(alpha, scale=beta)
            case ATTRICI_BERNOULLI_GAMMA: return (1.0 - p.p0) * p.p1;
            default: return p.p0;                                           // mu
        }
    }
    return a.which == 0 ? p.p0 : (a.which == 1 ? p.p1 : p.p2);
}

__global__ void __launch_bounds__(QMAP_THREADS) eval_parameter_kernel(EvalArgs a) {
    if (a.skip_if_fold && a.fold_bad[0] == 0) return;
    const int cell = blockIdx.x * QMAP_THREADS + threadIdx.x;
    if (cell >= a.C) return;
    const int t0 = blockIdx.y * a.chunk_len, t1 = min(a.T, t0 + a.chunk_len);
    CellCoef coef;
    load_coef(a.family, a.modes, a.params, cell, coef);
    const int n_dep = (a.family == ATTRICI_BERNOULLI_GAMMA) ? 2 : 1;
    for (int t = t0; t < t1; ++t) {
        const double* row = a.basis64 + (size_t)t * B_STRIDE;
        double r[1 + 2 * NH];
#pragma unroll
        for (int i = 0; i < 1 + 2 * NH; ++i) r[i] = __ldg(row + i);
        DayParams d;
        day_predictors(coef, r, n_dep, d);
        a.out[(size_t)t * a.ld_out + cell] = eval_field(a, d);
    }
}

// The same on a gap-free daily axis, in phase order (see fold.cu): the Fourier sums of the predictors are formed once
// per (cell, phase) - day_predictors with gmt = 0 gives base, with gmt = 1 base + trend - and so is everything that
// does not depend on gmt (the GMT-independent parameter, every counterfactual field).  A day then costs one FMA and
// the link of the one field that is asked for.
enum { EV_CONST = 0, EV_IDENT, EV_EXP, EV_INVLOGIT, EV_EXP_TIMES, EV_BG_MEAN };

__global__ void __launch_bounds__(QMAP_THREADS) eval_parameter_fold_kernel(EvalArgs a) {
    if (a.fold_bad[0] != 0) return;
    const int cell = blockIdx.x * QMAP_THREADS + threadIdx.x;
    if (cell >= a.C) return;
    const int d0 = blockIdx.y * a.ph_len, d1 = min(a.n_ph, d0 + a.ph_len);
    CellCoef coef;
    load_coef(a.family, a.modes, a.params, cell, coef);
    const int n_dep = (a.family == ATTRICI_BERNOULLI_GAMMA) ? 2 : 1;
    // which field: position of the GMT-independent parameter in the family's order, links of the dependent ones
    // (family_params, qmap_impl.cuh)
    const int indep = (a.family == ATTRICI_WEIBULL) ? 0 : (a.family == ATTRICI_BERNOULLI_GAMMA ? 2 : 1);
    int kind, slot = 0;
    if (a.which < 0) {
        kind = (a.family == ATTRICI_WEIBULL) ? EV_EXP_TIMES : (a.family == ATTRICI_BERNOULLI_GAMMA ? EV_BG_MEAN
               : (a.family == ATTRICI_NORMAL ? EV_IDENT : (a.family == ATTRICI_BETA ? EV_INVLOGIT : EV_EXP)));
    } else if (a.which == indep) {
        kind = EV_CONST;
    } else {
        slot = (a.family == ATTRICI_BERNOULLI_GAMMA && a.which == 1) ? 1 : 0;
        kind = (a.family == ATTRICI_NORMAL) ? EV_IDENT
               : ((a.family == ATTRICI_BETA || (a.family == ATTRICI_BERNOULLI_GAMMA && a.which == 0)) ? EV_INVLOGIT : EV_EXP);
    }
    // coefficients in shared memory, [coefficient][thread]: the dynamically indexed CellCoef lives in local memory,
    // and re-reading it for every phase (108 loads) is what the kernel would otherwise spend its time on
    __shared__ double cs[(2 * NA + NB) * QMAP_THREADS];
    {
        double* c = cs + threadIdx.x;
#pragma unroll
        for (int i = 0; i < NA; ++i) { c[i * QMAP_THREADS] = coef.a[0][i]; c[(NA + i) * QMAP_THREADS] = coef.a[1][i]; }
#pragma unroll
        for (int i = 0; i < NB; ++i) c[(2 * NA + i) * QMAP_THREADS] = coef.b[i];
    }
    const double* c0 = cs + threadIdx.x;                       // slot 0
    const double* c1 = cs + NA * QMAP_THREADS + threadIdx.x;   // slot 1 (pr's mu)
    const double* cb = cs + 2 * NA * QMAP_THREADS + threadIdx.x;
    for (int ph = d0; ph < d1; ++ph) {
        const double* row = a.basis64 + (size_t)ph * B_STRIDE;
        // base / trend of the dependent parameters and the independent predictor, in day_predictors' order
        double b0 = c0[0], tr0 = c0[QMAP_THREADS], b1 = c1[0], tr1 = c1[QMAP_THREADS], eb = cb[0];
#pragma unroll
        for (int k = 0; k < MAXM; ++k) {
            const double ck = __ldg(row + 1 + k), sk = __ldg(row + 1 + NH + k);
            b0 = fma(c0[(2 + k) * QMAP_THREADS], ck, b0);
            b0 = fma(c0[(2 + MAXM + k) * QMAP_THREADS], sk, b0);
            tr0 = fma(c0[(2 + 2 * MAXM + k) * QMAP_THREADS], ck, tr0);
            tr0 = fma(c0[(2 + 3 * MAXM + k) * QMAP_THREADS], sk, tr0);
            if (n_dep > 1) {
                b1 = fma(c1[(2 + k) * QMAP_THREADS], ck, b1);
                b1 = fma(c1[(2 + MAXM + k) * QMAP_THREADS], sk, b1);
                tr1 = fma(c1[(2 + 2 * MAXM + k) * QMAP_THREADS], ck, tr1);
                tr1 = fma(c1[(2 + 3 * MAXM + k) * QMAP_THREADS], sk, tr1);
            }
            eb = fma(cb[(1 + k) * QMAP_THREADS], ck, eb);
            eb = fma(cb[(1 + MAXM + k) * QMAP_THREADS], sk, eb);
        }
        // the counterfactual has gmt = 0 on every day (detrend.py:366-370): no trend
        if (a.counterfactual) { tr0 = 0.0; tr1 = 0.0; }
        const double vb = exp(eb);                                                   // the independent parameter
        const double mean_factor = (kind == EV_EXP_TIMES) ? tgamma(1.0 + 1.0 / vb) : 0.0;   // weibull_min.mean
        const double bs = slot ? b1 : b0, ts = slot ? tr1 : tr0;
        // the days of the phase in batches of independent gmt loads (each is a different cache line, and a serial
        // load -> store chain of ~2 900 L2 round trips per thread is what bounded the first version)
        constexpr int EU = 10;
        for (int t = ph; t < a.T; t += EU * NPHASE) {
            double g[EU];
#pragma unroll
            for (int u = 0; u < EU; ++u) {
                const int tt = t + u * NPHASE;
                g[u] = tt < a.T ? __ldg(a.basis64 + (size_t)tt * B_STRIDE) : 0.0;
            }
#pragma unroll
            for (int u = 0; u < EU; ++u) {
                const int tt = t + u * NPHASE;
                if (tt >= a.T) break;
                const double eta = fma(g[u], ts, bs);
                double v;
                switch (kind) {
                    case EV_CONST: v = vb; break;
                    case EV_IDENT: v = eta; break;
                    case EV_EXP: v = exp(eta); break;
                    case EV_INVLOGIT: v = invlogit(eta); break;
                    case EV_EXP_TIMES: v = exp(eta) * mean_factor; break;
                    default: v = (1.0 - invlogit(fma(g[u], tr0, b0))) * exp(fma(g[u], tr1, b1)); break;   // (1 - p) mu
                }
                a.out[(size_t)tt * a.ld_out + cell] = v;
            }
        }
    }
}

// One MT19937 stream per cell (np.random.seed(seed); np.random.rand(n): detrend.py:285, variables.py:458), one WARP
// per stream: the state (624 words) lives in shared memory, the twist runs 32 words at a time - word k of the new
// state needs the old words k, k + 1 and either the old word k + 397 (k < 227) or the NEW word k - 227, which an
// earlier 32-word step has already written - and the lanes temper and convert 32 doubles at a time.  (A thread per
// stream with the state in local memory, the first version, ran at 3 % occupancy: 17.9 ms per 10 000 cells x 43 464
// doubles, as long as a fifth of the whole precipitation step.)
constexpr int MT_WARPS = 16;          // streams (cells) per CTA
constexpr int MT_PITCH = MT_WARPS + 1; // doubles per row of the output stage

// The 312 doubles a state yields are staged per CTA as [312][16 cells] and written out row by row: 128 contiguous
// bytes per day and CTA instead of 8-byte stores 8 Cp bytes apart (measured 3.9 ms with the scattered stores)
__global__ void __launch_bounds__(32 * MT_WARPS) mt19937_kernel(const uint32_t* __restrict__ seeds, int C, int n,
                                                                double* __restrict__ out, int64_t ld) {
    constexpr int N = Mt19937::N, M = Mt19937::M, ND = N / 2;
    extern __shared__ __align__(16) unsigned char mt_smem[];
    uint32_t* state = reinterpret_cast<uint32_t*>(mt_smem);                          // [MT_WARPS][N]
    double* stage = reinterpret_cast<double*>(mt_smem + sizeof(uint32_t) * MT_WARPS * N);   // [ND][MT_PITCH]
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell0 = blockIdx.x * MT_WARPS;
    const int cell = cell0 + w;
    const bool live = cell < C;   // whole warps; idle warps still meet the barriers
    uint32_t* mt = state + w * N;
    if (live && lane == 0) {   // init_genrand: a serial recurrence
        uint32_t p = seeds[cell];
        mt[0] = p;
        for (int i = 1; i < N; ++i) {
            p = 1812433253u * (p ^ (p >> 30)) + (uint32_t)i;
            mt[i] = p;
        }
    }
    __syncwarp();
    const int n_cells = min(MT_WARPS, C - cell0);
    for (int produced = 0; produced < n; produced += ND) {
        if (live) {
            for (int k0 = 0; k0 < N; k0 += 32) {
                const int k = k0 + lane;
                const bool act = k < N - 1;
                uint32_t v = 0;
                if (act) {
                    const uint32_t y = (mt[k] & 0x80000000u) | (mt[k + 1] & 0x7fffffffu);
                    v = mt[k < N - M ? k + M : k + (M - N)] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
                }
                __syncwarp();   // every lane has read the old words of this step (lane 31's k + 1 belongs to the next)
                if (act) mt[k] = v;
                __syncwarp();
            }
            if (lane == 0) {
                const uint32_t y = (mt[N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
                mt[N - 1] = mt[M - 1] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            __syncwarp();
            // random_sample: consecutive word pairs (624 is even: no pair straddles two states)
            for (int i = lane; i < ND; i += 32) {
                const uint2 ab = *reinterpret_cast<const uint2*>(mt + 2 * i);
                stage[i * MT_PITCH + w] = Mt19937::to_double(Mt19937::temper(ab.x), Mt19937::temper(ab.y));
            }
        }
        __syncthreads();
        const int rows = min(ND, n - produced);
        for (int i = threadIdx.x; i < rows * MT_WARPS; i += 32 * MT_WARPS) {
            const int r = i / MT_WARPS, c = i - r * MT_WARPS;
            if (c < n_cells) out[(size_t)(produced + r) * ld + cell0 + c] = stage[r * MT_PITCH + c];
        }
        __syncthreads();
    }
}

int qmap_chunks(int T, int C, int* chunk_len) {
    int n_cb = (C + QMAP_THREADS - 1) / QMAP_THREADS;
    int want = (148 * 8 + n_cb - 1) / n_cb;   // ~8 CTAs of 128 threads per SM
    if (want < 1) want = 1;
    if (want > 256) want = 256;
    int len = (T + want - 1) / want;
    if (len < 16) len = 16;
    *chunk_len = len;
    return (T + len - 1) / len;
}

}  // namespace

int launch_mt19937(const uint32_t* seeds, int C, int n, double* out, int64_t ld, cudaStream_t s) {
    if (C <= 0 || n <= 0) return 0;
    const size_t smem = sizeof(uint32_t) * MT_WARPS * Mt19937::N + sizeof(double) * (Mt19937::N / 2) * MT_PITCH;
    ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(mt19937_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    mt19937_kernel<<<(C + MT_WARPS - 1) / MT_WARPS, 32 * MT_WARPS, smem, s>>>(seeds, C, n, out, ld);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

// ATTRICI_NO_PR_FOLD=1: precipitation keeps the day-ordered kernel (A/B measurements)
static bool pr_fold_off() {
    static const bool off = [] { const char* e = getenv("ATTRICI_NO_PR_FOLD"); return e && e[0] == '1'; }();
    return off;
}

int launch_qmap(Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled,
                int64_t ld, const double* params, const double* scaling, const uint32_t* seeds, double* cfact,
                uint8_t* replaced, double* cfact_scaled, int64_t ld_out, cudaStream_t s, float* cfact32) {
    if (cfact32) {
        // float32 counterfactual: written by the lean phase-ordered Normal kernel only (the caller has checked that
        // the axis is gap-free daily: attrici_quantile_map_ex / attrici_detrend_host_ex)
        return launch_qmap_fold(ws, var, y, y_scaled, ld, params, scaling, nullptr, replaced, cfact_scaled, ld_out, s, cfact32);
    }
    QmapArgs a;
    a.y = y; a.ys = y_scaled; a.ld = ld;
    a.T = ws.T; a.C = ws.C; a.Cp = ws.Cp;
    a.family = var.family; a.scaling = var.scaling; a.post_rule = var.post_rule; a.modes = var.modes;
    a.basis64 = ws.basis64;
    a.basis32 = ws.basis32;
    a.params = params; a.scaling_v = scaling;
    a.uniforms = nullptr;
    if (var.family == ATTRICI_BERNOULLI_GAMMA) {
        if (!seeds) { set_error("pr needs per-cell seeds (detrend.py:285)"); return ATTRICI_EINVAL; }
        if (!ws.uniforms) { set_error("workspace was sized for another family"); return ATTRICI_EWORKSPACE; }
        ATTRICI_TRY(launch_mt19937(seeds, ws.C, ws.T, ws.uniforms, ws.Cp, s));
        a.uniforms = ws.uniforms;
    }
    a.cfact = cfact; a.replaced = replaced; a.cfact_scaled = cfact_scaled; a.ld_out = ld_out;
    a.fold_bad = ws.fold_bad;
    a.want_fold = -1;
    a.ph_len = 1;
    a.has_lo = !(var.lower_threshold != var.lower_threshold);
    a.has_hi = !(var.upper_threshold != var.upper_threshold);
    a.lo = var.lower_threshold; a.hi = var.upper_threshold;
    const int n_cb = (ws.C + QMAP_THREADS - 1) / QMAP_THREADS;
    static const bool fold_on = [] { const char* e = getenv("ATTRICI_NO_FOLD"); return !(e && e[0] == '1'); }();
    if (fold_on && var.family == ATTRICI_NORMAL) {
        // device-side dispatch on the axis flag keeps this call asynchronous: of the phase-ordered kernel
        // (gap-free daily axis) and the day-ordered one below, the one that does not apply returns at once
        ATTRICI_TRY(launch_qmap_fold(ws, var, y, y_scaled, ld, params, scaling, cfact, replaced, cfact_scaled, ld_out, s));
        a.want_fold = 0;
    } else if (fold_on && y_scaled && (var.family == ATTRICI_GAMMA || var.family == ATTRICI_WEIBULL)) {
        ATTRICI_TRY(launch_qmap_fold_scale(ws, var, y, y_scaled, ld, params, scaling, cfact, replaced, cfact_scaled, ld_out, s));
        a.want_fold = 0;
    } else if (fold_on && y_scaled && var.family == ATTRICI_BERNOULLI_GAMMA && !pr_fold_off()) {
        ATTRICI_TRY(launch_qmap_fold_pr(ws, var, y, y_scaled, ld, params, scaling, a.uniforms, cfact, replaced, cfact_scaled, ld_out, s));
        a.want_fold = 0;
    }
    int n_chunks = qmap_chunks(ws.T, ws.C, &a.chunk_len);
    dim3 grid(n_cb, n_chunks);
    qmap_kernel<<<grid, QMAP_THREADS, 0, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_eval_parameter(Workspace& ws, const attrici_variable_t& var, const double* params, int which,
                          int counterfactual, double* out, int64_t ld_out, cudaStream_t s) {
    EvalArgs a;
    a.T = ws.T; a.C = ws.C; a.family = var.family; a.modes = var.modes; a.which = which;
    a.counterfactual = counterfactual;
    a.basis64 = ws.basis64; a.params = params; a.out = out; a.ld_out = ld_out;
    a.fold_bad = ws.fold_bad; a.skip_if_fold = 0;
    a.n_ph = ws.T < NPHASE ? ws.T : NPHASE; a.ph_len = 1;
    const int n_cb = (ws.C + QMAP_THREADS - 1) / QMAP_THREADS;
    static const bool fold_on = [] { const char* e = getenv("ATTRICI_NO_FOLD"); return !(e && e[0] == '1'); }();
    if (fold_on) {
        // device-side dispatch on the axis flag, as in launch_qmap: the kernel that does not apply returns at once
        int want = (148 * 8 + n_cb - 1) / n_cb;
        if (want > a.n_ph) want = a.n_ph;
        if (want < 1) want = 1;
        a.ph_len = (a.n_ph + want - 1) / want;
        eval_parameter_fold_kernel<<<dim3(n_cb, (a.n_ph + a.ph_len - 1) / a.ph_len), QMAP_THREADS, 0, s>>>(a);
        ATTRICI_LAUNCH_CHECK();
        a.skip_if_fold = 1;
    }
    int n_chunks = qmap_chunks(ws.T, ws.C, &a.chunk_len);
    eval_parameter_kernel<<<dim3(n_cb, n_chunks), QMAP_THREADS, 0, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
