// Quantile-map kernels: thread = cell (coalesced 128-byte lines of the time-major series), CTA = 128
// cells x one time chunk; FP64 throughout like the reference (scipy.stats cdf / ppf).
// Mathematics in qmap_impl.cuh; reference lines cited there.
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

__global__ void __launch_bounds__(QMAP_THREADS) qmap_kernel(QmapArgs a) {
    const int cell = blockIdx.x * QMAP_THREADS + threadIdx.x;
    if (cell >= a.C) return;
    const int t0 = blockIdx.y * a.chunk_len, t1 = min(a.T, t0 + a.chunk_len);
    CellCoef coef;
    load_coef(a.family, a.modes, a.params, cell, coef);
    const int n_dep = (a.family == ATTRICI_BERNOULLI_GAMMA) ? 2 : 1;
    const double datamin = a.scaling_v[2 * (size_t)cell], scale = a.scaling_v[2 * (size_t)cell + 1];
    const bool mask_in_place = (a.scaling == ATTRICI_SCALE_PERCENT || a.scaling == ATTRICI_SCALE_RANGE);
    const bool normal = (a.family == ATTRICI_NORMAL);
    float af[NA], bf[NB];
#pragma unroll
    for (int i = 0; i < NA; ++i) af[i] = (float)coef.a[0][i];
#pragma unroll
    for (int i = 0; i < NB; ++i) bf[i] = (float)coef.b[i];
    for (int t = t0; t < t1; ++t) {
        const double* row = a.basis64 + (size_t)t * B_STRIDE;
        float yv = a.y[(size_t)t * a.ld + cell];
        const float sv = a.ys[(size_t)t * a.ld + cell];
        // the caller-visible series after Inf -> NaN (detrend.py:311) and the in-place masks of
        // hurs / tasrange / sfcWind (variables.py:583, 723, 771)
        if (isinf(yv) || (mask_in_place && isnan(sv))) yv = __int_as_float(0x7fc00000);
        QmapOut o;
        double cs_fast;
        if (normal && normal_fast_path(coef, af, bf, a.basis32 + (size_t)t * B_STRIDE, row, sv, &cs_fast)) {
            o = qmap_finish(a.scaling, a.post_rule, cs_fast, yv, datamin, scale);
        } else {
            double r[1 + 2 * NH];
#pragma unroll
            for (int i = 0; i < 1 + 2 * NH; ++i) r[i] = __ldg(row + i);
            DayParams d;
            day_predictors(coef, r, n_dep, d);
            DistPar ref, cf;
            family_params(a.family, d, ref, cf);
            const double u = a.uniforms ? a.uniforms[(size_t)t * a.Cp + cell] : 0.0;
            o = qmap_day(a.family, a.scaling, a.post_rule, yv, sv, ref, cf, datamin, scale, u);
        }
        const size_t oi = (size_t)t * a.ld_out + cell;
        a.cfact[oi] = o.cfact;
        if (a.replaced) a.replaced[oi] = (uint8_t)o.replaced;
        if (a.cfact_scaled) a.cfact_scaled[oi] = o.cfact_scaled;
    }
}

struct EvalArgs {
    int T, C, chunk_len, family, modes, which, counterfactual;
    const double* basis64;
    const double* params;
    double* out;
    int64_t ld_out;
};

__global__ void __launch_bounds__(QMAP_THREADS) eval_parameter_kernel(EvalArgs a) {
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
        DistPar ref, cf;
        family_params(a.family, d, ref, cf);
        const DistPar& p = a.counterfactual ? cf : ref;
        a.out[(size_t)t * a.ld_out + cell] = a.which == 0 ? p.p0 : (a.which == 1 ? p.p1 : p.p2);
    }
}

// one MT19937 stream per cell; the state lives in (L1-cached, cell-interleaved) local memory
struct LocalState {
    uint32_t* p;
    __device__ uint32_t& operator()(int i) const { return p[i]; }
};

__global__ void __launch_bounds__(64) mt19937_kernel(const uint32_t* __restrict__ seeds, int C, int n,
                                                     double* __restrict__ out, int64_t ld) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= C) return;
    uint32_t mt[Mt19937::N];
    LocalState st{mt};
    Mt19937::seed(st, seeds[cell]);
    int pos = Mt19937::N;
    uint32_t pending = 0;
    bool have = false;
    int produced = 0;
    while (produced < n) {
        if (pos >= Mt19937::N) { Mt19937::twist(st); pos = 0; }
        uint32_t v = Mt19937::temper(mt[pos++]);
        if (!have) { pending = v; have = true; }
        else {
            out[(size_t)produced * ld + cell] = Mt19937::to_double(pending, v);
            ++produced;
            have = false;
        }
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
    mt19937_kernel<<<(C + 63) / 64, 64, 0, s>>>(seeds, C, n, out, ld);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_qmap(Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled,
                int64_t ld, const double* params, const double* scaling, const uint32_t* seeds, double* cfact,
                uint8_t* replaced, double* cfact_scaled, int64_t ld_out, cudaStream_t s) {
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
    int n_chunks = qmap_chunks(ws.T, ws.C, &a.chunk_len);
    dim3 grid((ws.C + QMAP_THREADS - 1) / QMAP_THREADS, n_chunks);
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
    int n_chunks = qmap_chunks(ws.T, ws.C, &a.chunk_len);
    dim3 grid((ws.C + QMAP_THREADS - 1) / QMAP_THREADS, n_chunks);
    eval_parameter_kernel<<<grid, QMAP_THREADS, 0, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
