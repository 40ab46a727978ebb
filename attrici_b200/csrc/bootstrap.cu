// Block bootstrap of the fitted trend: the pieces of attrici/detrend.py:479-612 that
// are not already the hot path.  A bootstrap sample of a cell is
//     y*_t = F_ref,t^-1( F_ref,s(y_s) ),   s = source day of t under a random permutation of whole calendar years
// (detrend.py:533-557), refitted like the original series (:559-572); the statistics over the samples' fitted
// expectations are the product (:574-612).  The refit IS the hot path: the samples of a batch of cells are laid
// out as extra cells (column v = r C + c of a [T x R C] series) and go through attrici_prep_scale_mask /
// attrici_fit_batch / attrici_eval_parameter unchanged.  This file adds
//   * the year-block choices from NumPy's legacy generator (np.random.randint = masked rejection on successive
//     MT19937 words; one stream per cell, seeded as detrend.py:285) - bit-identical indices,
//   * the factual scores of the observed series and the resampled series,
//   * mean / std / percentiles over the samples.
// Normal: the inverse cdf is ndtri(q) sigma_t + mu_t and the score stored per day is ndtri(ndtr(z)), so that the
// sample uses exactly the reference's round trip including its saturation; the other families store the quantile
// and invert it per sample-day with the functions of the quantile map (qmap_impl.cuh).  pr: a dry source day has
// no quantile (NaN) and a quantile below the target day's dry probability has no inverse (NaN), so the sample
// keeps the observed value there (detrend.py:553-557) - which may itself be dry; the caller skips the generator
// words pr's quantile map consumed before the bootstrap (np.random.rand(T), variables.py:458).
#include "common.cuh"
#include "special.cuh"
#include "qmap_impl.cuh"

namespace attrici {

namespace {

constexpr int BS_THREADS = 128;
constexpr int BS_MAX_SAMPLES = 256;   // samples per (cell, day) the statistics kernel sorts in local memory

__device__ void bs_load_coef(int family, int modes, const double* __restrict__ params, int cell, CellCoef& c) {
    const int P = family_num_params(family, modes);
#pragma unroll
    for (int i = 0; i < NA; ++i) { c.a[0][i] = 0.0; c.a[1][i] = 0.0; }
#pragma unroll
    for (int i = 0; i < NB; ++i) c.b[i] = 0.0;
    for (int j = 0; j < P; ++j) {
        int sl, cn;
        ref_to_canon(family, modes, j, &sl, &cn);
        const double v = params[(size_t)cell * P + j];
        if (cn < NA) c.a[sl][cn] = v;
        else c.b[cn - NA] = v;
    }
}

__device__ __forceinline__ void bs_day_params(int family, const CellCoef& coef, const double* __restrict__ basis_row,
                                              DistPar& ref) {
    double r[1 + 2 * NH];
#pragma unroll
    for (int i = 0; i < 1 + 2 * NH; ++i) r[i] = __ldg(basis_row + i);
    DayParams d;
    day_predictors(coef, r, family == ATTRICI_BERNOULLI_GAMMA ? 2 : 1, d);
    DistPar cf;
    family_params(family, d, ref, cf);
}

// inverse cdf of the non-Normal families, out of line: the batched loop of the resampling kernel is unrolled
__device__ __noinline__ double bs_invcdf(int family, double p0, double p1, double p2, double q) {
    DistPar p;
    p.p0 = p0; p.p1 = p1; p.p2 = p2;
    return dist_invcdf(family, p, q);
}

struct LocalMt {
    uint32_t* p;
    __device__ uint32_t& operator()(int i) const { return p[i]; }
};

// np.random.seed(s); [np.random.randint(0, n_blocks) for _ in range(n_draws)] per cell, after `skip_words`
// 32-bit outputs.  legacy randint(0, n): rng = n - 1, mask = next power of two - 1, draw words until
// (word & mask) <= rng (numpy/random/_bounded_integers: buffered_bounded_masked_uint32); n = 1 draws nothing.
__global__ void __launch_bounds__(64) bootstrap_choices_kernel(const uint32_t* __restrict__ seeds, int C, int n_draws,
                                                               int n_blocks, int skip_words, int16_t* __restrict__ choices) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (cell >= C) return;
    int16_t* out = choices + (size_t)cell * n_draws;
    const uint32_t rng = (uint32_t)(n_blocks - 1);
    if (rng == 0) {
        for (int i = 0; i < n_draws; ++i) out[i] = 0;
        return;
    }
    uint32_t mask = rng;
    mask |= mask >> 1; mask |= mask >> 2; mask |= mask >> 4; mask |= mask >> 8; mask |= mask >> 16;
    uint32_t mt[Mt19937::N];
    LocalMt st{mt};
    Mt19937::seed(st, seeds[cell]);
    int pos = Mt19937::N;
    auto next = [&]() {
        if (pos >= Mt19937::N) { Mt19937::twist(st); pos = 0; }
        return Mt19937::temper(mt[pos++]);
    };
    for (int i = 0; i < skip_words; ++i) next();
    for (int i = 0; i < n_draws;) {
        const uint32_t v = next() & mask;
        if (v <= rng) out[i++] = (int16_t)v;
    }
}

struct ScoreArgs {
    int T, C, chunk_len, family, modes;
    const double* basis64; const double* params;
    const float* ys; int64_t ld;
    double* scores; int64_t ld_s;
};

// scores[t][c]: what the resampling carries from the source day (detrend.py:534): the factual quantile of the
// observation, kept for the Normal family as ndtri(quantile)
__global__ void __launch_bounds__(BS_THREADS) bootstrap_scores_kernel(ScoreArgs a) {
    const int cell = blockIdx.x * BS_THREADS + threadIdx.x;
    if (cell >= a.C) return;
    const int t0 = blockIdx.y * a.chunk_len, t1 = min(a.T, t0 + a.chunk_len);
    CellCoef coef;
    bs_load_coef(a.family, a.modes, a.params, cell, coef);
    for (int t = t0; t < t1; ++t) {
        DistPar ref;
        bs_day_params(a.family, coef, a.basis64 + (size_t)t * B_STRIDE, ref);
        const double y = (double)a.ys[(size_t)t * a.ld + cell];
        double q = dist_cdf(a.family, ref, y);
        if (a.family == ATTRICI_NORMAL) q = ndtri(q);
        a.scores[(size_t)t * a.ld_s + cell] = q;
    }
}

struct ResampleArgs {
    int T, C, chunk_len, family, modes;
    const double* basis64; const double* params;
    const float* ys; int64_t ld;
    const double* scores; int64_t ld_s;
    const int16_t* choices; int n_samples, n_blocks, r0, R;
    const int32_t* block_start; const int16_t* block_of_day;
    float* y_new; int64_t ld_new;
    int32_t* replaced; int64_t ld_rc;
};

// y_new[t][r C + c] for the samples r0 .. r0 + R - 1 (detrend.py:541-557).  The source day of t: same offset
// within the randomly chosen year block; a chosen block shorter than the target year continues into the next
// block, or - if it is the last one - starts correspondingly earlier (get_random_block_indices, :500-531).
__global__ void __launch_bounds__(BS_THREADS) bootstrap_resample_kernel(ResampleArgs a) {
    const int cell = blockIdx.x * BS_THREADS + threadIdx.x;
    if (cell >= a.C) return;
    const int t0 = blockIdx.y * a.chunk_len, t1 = min(a.T, t0 + a.chunk_len);
    CellCoef coef;
    bs_load_coef(a.family, a.modes, a.params, cell, coef);
    const int16_t* ch = a.choices + (size_t)cell * a.n_samples * a.n_blocks;
    for (int t = t0; t < t1; ++t) {
        DistPar ref;
        bs_day_params(a.family, coef, a.basis64 + (size_t)t * B_STRIDE, ref);
        const float y0 = a.ys[(size_t)t * a.ld + cell];
        const int yb = a.block_of_day[t];
        const int b0 = a.block_start[yb];
        const int target = a.block_start[yb + 1] - b0;
        const int r_in = t - b0;
        int count = 0;
        // samples in batches: the choice -> block start -> score loads of one sample form a dependent chain of L2 round
        // trips, and the store that ends it may alias the next sample's loads as far as the compiler knows - so the
        // loads of RU samples are issued together before any of their stores
        constexpr int RU = 8;
        for (int r = 0; r < a.R; r += RU) {
            int cidx[RU], src[RU];
            double sc[RU];
#pragma unroll
            for (int u = 0; u < RU; ++u) cidx[u] = __ldg(ch + (size_t)(a.r0 + min(r + u, a.R - 1)) * a.n_blocks + yb);
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                int start = __ldg(a.block_start + cidx[u]);
                const int len = __ldg(a.block_start + cidx[u] + 1) - start;
                if (len < target && cidx[u] >= a.n_blocks - 1) start -= target - len;
                src[u] = start + r_in;
            }
#pragma unroll
            for (int u = 0; u < RU; ++u)
                sc[u] = (src[u] >= 0 && src[u] < a.T) ? __ldg(a.scores + (size_t)src[u] * a.ld_s + cell) : NAN;
#pragma unroll
            for (int u = 0; u < RU; ++u) {
                if (r + u < a.R) {
                    double v = sc[u];
                    if (!isnan(v)) v = (a.family == ATTRICI_NORMAL) ? sc[u] * ref.p1 + ref.p0 : bs_invcdf(a.family, ref.p0, ref.p1, ref.p2, sc[u]);
                    float out = (float)v;
                    if (isnan(v) || isinf(v)) { out = y0; ++count; }   // detrend.py:553-557
                    a.y_new[(size_t)t * a.ld_new + (size_t)(r + u) * a.C + cell] = out;
                }
            }
        }
        if (a.replaced) a.replaced[(size_t)t * a.ld_rc + cell] += count;
    }
}

struct StatsArgs {
    int T, C, N, n_levels, scaling, unity;
    const double* E; int64_t ld_e;
    const double* scaling_v; const double* levels;
    double* mean; double* sd; double* quant;
};

// per (day, cell): mean, standard deviation (ddof 0) and percentiles (NumPy's default linear interpolation) of the
// N rescaled expectations (detrend.py:574-612).  thread = cell, blockIdx.y = a run of consecutive days: the fitted
// expectation of a sample is a smooth function of the day, so the order of the samples changes slowly from one day
// to the next - the thread keeps the samples' permutation of the previous day and re-sorts by insertion, which
// costs O(N) on a nearly sorted sequence instead of O(N^2) (54 -> 4 ms for 100 cells x 100 samples x 43 464 days).
constexpr int BS_STATS_DAYS = 64;

// Variable.rescale (variables.py:114-131; Pr.rescale :483-486 maps non-positive scaled values to 0): multiply, then
// add - two roundings like NumPy
__device__ __forceinline__ double bs_rescale(double x, double mul, double add, bool is_pr) {
    const double r = __dadd_rn(__dmul_rn(x, mul), add);
    return (is_pr && x <= 0.0) ? 0.0 : r;
}

__global__ void __launch_bounds__(BS_THREADS) bootstrap_stats_kernel(StatsArgs a) {
    const int cell = blockIdx.x * BS_THREADS + threadIdx.x;
    if (cell >= a.C) return;
    const int t0 = blockIdx.y * BS_STATS_DAYS, t1 = min(a.T, t0 + BS_STATS_DAYS);
    const double datamin = a.scaling_v[2 * (size_t)cell], scale = a.scaling_v[2 * (size_t)cell + 1];
    const double mul = (a.scaling == ATTRICI_SCALE_NONE) ? 1.0 : scale;
    const bool is_pr = a.scaling == ATTRICI_SCALE_PR;
    const double add = a.unity ? datamin : (is_pr ? 0.0000011574 : 0.0);   // Pr.THRESHOLD, variables.py:350, 483-486
    double v[BS_MAX_SAMPLES];
    unsigned char perm[BS_MAX_SAMPLES];
    for (int r = 0; r < a.N; ++r) perm[r] = (unsigned char)r;
    for (int t = t0; t < t1; ++t) {
        const double* e = a.E + (size_t)t * a.ld_e + cell;
        // mean in sample order (numpy.mean over axis 0 adds the samples one after the other)
        double sum = 0.0;
        for (int r = 0; r < a.N; ++r) sum += bs_rescale(e[(size_t)r * a.C], mul, add, is_pr);
        const double mean = sum / (double)a.N;
        // values in yesterday's order, then insertion sort of (value, sample) pairs
        double ss = 0.0;
        for (int k = 0; k < a.N; ++k) {
            // rescale (variables.py:114-131): multiply, then add - two roundings like NumPy
            const double x = bs_rescale(e[(size_t)perm[k] * a.C], mul, add, is_pr);
            v[k] = x;
            const double d = x - mean;
            ss += d * d;
        }
        for (int i = 1; i < a.N; ++i) {
            const double x = v[i];
            if (!(v[i - 1] > x)) continue;
            const unsigned char px = perm[i];
            int j = i - 1;
            while (j >= 0 && v[j] > x) { v[j + 1] = v[j]; perm[j + 1] = perm[j]; --j; }
            v[j + 1] = x;
            perm[j + 1] = px;
        }
        const size_t o = (size_t)t * a.C + cell;
        a.mean[o] = mean;
        a.sd[o] = sqrt(ss / (double)a.N);
        for (int k = 0; k < a.n_levels; ++k) {
            // numpy.percentile, method "linear": virtual index q (N - 1), lerp between the neighbours
            const double vi = a.levels[k] * (double)(a.N - 1);
            int lo = (int)floor(vi);
            lo = max(0, min(lo, a.N - 1));
            const int hi = min(lo + 1, a.N - 1);
            const double g = vi - (double)lo;
            const double lo_v = v[lo], hi_v = v[hi], diff = hi_v - lo_v;
            double res = __dadd_rn(lo_v, __dmul_rn(diff, g));
            if (g >= 0.5) res = __dsub_rn(hi_v, __dmul_rn(diff, 1.0 - g));   // numpy _lerp
            if (diff == 0.0) res = lo_v;
            a.quant[((size_t)k * a.T + t) * a.C + cell] = res;
        }
    }
}

int bs_chunks(int T, int C, int* chunk_len) {
    const int n_cb = (C + BS_THREADS - 1) / BS_THREADS;
    int want = (148 * 8 + n_cb - 1) / n_cb;
    if (want < 1) want = 1;
    if (want > 1024) want = 1024;
    int len = (T + want - 1) / want;
    if (len < 8) len = 8;
    *chunk_len = len;
    return (T + len - 1) / len;
}

int bs_bind(Workspace& ws, void* workspace, size_t workspace_bytes, int T, int C, int family) {
    if (!workspace) { set_error("workspace is null"); return ATTRICI_EINVAL; }
    if (((uintptr_t)workspace & 255) != 0) { set_error("workspace must be 256-byte aligned"); return ATTRICI_EINVAL; }
    const size_t need = ws.layout(T, C, family, workspace);
    if (need > workspace_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
        return ATTRICI_EWORKSPACE;
    }
    return 0;
}

int bs_check_family(const attrici_variable_t* var) {
    if (!var) { set_error("var is null"); return ATTRICI_EINVAL; }
    if (var->family < ATTRICI_NORMAL || var->family > ATTRICI_BERNOULLI_GAMMA) { set_error("unknown family %d", var->family); return ATTRICI_EINVAL; }
    if (var->modes < 1 || var->modes > ATTRICI_MAX_MODES) { set_error("bad modes %d", var->modes); return ATTRICI_EINVAL; }
    return 0;
}

}  // namespace
}  // namespace attrici

using namespace attrici;

extern "C" {

int attrici_bootstrap_block_choices(const uint32_t* seeds, int C, int n_draws, int n_blocks, int skip_words,
                                    int16_t* choices, void* stream) {
    if (!seeds || !choices) { set_error("seeds / choices is null"); return ATTRICI_EINVAL; }
    if (C < 1 || n_draws < 1 || n_blocks < 1 || n_blocks > 32767 || skip_words < 0) {
        set_error("bad shape C=%d n_draws=%d n_blocks=%d skip_words=%d", C, n_draws, n_blocks, skip_words);
        return ATTRICI_EINVAL;
    }
    bootstrap_choices_kernel<<<(C + 63) / 64, 64, 0, (cudaStream_t)stream>>>(seeds, C, n_draws, n_blocks, skip_words, choices);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int attrici_bootstrap_scores(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                             const float* y_scaled, int64_t ld, int T, int C, const double* params, double* scores,
                             int64_t ld_s, void* stream) {
    ATTRICI_TRY(bs_check_family(var));
    if (!y_scaled || !params || !scores) { set_error("y_scaled / params / scores is null"); return ATTRICI_EINVAL; }
    if (T < 1 || C < 1 || ld < C || ld_s < C) { set_error("bad shape T=%d C=%d", T, C); return ATTRICI_EINVAL; }
    Workspace ws;
    ATTRICI_TRY(bs_bind(ws, workspace, workspace_bytes, T, C, var->family));
    ScoreArgs a;
    a.T = T; a.C = C; a.family = var->family; a.modes = var->modes;
    a.basis64 = ws.basis64; a.params = params; a.ys = y_scaled; a.ld = ld; a.scores = scores; a.ld_s = ld_s;
    const int n_chunks = bs_chunks(T, C, &a.chunk_len);
    bootstrap_scores_kernel<<<dim3((C + BS_THREADS - 1) / BS_THREADS, n_chunks), BS_THREADS, 0, (cudaStream_t)stream>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int attrici_bootstrap_resample(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                               const float* y_scaled, int64_t ld, int T, int C, const double* params,
                               const double* scores, int64_t ld_s, const int16_t* choices, int n_samples, int n_blocks,
                               int r0, int R, const int32_t* block_start, const int16_t* block_of_day, float* y_new,
                               int64_t ld_new, int32_t* replaced_count, int64_t ld_rc, void* stream) {
    ATTRICI_TRY(bs_check_family(var));
    if (!y_scaled || !params || !scores || !choices || !block_start || !block_of_day || !y_new) {
        set_error("a required pointer is null");
        return ATTRICI_EINVAL;
    }
    if (T < 1 || C < 1 || ld < C || ld_s < C || R < 1 || r0 < 0 || r0 + R > n_samples || n_blocks < 1 ||
        ld_new < (int64_t)R * C || (replaced_count && ld_rc < C)) {
        set_error("bad shape T=%d C=%d R=%d r0=%d n_samples=%d n_blocks=%d", T, C, R, r0, n_samples, n_blocks);
        return ATTRICI_EINVAL;
    }
    Workspace ws;
    ATTRICI_TRY(bs_bind(ws, workspace, workspace_bytes, T, C, var->family));
    ResampleArgs a;
    a.T = T; a.C = C; a.family = var->family; a.modes = var->modes;
    a.basis64 = ws.basis64; a.params = params; a.ys = y_scaled; a.ld = ld; a.scores = scores; a.ld_s = ld_s;
    a.choices = choices; a.n_samples = n_samples; a.n_blocks = n_blocks; a.r0 = r0; a.R = R;
    a.block_start = block_start; a.block_of_day = block_of_day;
    a.y_new = y_new; a.ld_new = ld_new; a.replaced = replaced_count; a.ld_rc = ld_rc;
    const int n_chunks = bs_chunks(T, C, &a.chunk_len);
    bootstrap_resample_kernel<<<dim3((C + BS_THREADS - 1) / BS_THREADS, n_chunks), BS_THREADS, 0, (cudaStream_t)stream>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int attrici_bootstrap_stats(const attrici_variable_t* var, const double* expectations, int64_t ld_e, int T, int C,
                            int n_samples, const double* scaling, const double* levels, int n_levels, double* mean,
                            double* std_dev, double* quantiles, void* stream) {
    ATTRICI_TRY(bs_check_family(var));
    if (!expectations || !scaling || !levels || !mean || !std_dev || !quantiles) {
        set_error("a required pointer is null");
        return ATTRICI_EINVAL;
    }
    if (T < 1 || C < 1 || n_samples < 1 || n_samples > BS_MAX_SAMPLES || n_levels < 1 || ld_e < (int64_t)n_samples * C) {
        set_error("bad shape T=%d C=%d n_samples=%d (max %d) n_levels=%d", T, C, n_samples, BS_MAX_SAMPLES, n_levels);
        return ATTRICI_EINVAL;
    }
    StatsArgs a;
    a.T = T; a.C = C; a.N = n_samples; a.n_levels = n_levels; a.scaling = var->scaling;
    a.unity = var->scaling == ATTRICI_SCALE_UNITY;
    a.E = expectations; a.ld_e = ld_e; a.scaling_v = scaling; a.levels = levels;
    a.mean = mean; a.sd = std_dev; a.quant = quantiles;
    bootstrap_stats_kernel<<<dim3((C + BS_THREADS - 1) / BS_THREADS, (T + BS_STATS_DAYS - 1) / BS_STATS_DAYS), BS_THREADS, 0,
                             (cudaStream_t)stream>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // extern "C"
