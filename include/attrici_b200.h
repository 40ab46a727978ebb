/*
 * attrici_b200 -- C ABI of the B200-native batched solver for ATTRICI's per-gridcell
 * detrending path (scale -> MAP fit -> factual/counterfactual parameters -> quantile map).
 *
 * This is the drop-in boundary.  The reference has no FFI (it is pure Python); each entry
 * point below names the reference code it replaces (paths relative to the reference repo).
 * The Python binding a maintainer adds is the ctypes stub in INTEGRATION.md
 * (attrici_b200/_lib.py is that stub, shipped).
 *
 * Conventions
 *   - every array pointer is a DEVICE pointer unless the parameter name ends in _host;
 *     the caller (torch / cudaMalloc) owns every buffer and the library never allocates device
 *     memory.  What it does create: attrici_detrend_host makes and destroys two CUDA streams per
 *     call, and one 256-byte pinned, mapped host mailbox per process (cudaHostAlloc on first use,
 *     kept) carries the small device -> host read-backs of the fit;
 *   - cell series are TIME-MAJOR: element (day t, cell c) of a [T x C] array lives at
 *     t*ld + c, ld >= C (the layout a (time, lat, lon) NetCDF variable already has after
 *     the reference's obs_data.stack(latlon=("lat","lon")), attrici/detrend.py:672);
 *   - per-cell vectors are [C]; per-cell parameter vectors are row-major [C x P] in the
 *     reference's theta layout (attrici/estimation/model_scipy.py:410-421, 104-113, 225-234);
 *   - all calls are asynchronous on `stream` (a cudaStream_t passed as void*), return 0 on
 *     success, a negative ATTRICI_E* code for invalid arguments, or a positive cudaError_t;
 *     attrici_last_error() gives a thread-local message;  numerical trouble in one cell is
 *     reported in status[c], never as a call failure;
 *   - the library is re-entrant per stream.  Process-wide state is limited to diagnostics and plumbing that
 *     do not influence results: the launch counter (atomic), the thread-local last-error text and fit
 *     profile, the mailbox above (mutex-guarded), and the ATTRICI_* tuning variables read once from the
 *     environment.
 */
#ifndef ATTRICI_B200_H
#define ATTRICI_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ATTRICI_ABI_VERSION 2
#if defined(__GNUC__)
#define ATTRICI_API __attribute__((visibility("default")))
#else
#define ATTRICI_API
#endif
#define ATTRICI_MAX_MODES 4

/* distributions (attrici/distributions.py:64-216; attrici/variables.py model table) */
enum {
    ATTRICI_NORMAL = 0,          /* tas ps rlds rsds tasskew : mu (identity, GMT-dep), sigma (exp)      */
    ATTRICI_GAMMA = 1,           /* tasrange                 : mu (exp, dep), nu (exp)                   */
    ATTRICI_BETA = 2,            /* hurs                     : mu (invlogit, dep), phi (exp)             */
    ATTRICI_WEIBULL = 3,         /* sfcWind / wind           : alpha (exp), beta (exp, dep); theta=[alpha|beta] */
    ATTRICI_BERNOULLI_GAMMA = 4  /* pr                       : p (invlogit, dep), mu (exp, dep), nu (exp) */
};

/* scaling rules (attrici/variables.py:92-111, 363-403, 574-591, 619-623, 722-732, 770-780) */
enum {
    ATTRICI_SCALE_UNITY = 0,   /* (y - min) / (max - min)                       tas ps rlds rsds     */
    ATTRICI_SCALE_NONE = 1,    /* mask a copy with thresholds, no scaling        tasskew              */
    ATTRICI_SCALE_PERCENT = 2, /* mask in place, y / 100                         hurs                 */
    ATTRICI_SCALE_RANGE = 3,   /* mask in place, y / (max - min)                 tasrange sfcWind     */
    ATTRICI_SCALE_PR = 4       /* y - THRESHOLD, <=0 -> dry (NaN), / gamma-fit scale   pr             */
};

/* post rule of the quantile map (attrici/variables.py:645-653, 691-698) */
enum { ATTRICI_POST_NONE = 0, ATTRICI_POST_LE0_NAN = 1, ATTRICI_POST_OUTSIDE01_NAN = 2 };

/* per-cell fit status */
enum {
    ATTRICI_STATUS_CONVERGED = 0,
    ATTRICI_STATUS_RUNNING = 1,   /* only seen if max_pass was hit */
    ATTRICI_STATUS_STALLED = 2,   /* damping exceeded its bound without an acceptable step */
    ATTRICI_STATUS_NONFINITE = 3, /* objective not finite at the starting point */
    ATTRICI_STATUS_NO_DATA = 4    /* no valid day in the fit window (reference skips: detrend.py:315-317) */
};

/* `replaced` codes (the reference stores NaN / +Inf / -Inf as float64, detrend.py:397-411) */
enum { ATTRICI_REPLACED_NONE = 0, ATTRICI_REPLACED_NAN = 1, ATTRICI_REPLACED_PINF = 2, ATTRICI_REPLACED_NINF = 3 };

enum {
    ATTRICI_EINVAL = -1,     /* bad argument */
    ATTRICI_EWORKSPACE = -2, /* workspace too small */
    ATTRICI_EUNSUPPORTED = -3
};

/* what create_variable() + Variable.create_model() decide for one variable name
 * (attrici/variables.py:801-834 and the create_model overrides) */
typedef struct attrici_variable {
    int32_t family;     /* ATTRICI_NORMAL ... */
    int32_t scaling;    /* ATTRICI_SCALE_* */
    int32_t post_rule;  /* ATTRICI_POST_* */
    int32_t modes;      /* 1..ATTRICI_MAX_MODES (Config.modes, attrici/detrend.py:66) */
    float lower_threshold; /* values <= this are masked; NaN = none (mask_thresholded, variables.py:40-68) */
    float upper_threshold; /* values >= this are masked; NaN = none */
} attrici_variable_t;

typedef struct attrici_fit_options {
    int32_t max_pass;     /* cap on likelihood passes per cell (default 60) */
    int32_t use_fp64;     /* 0: FP32 per-day arithmetic, FP64 reductions (production); 1: all FP64 */
    double tol_pred;      /* stop when the predicted decrease <= tol_pred * max(|f|,1) (default 1e-11) */
    double lambda0;       /* initial Levenberg-Marquardt damping (default 1e-3) */
    int32_t zero_init;    /* 1: start from theta = 0 like model_scipy.py:113,234; 0: moment-based start */
    int32_t profile;      /* 1: time every likelihood pass with CUDA events (see attrici_fit_profile) */
} attrici_fit_options_t;

ATTRICI_API const char* attrici_last_error(void);
ATTRICI_API int attrici_abi_version(void);

/* diagnostics (bench.py): kernels launched by this process so far; device time of the likelihood
 * passes of the profiled fits of this thread (options.profile = 1) with the work they covered
 * (cell_days = sum over launches of cells still iterating x days in the fit window) */
ATTRICI_API long long attrici_launch_count(void);
ATTRICI_API int attrici_fit_profile(int reset, double* pass_ms, long long* pass_launches, double* cell_days);

/* number of fitted coefficients of a family (model_scipy.py:113, 234, 410-421) */
ATTRICI_API int attrici_num_params(int family, int modes);

/* defaults of the fit options (max_pass 60, FP32 day arithmetic, tol_pred 1e-11, lambda0 1e-3) */
ATTRICI_API int attrici_default_fit_options(attrici_fit_options_t* opts);

/* bytes of device scratch needed by the calls below for a [T x C] batch of one family
 * (the Bernoulli-Gamma family also holds the [T x C] MT19937 uniforms of variables.py:458) */
ATTRICI_API int attrici_workspace_bytes(int family, int T, int C, size_t* bytes);

/* Fourier / GMT basis shared by all cells.
 * replaces attrici/util.py:85-104 (calc_oscillations) and the covariate construction of
 * attrici/estimation/model_scipy.py:180-197, 280-289.
 *   days[T]  int32 day offsets of the time axis (days since its first day; integers from the host)
 *   gmt[T]   float64 scaled predictor (attrici/detrend.py:698-707) */
ATTRICI_API int attrici_build_basis(void* workspace, size_t workspace_bytes, int family, const int32_t* days,
                        const double* gmt, int T, int C, void* stream);

/* scaling + validity / wet-dry masks.
 * replaces create_variable -> <Variable>.__init__ (attrici/variables.py:92-111, 354-403, 565-591,
 * 619-623, 712-732, 760-780) and the Inf -> NaN step of attrici/detrend.py:311.
 *   y[T x ld]         float32 observations
 *   y_scaled[T x ld]  float32 out; NaN = masked / dry / missing (bit-exact float32 arithmetic).
 *                     May be NULL for the Normal family: the scaled series is then not materialised (the fit
 *                     runs on the per-phase sums gathered here, the quantile map recomputes the values)
 *   scaling[C x 2]    float64 out: {datamin, scale} (datamin = 0 where the reference has none)
 * Also gathers, into the workspace, the per-cell start values / phase origins of the fit window [i0, i1) and,
 * for the Normal family, the per-phase sums {sum y, sum y^2, sum y gmt} (FP64) of the phase-folded likelihood
 * passes (used when the time axis is gap-free daily, as attrici_build_basis found). */
ATTRICI_API int attrici_prep_scale_mask(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                            const float* y, float* y_scaled, int64_t ld, int T, int C, int i0, int i1,
                            double* scaling, void* stream);

/* The same with flags.  With y_scaled == NULL (Normal family, gap-free daily axis, T >= 1461) the default gathers the
 * per-phase sums in ONE pass over the series, together with the min / max of variables.py:108-110: sums of
 * y - pivot, mapped to the scaled values by a per-cell affine map inside the fit.  Those are the sums of the
 * UNROUNDED scaled values, where the reference fits y_scaled rounded to float32: logp moves by ~2e-9 relative.
 * ATTRICI_PREP_EXACT_SUMS keeps the two-pass form (min / max, then the sums of the float32 y_scaled). */
#define ATTRICI_PREP_EXACT_SUMS 1
ATTRICI_API int attrici_prep_scale_mask_ex(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                               const float* y, float* y_scaled, int64_t ld, int T, int C, int i0, int i1,
                               double* scaling, int flags, void* stream);

/* log-posterior, gradient and Fisher information at given coefficients (known-answer entry).
 * replaces DistributionScipy.log_likelihood (model_scipy.py:311-331) summed as in :520; the
 * reference has no gradient (finite differences inside scipy's L-BFGS-B).
 *   theta[C x P] in ; nll[C] out (= -log posterior) ; grad[C x P] out ; fisher[C x P x P] out (nullable)
 * needs attrici_build_basis + attrici_prep_scale_mask on the same workspace first. */
ATTRICI_API int attrici_nll_grad(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                     const float* y_scaled, int64_t ld, int T, int C, int i0, int i1, int use_fp64,
                     const double* theta, double* nll, double* grad, double* fisher, void* stream);

/* batched MAP fit: Fisher-scoring Levenberg-Marquardt with per-cell convergence masks.
 * replaces ModelScipy.fit (model_scipy.py:505-527: scipy L-BFGS-B with finite differences).
 *   params[C x P] out, logp[C] out (log posterior, FP64 evaluation at the returned point),
 *   status[C] out, n_pass[C] out (nullable).  Synchronises `stream` internally: once for the Normal family on a
 *   gap-free daily axis (the whole iteration of a term is one persistent kernel), once per block of passes
 *   otherwise.  logp of a converged cell of that fused path is the objective of the last evaluated point, minus the
 *   predicted decrease of the final step when that step was applied unevaluated (<= 1e-8 |logp|, see step_impl.cuh).
 *   y_scaled may be NULL for the Normal family on a gap-free daily axis when attrici_prep_scale_mask ran on
 *   this workspace with the same window (the passes then read the per-phase sums); otherwise ATTRICI_EINVAL. */
ATTRICI_API int attrici_fit_batch(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                      const float* y_scaled, int64_t ld, int T, int C, int i0, int i1,
                      const attrici_fit_options_t* opts, double* params, double* logp, int32_t* status,
                      int32_t* n_pass, void* stream);

/* factual / counterfactual distribution parameters, quantile map, rescale, replacement.
 * replaces ModelScipy.estimate_distribution x2 (model_scipy.py:547-570; attrici/detrend.py:362-370),
 * Variable.quantile_mapping and overrides (attrici/variables.py:287-303, 422-480, 645-653, 691-698),
 * attrici/distributions.py:81-216, rescale (variables.py:71-89, 114-131, 483-486) and the NaN/Inf
 * replacement of attrici/detrend.py:392-411.
 *   y, y_scaled [T x ld]   ; params[C x P] ; scaling[C x 2]     (y_scaled may be NULL for the Normal family:
 *                            it is recomputed from y and scaling with the float32 arithmetic of the scaling step)
 *   seeds[C] uint32  (pr only; attrici/detrend.py:285, consumed by np.random.rand in variables.py:458)
 *   cfact[T x ld_out] float64 out ; replaced[T x ld_out] uint8 out (nullable) ;
 *   cfact_scaled[T x ld_out] float64 out (nullable) */
ATTRICI_API int attrici_quantile_map(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                         const float* y, const float* y_scaled, int64_t ld, int T, int C,
                         const double* params, const double* scaling, const uint32_t* seeds,
                         double* cfact, uint8_t* replaced, double* cfact_scaled, int64_t ld_out,
                         void* stream);

/* per-day distribution parameters for report_variables (attrici/detrend.py:462-467):
 *   which = index of the parameter in the family's order, or ATTRICI_WHICH_EXPECTATION for
 *   Distribution.expectation() (attrici/distributions.py:89-216); counterfactual != 0 -> predictor = 0 */
#define ATTRICI_WHICH_EXPECTATION (-1)
ATTRICI_API int attrici_eval_parameter(void* workspace, size_t workspace_bytes, const attrici_variable_t* var, int T,
                           int C, const double* params, int which, int counterfactual, double* out,
                           int64_t ld_out, void* stream);

/* NumPy-legacy MT19937 doubles, one stream per cell (np.random.seed(s); np.random.rand(n)).
 *   out[n x ld] float64, cell-minor like every other series */
ATTRICI_API int attrici_mt19937_uniform(const uint32_t* seeds, int C, int n, double* out, int64_t ld, void* stream);

/* SSA smoothing of the GMT predictor: first elementary component of the singular spectrum analysis of x[n]
 * (window W), i.e. row 0 of SingularSpectrumAnalysis(W).transform(x) as used by calc_gmt_by_ssa
 * (attrici/preprocessing.py:13-40; attrici/vendored/singularspectrumanalysis.py:161-207, 51-76) - what
 * `attrici ssa` (attrici/ssa.py:12-60) computes before `detrend` reads it with --gmt-file.
 *   x, out: device float64 [n]; workspace: attrici_ssa_workspace_bytes(n, W) bytes, 256-byte aligned;
 *   the leading eigenvector comes from a power iteration that stops at |G v - lambda v| <= tol lambda or after
 *   max_iter steps; info (device int32 [2], nullable) = {iterations, converged}, eigenvalue (device float64 [1],
 *   nullable) = lambda.  window outside [2, n] -> ATTRICI_EINVAL (pyts raises ValueError, :253-262) */
ATTRICI_API int attrici_ssa_workspace_bytes(int n, int window, size_t* bytes);
ATTRICI_API int attrici_ssa_first_component(const double* x, int n, int window, double* out, void* workspace,
                                size_t workspace_bytes, int max_iter, double tol, int32_t* info, double* eigenvalue,
                                void* stream);

/* ---- rolling-window model (`--window-size`; attrici/util.py:107-172 collect_windows, attrici/estimation/
 * model_pymc5.py:366-547; attrici/detrend.py:646-649 modes XOR window_size) - Normal family ----
 * Every day of the year k = 1..366 has its own weights: mu(t) = a_k + b_k gmt(t), sigma(t) = exp(c_k), k = dayofyear(t),
 * each with a N(0, 1) prior, fitted on the window_size days around every day t of the fit series [i0, i1) whose day of
 * the year is k (series mirrored at both ends; day 366 = the windows around the day after each day 365).
 *   run_start / run_first_doy / run_len [n_runs] (device int32): the fit series as runs of consecutive days with
 *     consecutive days of the year (calendar years): first day relative to i0, its day of the year (1-based), length.
 *     The integers come from the host (numpy / pandas `time.dt.dayofyear`), never from floating point on the device.
 *   params [C x 3 x 366] out: weights_mu_longterm_intercept, weights_mu_longterm_trend, weights_sigma_longterm_intercept
 *     (the reference's trace entries for the Normal family); logp [C] out: log posterior at the MAP;
 *   status [C] out, n_iter [C] out (largest Newton iteration count of the cell's 366 problems), both nullable.
 * Missing days are skipped inside the windows (the reference drops them before it windows by index).
 * attrici_window_quantile_map: estimate_distribution (:435-447, :524-538) by day of the year + quantile map, rescale and
 *   replacement as attrici_quantile_map; dayofyear [T] device int16 for the WHOLE axis. */
ATTRICI_API int attrici_window_workspace_bytes(int C, size_t* bytes);
ATTRICI_API int attrici_window_fit(void* workspace, size_t workspace_bytes, const attrici_variable_t* var, const float* y_scaled,
                       int64_t ld, int T, int C, int i0, int i1, const double* gmt, const int32_t* run_start,
                       const int32_t* run_first_doy, const int32_t* run_len, int n_runs, int window_size, double* params,
                       double* logp, int32_t* status, int32_t* n_iter, void* stream);
ATTRICI_API int attrici_window_quantile_map(const attrici_variable_t* var, const float* y, const float* y_scaled, int64_t ld, int T, int C,
                                const double* gmt, const int16_t* dayofyear, const double* params, const double* scaling,
                                double* cfact, uint8_t* replaced, double* cfact_scaled, int64_t ld_out, void* stream);

/* ---- block bootstrap of the fitted trend (attrici/detrend.py:479-612; all five families) ----
 * The samples of a batch of cells are refitted as extra cells of a [T x R C] series (column r C + c) through
 * attrici_prep_scale_mask / attrici_fit_batch / attrici_eval_parameter; these four entries provide the rest.
 *
 * attrici_bootstrap_block_choices: np.random.seed(seeds[c]); choices[c][i] = np.random.randint(0, n_blocks),
 *   i < n_draws = n_samples * n_blocks (sample-major, year-minor: the order of detrend.py:541-546), after
 *   skip_words 32-bit generator outputs.  choices: device int16 [C][n_draws].
 * attrici_bootstrap_scores: what a source day contributes (detrend.py:534): scores[t][c] = F_ref,t(y_scaled), kept
 *   as ndtri(F) for the Normal family (build_basis must have run on this workspace for (T, C)).
 * attrici_bootstrap_resample: samples r0 .. r0+R-1 -> y_new[t][r C + c] = F_ref,t^-1(score of the source day),
 *   invalid values replaced by y_scaled[t][c] and counted in replaced_count[t][c] (+=) (detrend.py:541-557,
 *   get_random_block_indices :500-531).  block_start: device int32 [n_blocks + 1] first day of each calendar
 *   year block (+ T), block_of_day: device int16 [T].
 * attrici_bootstrap_stats: mean / std (ddof 0) / numpy.percentile(linear) at `levels` of the rescaled
 *   expectations[t][r C + c] (detrend.py:574-612); mean, std_dev: [T x C], quantiles: [n_levels x T x C]. */
ATTRICI_API int attrici_bootstrap_block_choices(const uint32_t* seeds, int C, int n_draws, int n_blocks, int skip_words,
                                    int16_t* choices, void* stream);
ATTRICI_API int attrici_bootstrap_scores(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                             const float* y_scaled, int64_t ld, int T, int C, const double* params, double* scores,
                             int64_t ld_s, void* stream);
ATTRICI_API int attrici_bootstrap_resample(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                               const float* y_scaled, int64_t ld, int T, int C, const double* params,
                               const double* scores, int64_t ld_s, const int16_t* choices, int n_samples, int n_blocks,
                               int r0, int R, const int32_t* block_start, const int16_t* block_of_day, float* y_new,
                               int64_t ld_new, int32_t* replaced_count, int64_t ld_rc, void* stream);
ATTRICI_API int attrici_bootstrap_stats(const attrici_variable_t* var, const double* expectations, int64_t ld_e, int T, int C,
                            int n_samples, const double* scaling, const double* levels, int n_levels, double* mean,
                            double* std_dev, double* quantiles, void* stream);

/* derive-huss: specific humidity [kg kg-1] from relative humidity [%], air pressure [Pa] and temperature [K],
 * elementwise over n float64 values (device pointers): calc_huss_weedon2010, attrici/commands/derive_huss.py:27-99 */
ATTRICI_API int attrici_derive_huss(const double* hurs, const double* ps, const double* tas, size_t n, double* huss,
                        void* stream);

/* whole path for one batch with HOST buffers (pinned recommended): H2D of y, scale, fit,
 * quantile map, D2H of cfact / params / logp / status, cells processed in slabs of
 * `slab_cells` with copies and kernels overlapped on internal streams.
 * replaces the cell loop attrici/detrend.py:760-788 around fit_and_detrend_cell (:258-416, I/O excluded).
 *   device_scratch: device buffer of attrici_detrend_host_scratch_bytes() bytes */
ATTRICI_API int attrici_detrend_host_scratch_bytes(int family, int T, int slab_cells, size_t* bytes);
ATTRICI_API int attrici_detrend_host(const attrici_variable_t* var, const attrici_fit_options_t* opts,
                         const float* y_host, int64_t ld_host, int T, int C, int i0, int i1,
                         const int32_t* days_host, const double* gmt_host, const uint32_t* seeds_host,
                         double* cfact_host, int64_t ld_out_host, uint8_t* replaced_host,
                         double* params_host, double* logp_host, double* scaling_host,
                         int32_t* status_host, int32_t* n_pass_host,
                         void* device_scratch, size_t device_scratch_bytes, int slab_cells);

/* The same with options for what is carried across PCIe - the pipeline is bound by the download of the results:
 *   cfact_dtype    ATTRICI_F64 (the reference's output type) or ATTRICI_F32 (the FP64 result rounded once to the dtype
 *                  of the input files, 6e-8 relative; cfact_host is then float*): halves the download;
 *   replaced_mode  ATTRICI_FLAGS_DENSE: the [T x C] uint8 plane; ATTRICI_FLAGS_SPARSE: the non-zero entries as
 *                  (day, cell, code) int32 triples appended (in no particular order) to replaced_sparse_host, which
 *                  must be PINNED host memory of sparse_capacity triples; *n_sparse_host receives their number
 *                  (ATTRICI_EWORKSPACE if it exceeds the capacity); ATTRICI_FLAGS_OFF: not returned;
 *   slab_major     1: cfact_host (and the dense replaced_host) hold the cell slabs one after the other, slab k =
 *                  [T x n_cells_k] contiguous at element offset T * cell0_k (attrici_host_slab_range), so that each slab
 *                  comes down as one linear copy instead of a T-row pitched one; ld_out_host is ignored;
 *   prep_flags     ATTRICI_PREP_* of attrici_prep_scale_mask_ex. */
#define ATTRICI_F64 0
#define ATTRICI_F32 1
#define ATTRICI_FLAGS_DENSE 0
#define ATTRICI_FLAGS_SPARSE 1
#define ATTRICI_FLAGS_OFF 2
typedef struct {
    int32_t prep_flags;
    int32_t cfact_dtype;
    int32_t replaced_mode;
    int32_t slab_major;
    int64_t sparse_capacity;
} attrici_host_options_t;
ATTRICI_API int attrici_default_host_options(attrici_host_options_t* options);
/* slab k of the pipeline for (C, slab_cells): *cell0, *n_cells (k outside the range: untouched); returns the number of slabs */
ATTRICI_API int attrici_host_slab_range(int C, int slab_cells, int k, int* cell0, int* n_cells);
ATTRICI_API int attrici_detrend_host_ex(const attrici_variable_t* var, const attrici_fit_options_t* opts,
                            const attrici_host_options_t* host_options,
                            const float* y_host, int64_t ld_host, int T, int C, int i0, int i1,
                            const int32_t* days_host, const double* gmt_host, const uint32_t* seeds_host,
                            void* cfact_host, int64_t ld_out_host, uint8_t* replaced_host,
                            int32_t* replaced_sparse_host, int64_t* n_sparse_host,
                            double* params_host, double* logp_host, double* scaling_host,
                            int32_t* status_host, int32_t* n_pass_host,
                            void* device_scratch, size_t device_scratch_bytes, int slab_cells);

/* quantile map with a float32 counterfactual (device buffers): Normal variables without a post rule on a gap-free
 * daily axis, `replaced` required; anything else returns ATTRICI_EUNSUPPORTED.  Synchronises `stream` once. */
ATTRICI_API int attrici_quantile_map_f32(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                             const float* y, const float* y_scaled, int64_t ld, int T, int C,
                             const double* params, const double* scaling, float* cfact, uint8_t* replaced,
                             int64_t ld_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ATTRICI_B200_H */
