// C ABI of libattrici_b200.so (declared in include/attrici_b200.h): argument checking, workspace
// carving, the batched fit loop and the host-buffer pipeline.  No torch types; process-wide state is limited to
// diagnostics and plumbing (launch counter, last-error text, fit profile, the read-back mailbox, environment
// switches read once); every call is asynchronous on the caller's stream except where stated.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <atomic>
#include <mutex>
#include <vector>
#include "common.cuh"

namespace attrici {

static thread_local char g_error[512] = "";

// diagnostics: kernels launched by this process, and the pass-kernel timing of the last profiled fit
std::atomic<long long> g_launches{0};
struct FitProfile {
    double pass_ms;        // summed device time of the moment passes
    long long launches;    // number of moment-pass launches
    double cell_days;      // sum over launches of (cells still iterating) x (days in the fit window)
};
static thread_local FitProfile g_profile = {0.0, 0, 0.0};

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

namespace {

bool valid_family(int f) { return f >= ATTRICI_NORMAL && f <= ATTRICI_BERNOULLI_GAMMA; }

int check_var(const attrici_variable_t* var) {
    if (!var) { set_error("variable descriptor is null"); return ATTRICI_EINVAL; }
    if (!valid_family(var->family)) { set_error("unknown family %d", var->family); return ATTRICI_EINVAL; }
    if (var->scaling < ATTRICI_SCALE_UNITY || var->scaling > ATTRICI_SCALE_PR) {
        set_error("unknown scaling %d", var->scaling);
        return ATTRICI_EINVAL;
    }
    if (var->post_rule < ATTRICI_POST_NONE || var->post_rule > ATTRICI_POST_OUTSIDE01_NAN) {
        set_error("unknown post rule %d", var->post_rule);
        return ATTRICI_EINVAL;
    }
    if (var->modes < 1 || var->modes > ATTRICI_MAX_MODES) {
        set_error("modes must be in 1..%d, got %d", ATTRICI_MAX_MODES, var->modes);
        return ATTRICI_EINVAL;
    }
    if (var->scaling == ATTRICI_SCALE_PR && var->family != ATTRICI_BERNOULLI_GAMMA) {
        set_error("the pr scaling belongs to the Bernoulli-Gamma family (variables.py:341-486)");
        return ATTRICI_EINVAL;
    }
    return 0;
}

int check_shape(int T, int C, int64_t ld) {
    if (T < 2 || C < 1) { set_error("need T >= 2 and C >= 1, got T=%d C=%d", T, C); return ATTRICI_EINVAL; }
    if (ld < C) { set_error("leading dimension %lld < C=%d", (long long)ld, C); return ATTRICI_EINVAL; }
    return 0;
}

int check_window(int T, int i0, int i1) {
    if (i0 < 0 || i1 > T || i1 <= i0) {
        set_error("fit window [%d, %d) outside [0, %d)", i0, i1, T);
        return ATTRICI_EINVAL;
    }
    return 0;
}

int bind_workspace(Workspace& ws, void* workspace, size_t workspace_bytes, int T, int C, int family) {
    if (!workspace) { set_error("workspace is null"); return ATTRICI_EINVAL; }
    if (((uintptr_t)workspace & 255) != 0) { set_error("workspace must be 256-byte aligned"); return ATTRICI_EINVAL; }
    size_t need = ws.layout(T, C, family, workspace);
    if (need > workspace_bytes) {
        set_error("workspace too small: %zu bytes given, %zu needed", workspace_bytes, need);
        return ATTRICI_EWORKSPACE;
    }
    return 0;
}

attrici_fit_options_t default_options() {
    attrici_fit_options_t o;
    o.max_pass = 60;
    o.use_fp64 = 0;
    o.tol_pred = 1e-11;
    o.lambda0 = 1e-3;
    o.zero_init = 0;
    o.profile = 0;
    return o;
}

StepParams step_params(const attrici_variable_t& var, const attrici_fit_options_t& o, int slot) {
    StepParams sp;
    sp.term = family_term(var.family, slot);
    sp.modes = var.modes;
    sp.slot = slot;
    sp.max_pass = o.max_pass;
    sp.tol_pred = o.tol_pred;
    sp.lambda0 = o.lambda0;
    sp.zero_init = o.zero_init;
    return sp;
}

// Small device -> host read-backs (numbers of running cells, the fold flags) go through a pinned, mapped host
// mailbox written by a one-warp kernel, not through cudaMemcpyAsync: a copy would queue on the D2H copy engine
// behind the bulk download of the previous slab of attrici_detrend_host (tens of ms) and stall the fit.
// One mailbox per process (not per thread: ModelCuda.fit_cached runs every fit with a timeout on a fresh thread, a
// per-thread mailbox would leak one pinned page per cell), used under a mutex from the launch to the read.
struct Mailbox { int* host = nullptr; int* dev = nullptr; bool tried = false; };
std::mutex g_mailbox_mutex;
Mailbox& mailbox() {
    static Mailbox m;
    if (!m.tried) {
        m.tried = true;
        void* h = nullptr;
        void* d = nullptr;
        if (cudaHostAlloc(&h, 256, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess &&
            cudaHostGetDevicePointer(&d, h, 0) == cudaSuccess) {
            m.host = (int*)h;
            m.dev = (int*)d;
        } else {
            if (h) cudaFreeHost(h);
            cudaGetLastError();
        }
    }
    return m;
}

__global__ void post_ints_kernel(const int* __restrict__ src, int n, volatile int* dst) {
    if ((int)threadIdx.x < n) dst[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}

// FP64 counterfactual of one slab -> float32 (families whose quantile-map kernels store FP64 only)
__global__ void narrow_f64_f32_kernel(const double* __restrict__ src, float* __restrict__ dst, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = (float)src[i];
}

// `replaced` plane of one slab -> list of (day, cell, code) of its non-zero entries, appended to a pinned, mapped host
// buffer through the device counter `count` (entries beyond `capacity` are counted, not stored).  The plane is
// non-zero on << 1 % of the days, so the list replaces a byte per cell-day of download by a few KB.
__global__ void replaced_sparse_kernel(const uint8_t* __restrict__ plane, int T, int nc, int cell0, int32_t* __restrict__ entries,
                                       unsigned long long capacity, unsigned long long* __restrict__ count) {
    const size_t n = (size_t)T * nc;
    const size_t n16 = n / 16;
    const uint4* p16 = reinterpret_cast<const uint4*>(plane);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n16; i += (size_t)gridDim.x * blockDim.x) {
        uint32_t w[4] = {0u, 0u, 0u, 0u};
        size_t e0 = i * 16;
        int nb = 16;
        if (i < n16) {
            const uint4 v = __ldg(p16 + i);
            w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
            if (!(w[0] | w[1] | w[2] | w[3])) continue;
        } else {
            nb = (int)(n - e0);   // tail of fewer than 16 bytes
        }
        for (int b = 0; b < nb; ++b) {
            const uint32_t code = (i < n16) ? ((w[b >> 2] >> (8 * (b & 3))) & 0xffu) : (uint32_t)plane[e0 + b];
            if (code) {
                const size_t e = e0 + b;
                const unsigned long long k = atomicAdd(count, 1ull);
                if (k < capacity) {
                    entries[3 * k + 0] = (int32_t)(e / nc);
                    entries[3 * k + 1] = cell0 + (int32_t)(e % nc);
                    entries[3 * k + 2] = (int32_t)code;
                }
            }
        }
    }
}
int read_ints(const int* dev, int* host, int n, cudaStream_t s) {
    std::lock_guard<std::mutex> lock(g_mailbox_mutex);
    Mailbox& m = mailbox();
    if (m.host && n <= 32) {
        post_ints_kernel<<<1, 32, 0, s>>>(dev, n, m.dev);
        ATTRICI_LAUNCH_CHECK();
        ATTRICI_CUDA_CHECK(cudaStreamSynchronize(s));
        for (int i = 0; i < n; ++i) host[i] = ((volatile int*)m.host)[i];
        return 0;
    }
    ATTRICI_CUDA_CHECK(cudaMemcpyAsync(host, dev, sizeof(int) * n, cudaMemcpyDeviceToHost, s));
    ATTRICI_CUDA_CHECK(cudaStreamSynchronize(s));
    return 0;
}

int read_counter(const int* dev, int* host, cudaStream_t s) { return read_ints(dev, host, 1, s); }

// ATTRICI_NO_FOLD=1 keeps every family on the day-by-day kernels (A/B measurements only)
bool fold_enabled() {
    static const bool on = [] { const char* e = getenv("ATTRICI_NO_FOLD"); return !(e && e[0] == '1'); }();
    return on;
}

// ATTRICI_NO_FUSED=1 keeps the folded Normal fit on the per-pass kernels of fold.cu / step.cu (A/B measurements only)
bool fused_enabled() {
    static const bool on = [] { const char* e = getenv("ATTRICI_NO_FUSED"); return !(e && e[0] == '1'); }();
    return on;
}

// does the phase-folded pass apply?  Normal term on a gap-free daily axis (flag left by attrici_build_basis)
// whose per-phase sums were gathered by attrici_prep_scale_mask on this workspace.  Synchronises `s`.
// how a term's likelihood passes run: FOLD_SUMS = Normal term from the per-phase sums of prep (fold.cu),
// FOLD_STREAM = the other terms streamed in phase order, FOLD_NONE = day-by-day kernels
enum { FOLD_NONE = 0, FOLD_SUMS = 1, FOLD_STREAM = 2 };
int fold_applies(Workspace& ws, int term, int i0, int i1, cudaStream_t s, int* fold, bool* phase_major = nullptr) {
    *fold = FOLD_NONE;
    if (phase_major) *phase_major = true;
    if (!fold_enabled()) return 0;
    if (term == T_NORMAL && !ws.ystat) return 0;
    // {days off the gap-free axis, fit window the phase rows / sums were gathered for, sums also in fold.cu's layout}
    int h[4] = {1, -1, -1, 1};
    ATTRICI_TRY(read_ints(ws.fold_bad, h, 4, s));
    if (phase_major) *phase_major = h[3] != 0;
    if (h[0] == 0 && h[1] == i0 && h[2] == i1) {
        // measured (scripts/bench_families.py, two-stage form of fold.cu): the phase-ordered stream wins for every other
        // term - hurs fit 29.6 -> 13.9 ms, pr 12.6 -> 5.0 ms, tasrange 11.0 -> 3.1 ms, sfcWind 3.4 -> 2.2 ms per 1 024 cells
        if (term == T_NORMAL) *fold = FOLD_SUMS;
        else *fold = FOLD_STREAM;
    }
    return 0;
}

// one likelihood pass + cross-chunk reduction at ws.th_pass, by whichever kernel applies
// mode: PASS_NLL / PASS_GRAD / PASS_FULL (the day-by-day kernels have no gradient-only form: PASS_GRAD = PASS_FULL)
int run_pass(Workspace& ws, int fold, const int* list, int n_active, int term, bool fp64, int mode,
             bool exact, const float* y_scaled, int64_t ld, int i0, int i1, const int* status, cudaStream_t s,
             cudaEvent_t e0 = nullptr, cudaEvent_t e1 = nullptr) {
    if (fold == FOLD_STREAM && !fp64 && mode != PASS_NLL) {
        mode = PASS_FULL;
        PassShape ps = plan_pass(ws, list, n_active, 0, fold_phases(ws), false);
        if (e0) ATTRICI_CUDA_CHECK(cudaEventRecord(e0, s));
        ATTRICI_TRY(launch_fold_pass_generic(ws, ps, term, false, y_scaled, ld, i0, i1, status, s));
        if (e1) ATTRICI_CUDA_CHECK(cudaEventRecord(e1, s));
        return launch_reduce(ws, ps, false, PASS_FULL, s, status);
    }
    if (fold == FOLD_STREAM && fp64 && mode == PASS_NLL) {   // the FP64 log posterior of the returned point
        PassShape ps = plan_pass(ws, list, n_active, 0, fold_phases(ws), true);
        ATTRICI_TRY(launch_fold_pass_generic(ws, ps, term, true, y_scaled, ld, i0, i1, status, s));
        return launch_reduce(ws, ps, true, PASS_NLL, s, status);
    }
    if (fold == FOLD_SUMS) {
        PassShape ps = plan_pass(ws, list, n_active, 0, fold_items(ws, mode, exact), true);
        if (e0) ATTRICI_CUDA_CHECK(cudaEventRecord(e0, s));
        ATTRICI_TRY(launch_fold_pass(ws, ps, mode, exact, status, s));
        if (e1) ATTRICI_CUDA_CHECK(cudaEventRecord(e1, s));
        return launch_reduce(ws, ps, true, mode, s, status);
    }
    if (mode == PASS_GRAD || mode == PASS_FLAT) mode = PASS_FULL;
    PassShape ps = plan_pass(ws, list, n_active, i0, i1, fp64);
    if (e0) ATTRICI_CUDA_CHECK(cudaEventRecord(e0, s));
    ATTRICI_TRY(launch_pass(ws, ps, term, fp64, mode != PASS_NLL, y_scaled, ld, i0, i1, status, s));
    if (e1) ATTRICI_CUDA_CHECK(cudaEventRecord(e1, s));
    return launch_reduce(ws, ps, fp64, mode, s, status);
}

// Passes FISHER_PASSES .. FISHER_PASSES + CHORD_PASSES - 1 of the folded iteration reuse the information matrix of
// the accepted point instead of refreshing the Fisher moments: the matrix depends on the coefficients only through
// sigma(d), which has settled by then for a typical cell, and those are its last passes.  Cells that are still
// iterating afterwards (strong seasonal cycle of sigma) go back to fresh moments.
constexpr int FISHER_PASSES = 3;
constexpr int CHORD_PASSES = 2;

// Fisher-scoring iteration of one likelihood term over the whole batch.  Cells still iterating are kept in
// a compact list (ping-pong); once fewer than half of the batch is left the likelihood pass runs over the
// list (gathered columns, more time chunks per cell) instead of over all cells with a status mask.
int fit_term(Workspace& ws, const attrici_variable_t& var, const attrici_fit_options_t& o, int slot,
             const float* y_scaled, int64_t ld, int i0, int i1, cudaStream_t s) {
    StepParams sp = step_params(var, o, slot);
    const bool fp64 = o.use_fp64 != 0;
    int cur = 0, running = 0;
    int fold = FOLD_NONE;
    ATTRICI_TRY(launch_fit_init(ws, sp, ws.list[cur], ws.counters + cur, s));
    bool phase_major = true;
    ATTRICI_TRY(fold_applies(ws, sp.term, i0, i1, s, &fold, &phase_major));
    if (fold != FOLD_SUMS && !y_scaled) { set_error("y_scaled is needed: the phase-folded pass does not apply"); return ATTRICI_EINVAL; }
    if (fold == FOLD_SUMS && ws.T >= NPHASE && ((fused_enabled() && !o.profile) || !phase_major)) {
        // the whole iteration of the term in one persistent kernel (fit_fused.cu): no host round trip
        return launch_fit_fused(ws, sp, FISHER_PASSES, CHORD_PASSES, s);
    }
    ATTRICI_TRY(read_counter(ws.counters + cur, &running, s));
    const bool profile = o.profile != 0;
    std::vector<cudaEvent_t> ev;
    // a cell may spend passes on rejected trial points: the cap counts passes, not accepted steps.
    // The folded passes are cheap enough that reading the number of running cells back after every pass (a host
    // round trip with the GPU idle) costs more than a pass over cells that are already done: they are issued in
    // blocks (6, then 2 at a time) over all cells with the status mask, and the host synchronises once per block.
    for (int it = 0; running > 0 && it < o.max_pass;) {
        int block = fold == FOLD_SUMS ? (it == 0 ? 6 : 2) : 1;
        if (block > o.max_pass - it) block = o.max_pass - it;
        for (int b = 0; b < block; ++b) {
            const bool use_list = (block == 1) && (2 * running < ws.C);
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (profile) {
                ATTRICI_CUDA_CHECK(cudaEventCreate(&e0));
                ATTRICI_CUDA_CHECK(cudaEventCreate(&e1));
                ev.push_back(e0);
                ev.push_back(e1);
                g_profile.cell_days += (double)running * (double)(i1 - i0);
            }
            const bool fresh = fold != FOLD_SUMS || (it + b < FISHER_PASSES) || (it + b >= FISHER_PASSES + CHORD_PASSES);
            // pass 0 evaluates the point fit_init leaves: only the two intercepts are non-zero
            const int mode = (it + b == 0) ? PASS_FLAT : (fresh ? PASS_FULL : PASS_GRAD);
            ATTRICI_TRY(run_pass(ws, fold, use_list ? ws.list[cur] : nullptr, use_list ? running : ws.C, sp.term, fp64,
                                 mode, false, y_scaled, ld, i0, i1,
                                 use_list ? nullptr : ws.status, s, e0, e1));
            if (block == 1)
                ATTRICI_TRY(launch_fit_step(ws, sp, ws.list[cur], running, ws.list[cur ^ 1], ws.counters + (cur ^ 1), s, fresh));
            else   // every cell's warp looks at its own status
                ATTRICI_TRY(launch_fit_step(ws, sp, nullptr, ws.C, ws.list[cur ^ 1], ws.counters + (cur ^ 1), s, fresh));
            cur ^= 1;
        }
        it += block;
        ATTRICI_TRY(read_counter(ws.counters + cur, &running, s));
    }
    for (size_t k = 0; k + 1 < ev.size(); k += 2) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, ev[k], ev[k + 1]) == cudaSuccess) {
            g_profile.pass_ms += ms;
            g_profile.launches += 1;
        }
        cudaEventDestroy(ev[k]);
        cudaEventDestroy(ev[k + 1]);
    }
    // log posterior of the accepted point in FP64 (model_scipy.py:525: trace["logp"] = -res.fun)
    ATTRICI_TRY(launch_fit_final_point(ws, sp, s));
    ATTRICI_TRY(run_pass(ws, fold, nullptr, ws.C, sp.term, true, PASS_NLL, false, y_scaled, ld, i0, i1, nullptr, s));
    ATTRICI_TRY(launch_fit_store(ws, sp, s));
    return 0;
}

}  // namespace
}  // namespace attrici

using namespace attrici;

extern "C" {

const char* attrici_last_error(void) { return g_error; }

int attrici_abi_version(void) { return ATTRICI_ABI_VERSION; }

int attrici_num_params(int family, int modes) {
    if (!valid_family(family) || modes < 1 || modes > ATTRICI_MAX_MODES) {
        set_error("bad family %d / modes %d", family, modes);
        return ATTRICI_EINVAL;
    }
    return family_num_params(family, modes);
}

long long attrici_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int attrici_fit_profile(int reset, double* pass_ms, long long* pass_launches, double* cell_days) {
    if (pass_ms) *pass_ms = g_profile.pass_ms;
    if (pass_launches) *pass_launches = g_profile.launches;
    if (cell_days) *cell_days = g_profile.cell_days;
    if (reset) g_profile = FitProfile{0.0, 0, 0.0};
    return 0;
}

int attrici_default_fit_options(attrici_fit_options_t* opts) {
    if (!opts) { set_error("opts is null"); return ATTRICI_EINVAL; }
    *opts = default_options();
    return 0;
}

int attrici_workspace_bytes(int family, int T, int C, size_t* bytes) {
    if (!bytes) { set_error("bytes is null"); return ATTRICI_EINVAL; }
    if (!valid_family(family)) { set_error("unknown family %d", family); return ATTRICI_EINVAL; }
    ATTRICI_TRY(check_shape(T, C, C));
    Workspace ws;
    *bytes = ws.layout(T, C, family, nullptr);
    return 0;
}

int attrici_build_basis(void* workspace, size_t workspace_bytes, int family, const int32_t* days,
                        const double* gmt, int T, int C, void* stream) {
    if (!days || !gmt) { set_error("days / gmt is null"); return ATTRICI_EINVAL; }
    ATTRICI_TRY(check_shape(T, C, C));
    Workspace ws;
    ATTRICI_TRY(bind_workspace(ws, workspace, workspace_bytes, T, C, family));
    ATTRICI_TRY(launch_build_basis(ws, days, gmt, (cudaStream_t)stream));
    return launch_build_zmat(ws, (cudaStream_t)stream);
}

int attrici_prep_scale_mask(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                            const float* y, float* y_scaled, int64_t ld, int T, int C, int i0, int i1,
                            double* scaling, void* stream) {
    return attrici_prep_scale_mask_ex(workspace, workspace_bytes, var, y, y_scaled, ld, T, C, i0, i1, scaling, 0, stream);
}

int attrici_prep_scale_mask_ex(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                               const float* y, float* y_scaled, int64_t ld, int T, int C, int i0, int i1,
                               double* scaling, int flags, void* stream) {
    ATTRICI_TRY(check_var(var));
    ATTRICI_TRY(check_shape(T, C, ld));
    ATTRICI_TRY(check_window(T, i0, i1));
    if (!y || !scaling) { set_error("y / scaling is null"); return ATTRICI_EINVAL; }
    if (!y_scaled && var->family != ATTRICI_NORMAL) {
        set_error("y_scaled may only be omitted for the Normal family");
        return ATTRICI_EINVAL;
    }
    Workspace ws;
    ATTRICI_TRY(bind_workspace(ws, workspace, workspace_bytes, T, C, var->family));
    if (flags & ~ATTRICI_PREP_EXACT_SUMS) { set_error("unknown prep flags 0x%x", flags); return ATTRICI_EINVAL; }
    return launch_prep(ws, *var, y, y_scaled, ld, i0, i1, scaling, flags, (cudaStream_t)stream);
}

int attrici_nll_grad(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                     const float* y_scaled, int64_t ld, int T, int C, int i0, int i1, int use_fp64,
                     const double* theta, double* nll, double* grad, double* fisher, void* stream) {
    ATTRICI_TRY(check_var(var));
    ATTRICI_TRY(check_shape(T, C, ld));
    ATTRICI_TRY(check_window(T, i0, i1));
    if (!y_scaled || !theta || !nll) { set_error("y_scaled / theta / nll is null"); return ATTRICI_EINVAL; }
    Workspace ws;
    ATTRICI_TRY(bind_workspace(ws, workspace, workspace_bytes, T, C, var->family));
    cudaStream_t s = (cudaStream_t)stream;
    attrici_fit_options_t o = default_options();
    o.use_fp64 = use_fp64;
    const int nt = family_terms_count(var->family);
    for (int slot = 0; slot < nt; ++slot) {
        StepParams sp = step_params(*var, o, slot);
        ATTRICI_TRY(launch_phase_only(ws, sp, s));
        ATTRICI_TRY(launch_load_theta(ws, var->family, var->modes, slot, theta, s));
        ATTRICI_TRY(launch_rotate(ws, sp, s));
        int fold = FOLD_NONE;
        ATTRICI_TRY(fold_applies(ws, sp.term, i0, i1, s, &fold));
        ATTRICI_TRY(run_pass(ws, fold, nullptr, ws.C, sp.term, use_fp64 != 0, PASS_FULL, true, y_scaled, ld, i0, i1, nullptr, s));
        // canonical results go through idle fit-state buffers: f_acc, g_acc [Cp][NP], h_acc [Cp][NTRI]
        ATTRICI_TRY(launch_assemble(ws, sp, ws.f_acc, ws.g_acc, fisher ? ws.h_acc : nullptr, s));
        ATTRICI_TRY(launch_scatter_kat(ws, var->family, var->modes, slot, ws.f_acc, ws.g_acc,
                                       fisher ? ws.h_acc : nullptr, nll, grad, fisher, slot > 0, s));
    }
    return 0;
}

int attrici_fit_batch(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                      const float* y_scaled, int64_t ld, int T, int C, int i0, int i1,
                      const attrici_fit_options_t* opts, double* params, double* logp, int32_t* status,
                      int32_t* n_pass, void* stream) {
    ATTRICI_TRY(check_var(var));
    ATTRICI_TRY(check_shape(T, C, ld));
    ATTRICI_TRY(check_window(T, i0, i1));
    if (!params) { set_error("params is null"); return ATTRICI_EINVAL; }
    if (!y_scaled && var->family != ATTRICI_NORMAL) {
        set_error("y_scaled may only be omitted for the Normal family");
        return ATTRICI_EINVAL;
    }
    attrici_fit_options_t o = opts ? *opts : default_options();
    if (o.max_pass < 1 || !(o.tol_pred >= 0.0) || !(o.lambda0 >= 0.0)) {
        set_error("bad fit options (max_pass=%d tol_pred=%g lambda0=%g)", o.max_pass, o.tol_pred, o.lambda0);
        return ATTRICI_EINVAL;
    }
    Workspace ws;
    ATTRICI_TRY(bind_workspace(ws, workspace, workspace_bytes, T, C, var->family));
    cudaStream_t s = (cudaStream_t)stream;
    const int nt = family_terms_count(var->family);
    for (int slot = 0; slot < nt; ++slot) ATTRICI_TRY(fit_term(ws, *var, o, slot, y_scaled, ld, i0, i1, s));
    return launch_store_params(ws, var->family, var->modes, params, logp, status, n_pass, s);
}

int attrici_quantile_map(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                         const float* y, const float* y_scaled, int64_t ld, int T, int C,
                         const double* params, const double* scaling, const uint32_t* seeds,
                         double* cfact, uint8_t* replaced, double* cfact_scaled, int64_t ld_out,
                         void* stream) {
    ATTRICI_TRY(check_var(var));
    ATTRICI_TRY(check_shape(T, C, ld));
    if (ld_out < C) { set_error("ld_out %lld < C=%d", (long long)ld_out, C); return ATTRICI_EINVAL; }
    if (!y || !params || !scaling || !cfact) {
        set_error("y / params / scaling / cfact is null");
        return ATTRICI_EINVAL;
    }
    if (!y_scaled && var->family != ATTRICI_NORMAL) {
        set_error("y_scaled may only be omitted for the Normal family");
        return ATTRICI_EINVAL;
    }
    Workspace ws;
    ATTRICI_TRY(bind_workspace(ws, workspace, workspace_bytes, T, C, var->family));
    return launch_qmap(ws, *var, y, y_scaled, ld, params, scaling, seeds, cfact, replaced, cfact_scaled, ld_out,
                       (cudaStream_t)stream);
}

int attrici_eval_parameter(void* workspace, size_t workspace_bytes, const attrici_variable_t* var, int T,
                           int C, const double* params, int which, int counterfactual, double* out,
                           int64_t ld_out, void* stream) {
    ATTRICI_TRY(check_var(var));
    ATTRICI_TRY(check_shape(T, C, ld_out));
    const int n_par = var->family == ATTRICI_BERNOULLI_GAMMA ? 3 : 2;
    if (which < ATTRICI_WHICH_EXPECTATION || which >= n_par) {
        set_error("parameter index %d outside -1..%d", which, n_par - 1);
        return ATTRICI_EINVAL;
    }
    if (!params || !out) { set_error("params / out is null"); return ATTRICI_EINVAL; }
    Workspace ws;
    ATTRICI_TRY(bind_workspace(ws, workspace, workspace_bytes, T, C, var->family));
    return launch_eval_parameter(ws, *var, params, which, counterfactual, out, ld_out, (cudaStream_t)stream);
}

int attrici_window_workspace_bytes(int C, size_t* bytes) {
    if (!bytes || C < 1) { set_error("bytes is null or C < 1"); return ATTRICI_EINVAL; }
    *bytes = window_workspace_bytes(C);
    return 0;
}

int attrici_window_fit(void* workspace, size_t workspace_bytes, const attrici_variable_t* var, const float* y_scaled,
                       int64_t ld, int T, int C, int i0, int i1, const double* gmt, const int32_t* run_start,
                       const int32_t* run_first_doy, const int32_t* run_len, int n_runs, int window_size, double* params,
                       double* logp, int32_t* status, int32_t* n_iter, void* stream) {
    ATTRICI_TRY(check_var(var));
    ATTRICI_TRY(check_shape(T, C, ld));
    ATTRICI_TRY(check_window(T, i0, i1));
    if (var->family != ATTRICI_NORMAL) {
        set_error("the rolling-window model is implemented for the Normal family (tas, ps, rlds, rsds, tasskew)");
        return ATTRICI_EUNSUPPORTED;
    }
    if (window_size < 1 || window_size % 2 == 0 || window_size / 2 + 2 > i1 - i0) {
        set_error("window_size must be odd and shorter than the fit series, got %d", window_size);
        return ATTRICI_EINVAL;
    }
    if (!y_scaled || !gmt || !run_start || !run_first_doy || !run_len || n_runs < 1 || !params || !logp) {
        set_error("y_scaled / gmt / runs / params / logp is null");
        return ATTRICI_EINVAL;
    }
    if (!workspace || ((uintptr_t)workspace & 255) != 0 || workspace_bytes < window_workspace_bytes(C)) {
        set_error("window workspace: %zu bytes given, %zu needed (256-byte aligned)", workspace_bytes, window_workspace_bytes(C));
        return ATTRICI_EWORKSPACE;
    }
    return launch_window_fit(workspace, y_scaled, ld, T, C, i0, i1, gmt, run_start, run_first_doy, run_len, n_runs, window_size,
                             params, logp, status, n_iter, (cudaStream_t)stream);
}

int attrici_window_quantile_map(const attrici_variable_t* var, const float* y, const float* y_scaled, int64_t ld, int T, int C,
                                const double* gmt, const int16_t* dayofyear, const double* params, const double* scaling,
                                double* cfact, uint8_t* replaced, double* cfact_scaled, int64_t ld_out, void* stream) {
    ATTRICI_TRY(check_var(var));
    ATTRICI_TRY(check_shape(T, C, ld));
    if (var->family != ATTRICI_NORMAL) {
        set_error("the rolling-window model is implemented for the Normal family (tas, ps, rlds, rsds, tasskew)");
        return ATTRICI_EUNSUPPORTED;
    }
    if (ld_out < C) { set_error("ld_out %lld < C=%d", (long long)ld_out, C); return ATTRICI_EINVAL; }
    if (!y || !y_scaled || !gmt || !dayofyear || !params || !scaling || !cfact) {
        set_error("y / y_scaled / gmt / dayofyear / params / scaling / cfact is null");
        return ATTRICI_EINVAL;
    }
    return launch_window_qmap(*var, y, y_scaled, ld, T, C, gmt, dayofyear, params, scaling, cfact, replaced, cfact_scaled, ld_out,
                              (cudaStream_t)stream);
}

int attrici_mt19937_uniform(const uint32_t* seeds, int C, int n, double* out, int64_t ld, void* stream) {
    if (!seeds || !out) { set_error("seeds / out is null"); return ATTRICI_EINVAL; }
    if (C < 1 || n < 1 || ld < C) { set_error("bad shape C=%d n=%d ld=%lld", C, n, (long long)ld); return ATTRICI_EINVAL; }
    return launch_mt19937(seeds, C, n, out, ld, (cudaStream_t)stream);
}

// ---------------------------------------------------------------------------------------------
// whole path with HOST buffers: slabs of cells, copies and kernels overlapped on two streams
// ---------------------------------------------------------------------------------------------
namespace {
bool contiguous_host(const int32_t* days, int T) {
    for (int t = 0; t < T; ++t)
        if (days[t] - days[0] != t) return false;
    return true;
}

struct SlabBuffers {
    float* y; float* ys; double* cfact; float* cfact32; uint8_t* replaced; double* params; double* logp; double* scaling;
    int32_t* status; int32_t* npass; uint32_t* seeds; void* ws; size_t ws_bytes;
};

size_t slab_layout(int family, int T, int slab, int P, char* base, SlabBuffers* b) {
    size_t off = 0;
    auto take = [&](size_t n) {
        size_t o = off;
        off = Workspace::align(off + n);
        return base ? (void*)(base + o) : (void*)nullptr;
    };
    size_t ts = (size_t)T * slab;
    SlabBuffers tmp;
    tmp.y = (float*)take(4 * ts);
    tmp.ys = (float*)take(4 * ts);     // y_scaled, or the float32 counterfactual of families that do not keep it
    tmp.cfact32 = tmp.ys;
    tmp.cfact = (double*)take(8 * ts);
    tmp.replaced = (uint8_t*)take(ts);
    tmp.params = (double*)take(8 * (size_t)slab * P);
    tmp.logp = (double*)take(8 * (size_t)slab);
    tmp.scaling = (double*)take(16 * (size_t)slab);
    tmp.status = (int32_t*)take(4 * (size_t)slab);
    tmp.npass = (int32_t*)take(4 * (size_t)slab);
    tmp.seeds = (uint32_t*)take(4 * (size_t)slab);
    Workspace w;
    tmp.ws_bytes = w.layout(T, slab, family, nullptr);
    tmp.ws = take(tmp.ws_bytes);
    if (b) *b = tmp;
    return off;
}

// cell ranges of the slabs.  The pipeline is bound by the D2H copies, which can only start once the first slab has
// been uploaded and detrended: a short first slab shortens that fill time.
std::vector<int> slab_starts(int C, int slab_cells) {
    std::vector<int> start;
    if (slab_cells > C) slab_cells = C;
    int first = slab_cells / 4;
    if (first < 256) first = slab_cells;
    int c = 0;
    if (first < slab_cells && C > slab_cells) { start.push_back(0); c = first; }
    for (; c < C; c += slab_cells) start.push_back(c);
    start.push_back(C);
    return start;
}

}  // namespace

int attrici_detrend_host_scratch_bytes(int family, int T, int slab_cells, size_t* bytes) {
    if (!bytes) { set_error("bytes is null"); return ATTRICI_EINVAL; }
    if (!valid_family(family)) { set_error("unknown family %d", family); return ATTRICI_EINVAL; }
    ATTRICI_TRY(check_shape(T, slab_cells, slab_cells));
    const int P = family_num_params(family, ATTRICI_MAX_MODES);
    // two slab buffer sets (double buffering) + the day / gmt vectors + the counter of the sparse `replaced` list
    *bytes = 2 * slab_layout(family, T, slab_cells, P, nullptr, nullptr) + Workspace::align(4 * (size_t)T) +
             Workspace::align(8 * (size_t)T) + 256;
    return 0;
}

int attrici_default_host_options(attrici_host_options_t* o) {
    if (!o) { set_error("options is null"); return ATTRICI_EINVAL; }
    o->prep_flags = 0;
    o->cfact_dtype = ATTRICI_F64;
    o->replaced_mode = ATTRICI_FLAGS_DENSE;
    o->slab_major = 0;
    o->sparse_capacity = 0;
    return 0;
}

int attrici_host_slab_range(int C, int slab_cells, int k, int* cell0, int* n_cells) {
    if (C < 1 || slab_cells < 1) { set_error("need C >= 1 and slab_cells >= 1"); return ATTRICI_EINVAL; }
    const std::vector<int> start = slab_starts(C, slab_cells);
    const int n = (int)start.size() - 1;
    if (k >= 0 && k < n) {
        if (cell0) *cell0 = start[k];
        if (n_cells) *n_cells = start[k + 1] - start[k];
    }
    return n;
}

int attrici_detrend_host(const attrici_variable_t* var, const attrici_fit_options_t* opts,
                         const float* y_host, int64_t ld_host, int T, int C, int i0, int i1,
                         const int32_t* days_host, const double* gmt_host, const uint32_t* seeds_host,
                         double* cfact_host, int64_t ld_out_host, uint8_t* replaced_host,
                         double* params_host, double* logp_host, double* scaling_host,
                         int32_t* status_host, int32_t* n_pass_host,
                         void* device_scratch, size_t device_scratch_bytes, int slab_cells) {
    attrici_host_options_t ho;
    attrici_default_host_options(&ho);
    if (!replaced_host) ho.replaced_mode = ATTRICI_FLAGS_OFF;
    return attrici_detrend_host_ex(var, opts, &ho, y_host, ld_host, T, C, i0, i1, days_host, gmt_host, seeds_host, cfact_host,
                                   ld_out_host, replaced_host, nullptr, nullptr, params_host, logp_host, scaling_host,
                                   status_host, n_pass_host, device_scratch, device_scratch_bytes, slab_cells);
}

int attrici_detrend_host_ex(const attrici_variable_t* var, const attrici_fit_options_t* opts,
                            const attrici_host_options_t* hopts,
                            const float* y_host, int64_t ld_host, int T, int C, int i0, int i1,
                            const int32_t* days_host, const double* gmt_host, const uint32_t* seeds_host,
                            void* cfact_host, int64_t ld_out_host, uint8_t* replaced_host,
                            int32_t* replaced_sparse_host, int64_t* n_sparse_host,
                            double* params_host, double* logp_host, double* scaling_host,
                            int32_t* status_host, int32_t* n_pass_host,
                            void* device_scratch, size_t device_scratch_bytes, int slab_cells) {
    ATTRICI_TRY(check_var(var));
    ATTRICI_TRY(check_shape(T, C, ld_host));
    ATTRICI_TRY(check_window(T, i0, i1));
    if (slab_cells < 1) { set_error("slab_cells must be >= 1"); return ATTRICI_EINVAL; }
    if (!y_host || !days_host || !gmt_host || !cfact_host || !params_host) {
        set_error("y / days / gmt / cfact / params host pointer is null");
        return ATTRICI_EINVAL;
    }
    attrici_host_options_t ho;
    attrici_default_host_options(&ho);
    if (hopts) ho = *hopts;
    if (ho.cfact_dtype != ATTRICI_F64 && ho.cfact_dtype != ATTRICI_F32) { set_error("unknown cfact_dtype %d", ho.cfact_dtype); return ATTRICI_EINVAL; }
    if (ho.replaced_mode < ATTRICI_FLAGS_DENSE || ho.replaced_mode > ATTRICI_FLAGS_OFF) {
        set_error("unknown replaced_mode %d", ho.replaced_mode);
        return ATTRICI_EINVAL;
    }
    if (ho.prep_flags & ~ATTRICI_PREP_EXACT_SUMS) { set_error("unknown prep flags 0x%x", ho.prep_flags); return ATTRICI_EINVAL; }
    if (!ho.slab_major && ld_out_host < C) { set_error("ld_out_host %lld < C=%d", (long long)ld_out_host, C); return ATTRICI_EINVAL; }
    if (ho.replaced_mode == ATTRICI_FLAGS_DENSE && !replaced_host) { set_error("replaced_host is null"); return ATTRICI_EINVAL; }
    int32_t* sparse_dev = nullptr;
    if (ho.replaced_mode == ATTRICI_FLAGS_SPARSE) {
        if (!replaced_sparse_host || !n_sparse_host || ho.sparse_capacity < 1) {
            set_error("the sparse `replaced` list needs replaced_sparse_host, n_sparse_host and sparse_capacity >= 1");
            return ATTRICI_EINVAL;
        }
        void* d = nullptr;
        if (cudaHostGetDevicePointer(&d, replaced_sparse_host, 0) != cudaSuccess || !d) {
            cudaGetLastError();
            set_error("replaced_sparse_host must be pinned (page-locked) host memory: the kernel appends to it directly");
            return ATTRICI_EINVAL;
        }
        sparse_dev = (int32_t*)d;
    }
    if (var->family == ATTRICI_BERNOULLI_GAMMA && !seeds_host) {
        set_error("pr needs per-cell seeds (detrend.py:285)");
        return ATTRICI_EINVAL;
    }
    if (slab_cells > C) slab_cells = C;
    size_t need = 0;
    ATTRICI_TRY(attrici_detrend_host_scratch_bytes(var->family, T, slab_cells, &need));
    if (!device_scratch || device_scratch_bytes < need) {
        set_error("device scratch too small: %zu given, %zu needed", device_scratch_bytes, need);
        return ATTRICI_EWORKSPACE;
    }
    if (((uintptr_t)device_scratch & 255) != 0) { set_error("device scratch must be 256-byte aligned"); return ATTRICI_EINVAL; }
    const int P = family_num_params(var->family, var->modes);
    const int Pmax = family_num_params(var->family, ATTRICI_MAX_MODES);
    char* base = (char*)device_scratch;
    SlabBuffers buf[2];
    size_t one = slab_layout(var->family, T, slab_cells, Pmax, base, &buf[0]);
    slab_layout(var->family, T, slab_cells, Pmax, base + one, &buf[1]);
    int32_t* d_days = (int32_t*)(base + 2 * one);
    double* d_gmt = (double*)(base + 2 * one + Workspace::align(4 * (size_t)T));
    unsigned long long* d_count = (unsigned long long*)(base + 2 * one + Workspace::align(4 * (size_t)T) + Workspace::align(8 * (size_t)T));
    const bool f32 = ho.cfact_dtype == ATTRICI_F32;
    const size_t csize = f32 ? 4 : 8;

    cudaStream_t st[2];
    ATTRICI_CUDA_CHECK(cudaStreamCreateWithFlags(&st[0], cudaStreamNonBlocking));
    ATTRICI_CUDA_CHECK(cudaStreamCreateWithFlags(&st[1], cudaStreamNonBlocking));
    int rc = 0;
    auto run = [&]() -> int {
        ATTRICI_CUDA_CHECK(cudaMemcpyAsync(d_days, days_host, 4 * (size_t)T, cudaMemcpyHostToDevice, st[0]));
        ATTRICI_CUDA_CHECK(cudaMemcpyAsync(d_gmt, gmt_host, 8 * (size_t)T, cudaMemcpyHostToDevice, st[0]));
        ATTRICI_CUDA_CHECK(cudaMemsetAsync(d_count, 0, 8, st[0]));
        ATTRICI_CUDA_CHECK(cudaStreamSynchronize(st[0]));
        const std::vector<int> start = slab_starts(C, slab_cells);
        const int n_slabs = (int)start.size() - 1;
        // upload of slab k is issued before the (host-synchronising) fit of slab k-1 starts, on the
        // stream that owns its buffer set: stream order serialises the reuse of that set, and the copy
        // overlaps the other stream's kernels and downloads
        auto upload = [&](int k) -> int {
            const int bb = k & 1;
            const int cc0 = start[k];
            const int ncc = start[k + 1] - cc0;
            ATTRICI_CUDA_CHECK(cudaMemcpy2DAsync(buf[bb].y, 4 * (size_t)ncc, y_host + cc0, 4 * (size_t)ld_host,
                                                 4 * (size_t)ncc, T, cudaMemcpyHostToDevice, st[bb]));
            if (seeds_host)
                ATTRICI_CUDA_CHECK(cudaMemcpyAsync(buf[bb].seeds, seeds_host + cc0, 4 * (size_t)ncc,
                                                   cudaMemcpyHostToDevice, st[bb]));
            return 0;
        };
        ATTRICI_TRY(upload(0));
        const bool gap_free = contiguous_host(days_host, T);
        const bool skip_scaled = var->family == ATTRICI_NORMAL && fold_enabled() && gap_free;
        for (int is = 0; is < n_slabs; ++is) {
            const int b = is & 1;
            cudaStream_t s = st[b];
            SlabBuffers& B = buf[b];
            const int c0 = start[is];
            const int nc = start[is + 1] - c0;
            if (is + 1 < n_slabs) ATTRICI_TRY(upload(is + 1));
            ATTRICI_TRY(attrici_build_basis(B.ws, B.ws_bytes, var->family, d_days, d_gmt, T, nc, s));
            // Normal family on a gap-free daily axis: the fit runs on the per-phase sums and the quantile map
            // recomputes y_scaled from y, so the scaled series is never materialised
            float* ys = skip_scaled ? nullptr : B.ys;
            ATTRICI_TRY(attrici_prep_scale_mask_ex(B.ws, B.ws_bytes, var, B.y, ys, nc, T, nc, i0, i1, B.scaling, ho.prep_flags, s));
            ATTRICI_TRY(attrici_fit_batch(B.ws, B.ws_bytes, var, ys, nc, T, nc, i0, i1, opts, B.params, B.logp,
                                          B.status, B.npass, s));
            const bool want_plane = ho.replaced_mode != ATTRICI_FLAGS_OFF;
            // float32 counterfactual: stored directly by the lean Normal kernel (which needs the plane), narrowed from
            // the FP64 result for every other case
            const bool direct32 = f32 && skip_scaled && qmap_fold_lean_applies(*var, nullptr, B.replaced);
            const void* cfact_dev = B.cfact;
            if (direct32) {
                Workspace w;
                ATTRICI_TRY(bind_workspace(w, B.ws, B.ws_bytes, T, nc, var->family));
                ATTRICI_TRY(launch_qmap(w, *var, B.y, nullptr, nc, B.params, B.scaling, nullptr, nullptr, B.replaced, nullptr, nc, s,
                                        B.cfact32));
                cfact_dev = B.cfact32;
            } else {
                ATTRICI_TRY(attrici_quantile_map(B.ws, B.ws_bytes, var, B.y, ys, nc, T, nc, B.params, B.scaling,
                                                 seeds_host ? B.seeds : nullptr, B.cfact, want_plane ? B.replaced : nullptr,
                                                 nullptr, nc, s));
                if (f32) {
                    // y_scaled is dead after the quantile map: its buffer takes the narrowed counterfactual
                    narrow_f64_f32_kernel<<<148 * 8, 256, 0, s>>>(B.cfact, B.cfact32, (size_t)T * nc);
                    ATTRICI_LAUNCH_CHECK();
                    cfact_dev = B.cfact32;
                }
            }
            if (ho.slab_major) {
                // the host array holds the slabs one after the other, each [T x nc] contiguous: one linear copy
                ATTRICI_CUDA_CHECK(cudaMemcpyAsync((char*)cfact_host + csize * (size_t)T * c0, cfact_dev, csize * (size_t)T * nc,
                                                   cudaMemcpyDeviceToHost, s));
            } else {
                ATTRICI_CUDA_CHECK(cudaMemcpy2DAsync((char*)cfact_host + csize * c0, csize * (size_t)ld_out_host, cfact_dev,
                                                     csize * (size_t)nc, csize * (size_t)nc, T, cudaMemcpyDeviceToHost, s));
            }
            if (ho.replaced_mode == ATTRICI_FLAGS_DENSE) {
                if (ho.slab_major)
                    ATTRICI_CUDA_CHECK(cudaMemcpyAsync(replaced_host + (size_t)T * c0, B.replaced, (size_t)T * nc,
                                                       cudaMemcpyDeviceToHost, s));
                else
                    ATTRICI_CUDA_CHECK(cudaMemcpy2DAsync(replaced_host + c0, (size_t)ld_out_host, B.replaced, (size_t)nc,
                                                         (size_t)nc, T, cudaMemcpyDeviceToHost, s));
            } else if (ho.replaced_mode == ATTRICI_FLAGS_SPARSE) {
                replaced_sparse_kernel<<<148 * 8, 256, 0, s>>>(B.replaced, T, nc, c0, sparse_dev,
                                                              (unsigned long long)ho.sparse_capacity, d_count);
                ATTRICI_LAUNCH_CHECK();
            }
            ATTRICI_CUDA_CHECK(cudaMemcpyAsync(params_host + (size_t)c0 * P, B.params, 8 * (size_t)nc * P,
                                               cudaMemcpyDeviceToHost, s));
            if (logp_host)
                ATTRICI_CUDA_CHECK(cudaMemcpyAsync(logp_host + c0, B.logp, 8 * (size_t)nc, cudaMemcpyDeviceToHost, s));
            if (scaling_host)
                ATTRICI_CUDA_CHECK(cudaMemcpyAsync(scaling_host + 2 * (size_t)c0, B.scaling, 16 * (size_t)nc,
                                                   cudaMemcpyDeviceToHost, s));
            if (status_host)
                ATTRICI_CUDA_CHECK(cudaMemcpyAsync(status_host + c0, B.status, 4 * (size_t)nc, cudaMemcpyDeviceToHost, s));
            if (n_pass_host)
                ATTRICI_CUDA_CHECK(cudaMemcpyAsync(n_pass_host + c0, B.npass, 4 * (size_t)nc, cudaMemcpyDeviceToHost, s));
        }
        ATTRICI_CUDA_CHECK(cudaStreamSynchronize(st[0]));
        ATTRICI_CUDA_CHECK(cudaStreamSynchronize(st[1]));
        if (ho.replaced_mode == ATTRICI_FLAGS_SPARSE) {
            unsigned long long n = 0;
            ATTRICI_CUDA_CHECK(cudaMemcpy(&n, d_count, 8, cudaMemcpyDeviceToHost));
            *n_sparse_host = (int64_t)n;
            if ((int64_t)n > ho.sparse_capacity) {
                set_error("the sparse `replaced` list holds %lld of %llu entries: raise sparse_capacity", (long long)ho.sparse_capacity, n);
                return ATTRICI_EWORKSPACE;
            }
        }
        return 0;
    };
    rc = run();
    if (rc != 0) { cudaStreamSynchronize(st[0]); cudaStreamSynchronize(st[1]); }
    cudaStreamDestroy(st[0]);
    cudaStreamDestroy(st[1]);
    return rc;
}

int attrici_quantile_map_f32(void* workspace, size_t workspace_bytes, const attrici_variable_t* var,
                             const float* y, const float* y_scaled, int64_t ld, int T, int C,
                             const double* params, const double* scaling, float* cfact, uint8_t* replaced,
                             int64_t ld_out, void* stream) {
    ATTRICI_TRY(check_var(var));
    ATTRICI_TRY(check_shape(T, C, ld));
    if (ld_out < C) { set_error("ld_out %lld < C=%d", (long long)ld_out, C); return ATTRICI_EINVAL; }
    if (!y || !params || !scaling || !cfact || !replaced) { set_error("y / params / scaling / cfact / replaced is null"); return ATTRICI_EINVAL; }
    if (!qmap_fold_lean_applies(*var, nullptr, replaced)) {
        set_error("a float32 counterfactual is available for Normal variables without a post rule (tas, ps, rlds, ...)");
        return ATTRICI_EUNSUPPORTED;
    }
    Workspace ws;
    ATTRICI_TRY(bind_workspace(ws, workspace, workspace_bytes, T, C, var->family));
    int h[1] = {1};
    ATTRICI_TRY(read_ints(ws.fold_bad, h, 1, (cudaStream_t)stream));
    if (h[0] != 0 || !fold_enabled()) {
        set_error("a float32 counterfactual needs a gap-free daily time axis (the phase-ordered kernel)");
        return ATTRICI_EUNSUPPORTED;
    }
    return launch_qmap(ws, *var, y, y_scaled, ld, params, scaling, nullptr, nullptr, replaced, nullptr, ld_out,
                       (cudaStream_t)stream, cfact);
}

}  // extern "C"
