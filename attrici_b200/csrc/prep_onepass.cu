// One pass over the series for the Normal family on a gap-free daily axis: min / max AND the per-phase sums.
//
// prep.cu reads the series twice before the fit - a min / max pass, then the scaling pass that gathers the
// per-phase sums of y_scaled = RN_f32((y - min) / (max - min)) (variables.py:92-111).  The sums do not need the
// scaled values: with u = y - pivot (pivot = the cell's first value, so that u stays of the size of the range and
// the later cancellations cost no digits; exact in FP64) and
//     y_scaled = (u - m) / s ,   m = datamin - pivot ,  s = scale
// the fused fit (fit_fused.cu) works on  sum u, sum u^2, sum u g  per phase and maps its predictors into u-space.
// This kernel therefore gathers those sums while it takes the min / max, and the series is read from HBM once
// before the quantile map instead of twice.
//
// What differs from the two-pass path: the sums are those of the UNROUNDED scaled values (the reference fits the
// float32-rounded ones).  Measured effect on logp: ~2e-9 relative, on the coefficients below 1e-7 - far inside the
// tolerances of BASELINE.json (1e-6 on the NLL); the two-pass path stays available (ATTRICI_PREP_EXACT_SUMS) and a
// GPU test compares the two.  y_scaled itself is not produced here: the quantile map recomputes it from y with the
// same float32 operations (bit-exact, qmap_fold.cu).
//
// Mapping: thread = V cells, CTA = 128 threads x up to 16 phases, days of a phase in batches of independent loads
// (as prep_scale_fold_kernel).  Results are stored cell-major ([cell][3][FUSED_SP]) through a shared-memory stage of
// four phases, so that every global store is a full 32-byte sector.  Cells with missing days: the valid count of
// every (cell, phase) is stored (one byte); a phase that lost days also gets its own {n, sum g, sum g^2}.
#include <stdlib.h>
#include "common.cuh"
#include "bulk.cuh"

namespace attrici {

namespace {

constexpr int OP_THREADS = 128;
constexpr int OP_MAXPH = 128;  // phases per CTA at most (bounds the staged gmt rows: 128 x 30 x 8 bytes)
constexpr int OP_U = 10;
constexpr int OP_STAGE = 4;    // phases per store group

struct OnePassArgs {
    const float* y;
    int64_t ld;
    int T, C, Cp, i0, i1;
    int ph_len, n_chunks;
    int scaling;
    int prefetch;            // request the batch after next into L2 (ATTRICI_PREP_PREFETCH=0 turns it off)
    float in_lo, in_hi;      // in-place input masks (hurs / tasrange / wind), -Inf / +Inf = none (Inf -> NaN, detrend.py:311)
    float fit_lo, fit_hi;    // values excluded from the fit only (tasskew)
    const double* basis64;
    // per (chunk, cell) partials
    float* p_min; float* p_max;
    double* p_n; double* p_su; double* p_suu; int* p_first;
    // results
    double* ysumT; double* gsumT; uint8_t* cntT;
    float* pivot;            // [Cp]
    float* cmin; float* cscale;
    double* aff; int* miss;
    double* st_cnt; double* st_sum; double* st_sq; double* st_ln; int* st_first;
    double* scaling_out;
};

__device__ __forceinline__ float op_nan() { return __int_as_float(0x7fc00000); }

// ---- bulk asynchronous copies (TMA engine, cp.async.bulk) completing on an mbarrier ----------------------------

// barrier of the OP_THREADS consumer threads only (the producer warp of the TMA variant has left by then)
__device__ __forceinline__ void op_consumer_sync() { asm volatile("bar.sync 1, %0;" ::"n"(OP_THREADS) : "memory"); }

template <int V> __device__ __forceinline__ void op_load(const float* p, float (&v)[V]);
template <> __device__ __forceinline__ void op_load<1>(const float* p, float (&v)[1]) { v[0] = __ldg(p); }
template <> __device__ __forceinline__ void op_load<2>(const float* p, float (&v)[2]) {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
}

// MASKS: the variable has threshold masks (in place: hurs / tasrange / wind; fit only: tasskew); without them the only
// invalid values are NaN and +-Inf, one comparison per value
// TMA: the CTA's rows are brought into a shared-memory ring by bulk asynchronous copies (cp.async.bulk, one 1 KB row
// segment per day, issued by one thread and completing on an mbarrier per stage) instead of per-thread loads into a
// register double buffer: OP_NS batches of OP_U days ahead of the arithmetic, no load / address instruction and no
// register copy in the issue stream.  Needs 16-byte aligned row segments (leading dimension and C multiples of 4).
template <int V, bool MASKS, int TMA>
__global__ void __launch_bounds__(OP_THREADS + 32) prep_onepass_kernel(OnePassArgs a, int nj_max) {
    constexpr int CW = OP_THREADS * V;
    constexpr int OP_NS = TMA > 0 ? TMA : 1;   // stages of the input ring of the TMA variant (each one batch of OP_U days)
    extern __shared__ __align__(16) unsigned char op_smem[];
    double* g_s = reinterpret_cast<double*>(op_smem);                       // [ph_len][nj_max]
    double* stage = g_s + a.ph_len * nj_max;                                // [OP_STAGE][3][CW]
    int* stage_n = reinterpret_cast<int*>(stage + OP_STAGE * 3 * CW);       // [OP_STAGE][CW]
    float* ring = reinterpret_cast<float*>(stage_n + OP_STAGE * CW);        // [OP_NS][OP_U][CW]   (TMA)
    unsigned long long* full = reinterpret_cast<unsigned long long*>(ring + (TMA ? OP_NS * OP_U * CW : 0));   // [OP_NS]
    unsigned long long* empty = full + OP_NS;                                                                  // [OP_NS]

    const int tid = threadIdx.x;
    const int cta_cell0 = blockIdx.x * CW;
    const int cell0 = cta_cell0 + tid * V;
    const int col0 = min(cell0, ((a.C - 1) / V) * V);   // threads beyond the batch read valid columns, store nothing
    const int chunk = blockIdx.y;
    const int n_ph = NPHASE;
    const int d0 = chunk * a.ph_len, d1 = min(n_ph, d0 + a.ph_len);
    const int nph = d1 - d0;
    for (int i = tid; i < nph * nj_max; i += (int)blockDim.x) {
        const int p = i / nj_max, j = i - p * nj_max;
        const int t = d0 + p + j * NPHASE;
        g_s[i] = (t < a.T) ? a.basis64[(size_t)t * B_STRIDE] : 0.0;
    }
    // pivot: the cell's first value (0 if that is missing)
    float pivf[V];
    double piv[V];
    {
        float first[V];
        op_load<V>(a.y + col0, first);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float x = first[v];
            if (!(fabsf(x) <= 3.0e38f)) x = 0.0f;
            pivf[v] = x; piv[v] = (double)x;
            if (chunk == 0 && cell0 + v < a.C) a.pivot[cell0 + v] = x;
        }
    }
    __syncthreads();

    const float in_lo = a.in_lo, in_hi = a.in_hi, fit_lo = a.fit_lo, fit_hi = a.fit_hi;
    float mn[V], mx[V];
    double tot_u[V], tot_uu[V];
    int tot_n[V], first_day[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { mn[v] = INFINITY; mx[v] = -INFINITY; tot_u[v] = tot_uu[v] = 0.0; tot_n[v] = 0; first_day[v] = 0x7fffffff; }
    const size_t jstride = (size_t)NPHASE * a.ld;

    auto days_of = [&](int p) { return (a.T - (d0 + p) + NPHASE - 1) / NPHASE; };
    auto load_batch = [&](int p, int j0, float (&vv)[OP_U][V]) {
        const size_t row0 = (size_t)(d0 + p + j0 * NPHASE) * a.ld + col0;
        const int last = days_of(p) - 1 - j0;
#pragma unroll
        for (int u = 0; u < OP_U; ++u) op_load<V>(a.y + row0 + (size_t)min(u, last) * jstride, vv[u]);
    };
    // the batch after the next one is requested into L2 while the next one is on its way into registers: its loads then
    // find their lines in L2 (the kernel waits on its loads - long scoreboard 3 per issued instruction - and has no
    // registers left for a second batch in flight)
    auto prefetch_batch = [&](int p, int j0) {
        const size_t row0 = (size_t)(d0 + p + j0 * NPHASE) * a.ld + col0;
        const int last = days_of(p) - 1 - j0;
#pragma unroll
        for (int u = 0; u < OP_U; ++u)
            if (u <= last) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.y + row0 + (size_t)u * jstride));
    };
    // TMA variant: warp 4 is the producer.  Its lane 0 walks the CTA's batches, waits until the consumer warps have
    // released the stage (empty barrier, one arrival per consumer warp), arms the full barrier with the byte count and
    // issues one bulk copy per day; the four consumer warps never synchronise with each other for the input.
    if (TMA) {
        if (tid == 0) {
            for (int st = 0; st < OP_NS; ++st) { bulk::mbar_init(full + st, 1); bulk::mbar_init(empty + st, OP_THREADS / 32); }
            bulk::mbar_init_fence();
        }
        __syncthreads();
        if (tid >= OP_THREADS) {
            if (tid == OP_THREADS) {
                const unsigned row_bytes = 4u * (unsigned)min(CW, a.C - cta_cell0);
                int b = 0;
                for (int pp = 0; pp < nph; ++pp) {
                    const int njp = days_of(pp);
                    for (int pj = 0; pj < njp; pj += OP_U, ++b) {
                        const int st = b % OP_NS;
                        if (b >= OP_NS) bulk::mbar_wait(empty + st, (unsigned)(((b / OP_NS) - 1) & 1));
                        const int nu = min(OP_U, njp - pj);
                        bulk::mbar_expect_tx(full + st, row_bytes * (unsigned)nu);
                        const float* src = a.y + (size_t)(d0 + pp + pj * NPHASE) * a.ld + cta_cell0;
                        for (int u = 0; u < nu; ++u)
                            bulk::g2s(ring + ((size_t)st * OP_U + u) * CW, src + (size_t)u * jstride, row_bytes, full + st);
                    }
                }
            }
            return;
        }
    }

    double su[V], suu[V], sug[V];
    int cnt[V];
    float va[OP_U][V], vb[OP_U][V];
    int p = 0, j0 = 0, ja = 0, jb = 0, batch = 0;
    if (!TMA) load_batch(0, 0, va);
    while (true) {
        const int d = d0 + p;
        const int nj = days_of(p);
        int pn = p, jn = j0 + OP_U;
        if (jn >= nj) { pn = p + 1; jn = 0; }
        const bool has_next = pn < nph;
        if (TMA) {
            // this batch's rows have landed in stage batch % OP_NS; each thread takes its V columns of the OP_U days
            const int st = batch % OP_NS;
            bulk::mbar_wait(full + st, (unsigned)((batch / OP_NS) & 1));
            const float* rp = ring + (size_t)st * OP_U * CW + tid * V;
#pragma unroll
            for (int u = 0; u < OP_U; ++u) {
                if (V == 2) {
                    const float2 t2 = *reinterpret_cast<const float2*>(rp + u * CW);
                    va[u][0] = t2.x; va[u][V - 1] = t2.y;
                } else {
                    va[u][0] = rp[u * CW];
                }
            }
        } else if (has_next) {
            load_batch(pn, jn, vb);
            if (a.prefetch) {
                int p2 = pn, j2 = jn + OP_U;
                if (j2 >= days_of(pn)) { p2 = pn + 1; j2 = 0; }
                if (p2 < nph) prefetch_batch(p2, j2);
            }
        }
        if (j0 == 0) {
            ja = a.i0 > d ? (a.i0 - d + NPHASE - 1) / NPHASE : 0;
            jb = a.i1 > d ? min(nj, (a.i1 - d + NPHASE - 1) / NPHASE) : 0;
#pragma unroll
            for (int v = 0; v < V; ++v) { su[v] = suu[v] = sug[v] = 0.0; cnt[v] = 0; }
        }
        const double* gp = g_s + p * nj_max + j0;
        const int last = nj - 1 - j0;
        const bool whole = last >= OP_U - 1 && j0 >= ja && j0 + OP_U <= jb;   // batch inside the phase and the window
        if (whole) {
#pragma unroll
            for (int u = 0; u < OP_U; ++u) {
                const double g = gp[u];
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float x = va[u][v];
                    bool in_range, used;
                    if (MASKS) {
                        in_range = !(x <= in_lo || x >= in_hi) && (x == x);
                        used = in_range && !(x <= fit_lo || x >= fit_hi);
                    } else {
                        in_range = used = fabsf(x) < INFINITY;          // false for NaN and +-Inf (detrend.py:311)
                    }
                    const float xm = in_range ? x : op_nan();
                    mn[v] = fminf(mn[v], xm); mx[v] = fmaxf(mx[v], xm);    // NaN-skipping (variables.py:108-110)
                    const float xs = used ? x : pivf[v];                   // unused days contribute u = 0
                    const double uu = (double)xs - piv[v];
                    su[v] += uu; suu[v] = fma(uu, uu, suu[v]); sug[v] = fma(uu, g, sug[v]);
                    cnt[v] += used ? 1 : 0;
                }
            }
        } else {
#pragma unroll
            for (int u = 0; u < OP_U; ++u) {
                const bool in_phase = u <= last;
                const bool in_win = in_phase && j0 + u >= ja && j0 + u < jb;
                const double g = gp[min(u, last)];
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    float x = va[u][v];
                    if (x <= in_lo || x >= in_hi) x = op_nan();     // detrend.py:311, mask_thresholded (in place)
                    if (in_phase) { mn[v] = fminf(mn[v], x); mx[v] = fmaxf(mx[v], x); }
                    const bool used = in_win && (x == x) && !(x <= fit_lo || x >= fit_hi);
                    const float xs = used ? x : pivf[v];
                    const double uu = (double)xs - piv[v];
                    su[v] += uu; suu[v] = fma(uu, uu, suu[v]); sug[v] = fma(uu, g, sug[v]);
                    cnt[v] += used ? 1 : 0;
                }
            }
        }
        if (TMA) {
            // The stage is released only here, after the arithmetic has consumed every value read from it.  Releasing
            // it right after the shared-memory loads were ISSUED is a race: mbarrier.arrive does not wait for the
            // warp's loads in flight, and the refill (three batches ahead, often an L2 hit) can land first - measured
            // as a few hundred cells per 10 000 whose sums changed from run to run (scripts/check_race.py).
            __syncwarp();
            if ((tid & 31) == 0) bulk::mbar_arrive(empty + (batch % OP_NS));
        }
        if (pn != p) {
            // phase complete
            const int expected = jb - ja;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                tot_u[v] += su[v]; tot_uu[v] += suu[v]; tot_n[v] += cnt[v];
                if (cnt[v] == expected) {
                    if (expected > 0) first_day[v] = min(first_day[v], d + ja * NPHASE);
                } else if (cell0 + v < a.C) {
                    // days are missing in this phase: its own {n, sum g, sum g^2} and first used day (rare)
                    double n = 0.0, sg = 0.0, sgg = 0.0;
                    const float* yp = a.y + (size_t)d * a.ld + min(col0 + v, a.C - 1);
                    const double* gq = g_s + p * nj_max;
                    int fj = 0x7fffffff;
                    for (int j = ja; j < jb; ++j) {
                        float x = __ldg(yp + (size_t)j * jstride);
                        if (x <= in_lo || x >= in_hi) x = op_nan();
                        if ((x == x) && !(x <= fit_lo || x >= fit_hi)) {
                            const double g = gq[j];
                            n += 1.0; sg += g; sgg = fma(g, g, sgg);
                            fj = min(fj, j);
                        }
                    }
                    if (fj != 0x7fffffff) first_day[v] = min(first_day[v], d + fj * NPHASE);
                    if (cell0 + v < a.C) {
                        double* go = a.gsumT + (size_t)(cell0 + v) * 3 * FUSED_SP + d;
                        go[0] = n; go[FUSED_SP] = sg; go[2 * FUSED_SP] = sgg;
                    }
                }
                const int slot = p & (OP_STAGE - 1);
                stage[(slot * 3 + 0) * CW + tid * V + v] = su[v];
                stage[(slot * 3 + 1) * CW + tid * V + v] = suu[v];
                stage[(slot * 3 + 2) * CW + tid * V + v] = sug[v];
                stage_n[slot * CW + tid * V + v] = cnt[v];
            }
            if ((p & (OP_STAGE - 1)) == OP_STAGE - 1 || !has_next) {
                // store the group of (up to) four phases cell-major: 32 bytes per (cell, sum) - one full sector
                const int g0 = p & ~(OP_STAGE - 1);
                const int ng = p - g0 + 1;
                const int dg = d0 + g0;
                if (TMA) op_consumer_sync(); else __syncthreads();
                for (int i = tid; i < 3 * CW; i += OP_THREADS) {
                    const int k = i / CW, c = i - k * CW;
                    if (cta_cell0 + c < a.C) {
                        double* dst = a.ysumT + ((size_t)(cta_cell0 + c) * 3 + k) * FUSED_SP + dg;
                        if (ng == OP_STAGE) {
                            const double2 lo = make_double2(stage[(0 * 3 + k) * CW + c], stage[(1 * 3 + k) * CW + c]);
                            const double2 hi = make_double2(stage[(2 * 3 + k) * CW + c], stage[(3 * 3 + k) * CW + c]);
                            reinterpret_cast<double2*>(dst)[0] = lo;
                            reinterpret_cast<double2*>(dst)[1] = hi;
                        } else {
                            for (int s = 0; s < ng; ++s) dst[s] = stage[(s * 3 + k) * CW + c];
                        }
                        // the padding of the rows (phases 1461 .. FUSED_SP - 1) reads as zero (fit_fused.cu)
                        if (dg + ng == NPHASE)
                            for (int s = NPHASE; s < FUSED_SP; ++s) dst[s - dg] = 0.0;
                    }
                }
                for (int c = tid; c < CW; c += OP_THREADS) {
                    if (cta_cell0 + c < a.C) {
                        uint8_t* dst = a.cntT + (size_t)(cta_cell0 + c) * FUSED_SP + dg;
                        if (ng == OP_STAGE) {
                            const uint32_t w = (uint32_t)stage_n[0 * CW + c] | ((uint32_t)stage_n[1 * CW + c] << 8) |
                                               ((uint32_t)stage_n[2 * CW + c] << 16) | ((uint32_t)stage_n[3 * CW + c] << 24);
                            *reinterpret_cast<uint32_t*>(dst) = w;
                        } else {
                            for (int s = 0; s < ng; ++s) dst[s] = (uint8_t)stage_n[s * CW + c];
                        }
                        if (dg + ng == NPHASE)
                            for (int s = NPHASE; s < FUSED_SP; ++s) dst[s - dg] = 0;
                    }
                }
                if (TMA) op_consumer_sync(); else __syncthreads();
            }
        }
        if (!has_next) break;
        if (!TMA) {
#pragma unroll
            for (int u = 0; u < OP_U; ++u)
#pragma unroll
                for (int v = 0; v < V; ++v) va[u][v] = vb[u][v];
        }
        p = pn; j0 = jn; ++batch;
    }
#pragma unroll
    for (int v = 0; v < V; ++v) {
        if (cell0 + v < a.C) {
            const size_t idx = (size_t)chunk * a.Cp + cell0 + v;
            a.p_min[idx] = mn[v]; a.p_max[idx] = mx[v];
            a.p_n[idx] = (double)tot_n[v]; a.p_su[idx] = tot_u[v]; a.p_suu[idx] = tot_uu[v]; a.p_first[idx] = first_day[v];
        }
    }
}

// per cell: datamin / scale (float32 arithmetic of variables.py:108-110), the affine map of the sums, the flag of
// missing days and the start statistics of the fit window (in units of the unrounded scaled values)
__global__ void prep_onepass_finalize_kernel(OnePassArgs a) {
    const int cell = blockIdx.x * OP_THREADS + threadIdx.x;
    if (cell >= a.C) return;
    float mn = INFINITY, mx = -INFINITY;
    double n = 0.0, su = 0.0, suu = 0.0;
    int first = 0x7fffffff;
    for (int ch = 0; ch < a.n_chunks; ++ch) {
        const size_t o = (size_t)ch * a.Cp + cell;
        mn = fminf(mn, a.p_min[o]); mx = fmaxf(mx, a.p_max[o]);
        n += a.p_n[o]; su += a.p_su[o]; suu += a.p_suu[o];
        first = min(first, a.p_first[o]);
    }
    const bool any = mx >= mn;
    float datamin = 0.0f, scale = 1.0f;
    switch (a.scaling) {
        case ATTRICI_SCALE_UNITY:
            datamin = any ? mn : op_nan();
            scale = any ? __fsub_rn(mx, mn) : op_nan();
            break;
        case ATTRICI_SCALE_RANGE: scale = any ? __fsub_rn(mx, mn) : op_nan(); break;
        case ATTRICI_SCALE_PERCENT: scale = 100.0f; break;
        default: scale = 1.0f; break;   // ATTRICI_SCALE_NONE
    }
    const bool ok = scale > 0.0f && (a.scaling != ATTRICI_SCALE_UNITY || datamin == datamin);
    const double m = (a.scaling == ATTRICI_SCALE_UNITY ? (double)datamin : 0.0) - (double)a.pivot[cell];
    const double s = ok ? (double)scale : 1.0;
    a.aff[2 * (size_t)cell] = m;
    a.aff[2 * (size_t)cell + 1] = s;
    a.miss[cell] = (n < (double)(a.i1 - a.i0)) ? 1 : 0;
    a.cmin[cell] = datamin;
    a.cscale[cell] = scale;
    a.scaling_out[2 * (size_t)cell] = (double)datamin;
    a.scaling_out[2 * (size_t)cell + 1] = (double)scale;
    // a degenerate scale turns every scaled value into NaN / Inf: no data for the fit
    const double nn = ok ? n : 0.0;
    const double is = 1.0 / s;
    a.st_cnt[cell] = nn;
    a.st_sum[cell] = (su - n * m) * is;
    a.st_sq[cell] = (suu - 2.0 * m * su + n * m * m) * is * is;
    a.st_ln[cell] = 0.0;
    a.st_first[cell] = (first == 0x7fffffff || !ok) ? -1 : first;
    const size_t c1 = (size_t)a.Cp + cell;
    a.st_cnt[c1] = 0.0; a.st_sum[c1] = 0.0; a.st_sq[c1] = 0.0; a.st_ln[c1] = 0.0; a.st_first[c1] = -1;
}

}  // namespace

bool prep_onepass_applies(const Workspace& ws, const attrici_variable_t& var, const float* y, const float* y_scaled, int64_t ld,
                          int flags) {
    static const bool off = [] { const char* e = getenv("ATTRICI_PREP_EXACT_SUMS"); return e && e[0] == '1'; }();
    if (off || (flags & ATTRICI_PREP_EXACT_SUMS) || var.family != ATTRICI_NORMAL || y_scaled != nullptr || ws.T < NPHASE || !ws.ysumT) return false;
    if ((ws.T + NPHASE - 1) / NPHASE > 255) return false;   // per-phase day counts are stored as bytes
    (void)y; (void)ld;
    return true;
}

int launch_prep_onepass(Workspace& ws, const attrici_variable_t& var, const float* y, int64_t ld, int i0, int i1,
                        double* scaling, cudaStream_t s) {
    OnePassArgs a;
    a.y = y; a.ld = ld; a.T = ws.T; a.C = ws.C; a.Cp = ws.Cp; a.i0 = i0; a.i1 = i1;
    a.scaling = var.scaling;
    static const bool prefetch_on = [] { const char* e = getenv("ATTRICI_PREP_PREFETCH"); return !(e && e[0] == '0'); }();
    a.prefetch = prefetch_on ? 1 : 0;
    const bool has_lo = !(var.lower_threshold != var.lower_threshold);
    const bool has_hi = !(var.upper_threshold != var.upper_threshold);
    const bool in_place = (var.scaling == ATTRICI_SCALE_PERCENT || var.scaling == ATTRICI_SCALE_RANGE);
    const bool only_fit = (var.scaling == ATTRICI_SCALE_NONE);
    a.in_lo = (in_place && has_lo) ? var.lower_threshold : -INFINITY;
    a.in_hi = (in_place && has_hi) ? var.upper_threshold : INFINITY;
    a.fit_lo = (only_fit && has_lo) ? var.lower_threshold : -INFINITY;
    a.fit_hi = (only_fit && has_hi) ? var.upper_threshold : INFINITY;
    a.basis64 = ws.basis64;
    static const bool force_v1 = [] { const char* e = getenv("ATTRICI_ONEPASS_V"); return e && e[0] == '1'; }();   // tuning only
    const bool vec2 = (ld % 2 == 0) && ((uintptr_t)y % 8 == 0) && ws.C >= 2 && !force_v1;
    const int V = vec2 ? 2 : 1;
    const int n_cb = (ws.C + OP_THREADS * V - 1) / (OP_THREADS * V);
    // partials live in the (otherwise idle) moment-partial buffer: 6 arrays of [chunk][Cp] 8-byte slots
    const size_t part_cap = ws.part_acc_cap / (6 * (size_t)ws.Cp);
    int want = (148 * 16 + n_cb - 1) / n_cb;
    if ((size_t)want > part_cap) want = (int)part_cap;
    if (want > NPHASE) want = NPHASE;
    if (want < 1) want = 1;
    int ph_len = (NPHASE + want - 1) / want;
    ph_len = (ph_len + OP_STAGE - 1) / OP_STAGE * OP_STAGE;   // store groups start at multiples of four phases
    if (ph_len > OP_MAXPH) ph_len = OP_MAXPH;
    a.ph_len = ph_len;
    a.n_chunks = (NPHASE + ph_len - 1) / ph_len;
    if ((size_t)a.n_chunks > part_cap) { set_error("workspace too small for the one-pass scaling kernel"); return ATTRICI_EWORKSPACE; }
    // 16 CTAs per SM are wanted (loads in flight), but not more than ~96 chunks: the per-chunk partials cost a
    // finalisation pass over [chunks x cells]
    if (a.n_chunks > 96) {
        a.ph_len = ((NPHASE + 95) / 96 + OP_STAGE - 1) / OP_STAGE * OP_STAGE;
        a.n_chunks = (NPHASE + a.ph_len - 1) / a.ph_len;
    }
    const size_t stride = part_cap * ws.Cp;
    double* base = reinterpret_cast<double*>(ws.part_acc);
    a.p_n = base; a.p_su = base + stride; a.p_suu = base + 2 * stride;
    a.p_min = reinterpret_cast<float*>(base + 3 * stride);
    a.p_max = reinterpret_cast<float*>(base + 4 * stride);
    a.p_first = reinterpret_cast<int*>(base + 5 * stride);
    a.ysumT = ws.ysumT; a.gsumT = ws.gsumT; a.cntT = ws.cntT; a.pivot = ws.pivot;
    a.cmin = ws.cmin; a.cscale = ws.cscale; a.aff = ws.aff; a.miss = ws.miss;
    a.st_cnt = ws.st_cnt; a.st_sum = ws.st_sum; a.st_sq = ws.st_sq; a.st_ln = ws.st_ln; a.st_first = ws.st_first;
    a.scaling_out = scaling;
    const int nj_max = (ws.T + NPHASE - 1) / NPHASE;
    // bulk asynchronous copies need 16-byte aligned row segments
    // measured (10 000 cells): 0.752 / 0.748 / 0.798 ms with 3 / 4 / 5 stages against 0.719 ms for the per-thread loads
    // once a stage is released after use (the early release that made it 0.66 ms is a race, see the kernel): more bytes
    // in flight do not help, the kernel is bound by the latency of its FP64 accumulation chains at 12 - 16 warps per SM.
    // The variant is opt-in: ATTRICI_TMA_PREP = number of stages of the ring (3 or 4; 1 = 3)
    static const int tma_ns = [] {
        const char* e = getenv("ATTRICI_NO_TMA"); const char* f = getenv("ATTRICI_TMA_PREP");
        if ((e && e[0] == '1') || !f) return 0;
        const int n = atoi(f);
        return n == 1 ? 3 : (n == 3 || n == 4 ? n : 0);
    }();
    const int OP_NS = tma_ns;
    const bool tma = tma_ns > 0 && V == 2 && (ld % 4 == 0) && ((uintptr_t)y % 16 == 0) && (ws.C % 4 == 0);
    const size_t smem = sizeof(double) * ((size_t)a.ph_len * nj_max + (size_t)OP_STAGE * 3 * OP_THREADS * V) +
                        sizeof(int) * (size_t)OP_STAGE * OP_THREADS * V +
                        (tma ? sizeof(float) * (size_t)OP_NS * OP_U * OP_THREADS * V + 16 * OP_NS : 0);
    if (smem > 200 * 1024) { set_error("series too long for the one-pass scaling kernel (T=%d)", ws.T); return ATTRICI_EUNSUPPORTED; }
    dim3 grid(n_cb, a.n_chunks);
    const bool masks = (in_place || only_fit) && (has_lo || has_hi);
    const int threads = tma ? OP_THREADS + 32 : OP_THREADS;   // the TMA variant adds its producer warp
    auto launch = [&](auto kernel) -> int {
        ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<grid, threads, smem, s>>>(a, nj_max);
        return 0;
    };
    if (tma && OP_NS == 3) ATTRICI_TRY(masks ? launch(prep_onepass_kernel<2, true, 3>) : launch(prep_onepass_kernel<2, false, 3>));
    else if (tma) ATTRICI_TRY(masks ? launch(prep_onepass_kernel<2, true, 4>) : launch(prep_onepass_kernel<2, false, 4>));
    else if (V == 2) ATTRICI_TRY(masks ? launch(prep_onepass_kernel<2, true, 0>) : launch(prep_onepass_kernel<2, false, 0>));
    else ATTRICI_TRY(masks ? launch(prep_onepass_kernel<1, true, 0>) : launch(prep_onepass_kernel<1, false, 0>));
    ATTRICI_LAUNCH_CHECK();
    prep_onepass_finalize_kernel<<<(ws.C + OP_THREADS - 1) / OP_THREADS, OP_THREADS, 0, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
