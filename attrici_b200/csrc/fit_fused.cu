// Whole Fisher-scoring iteration of the phase-folded Normal fit in ONE persistent kernel.
//
// fold.cu evaluates a likelihood pass with thread = cell: 129 moment accumulators per thread (255 registers,
// 8 warps per SM), partial sums per chunk of phases, a reduction kernel, a step kernel and a host round trip per
// block of passes.  Here a warp owns a cell for its whole iteration:
//
//   * lane = phase pair (d, 1461 - d): 23 sweeps of 32 pairs cover the 731 pairs of a cell.  The cell's per-phase
//     sums are stored cell-major ([cell][3][SP], prep.cu), so a sweep reads three (+ three mirrored) contiguous
//     256-byte segments; after the first pass they come from L2 (1 776 cells in flight x 35 KB = 62 MB < 126 MB).
//   * the shared columns (cos / sin 1..4 in FP64, {n, sum g, sum g^2} of both phases of a pair, cos / sin 1..8 as
//     TF32) are tables indexed by the pair, read through L1.
//   * gradient moments (27 per cell) are accumulated per lane in FP64 and all-reduced once per pass.
//   * Fisher moments (4 weight streams x 17 trig columns) are the contraction [weights x pairs] . [pairs x trig]:
//     they run on the tensor cores (mma.sync m16n8k8 TF32, FP32 accumulate).  A rows = {4 cosine-type weights,
//     4 sine-type weights} as TF32 high parts and the same 8 as TF32 low parts (the weights carry e^{-2c}, whose
//     range over cells is wide; the trig columns are rounded once), B = the TF32 trig table.  The accumulator
//     fragments replace 68 per-thread registers, and the cross-lane sum over pairs is part of the MMA.  The matrix
//     only steers the damped step; the gradient, the objective and the solve stay FP64.
//   * the accept / damp / Cholesky / propose step of step_impl.cuh runs in the same warp right after the pass, the
//     moments never leave the SM except through the L1-resident momd rows the step code reads.
//
// No partial buffers, no reduction kernel, no host synchronisation inside the fit: one launch per term.
//
// Sums may be those of the scaled series (aff = (0, 1)) or of u = y - pivot with y_scaled = (u - m) / s per cell
// (one-pass prep): the predictors are mapped to u-space per phase (a' = m + s a, b' = s b) and the residual sums
// scaled back (R / s, Q / s^2), which is the same function of the coefficients.
//
// Replaces DistributionScipy.log_likelihood (attrici/estimation/model_scipy.py:311-331) and the L-BFGS-B /
// finite-difference loop of ModelScipy.fit (model_scipy.py:505-527), like fold.cu + step.cu.
#include <stdlib.h>
#include "common.cuh"
#include "step_warp.cuh"

namespace attrici {

namespace {

// one CTA per SM (the pair tables, 127 KB, live in shared memory once).  12 warps leave 168 registers per thread, with
// which the sweep spills a few loop-invariant values; registers are allocated for multiples of four warps, so the
// alternative without spills is 8 warps (255 registers) - ATTRICI_FUSED_WARPS=8 selects it (A/B measurements)
constexpr int FF_WARPS_MAX = 12;
constexpr int FF_WPITCH = 36;          // floats per row of the weight transposition buffer (36 = 4 mod 32: conflict-free fragments)
constexpr int FF_SWEEPS = FUSED_TP / 32;   // 23

struct FusedArgs {
    StepArgs st;
    const HEntry* tab_g;
    const double* ysumT;     // [Cp][3][FUSED_SP] per-phase sums {u, u^2, u g}, cell-major
    const double* gsumT;     // [Cp][3][FUSED_SP] per-phase {n, sum g, sum g^2} of cells with missing days
    const double* aff;       // [Cp][2] {m, s}: y_scaled = (u - m) / s
    const int* miss;         // [Cp]
    const uint8_t* cntT;     // [Cp][FUSED_SP] valid window days per phase
    const double* tab_d;     // [FUSED_ND][FUSED_TP]
    const float* tab_f;      // [16][FUSED_FP] TF32-rounded cos 1..8, sin 1..8 per pair
    int* work;               // next cell to hand out
    int fisher_passes, chord_passes;
};

__device__ __forceinline__ uint32_t ff_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void ff_mma(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                       uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ double ff_wsum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One likelihood pass of `cell` at th (shared-frame coefficients, 27 doubles in shared memory); results go to the
// cell's momd rows (global, read back by the step code of this warp).
//   mode PASS_NLL: row NACC only; PASS_GRAD: + gradient moments; PASS_FULL / PASS_FLAT: + Fisher moments
//   (PASS_FLAT: only the two intercepts are non-zero - no predictor chain, one exp)
// shared-memory load the compiler may not merge with an earlier one of the same address (RELOAD: the trig values of the
// pair are read again for the gradient moments instead of staying in 16 registers across the exp / residual section)
__device__ __forceinline__ double ff_lds(const double* p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

template <bool RELOAD, bool PREFETCH>
__device__ __forceinline__ void fused_pass(const FusedArgs& a, const double* tab_d, const float* tab_f, int cell, int mode,
                                           const double* th, uint32_t* W, double* mom) {
    const int lane = threadIdx.x & 31;
    const int gid = lane >> 2, tig = lane & 3;
    const bool flat = mode == PASS_FLAT;
    const bool fisher = mode == PASS_FULL || mode == PASS_FLAT;
    const bool grad = mode != PASS_NLL;
    const double* ys = a.ysumT + (size_t)cell * 3 * FUSED_SP;
    const bool own = a.miss[cell] != 0;
    const double* gs = a.gsumT + (size_t)cell * 3 * FUSED_SP;
    const double am_ = a.aff[2 * (size_t)cell], as_ = a.aff[2 * (size_t)cell + 1];
    const double is1 = 1.0 / as_, is2 = is1 * is1;

    double GA[2 * NB], GB[NB];
#pragma unroll
    for (int i = 0; i < 2 * NB; ++i) GA[i] = 0.0;
#pragma unroll
    for (int i = 0; i < NB; ++i) GB[i] = 0.0;
    float cc[4] = {0.f, 0.f, 0.f, 0.f}, cs[4] = {0.f, 0.f, 0.f, 0.f};   // accumulator fragments: cos / sin tiles
    float fc[4] = {0.f, 0.f, 0.f, 0.f};                                  // constant moments of the 4 cosine-type weights
    double nll = 0.0;
    const double e2_flat = flat ? exp(-2.0 * th[NA]) : 0.0;

    // the cell's sums of the next sweep are requested before the current one is evaluated.  Lanes beyond the last
    // pair, and the mirror of pair 0, read the zero padding of the rows (phases 1461 .. FUSED_SP - 1): no selects
    // PREFETCH: the next sweep's six segments are requested with prefetch.global.L1 (no registers held across the
    // sweep); otherwise they are loaded one sweep ahead into registers
    double nx[6];
    auto index_of = [&](int sweep, int& qi, int& qm) {
        const int q = sweep * 32 + lane;
        qi = q < FUSED_NPAIR ? q : NPHASE;
        qm = (q < FUSED_NPAIR && q > 0) ? NPHASE - q : NPHASE;
    };
    auto request = [&](int sweep, double (&v)[6]) {
        int qi, qm;
        index_of(sweep, qi, qm);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            v[k] = __ldg(ys + k * FUSED_SP + qi);
            v[3 + k] = __ldg(ys + k * FUSED_SP + qm);
        }
    };
    auto prefetch = [&](int sweep) {
        int qi, qm;
        index_of(sweep, qi, qm);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ys + k * FUSED_SP + qi));
            asm volatile("prefetch.global.L1 [%0];" ::"l"(ys + k * FUSED_SP + qm));
        }
    };
    if (PREFETCH) prefetch(0); else request(0, nx);
#pragma unroll 1
    for (int sweep = 0; sweep < FF_SWEEPS; ++sweep) {
        const int q = sweep * 32 + lane;
        if (PREFETCH) {
            if (sweep + 1 < FF_SWEEPS) prefetch(sweep + 1);
            request(sweep, nx);
        }
        const double sy = nx[0], syy = nx[1], syg = nx[2], sym = nx[3], syym = nx[4], sygm = nx[5];
        if (!PREFETCH && sweep + 1 < FF_SWEEPS) request(sweep + 1, nx);
        double n, sg, sgg, nm, sgm, sggm;
        if (own) {
            // a cell with missing days: its own day count per phase, and its own gmt sums where days are missing
            const int qi = q < FUSED_NPAIR ? q : NPHASE;
            const int qm = (q < FUSED_NPAIR && q > 0) ? NPHASE - q : NPHASE;
            const uint8_t* cn = a.cntT + (size_t)cell * FUSED_SP;
            const double* t = tab_d + q;
            n = (double)cn[qi]; nm = (double)cn[qm];
            const bool lost = n != t[8 * FUSED_TP], lostm = nm != t[11 * FUSED_TP];
            sg = lost ? gs[FUSED_SP + qi] : t[9 * FUSED_TP]; sgg = lost ? gs[2 * FUSED_SP + qi] : t[10 * FUSED_TP];
            sgm = lostm ? gs[FUSED_SP + qm] : t[12 * FUSED_TP]; sggm = lostm ? gs[2 * FUSED_SP + qm] : t[13 * FUSED_TP];
        } else {
            const double* t = tab_d + q;
            n = t[8 * FUSED_TP]; sg = t[9 * FUSED_TP]; sgg = t[10 * FUSED_TP];
            nm = t[11 * FUSED_TP]; sgm = t[12 * FUSED_TP]; sggm = t[13 * FUSED_TP];
        }
        double ck[MAXM], sk[MAXM];
#pragma unroll
        for (int k = 0; k < MAXM; ++k) {
            ck[k] = tab_d[k * FUSED_TP + q];
            sk[k] = tab_d[(MAXM + k) * FUSED_TP + q];
        }
        double av, bv, cv, amv, bmv, cmv, e2, e2m;
        if (flat) {
            av = amv = th[0]; bv = bmv = 0.0; cv = cmv = th[NA]; e2 = e2m = e2_flat;
        } else {
            // cosine and sine halves of the three predictors: phase d takes C + S, phase 1461 - d takes C - S
            double ca = th[0], sa = 0.0, cb = th[1], sb = 0.0, c0 = th[NA], s0 = 0.0;
#pragma unroll
            for (int k = 0; k < MAXM; ++k) {
                ca = fma(th[2 + k], ck[k], ca);
                sa = fma(th[2 + MAXM + k], sk[k], sa);
                cb = fma(th[2 + 2 * MAXM + k], ck[k], cb);
                sb = fma(th[2 + 3 * MAXM + k], sk[k], sb);
                c0 = fma(th[NA + 1 + k], ck[k], c0);
                s0 = fma(th[NA + 1 + MAXM + k], sk[k], s0);
            }
            av = ca + sa; bv = cb + sb; cv = c0 + s0;
            amv = ca - sa; bmv = cb - sb; cmv = c0 - s0;
            e2 = exp(-2.0 * cv); e2m = exp(-2.0 * cmv);
        }
        // predictors in u-space, residual sums scaled back
        const double au = fma(as_, av, am_), bu = as_ * bv, aum = fma(as_, amv, am_), bum = as_ * bmv;
        const double R0 = sy - au * n - bu * sg, R0m = sym - aum * nm - bum * sgm;
        const double R1 = syg - au * sg - bu * sgg, R1m = sygm - aum * sgm - bum * sggm;
        const double Q = (syy - au * sy - bu * syg - au * R0 - bu * R1) * is2;
        const double Qm = (syym - aum * sym - bum * sygm - aum * R0m - bum * R1m) * is2;
        nll += n * (cv + 0.91893853320467274178) + 0.5 * e2 * Q + nm * (cmv + 0.91893853320467274178) + 0.5 * e2m * Qm;
        if (grad) {
            const double g0 = -e2 * R0 * is1, g0m = -e2m * R0m * is1, g1 = -e2 * R1 * is1, g1m = -e2m * R1m * is1;
            const double gb = n - e2 * Q, gbm = nm - e2m * Qm;
            const double w0c = g0 + g0m, w0s = g0 - g0m, w1c = g1 + g1m, w1s = g1 - g1m, wbc = gb + gbm, wbs = gb - gbm;
            GA[0] += w0c; GA[NB] += w1c; GB[0] += wbc;
#pragma unroll
            for (int k = 0; k < MAXM; ++k) {
                const double c_k = RELOAD ? ff_lds(tab_d + k * FUSED_TP + q) : ck[k];
                const double s_k = RELOAD ? ff_lds(tab_d + (MAXM + k) * FUSED_TP + q) : sk[k];
                GA[1 + k] = fma(w0c, c_k, GA[1 + k]);
                GA[1 + MAXM + k] = fma(w0s, s_k, GA[1 + MAXM + k]);
                GA[NB + 1 + k] = fma(w1c, c_k, GA[NB + 1 + k]);
                GA[NB + 1 + MAXM + k] = fma(w1s, s_k, GA[NB + 1 + MAXM + k]);
                GB[1 + k] = fma(wbc, c_k, GB[1 + k]);
                GB[1 + MAXM + k] = fma(wbs, s_k, GB[1 + MAXM + k]);
            }
        }
        if (fisher) {
            const double w0 = e2 * n, w0m = e2m * nm, w1 = e2 * sg, w1m = e2m * sgm, w2 = e2 * sgg, w2m = e2m * sggm;
            float wv[8];
            wv[0] = (float)(w0 + w0m); wv[1] = (float)(w1 + w1m); wv[2] = (float)(w2 + w2m); wv[3] = (float)(2.0 * (n + nm));
            wv[4] = (float)(w0 - w0m); wv[5] = (float)(w1 - w1m); wv[6] = (float)(w2 - w2m); wv[7] = (float)(2.0 * (n - nm));
#pragma unroll
            for (int i = 0; i < 4; ++i) fc[i] += wv[i];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t hi = ff_tf32(wv[i]);
                const uint32_t lo = ff_tf32(wv[i] - __uint_as_float(hi));
                W[i * FF_WPITCH + lane] = hi;
                W[(8 + i) * FF_WPITCH + lane] = lo;
            }
            __syncwarp();
            const float* tf = tab_f + sweep * 32 + tig;
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const uint32_t a0 = W[gid * FF_WPITCH + 8 * s + tig], a1 = W[(gid + 8) * FF_WPITCH + 8 * s + tig];
                const uint32_t a2 = W[gid * FF_WPITCH + 8 * s + tig + 4], a3 = W[(gid + 8) * FF_WPITCH + 8 * s + tig + 4];
                const uint32_t bc0 = __float_as_uint(tf[gid * FUSED_FP + 8 * s]);
                const uint32_t bc1 = __float_as_uint(tf[gid * FUSED_FP + 8 * s + 4]);
                const uint32_t bs0 = __float_as_uint(tf[(8 + gid) * FUSED_FP + 8 * s]);
                const uint32_t bs1 = __float_as_uint(tf[(8 + gid) * FUSED_FP + 8 * s + 4]);
                ff_mma(cc, a0, a1, a2, a3, bc0, bc1);
                ff_mma(cs, a0, a1, a2, a3, bs0, bs1);
            }
            __syncwarp();
        }
    }

    // ---- results -> the warp's moment scratch (layout of pass.cu: ACC_AA / ACC_AB / ACC_BB / ACC_GA / ACC_GB, [NACC] = nll)
    nll = ff_wsum(nll);
    if (lane == 0) mom[NACC] = nll;
    if (grad) {
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i) GA[i] = ff_wsum(GA[i]);
#pragma unroll
        for (int i = 0; i < NB; ++i) GB[i] = ff_wsum(GB[i]);
        // every lane holds the totals: lane i stores entry i (static indexing keeps the arrays in registers)
#pragma unroll
        for (int i = 0; i < 2 * NB; ++i)
            if (lane == i) mom[ACC_GA + i] = GA[i];
#pragma unroll
        for (int i = 0; i < NB; ++i)
            if (lane == 2 * NB + i) mom[ACC_GB + i] = GB[i];
    }
    if (fisher) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float v = fc[i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            fc[i] = v;
        }
        // weight stream p = 0..2: AA block of gmt power p, p = 3: BB block
        const int p = gid & 3;
        const int base = p < 3 ? ACC_AA + p * NTRIG : ACC_BB;
        if (gid < 4) {   // rows 0..3 (+ low parts in rows 8..11) of the cos tile: cos moments 2 tig + 1, 2 tig + 2
            mom[base + 1 + 2 * tig] = (double)cc[0] + (double)cc[2];
            mom[base + 2 + 2 * tig] = (double)cc[1] + (double)cc[3];
        } else {         // rows 4..7 (+ 12..15) of the sin tile
            mom[base + NH + 1 + 2 * tig] = (double)cs[0] + (double)cs[2];
            mom[base + NH + 2 + 2 * tig] = (double)cs[1] + (double)cs[3];
        }
        if (lane < 4) {
            const int b0 = lane < 3 ? ACC_AA + lane * NTRIG : ACC_BB;
            mom[b0] = (double)(lane == 0 ? fc[0] : (lane == 1 ? fc[1] : (lane == 2 ? fc[2] : fc[3])));
        }
        // the mu x sigma block of the Normal information matrix is zero
        for (int i = lane; i < 2 * NTRIG; i += 32) mom[ACC_AB + i] = 0.0;
    }
    __syncwarp();
}

using FusedScratch = CellScratchT<true>;

// dynamic shared memory of one CTA: the pair tables, then per warp {scratch, weight buffer, coefficients}
constexpr size_t FF_TABD_BYTES = sizeof(double) * FUSED_ND * FUSED_TP;
constexpr size_t FF_TABF_BYTES = sizeof(float) * 16 * FUSED_FP;
constexpr size_t FF_HTAB_BYTES = (sizeof(HEntry) * NTRI + 15) / 16 * 16;
constexpr size_t FF_WARP_BYTES = (sizeof(FusedScratch) + sizeof(uint32_t) * 16 * FF_WPITCH + sizeof(double) * 32 + 15) / 16 * 16;
constexpr size_t ff_smem_bytes(int warps) { return FF_TABD_BYTES + FF_TABF_BYTES + FF_HTAB_BYTES + warps * FF_WARP_BYTES; }
static_assert(ff_smem_bytes(FF_WARPS_MAX) <= 227 * 1024, "fused fit: shared memory of one CTA");

template <int FF_WARPS, bool RELOAD, bool PREFETCH>
__global__ void __launch_bounds__(32 * FF_WARPS, 1) fit_fused_kernel(FusedArgs a) {
    constexpr int FF_THREADS = 32 * FF_WARPS;
    extern __shared__ __align__(16) unsigned char ff_smem[];
    double* tab_d = reinterpret_cast<double*>(ff_smem);
    float* tab_f = reinterpret_cast<float*>(ff_smem + FF_TABD_BYTES);
    HEntry* tab = reinterpret_cast<HEntry*>(ff_smem + FF_TABD_BYTES + FF_TABF_BYTES);
    for (int i = threadIdx.x; i < FUSED_ND * FUSED_TP; i += FF_THREADS) tab_d[i] = __ldg(a.tab_d + i);
    for (int i = threadIdx.x; i < 16 * FUSED_FP; i += FF_THREADS) tab_f[i] = __ldg(a.tab_f + i);
    static_assert(sizeof(HEntry) == 12, "HEntry is copied as three words");
    for (int i = threadIdx.x; i < NTRI * 3; i += FF_THREADS)
        reinterpret_cast<int*>(tab)[i] = __ldg(reinterpret_cast<const int*>(a.tab_g) + i);
    __syncthreads();
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* mine = ff_smem + FF_TABD_BYTES + FF_TABF_BYTES + FF_HTAB_BYTES + (size_t)w * FF_WARP_BYTES;
    FusedScratch& Sw = *reinterpret_cast<FusedScratch*>(mine);
    uint32_t* Wb = reinterpret_cast<uint32_t*>(mine + sizeof(FusedScratch));
    double* th = reinterpret_cast<double*>(mine + sizeof(FusedScratch) + sizeof(uint32_t) * 16 * FF_WPITCH);
    WarpCtxT<true> ctx;   // the Normal information matrix has no mu x sigma block
    const size_t cp = a.st.Cp;
    for (;;) {
        int cell = 0;
        if (lane == 0) cell = atomicAdd(a.work, 1);
        cell = __shfl_sync(0xffffffffu, cell, 0);
        if (cell >= a.st.C) break;
        bool run = a.st.status[cell] == ATTRICI_STATUS_RUNNING;
        bool final_pass = false, evaluated = false;
        for (int it = 0;; ++it) {
            // pass 0 evaluates the point fit_init leaves (flat); passes fisher_passes .. + chord_passes - 1 reuse the
            // information matrix of the accepted point (see api.cu).  Cells that did not converge get one more pass
            // without moments for the log posterior of the returned point (model_scipy.py:525: trace["logp"] =
            // -res.fun); converged cells have it from the last evaluated point (cell_store, use_f_acc).
            const bool fresh = it < a.fisher_passes || it >= a.fisher_passes + a.chord_passes;
            int mode = it == 0 ? PASS_FLAT : (fresh ? PASS_FULL : PASS_GRAD);
            if (!run) {
                evaluated = it > 0 && a.st.status[cell] == ATTRICI_STATUS_CONVERGED;
                if (evaluated) break;
                cell_final_point(a.st, cell, Sw, ctx);
                mode = PASS_NLL;
                final_pass = true;
            }
            if (lane < NP) th[lane] = a.st.th_pass[(size_t)lane * cp + cell];
            __syncwarp();
            fused_pass<RELOAD, PREFETCH>(a, tab_d, tab_f, cell, mode, th, Wb, Sw.mom);
            if (final_pass) {   // cell_store takes the sum of the day terms from the momd row
                if (lane == 0) const_cast<double*>(a.st.momd)[(size_t)NACC * cp + cell] = Sw.mom[NACC];
                __syncwarp();
                break;
            }
            run = cell_step(a.st, cell, Sw, ctx, tab, fresh, true);
            __syncwarp();
        }
        cell_store(a.st, cell, Sw, ctx, evaluated);
        __syncwarp();
    }
}

// ---- tables and the cell-major copy of the per-phase sums ---------------------------------------------------

// pair tables from the phase rows pb64 ({n, cos 1..8, sin 1..8, sum g, sum g^2}) of the fit window
__global__ void fused_tables_kernel(const double* __restrict__ pb64, double* __restrict__ tab_d, float* __restrict__ tab_f) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= FUSED_FP) return;
    const bool valid = q < FUSED_NPAIR;
    const double* row = pb64 + (size_t)(valid ? q : 0) * B_STRIDE;
    const double* rm = pb64 + (size_t)((valid && q > 0) ? NPHASE - q : 0) * B_STRIDE;
    const bool has_m = valid && q > 0;
    if (q < FUSED_TP) {
        for (int k = 0; k < MAXM; ++k) {
            tab_d[(size_t)k * FUSED_TP + q] = valid ? row[1 + k] : 0.0;
            tab_d[(size_t)(MAXM + k) * FUSED_TP + q] = valid ? row[1 + NH + k] : 0.0;
        }
        tab_d[(size_t)8 * FUSED_TP + q] = valid ? row[0] : 0.0;
        tab_d[(size_t)9 * FUSED_TP + q] = valid ? row[1 + 2 * NH] : 0.0;
        tab_d[(size_t)10 * FUSED_TP + q] = valid ? row[2 + 2 * NH] : 0.0;
        tab_d[(size_t)11 * FUSED_TP + q] = has_m ? rm[0] : 0.0;
        tab_d[(size_t)12 * FUSED_TP + q] = has_m ? rm[1 + 2 * NH] : 0.0;
        tab_d[(size_t)13 * FUSED_TP + q] = has_m ? rm[2 + 2 * NH] : 0.0;
    }
    for (int k = 0; k < 2 * NH; ++k) {
        const float v = valid ? (float)row[1 + k] : 0.f;
        tab_f[(size_t)k * FUSED_FP + q] = __uint_as_float(ff_tf32(v));
    }
}

// [d][3][Cp] -> [cell][3][FUSED_SP] (tiles of 32 phases x 32 cells through shared memory); entries d >= n_ph are zero
__global__ void fused_transpose_kernel(const double* __restrict__ src, double* __restrict__ dst, int n_ph, int C, int Cp) {
    __shared__ double tile[32][33];
    const int k = blockIdx.z;
    const int d0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const int tx = threadIdx.x, ty = threadIdx.y;   // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int d = d0 + r, c = c0 + tx;
        tile[r][tx] = (d < n_ph && c < C) ? src[((size_t)d * 3 + k) * Cp + c] : 0.0;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int c = c0 + r, d = d0 + tx;
        if (c < C && d < FUSED_SP) dst[((size_t)c * 3 + k) * FUSED_SP + d] = tile[tx][r];
    }
}

__global__ void fused_unit_affine_kernel(double* aff, int C) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < C) { aff[2 * (size_t)c] = 0.0; aff[2 * (size_t)c + 1] = 1.0; }
}

// day counts of the cells with missing days from the n row of their (transposed) own sums; one warp per cell
__global__ void fused_counts_kernel(const double* __restrict__ gsumT, const int* __restrict__ miss, uint8_t* __restrict__ cntT, int C) {
    const int cell = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (cell >= C || miss[cell] == 0) return;
    for (int d = threadIdx.x & 31; d < FUSED_SP; d += 32)
        cntT[(size_t)cell * FUSED_SP + d] = d < NPHASE ? (uint8_t)gsumT[(size_t)cell * 3 * FUSED_SP + d] : 0;
}

}  // namespace

// pair tables of the fit window; after fold_phase_basis_kernel (prep.cu)
int launch_fused_tables(Workspace& ws, cudaStream_t s) {
    fused_tables_kernel<<<(FUSED_FP + 127) / 128, 128, 0, s>>>(ws.pb64, ws.ftab_d, ws.ftab_f);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

// per-phase sums of the scaled series in the phase-major layout of fold.cu -> cell-major, unit affine map
int launch_fused_from_phase_major(Workspace& ws, cudaStream_t s) {
    const int n_ph = ws.T < NPHASE ? ws.T : NPHASE;
    dim3 grid((FUSED_SP + 31) / 32, (ws.C + 31) / 32, 3), block(32, 8);
    fused_transpose_kernel<<<grid, block, 0, s>>>(ws.ystat, ws.ysumT, n_ph, ws.C, ws.Cp);
    ATTRICI_LAUNCH_CHECK();
    fused_transpose_kernel<<<grid, block, 0, s>>>(ws.gstat, ws.gsumT, n_ph, ws.C, ws.Cp);
    ATTRICI_LAUNCH_CHECK();
    fused_unit_affine_kernel<<<(ws.C + 127) / 128, 128, 0, s>>>(ws.aff, ws.C);
    ATTRICI_LAUNCH_CHECK();
    fused_counts_kernel<<<(ws.C + 3) / 4, 128, 0, s>>>(ws.gsumT, ws.miss, ws.cntT, ws.C);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

// whole iteration of the folded Normal term (after launch_fit_init); asynchronous on s
int launch_fit_fused(Workspace& ws, const StepParams& sp, int fisher_passes, int chord_passes, cudaStream_t s) {
    if (!ws.ysumT) { set_error("workspace was sized for another family"); return ATTRICI_EWORKSPACE; }
    FusedArgs a;
    a.st = make_step_args(ws, sp);
    a.tab_g = reinterpret_cast<const HEntry*>(ws.h_tab);
    a.ysumT = ws.ysumT; a.gsumT = ws.gsumT; a.aff = ws.aff; a.miss = ws.miss; a.cntT = ws.cntT;
    a.tab_d = ws.ftab_d; a.tab_f = ws.ftab_f;
    a.work = ws.counters + 8;
    a.fisher_passes = fisher_passes; a.chord_passes = chord_passes;
    ATTRICI_CUDA_CHECK(cudaMemsetAsync(a.work, 0, sizeof(int), s));
    static const int warps = [] {
        const char* e = getenv("ATTRICI_FUSED_WARPS");
        const int w = e ? atoi(e) : 12;
        return w == 8 ? 8 : 12;   // (variant 3 below runs 8 warps by itself)
    }();
    int ctas = 148;   // persistent: one CTA per SM, cells handed out through a.work
    const int need = (ws.C + warps - 1) / warps;
    if (ctas > need) ctas = need;
    auto launch = [&](auto kernel, int w) -> int {
        ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ff_smem_bytes(w)));
        kernel<<<ctas, 32 * w, ff_smem_bytes(w), s>>>(a);
        return 0;
    };
    // measured (10 000 cells): trig values re-read for the gradient moments 1.70 -> 1.52 ms (fewer spills at 168 registers)
    static const int variant = [] { const char* e = getenv("ATTRICI_FUSED_VARIANT"); return e ? atoi(e) : 0; }();
    // and the next sweep's sums requested by prefetch.global.L1 instead of held in registers 1.52 -> 1.39 ms; 8 warps
    // (no spills at all, a third fewer warps) 1.59 ms
    if (warps == 8) ATTRICI_TRY(launch(fit_fused_kernel<8, true, true>, 8));
    else if (variant == 1) ATTRICI_TRY(launch(fit_fused_kernel<12, false, false>, 12));
    else if (variant == 2) ATTRICI_TRY(launch(fit_fused_kernel<12, true, false>, 12));
    else ATTRICI_TRY(launch(fit_fused_kernel<12, true, true>, 12));
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
