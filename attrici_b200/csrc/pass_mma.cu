// Likelihood pass of the Normal family on the tensor cores (TF32 mma.sync m16n8k8, 3-term split products
// that carry FP32-level accuracy), replacing the FMA chains of pass.cu's thread-per-cell kernel.
//
// Both contractions of a pass are GEMM-shaped because the basis is shared by all cells:
//   stage 1   eta[cells x days]  = Theta[cells x K] . Z[days x K]^T          (linear predictors of mu, log sigma)
//   (elementwise, registers)      e = exp(-eta_s), z = (y - eta_mu) e, nll += z^2/2 + eta_s,
//                                 r_mu = z e, r_s = 1 - z^2, w = e^2, v = valid
//   stage 3   M[cells x cols]   += Q[cells x days] . Z[days x cols]          Q in {r_mu, r_s, w, v}
// Z holds the 51 products gmt^p x {1, cos k, sin k} (k <= 8); its first 16 columns (cos / sin 1..4 and their
// products with gmt) are the design columns stage 1 runs on the tensor cores, the intercept and gmt columns
// (16, 17) are added in FP32 (products of Fourier columns are Fourier columns: see pass.cu).  A warp owns 16 cells;
// the accumulator fragment of stage 1 is reused as the A fragment of stage 3 by pairing MMA k-index t with
// day 2t and t+4 with day 2t+1, so no shuffles are needed (the FlashAttention-2 register trick).
//
// Accuracy: every operand is split x = hi + lo with hi = tf32_rn(x), lo = tf32_rn(x - hi) and the product
// is hi.hi + lo.hi + hi.lo (exact TF32 products, FP32 accumulation): relative error ~2^-22 per term, the
// level of an FP32 FMA chain of this length.  Fisher moments (w, v) only steer the step: single TF32.
//
// Replaces the same reference code as pass.cu (DistributionScipy.log_likelihood,
// attrici/estimation/model_scipy.py:311-331 with np.dot(covariates, weights) + scipy.stats.norm.logpdf).
#include "common.cuh"
#include "pass_impl.cuh"

namespace attrici {

namespace {

constexpr int MMA_WARPS = 4;                    // warps per CTA (3 CTAs per SM)
constexpr int MMA_THREADS = 32 * MMA_WARPS;
constexpr int MMA_CELLS = 16 * MMA_WARPS;       // cells per CTA
constexpr int ZC_HI = 56;                       // 51 hi columns padded to 7 n-tiles
constexpr int ZC_LO = 24;                       // lo parts of the first 18 columns, 3 n-tiles

struct MmaArgs {
    const float* ys;
    int64_t ld;
    int Cp;
    const int* list;
    int n_active, Pp;
    int i0, i1, chunk_len;
    const float* zmat;       // [T + pad][ZM_STRIDE]
    const double* th_pass;   // [NP][Cp]
    const int* status;
    double* part_nll;        // [chunk][Pp]
    float* part_acc;         // [chunk][NACC][Pp]
};

__device__ __forceinline__ uint32_t tf32_rn(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return r;
}

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = tf32_rn(x);
    lo = tf32_rn(x - __uint_as_float(hi));
}

__device__ __forceinline__ void mma_tf32(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gmem));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// Column order of Z (power of gmt p, trig17 slot t = {0: 1, 1..8: cos k, 9..16: sin k}):
//   0..7   p0 x {cos 1..4, sin 1..4}      8..15  p1 x same          16, 17  p0 x 1, p1 x 1
//   18..25 p0 x {cos 5..8, sin 5..8}      26..33 p1 x same          34..50  p2 x trig17
__host__ __device__ __forceinline__ void zcol_decode(int c, int& p, int& t) {
    if (c < 16) { p = c >> 3; int j = c & 7; t = j < 4 ? 1 + j : NH + 1 + (j - 4); }
    else if (c < 18) { p = c - 16; t = 0; }
    else if (c < 34) { p = (c - 18) >> 3; int j = (c - 18) & 7; t = j < 4 ? 5 + j : NH + 5 + (j - 4); }
    else { p = 2; t = c - 34; }
}

// Z column -> index in the accumulator layout of pass.cu (ACC_AA block) ; -1 = padding
__device__ __forceinline__ int zcol_to_aa(int c) {
    if (c >= ZM_COLS) return -1;
    int p, t;
    zcol_decode(c, p, t);
    return ACC_AA + p * NTRIG + t;
}

// the 17 power-0 columns (moments of the sigma block): 0..7, 16, 18..25
__device__ __forceinline__ int bb_zcol(int j) { return j < 8 ? j : (j == 8 ? 16 : j + 9); }

constexpr int BB_COLS = 17;                     // Z columns of power 0 (0..8 and 18..25): moments of the sigma block
constexpr int MT = 32;                          // days per shared-memory tile
constexpr int NSTAGE = 3;                       // tiles in flight (cp.async ring)
constexpr int Y_STRIDE = MMA_CELLS + 4;         // floats per day of the observation tile; 2 * 68 = 8 (mod 32)
constexpr int STAGE_FLOATS = MT * ZM_STRIDE + MT * Y_STRIDE;

// a missing day inside the chunk: take its Z row out of the cell's sigma-block moments (rare)
__device__ __noinline__ void bb_correct(float* corr, const float* zr) {
    for (int j = 0; j < BB_COLS; ++j) atomicAdd(&corr[j], -zr[bb_zcol(j)]);
}

__global__ void __launch_bounds__(MMA_THREADS, 3) pass_mma_kernel(MmaArgs a) {
    extern __shared__ __align__(16) float smem[];   // [NSTAGE]{ Z tile [MT][ZM_STRIDE], y tile [MT][Y_STRIDE] }
    // sigma block: w_BB = 2 on every valid day, so its moments are 2 (column sums of Z over the chunk
    // - rows of Z on the cell's missing days): one shared column sum per CTA + a rare per-cell correction
    __shared__ float bb_sum[32];
    __shared__ float bb_corr[MMA_WARPS][16][BB_COLS];
    __shared__ int row_cell[MMA_CELLS];             // series column of each row of the CTA, -1 = not evaluated
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gid = lane >> 2, t4 = lane & 3;
    const int chunk = blockIdx.y;
    const int t_begin = a.i0 + chunk * a.chunk_len;
    const int t_end = min(a.i1, t_begin + a.chunk_len);

    // the two cells (rows gid, gid + 8 of the warp's 16) this thread holds fragments of
    int pos[2], cell[2];
    bool act[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        pos[h] = blockIdx.x * MMA_CELLS + warp * 16 + gid + 8 * h;
        const bool in_range = pos[h] < a.n_active;
        cell[h] = a.list ? a.list[in_range ? pos[h] : 0] : (in_range ? pos[h] : a.n_active - 1);
        act[h] = in_range;
        if (act[h] && a.status) act[h] = (a.status[cell[h]] == ATTRICI_STATUS_RUNNING);
        if (t4 == 0) row_cell[warp * 16 + gid + 8 * h] = act[h] ? cell[h] : -1;
    }
    if (__syncthreads_or(act[0] || act[1]) == 0 || t_begin >= t_end) return;   // uniform per CTA
    const bool warp_active = __any_sync(0xffffffffu, act[0] || act[1]);

    // coefficient fragments (A operands of stage 1): Z column k < 16 <-> canonical coefficient k + 2 of the
    // mu block; k < 8 <-> NA + 1 + k of the sigma block.  Intercepts and the gmt slope stay in FP32.
    uint32_t thA_h[2][4], thA_l[2][4], thB_h[4], thB_l[4];
    float thA0[2], thA1[2], thB0[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        thA0[h] = act[h] ? (float)a.th_pass[(size_t)0 * a.Cp + cell[h]] : 0.f;
        thA1[h] = act[h] ? (float)a.th_pass[(size_t)1 * a.Cp + cell[h]] : 0.f;
        thB0[h] = act[h] ? (float)a.th_pass[(size_t)NA * a.Cp + cell[h]] : 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int h = r & 1;                       // a0,a2: row gid ; a1,a3: row gid + 8
            const int k = 8 * ks + t4 + 4 * (r >> 1);
            float va = 0.f, vb = 0.f;
            if (act[h]) {
                va = (float)a.th_pass[(size_t)(k + 2) * a.Cp + cell[h]];
                if (ks == 0) vb = (float)a.th_pass[(size_t)(NA + 1 + k) * a.Cp + cell[h]];
            }
            split_tf32(va, thA_h[ks][r], thA_l[ks][r]);
            if (ks == 0) split_tf32(vb, thB_h[r], thB_l[r]);
        }
    }

    float accAA[7][4], accGA[3][4], accGB[2][4];
#pragma unroll
    for (int i = 0; i < 7; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) accAA[i][r] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) accGA[i][r] = 0.f;
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int r = 0; r < 4; ++r) accGB[i][r] = 0.f;
    double nll[2] = {0.0, 0.0};
    for (int i = threadIdx.x; i < MMA_WARPS * 16 * BB_COLS; i += MMA_THREADS) (&bb_corr[0][0][0])[i] = 0.f;
    // Z column summed by this thread (threads 0..16)
    const int bb_col = bb_zcol(threadIdx.x < BB_COLS ? threadIdx.x : 0);
    float bb_acc = 0.f;

    const int n_tiles = (t_end - t_begin + MT - 1) / MT;
    // tile `it` -> ring slot it % NSTAGE: the Z rows (16-byte copies) and the observations of the CTA's 64
    // cells (4-byte copies: columns are arbitrary when the cells come from the active list).  Days beyond the
    // chunk and rows that are not evaluated are not copied; the consumer masks them.
    auto issue_tile = [&](int it) {
        if (it < n_tiles) {
            float* zdst = smem + (size_t)(it % NSTAGE) * STAGE_FLOATS;
            float* ydst = zdst + MT * ZM_STRIDE;
            const int t_tile = t_begin + it * MT;
            const float4* src = reinterpret_cast<const float4*>(a.zmat + (size_t)t_tile * ZM_STRIDE);
            for (int v = threadIdx.x; v < MT * ZM_STRIDE / 4; v += MMA_THREADS)
                cp_async16(reinterpret_cast<float4*>(zdst) + v, src + v);
            const int row = threadIdx.x % MMA_CELLS;
            const int col = row_cell[row];
            if (col >= 0) {
                for (int d = threadIdx.x / MMA_CELLS; d < MT; d += MMA_THREADS / MMA_CELLS) {
                    const int t = t_tile + d;
                    if (t < t_end) cp_async4(ydst + d * Y_STRIDE + row, a.ys + (size_t)t * a.ld + col);
                }
            }
        }
        cp_async_commit();   // one group per call (possibly empty) keeps the wait counts uniform
    };
    issue_tile(0);
    issue_tile(1);

    for (int it = 0; it < n_tiles; ++it) {
        issue_tile(it + 2);
        cp_async_wait<2>();
        __syncthreads();
        const float* tile = smem + (size_t)(it % NSTAGE) * STAGE_FLOATS;
        const float* ytile = tile + MT * ZM_STRIDE;
        const int t_tile = t_begin + it * MT;
        if (threadIdx.x < BB_COLS) {
            const float* tl = tile + bb_col;
            const int nd = min(MT, t_end - t_tile);
            float s0 = 0.f, s1 = 0.f;
            for (int d = 0; d + 1 < nd; d += 2) { s0 += tl[(size_t)d * ZM_STRIDE]; s1 += tl[(size_t)(d + 1) * ZM_STRIDE]; }
            if (nd & 1) s0 += tl[(size_t)(nd - 1) * ZM_STRIDE];
            bb_acc += s0 + s1;
        }
        if (warp_active) {
#pragma unroll 1
            for (int s8 = 0; s8 < MT / 8; ++s8) {
                const int t0 = t_tile + 8 * s8;
                if (t0 >= t_end) break;
                const float* zt = tile + (size_t)(8 * s8) * ZM_STRIDE;
                // observations y[row h][day 2 t4 + d] of this step, index h * 2 + d
                float yv4[4];
                bool ok4[4];
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int d = 0; d < 2; ++d) {
                        const int dd = 8 * s8 + 2 * t4 + d;
                        const bool in = act[h] && (t_tile + dd < t_end);
                        const float v = in ? ytile[dd * Y_STRIDE + warp * 16 + gid + 8 * h] : 0.f;
                        ok4[h * 2 + d] = in && (v == v);
                        yv4[h * 2 + d] = ok4[h * 2 + d] ? v : 0.f;
                        if (in && !(v == v)) bb_correct(bb_corr[warp][gid + 8 * h], zt + (size_t)(2 * t4 + d) * ZM_STRIDE);
                    }
                // ---- stage 1: eta_mu, eta_sigma for 16 cells x 8 days (B operand: k = Z column, n = day)
                // three independent accumulation chains per predictor (hi.hi, lo.hi, hi.lo) keep the
                // dependent-MMA depth at 3 instead of 9
                float etaA[4] = {0.f, 0.f, 0.f, 0.f}, etaB[4] = {0.f, 0.f, 0.f, 0.f};
                float eA1[4] = {0.f, 0.f, 0.f, 0.f}, eA2[4] = {0.f, 0.f, 0.f, 0.f};
                float eB1[4] = {0.f, 0.f, 0.f, 0.f}, eB2[4] = {0.f, 0.f, 0.f, 0.f};
                const float* zrow = zt + (size_t)gid * ZM_STRIDE;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
                    const uint32_t bh0 = __float_as_uint(zrow[8 * ks + t4]), bh1 = __float_as_uint(zrow[8 * ks + t4 + 4]);
                    const uint32_t bl0 = __float_as_uint(zrow[ZC_HI + 8 * ks + t4]),
                                   bl1 = __float_as_uint(zrow[ZC_HI + 8 * ks + t4 + 4]);
                    mma_tf32(etaA, thA_h[ks], bh0, bh1);
                    mma_tf32(eA1, thA_l[ks], bh0, bh1);
                    mma_tf32(eA2, thA_h[ks], bl0, bl1);
                    if (ks == 0) {
                        mma_tf32(etaB, thB_h, bh0, bh1);
                        mma_tf32(eB1, thB_l, bh0, bh1);
                        mma_tf32(eB2, thB_h, bl0, bl1);
                    }
                }
                // intercepts and the gmt slope in FP32: gmt of the thread's two days = hi + lo of Z column 17
                float gday[2];
#pragma unroll
                for (int d = 0; d < 2; ++d) {
                    const float* zr = zt + (size_t)(2 * t4 + d) * ZM_STRIDE;
                    gday[d] = zr[17] + zr[ZC_HI + 17];
                }
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int h = c >> 1, d = c & 1;
                    etaA[c] = (etaA[c] + (eA1[c] + eA2[c])) + fmaf(thA1[h], gday[d], thA0[h]);
                    etaB[c] = (etaB[c] + (eB1[c] + eB2[c])) + thB0[h];
                }
                // ---- elementwise: accumulator fragment c{0,1,2,3} = (row gid | gid + 8) x (day 2 t4 | 2 t4 + 1)
                // -> A fragment a{0,1,2,3} = (row gid, k t4), (row gid + 8, k t4), (row gid, k t4 + 4), (gid + 8, t4 + 4)
                //    with k = t4 <-> day 2 t4 and k = t4 + 4 <-> day 2 t4 + 1
                uint32_t qAh[4], qAl[4], qBh[4], qBl[4], qW[4];
                float nl[2] = {0.f, 0.f};
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int h = c >> 1, d = c & 1;        // c0,c1: row gid ; c2,c3: row gid + 8
                    const bool valid = ok4[h * 2 + d];
                    const float e = expf(-etaB[c]);   // full precision: __expf's error grows with |x| and is not zero-mean
                    const float z = (yv4[h * 2 + d] - etaA[c]) * e;
                    const float z2 = z * z;
                    const float rA = valid ? -z * e : 0.f;
                    const float rB = valid ? 1.f - z2 : 0.f;
                    const float w = valid ? e * e : 0.f;
                    nl[h] += valid ? fmaf(0.5f, z2, etaB[c]) + 0.91893853320467274178f : 0.f;
                    const int ai = h + 2 * d;               // position in the A fragment
                    split_tf32(rA, qAh[ai], qAl[ai]);
                    split_tf32(rB, qBh[ai], qBl[ai]);
                    qW[ai] = tf32_rn(w);
                }
                nll[0] += (double)nl[0];
                nll[1] += (double)nl[1];
                // ---- stage 3: moments (B operand: k = day, n = Z column).  Issue order: all hi.hi products
                // first, then lo.hi, then hi.lo, so that the three updates of one accumulator are far apart.
                const float* zd0 = zt + (size_t)(2 * t4) * ZM_STRIDE + gid;
                const float* zd1 = zd0 + ZM_STRIDE;
                uint32_t zh0[7], zh1[7];
#pragma unroll
                for (int nt = 0; nt < 7; ++nt) {
                    zh0[nt] = __float_as_uint(zd0[8 * nt]);
                    zh1[nt] = __float_as_uint(zd1[8 * nt]);
                }
#pragma unroll
                for (int nt = 0; nt < 7; ++nt) {
                    mma_tf32(accAA[nt], qW, zh0[nt], zh1[nt]);
                    if (nt < 3) mma_tf32(accGA[nt], qAh, zh0[nt], zh1[nt]);
                    // sigma block: columns 0..7 (n-tile 0) and the intercept column 16 (n-tile 2)
                    if (nt == 0 || nt == 2) mma_tf32(accGB[nt >> 1], qBh, zh0[nt], zh1[nt]);
                }
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    mma_tf32(accGA[nt], qAl, zh0[nt], zh1[nt]);
                    if (nt != 1) mma_tf32(accGB[nt >> 1], qBl, zh0[nt], zh1[nt]);
                }
#pragma unroll
                for (int nt = 0; nt < 3; ++nt) {
                    const uint32_t bl0 = __float_as_uint(zd0[ZC_HI + 8 * nt]), bl1 = __float_as_uint(zd1[ZC_HI + 8 * nt]);
                    mma_tf32(accGA[nt], qAh, bl0, bl1);
                    if (nt != 1) mma_tf32(accGB[nt >> 1], qBh, bl0, bl1);
                }
            }
        }
        __syncthreads();
    }

    // ---- write the chunk's partial sums in the accumulator layout of pass.cu (indexed by position)
    // accumulator fragment: c0,c1 -> row gid, columns 2 t4, 2 t4 + 1 ; c2,c3 -> row gid + 8
    float* out = a.part_acc + (size_t)chunk * NACC * a.Pp;
    if (threadIdx.x < BB_COLS) bb_sum[threadIdx.x] = bb_acc;
    __syncthreads();
    // sigma-block moments: rows of the CTA x 17 columns, written by all threads
    for (int i = threadIdx.x; i < MMA_CELLS * BB_COLS; i += MMA_THREADS) {
        const int row = i % MMA_CELLS, j = i / MMA_CELLS;
        const int p = blockIdx.x * MMA_CELLS + row;
        if (p < a.n_active) {
            const int ia = zcol_to_aa(bb_zcol(j));   // power 0: index = trig17 slot
            out[(size_t)(ACC_BB + (ia - ACC_AA)) * a.Pp + p] = 2.f * (bb_sum[j] + bb_corr[row >> 4][row & 15][j]);
        }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        // per-cell nll: sum over the 4 threads sharing the row
        double s = nll[h];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const bool in_range = pos[h] < a.n_active;
        if (!in_range) continue;
        const size_t p = pos[h];
        if (t4 == 0) a.part_nll[(size_t)chunk * a.Pp + p] = s;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
            const int r = 2 * h + d;
#pragma unroll
            for (int nt = 0; nt < 7; ++nt) {
                const int c = 8 * nt + 2 * t4 + d;
                const int ia = zcol_to_aa(c);
                if (ia >= 0) out[(size_t)ia * a.Pp + p] = accAA[nt][r];
                if (nt < 3 && c < 18) {
                    // gradient slots [power][{1, cos 1..4, sin 1..4}]: Z columns 0..7 / 8..15 are the Fourier
                    // columns of power 0 / 1, columns 16 / 17 the intercept and gmt columns
                    const int pw = c < 16 ? (c >> 3) : c - 16;
                    const int slot = c < 16 ? 1 + (c & 7) : 0;
                    out[(size_t)(ACC_GA + pw * NB + slot) * a.Pp + p] = accGA[nt][r];
                    if (pw == 0) out[(size_t)(ACC_GB + slot) * a.Pp + p] = accGB[nt >> 1][r];
                }
            }
        }
    }
    // the cross moments of the two blocks vanish in expectation for the Normal family
    for (int i = threadIdx.x; i < 2 * NTRIG * MMA_CELLS; i += MMA_THREADS) {
        const int row = i % MMA_CELLS, k = i / MMA_CELLS;
        const int p = blockIdx.x * MMA_CELLS + row;
        if (p < a.n_active) out[(size_t)(ACC_AB + k) * a.Pp + p] = 0.f;
    }
}

// hi / lo TF32 parts of the moment matrix, from the FP64 basis: row t =
//   [ hi of the 51 columns | pad to 56 | lo of the first 18 columns | pad to 24 | pad to ZM_STRIDE ]
// column order: see zcol_decode
__global__ void build_zmat_kernel(const double* __restrict__ b64, int T_pad, float* __restrict__ zmat) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T_pad) return;
    const double* row = b64 + (size_t)t * B_STRIDE;
    const double g = row[0];
    float* z = zmat + (size_t)t * ZM_STRIDE;
    for (int c = 0; c < ZM_STRIDE; ++c) z[c] = 0.f;
    // rows beyond the series are all-zero in basis64 (cos columns included): keep them zero here
    bool pad_row = true;
    for (int i = 1; i <= NH; ++i) pad_row = pad_row && (row[i] == 0.0) && (row[NH + i] == 0.0);
    if (pad_row) return;
    for (int c = 0; c < ZM_COLS; ++c) {
        int p, tr;
        zcol_decode(c, p, tr);
        double v = (tr == 0) ? 1.0 : row[tr];
        for (int q = 0; q < p; ++q) v *= g;
        const float vf = (float)v;
        uint32_t hi, lo;
        split_tf32(vf, hi, lo);
        z[c] = __uint_as_float(hi);
        if (c < 18) z[ZC_HI + c] = __uint_as_float(lo);
    }
}

}  // namespace

int launch_build_zmat(Workspace& ws, cudaStream_t s) {
    int T_pad = ws.T + 2 * PASS_TILE;
    build_zmat_kernel<<<(T_pad + 127) / 128, 128, 0, s>>>(ws.basis64, T_pad, ws.zmat);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

int launch_pass_mma(Workspace& ws, const PassShape& ps, const float* y_scaled, int64_t ld, int i0, int i1,
                    const int* status, cudaStream_t s) {
    MmaArgs a;
    a.ys = y_scaled;
    a.ld = ld;
    a.Cp = ws.Cp;
    a.list = ps.list;
    a.n_active = ps.n_active;
    a.Pp = ps.Pp;
    a.i0 = i0;
    a.i1 = i1;
    a.chunk_len = ps.chunk_len;
    a.zmat = ws.zmat;
    a.th_pass = ws.th_pass;
    a.status = status;
    a.part_nll = ws.part_nll;
    a.part_acc = (float*)ws.part_acc;
    const size_t smem = (size_t)NSTAGE * STAGE_FLOATS * sizeof(float);
    ATTRICI_CUDA_CHECK(cudaFuncSetAttribute(pass_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((ps.n_active + MMA_CELLS - 1) / MMA_CELLS, ps.n_chunks);
    pass_mma_kernel<<<grid, MMA_THREADS, smem, s>>>(a);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}

}  // namespace attrici
