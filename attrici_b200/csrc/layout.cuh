// Compile-time layout constants and the coefficient-index algebra shared by every kernel.
// No CUDA runtime dependency: tests/hostmath compiles this header with g++ to unit-test the
// per-cell mathematics on the CPU (test scaffolding only; the product has no CPU path).
#pragma once
#include <stdint.h>
#include <stddef.h>
#include "../../include/attrici_b200.h"
#include "special.cuh"

namespace attrici {

constexpr int MAXM = ATTRICI_MAX_MODES;      // 4 Fourier modes
constexpr int NA = 2 + 4 * MAXM;             // 18 coefficients of a GMT-dependent parameter
constexpr int NB = 1 + 2 * MAXM;             // 9 coefficients of an independent parameter
constexpr int NP = NA + NB;                  // 27 coefficients of one likelihood term (canonical layout)
constexpr int NTRI = NP * (NP + 1) / 2;      // 378 packed lower-triangular entries
constexpr int NH = 2 * MAXM;                 // 8 harmonics carried for the moments
constexpr int NTRIG = 1 + 2 * NH;            // 17 trig columns {1, cos 1..8, sin 1..8}
constexpr int B_STRIDE = 20;                 // values per day in basis32 / basis64
constexpr int B_TAU = 17;                    // column holding tau = (day - day0) / 365.25

// accumulator layout of one likelihood pass (per cell)
constexpr int ACC_AA = 0;                    // 3 powers of gmt x 17 trig
constexpr int ACC_AB = ACC_AA + 3 * NTRIG;   // 2 x 17
constexpr int ACC_BB = ACC_AB + 2 * NTRIG;   // 17
constexpr int ACC_GA = ACC_BB + NTRIG;       // 2 powers x {1, cos 1..4, sin 1..4}
constexpr int ACC_GB = ACC_GA + 2 * NB;      // 9
constexpr int NACC = ACC_GB + NB;            // 129

constexpr int PASS_WARPS = 4;                // warps per CTA of the pass kernel
constexpr int PASS_CELLS = 32 * PASS_WARPS;  // cells per CTA
constexpr int PASS_TILE = 64;                // days per shared-memory basis tile
constexpr int MAX_CHUNKS = 32;
constexpr int ZM_COLS = 51;                  // moment-matrix columns gmt^p x trig17 (tensor-core pass)
constexpr int NPHASE = 1461;                // days after which every Fourier column repeats (4 x 365.25)
constexpr int FUSED_NPAIR = NPHASE / 2 + 1;  // 731 items of the fused fit: phase 0 alone, then the pairs (d, 1461 - d)
constexpr int FUSED_TP = 736;                // pairs padded to 23 sweeps of 32 lanes (pitch of the FP64 pair tables)
constexpr int FUSED_FP = 740;                // pitch of the TF32 pair table (= 4 mod 32: conflict-free B fragments)
constexpr int FUSED_SP = 1464;               // phases per row of the cell-major per-phase sums (1461 padded)
constexpr int FUSED_ND = 14;                 // rows of the FP64 pair table: cos 1..4, sin 1..4, {n, sum g, sum g^2} of d and of 1461 - d
constexpr int ZM_STRIDE = 100;               // floats per day of the moment matrix: 56 hi + 24 lo + pad; = 4 (mod 32)

// likelihood term kinds (a Bernoulli-Gamma fit is a Bernoulli term plus a Gamma term that share
// no coefficient: model_scipy.py:433-446)
enum Term : int { T_NORMAL = 0, T_GAMMA = 1, T_BETA = 2, T_WEIBULL = 3, T_BERNOULLI = 4 };

ATTRICI_HD bool term_has_b(int term) { return term != T_BERNOULLI; }
ATTRICI_HD bool term_has_ab(int term) { return term == T_WEIBULL || term == T_BETA; }

// number of likelihood terms of a family and the term fitted in each slot
ATTRICI_HD int family_terms_count(int family) { return family == ATTRICI_BERNOULLI_GAMMA ? 2 : 1; }
ATTRICI_HD int family_term(int family, int slot) {
    switch (family) {
        case ATTRICI_NORMAL: return T_NORMAL;
        case ATTRICI_GAMMA: return T_GAMMA;
        case ATTRICI_BETA: return T_BETA;
        case ATTRICI_WEIBULL: return T_WEIBULL;
        default: return slot == 0 ? T_BERNOULLI : T_GAMMA;
    }
}

ATTRICI_HD int family_num_params(int family, int modes) {
    int a = 2 + 4 * modes, b = 1 + 2 * modes;
    return family == ATTRICI_BERNOULLI_GAMMA ? 2 * a + b : a + b;
}

// canonical coefficient index -> (block, power of gmt, trig type 0 = cos / 1 = sin, harmonic)
struct Col { int blk, pw, sn, k; };
ATTRICI_HD Col col_of(int i) {
    Col c;
    if (i < NA) {
        c.blk = 0;
        if (i < 2) { c.pw = i; c.sn = 0; c.k = 0; }
        else {
            int j = i - 2;                  // 0..15: cos, sin, g*cos, g*sin blocks of MAXM
            c.pw = j / (2 * MAXM);
            c.sn = (j / MAXM) & 1;
            c.k = j % MAXM + 1;
        }
    } else {
        int j = i - NA;
        c.blk = 1; c.pw = 0;
        if (j == 0) { c.sn = 0; c.k = 0; }
        else { c.sn = (j - 1) / MAXM; c.k = (j - 1) % MAXM + 1; }
    }
    return c;
}

ATTRICI_HD bool col_active(const Col& c, int term, int modes) {
    if (c.blk == 1 && !term_has_b(term)) return false;
    return c.k <= modes;
}

// Reference theta layout (model_scipy.py:104-113, 225-234, 410-421) ->
// (term slot, canonical index).  dependent block: [w0, w1, cos 1..m, sin 1..m, T cos 1..m, T sin 1..m];
// independent block: [w0, cos 1..m, sin 1..m].
//   normal / gamma / beta : [dep | indep]      weibull : [indep (alpha) | dep (beta)]
//   bernoulli-gamma       : [p dep (slot 0) | mu dep (slot 1) | nu indep (slot 1)]
ATTRICI_HD void ref_to_canon(int family, int modes, int j, int* slot, int* canon) {
    const int a = 2 + 4 * modes, b = 1 + 2 * modes;
    int is_dep, local;
    *slot = 0;
    if (family == ATTRICI_BERNOULLI_GAMMA) {
        if (j < a) { is_dep = 1; local = j; }
        else if (j < 2 * a) { is_dep = 1; local = j - a; *slot = 1; }
        else { is_dep = 0; local = j - 2 * a; *slot = 1; }
    } else if (family == ATTRICI_WEIBULL) {
        if (j < b) { is_dep = 0; local = j; }
        else { is_dep = 1; local = j - b; }
    } else {
        if (j < a) { is_dep = 1; local = j; }
        else { is_dep = 0; local = j - a; }
    }
    if (is_dep) {
        if (local < 2) *canon = local;
        else { int q = local - 2; *canon = 2 + (q / modes) * MAXM + q % modes; }
    } else {
        if (local == 0) *canon = NA;
        else { int q = local - 1; *canon = NA + 1 + (q / modes) * MAXM + q % modes; }
    }
}

ATTRICI_HD int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

}  // namespace attrici
