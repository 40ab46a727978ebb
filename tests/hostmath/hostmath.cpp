// CPU build of attrici_b200/csrc/special.cuh + families.cuh for UNIT TESTS of the device math
// against scipy (tests/test_hostmath.py).  Test scaffolding only: nothing in the product loads
// this library, and the product has no CPU path.
#include <cstdint>
#include "../../attrici_b200/csrc/special.cuh"
#ifdef WITH_FAMILIES
#include "../../attrici_b200/csrc/families.cuh"
#endif
using namespace attrici;
extern "C" {
void hm_unary(int which, const double* x, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        switch (which) {
            case 0: out[i] = digamma<double>(x[i]); break;
            case 1: out[i] = trigamma<double>(x[i]); break;
            case 2: out[i] = ndtr(x[i]); break;
            case 3: out[i] = ndtri(x[i]); break;
            case 4: out[i] = (double)digamma<float>((float)x[i]); break;
            case 5: out[i] = (double)trigamma<float>((float)x[i]); break;
        }
    }
}
void hm_binary(int which, const double* a, const double* x, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        switch (which) {
            case 0: out[i] = gamma_p(a[i], x[i]); break;
            case 1: out[i] = gamma_q(a[i], x[i]); break;
            case 2: out[i] = gamma_p_inv(a[i], x[i]); break;
        }
    }
}
void hm_ternary(int which, const double* a, const double* b, const double* x, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        switch (which) {
            case 0: out[i] = beta_i(a[i], b[i], x[i]); break;
            case 1: out[i] = beta_i_inv(a[i], b[i], x[i]); break;
        }
    }
}
#ifdef WITH_FAMILIES
// kind, y, etaA, etaB -> out[6] = nll, rA, rB, wAA, wAB, wBB (Fisher weights)
void hm_family_f64(int kind, const double* y, const double* ea, const double* eb, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        DayTerms<double> d = family_day<double>(kind, y[i], ea[i], eb[i]);
        out[6 * i + 0] = d.nll; out[6 * i + 1] = d.rA; out[6 * i + 2] = d.rB;
        out[6 * i + 3] = d.wAA; out[6 * i + 4] = d.wAB; out[6 * i + 5] = d.wBB;
    }
}
void hm_family_f32(int kind, const double* y, const double* ea, const double* eb, double* out, int n) {
    for (int i = 0; i < n; ++i) {
        DayTerms<float> d = family_day<float>(kind, (float)y[i], (float)ea[i], (float)eb[i]);
        out[6 * i + 0] = d.nll; out[6 * i + 1] = d.rA; out[6 * i + 2] = d.rB;
        out[6 * i + 3] = d.wAA; out[6 * i + 4] = d.wAB; out[6 * i + 5] = d.wBB;
    }
}
#endif
}
