// derive-huss: specific humidity from (counterfactual) relative humidity, air pressure and temperature.
// Replaces calc_huss_weedon2010 (attrici/commands/derive_huss.py:27-99: Buck 1981 saturation vapour pressure over
// water / ice with the enhancement factors of Weedon et al. 2010, clipped to [1e-7, 1e-1]); SURVEY.md 8f row 4.
// Elementwise FP64 like the reference (NumPy float64): 24 B in, 8 B out per value - an HBM-bound stream; two values
// per thread as 16-byte accesses when the four pointers allow it.
#include "common.cuh"

namespace attrici {

namespace {

__device__ __forceinline__ double huss_value(double hurs_in, double ps_in, double tas_in) {
    const double hurs = hurs_in / 100.0;    // relative humidity [1]
    const double ps = ps_in / 100.0;        // air pressure [mb]
    const double tas = tas_in - 273.15;     // temperature [degC]
    const bool water = tas >= 0.0;
    // Buck (1981) curves e_w4, e_i3 with f_w4, f_i4 (derive_huss.py:57-80); same operation order as NumPy
    const double a = water ? 6.1121 : 6.1115;
    const double b = water ? 18.729 : 23.036;
    const double c = water ? 257.87 : 279.82;
    const double d = water ? 227.3 : 333.7;
    const double x = water ? 7.2e-4 : 2.2e-4;
    const double y = water ? 3.20e-6 : 3.83e-6;
    const double z = water ? 5.9e-10 : 6.4e-10;
    const double arg = __ddiv_rn(__dmul_rn(__dsub_rn(b, __ddiv_rn(tas, d)), tas), __dadd_rn(tas, c));
    const double enh = __dadd_rn(__dadd_rn(1.0, x), __dmul_rn(ps, __dadd_rn(y, __dmul_rn(z, __dmul_rn(tas, tas)))));
    const double svp = __dmul_rn(__dmul_rn(a, exp(arg)), enh);
    const double rd_rv = 0.62198;
    const double den = __dsub_rn(__dadd_rn(__ddiv_rn(ps, svp), rd_rv), 1.0);
    double huss = __ddiv_rn(__dmul_rn(hurs, rd_rv), den);
    // np.minimum(np.maximum(huss, 1e-7), 1e-1): NaN propagates
    if (huss < 1e-7) huss = 1e-7;
    if (huss > 1e-1) huss = 1e-1;
    return huss;
}

template <int V>
__global__ void __launch_bounds__(256) huss_kernel(const double* __restrict__ hurs, const double* __restrict__ ps,
                                                   const double* __restrict__ tas, size_t n, double* __restrict__ out) {
    const size_t stride = (size_t)gridDim.x * blockDim.x * V;
    for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * V; i < n; i += stride) {
        if (V == 2 && i + 1 < n) {
            const double2 h = *reinterpret_cast<const double2*>(hurs + i);
            const double2 p = *reinterpret_cast<const double2*>(ps + i);
            const double2 t = *reinterpret_cast<const double2*>(tas + i);
            *reinterpret_cast<double2*>(out + i) = make_double2(huss_value(h.x, p.x, t.x), huss_value(h.y, p.y, t.y));
        } else {
            out[i] = huss_value(hurs[i], ps[i], tas[i]);
            if (V == 2 && i + 1 < n) out[i + 1] = huss_value(hurs[i + 1], ps[i + 1], tas[i + 1]);
        }
    }
}

}  // namespace
}  // namespace attrici

using namespace attrici;

extern "C" int attrici_derive_huss(const double* hurs, const double* ps, const double* tas, size_t n, double* huss,
                                   void* stream) {
    if (!hurs || !ps || !tas || !huss) { set_error("hurs / ps / tas / huss is null"); return ATTRICI_EINVAL; }
    if (n == 0) return 0;
    const bool vec = (((uintptr_t)hurs | (uintptr_t)ps | (uintptr_t)tas | (uintptr_t)huss) & 15) == 0;
    const size_t per_block = 256 * 2;
    size_t blocks = (n + per_block - 1) / per_block;
    if (blocks > 148 * 32) blocks = 148 * 32;
    if (vec) huss_kernel<2><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(hurs, ps, tas, n, huss);
    else huss_kernel<1><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(hurs, ps, tas, n, huss);
    ATTRICI_LAUNCH_CHECK();
    return 0;
}
