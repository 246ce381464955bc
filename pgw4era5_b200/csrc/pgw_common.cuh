// Shared device helpers for libpgw_b200 (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "pgw_b200.h"

// host-side error plumbing (pgw_misc.cu)
int pgw_check_launch(const char *what);
void pgw_set_error(const char *fmt, ...);

namespace pgw {

// constants.py:3-7 of the reference
constexpr double kRd = 287.05;
constexpr double kG = 9.80665;
constexpr double kMwMd = 0.622;

// Rd * Tv the way the reference forms it: tav = ta * (1 + 0.61 * hus) (functions.py:144) and
// CON_RD * tav (:151, :177) are numpy products in the dtype of ta and hus -- float32, each product
// rounded, for the float32 T and QV of an ERA5 file; float64 for the float64 PGW state.  Found by
// executing the reference (oracle/make_golden_glue.py); it shifts phi(p_ref) of the ERA state by up to
// ~4e-3 m2/s2, i.e. ps by ~7e-3 Pa, against an all-float64 evaluation.
__device__ __forceinline__ double rd_tv(float t, float q) {
    return (double)__fmul_rn(287.05f, __fmul_rn(t, __fadd_rn(1.0f, __fmul_rn(0.61f, q))));
}
__device__ __forceinline__ double rd_tv(double t, double q) { return kRd * (t * (1.0 + 0.61 * q)); }

// ---------------------------------------------------------------------------
// IFS saturation vapour pressure, functions.py:74-105.  Generic (exact formula
// in the working precision; used by the standalone conversions).
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T esat_generic(T ta) {
    const T T0 = T(273.16), Ti = T(250.16);
    const T ew = T(611.21) * exp(T(17.502) * (ta - T0) / (ta - T(32.19)));
    const T ei = T(611.21) * exp(T(22.587) * (ta - T0) / (ta - T(-0.7)));
    T alpha;
    if (ta >= T0) alpha = T(1);
    else if (ta <= Ti) alpha = T(0);
    else if (ta < T0 && ta > Ti) { const T r = (ta - Ti) / (T0 - Ti); alpha = r * r; }
    else alpha = T(NAN);
    return alpha * ew + (T(1) - alpha) * ei;
}

// ---------------------------------------------------------------------------
// Fast float32 variant for the fused pass.  The temperature is handed over as
// (tm273 = T_era - 273, exact in fp32 for 136.5 K < T < 546 K) plus a delta so
// that T_pgw - 273.16 is formed without first rounding T_pgw to fp32; only the
// branch that contributes is evaluated (alpha is exactly 0 or 1 outside the
// 250.16..273.16 K mixed-phase band, so skipping the other exp is exact).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float esat_fast(float tm273, float d) {
    const float dT = tm273 + (d - 0.16f);      // T - 273.16
    const float tk = tm273 + d;                 // T - 273
    if (dT >= 0.0f) {
        return 611.21f * __expf(__fdividef(17.502f * dT, tk + (273.0f - 32.19f)));
    } else if (dT <= -23.0f) {
        return 611.21f * __expf(__fdividef(22.587f * dT, tk + (273.0f + 0.7f)));
    } else {
        const float ew = 611.21f * __expf(__fdividef(17.502f * dT, tk + (273.0f - 32.19f)));
        const float ei = 611.21f * __expf(__fdividef(22.587f * dT, tk + (273.0f + 0.7f)));
        const float r = (dT + 23.0f) * (1.0f / 23.0f);
        const float alpha = r * r;               // NaN temperature lands here -> NaN
        return alpha * ew + (1.0f - alpha) * ei;
    }
}

// ---------------------------------------------------------------------------
// ln(pb/pt) in float64 for pb >= pt > 0 without a float64 log or divide in the
// common case: with s = (pb-pt)/(pb+pt), ln(pb/pt) = 2 atanh(s)
//   = 2 s (1 + s^2/3 + s^4/5 + ... ).  Adjacent ERA5 half levels below
// ~100 hPa have s < 0.05; six terms leave a relative truncation error below
// 2e-16 for s < 0.06.  The reciprocal is an fp32 MUFU seed plus one Newton
// step in fp64 (relative error ~4e-15).
// ---------------------------------------------------------------------------
__device__ __forceinline__ double log_ratio(double pb, double pt) {
    const double d = pb - pt;
    const double sm = pb + pt;
    double r = (double)__frcp_rn((float)sm);
    r = fma(r, fma(-sm, r, 1.0), r);
    const double s = d * r;
    if (s < 0.06) {
        const double s2 = s * s;
        double poly = fma(s2, 1.0 / 11.0, 1.0 / 9.0);
        poly = fma(s2, poly, 1.0 / 7.0);
        poly = fma(s2, poly, 1.0 / 5.0);
        poly = fma(s2, poly, 1.0 / 3.0);
        poly = fma(s2, poly, 1.0);
        return (s + s) * poly;
    }
    return log(pb / pt);
}

__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float warp_min(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Streaming loads/stores: every ERA5 value is touched exactly once, so keep it
// out of L1 and mark it evict-first in L2.
__device__ __forceinline__ float ld_stream(const float *p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float *p, float v) { __stcs(p, v); }

}  // namespace pgw
