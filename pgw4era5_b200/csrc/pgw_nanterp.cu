// NaN-ignoring Gaussian-kernel regridding of the ocean deltas (tos, siconc) for libpgw_b200 (sm_100a):
// nan_ignoring_interp, functions.py:900-1060.
//
// The reference maps (lat, lon) of the curvilinear GCM ocean grid and of the ERA5 grid to signed
// "meter" coordinates with three WGS84 geodesic distances per point (pyproj Geod.inv, :958-973,
// :1011-1022): the meridian arc from the equator, the geodesic to the point of EQUAL latitude on the
// Greenwich meridian, and the distance to the point 180 degrees away (used to replicate the point
// cloud east and west, :978-988).  It then lets pyvista/VTK interpolate with a Gaussian kernel over all
// points within a radius (:1038-1048).  Neither package is vendored with the reference; their
// published algorithms are implemented here (PARITY UNPINNED, see oracle/pgw_oracle.py):
//
//  * geodesics from the exact integrals of Karney (2013), eqs 7-8, by 16-point Gauss-Legendre
//    quadrature; the equal-latitude geodesic is symmetric about its vertex, so its azimuth at the
//    equator is found by bisection on the longitude integral;
//  * vtkGaussianKernel: w = exp(-(sharpness/radius)^2 d^2) over the points with d <= radius,
//    normalised; an exact hit (d^2 < 256 eps) takes that point's value; no point -> NaN.
//
// The sources are sorted by their meridional coordinate on the host side (torch.sort on the device),
// so each target scans only the band |dlat_m| <= radius; all fields (months) of a call share the
// weights, a NaN source value is "absent" for that field only (the reference drops NaNs per month).
#include "pgw_common.cuh"

#define PGW_REQUIRE(cond) do { if (!(cond)) return PGW_E_INVALID; } while (0)

namespace pgw {

constexpr double kWgsA = 6378137.0;
constexpr double kWgsF = 1.0 / 298.257223563;
constexpr double kPi = 3.14159265358979323846;

// 16-point Gauss-Legendre, positive half
__constant__ double c_glx[8] = {0.095012509837637441, 0.28160355077925892, 0.45801677765722737, 0.61787624440264377,
                                0.755404408355003,    0.86563120238783176, 0.9445750230732326,  0.98940093499164994};
__constant__ double c_glw[8] = {0.18945061045506864, 0.18260341504492364, 0.16915651939500265,  0.14959598881657671,
                                0.12462897125553407, 0.095158511682492605, 0.062253523938647456, 0.027152459411754176};

// integral over [lo, hi] of sqrt(1 + k2 sin^2 s)               (DIST, Karney eq. 7)
//                     or of (2 - f) / (1 + (1 - f) sqrt(...))   (!DIST, eq. 8)
template <bool DIST>
__device__ double geod_quad(double lo, double hi, double k2) {
    const double mid = 0.5 * (hi + lo), half = 0.5 * (hi - lo);
    double sum = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int sgn = -1; sgn <= 1; sgn += 2) {
            const double sn = sin(mid + sgn * half * c_glx[i]);
            const double r = sqrt(fma(k2 * sn, sn, 1.0));
            sum += c_glw[i] * (DIST ? r : (2.0 - kWgsF) / fma(1.0 - kWgsF, r, 1.0));
        }
    }
    return half * sum;
}

__device__ double meridian_arc(double abs_lat_deg) {
    const double b = kWgsA * (1.0 - kWgsF);
    const double ep2 = kWgsF * (2.0 - kWgsF) / ((1.0 - kWgsF) * (1.0 - kWgsF));
    const double beta = atan((1.0 - kWgsF) * tan(abs_lat_deg * (kPi / 180.0)));
    return b * geod_quad<true>(0.0, beta, ep2);
}

__device__ double same_lat_distance(double abs_lat_deg, double abs_dlon_deg) {
    if (abs_dlon_deg == 0.0) return 0.0;
    const double b = kWgsA * (1.0 - kWgsF);
    const double ep2 = kWgsF * (2.0 - kWgsF) / ((1.0 - kWgsF) * (1.0 - kWgsF));
    const double beta = atan((1.0 - kWgsF) * tan(abs_lat_deg * (kPi / 180.0)));
    const double half = 0.5 * abs_dlon_deg * (kPi / 180.0);
    if (beta == 0.0 && half <= (1.0 - kWgsF) * 0.5 * kPi) return kWgsA * 2.0 * half;     // along the equator
    const double sb = sin(beta);
    double lo = 0.0, hi = 0.5 * kPi - beta;
    double sig1 = 0.0, k2 = 0.0;
    for (int it = 0; it <= 64; ++it) {
        const double a0 = 0.5 * (lo + hi);
        double sa, ca;
        sincos(a0, &sa, &ca);
        k2 = ep2 * ca * ca;
        sig1 = asin(fmin(fmax(sb / ca, -1.0), 1.0));
        if (it == 64) break;
        double s1, c1;
        sincos(sig1, &s1, &c1);
        const double omega1 = atan2(sa * s1, c1);
        const double h = (0.5 * kPi - omega1) - kWgsF * sa * geod_quad<false>(sig1, 0.5 * kPi, k2);
        if (h > half) lo = a0; else hi = a0;            // the half-longitude decreases with alpha0
    }
    return 2.0 * b * geod_quad<true>(sig1, 0.5 * kPi, k2);
}

// functions.py:946-973 / :1006-1022: degrees -> signed meter coordinates (+ the half-turn distance)
__global__ void __launch_bounds__(128)
geod_to_meter_kernel(const double *__restrict__ lat_deg, const double *__restrict__ lon_deg,
                     double *__restrict__ lat_m, double *__restrict__ lon_m, double *__restrict__ half_turn,
                     long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lat = lat_deg[i];
    double lon = lon_deg[i];
    if (lon > 180.0) lon -= 360.0;                                   // :946-948
    const double sg_lat = (lat > 0.0) - (lat < 0.0), sg_lon = (lon > 0.0) - (lon < 0.0);   // np.sign
    const double m = meridian_arc(fabs(lat));
    lat_m[i] = m * sg_lat;
    lon_m[i] = same_lat_distance(fabs(lat), fabs(lon)) * sg_lon;
    if (half_turn) half_turn[i] = 2.0 * (meridian_arc(90.0) - m);     // geod.inv(0, lat, 180, lat): over the pole
}

constexpr int kMaxFields = 12;

// vtkPointInterpolator with a vtkGaussianKernel, radius footprint, null value NaN (functions.py:1038-1048),
// followed by the land mask (:1031, :1055).  src_lat_m ascending.
__global__ void __launch_bounds__(128)
gauss_interp_kernel(const double *__restrict__ src_lat_m, const double *__restrict__ src_lon_m,
                    const double *__restrict__ src_val, long long nsrc, int nfield,
                    const double *__restrict__ dst_lat_m, const double *__restrict__ dst_lon_m,
                    const float *__restrict__ land_fr, double *__restrict__ out, long long ndst,
                    double radius, double f2) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ndst) return;
    const double x = dst_lat_m[t], y = dst_lon_m[t];
    const double r2 = radius * radius, eps = 256.0 * 2.220446049250313e-16;
    double sw[kMaxFields], swv[kMaxFields];
    unsigned locked = 0;
#pragma unroll
    for (int f = 0; f < kMaxFields; ++f) { sw[f] = 0.0; swv[f] = 0.0; }
    long long lo = 0, hi = nsrc;                       // first source with lat_m >= x - radius
    const double xmin = x - radius, xmax = x + radius;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (src_lat_m[mid] < xmin) lo = mid + 1; else hi = mid;
    }
    for (long long i = lo; i < nsrc; ++i) {
        const double sx = src_lat_m[i];
        if (sx > xmax) break;
        const double dx = x - sx, dy = y - src_lon_m[i];
        const double d2 = fma(dx, dx, dy * dy);
        if (d2 > r2) continue;
        const bool hit = d2 < eps;                     // precise hit on an existing point
        const double w = exp(-f2 * d2);
#pragma unroll
        for (int f = 0; f < kMaxFields; ++f) {
            if (f < nfield && !((locked >> f) & 1u)) {
                const double v = src_val[(long long)f * nsrc + i];
                if (!isnan(v)) {
                    if (hit) { sw[f] = 1.0; swv[f] = v; locked |= 1u << f; }
                    else { sw[f] += w; swv[f] = fma(w, v, swv[f]); }
                }
            }
        }
    }
    const bool land = land_fr && land_fr[t] > 0.7f;
#pragma unroll
    for (int f = 0; f < kMaxFields; ++f)
        if (f < nfield) out[(long long)f * ndst + t] = (land || !(sw[f] > 0.0)) ? NAN : swv[f] / sw[f];
}

}  // namespace pgw

int pgw_geod_to_meter_f64(const double *lat_deg, const double *lon_deg, double *lat_m, double *lon_m,
                          double *half_turn, long long n, void *stream) {
    PGW_REQUIRE(lat_deg && lon_deg && lat_m && lon_m && n > 0);
    pgw::geod_to_meter_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        lat_deg, lon_deg, lat_m, lon_m, half_turn, n);
    return pgw_check_launch("geod_to_meter_kernel");
}

int pgw_gauss_interp_f64(const double *src_lat_m, const double *src_lon_m, const double *src_val,
                         long long nsrc, int nfield, const double *dst_lat_m, const double *dst_lon_m,
                         const float *land_fr, double *out, long long ndst, double radius, double sharpness,
                         void *stream) {
    PGW_REQUIRE(src_lat_m && src_lon_m && src_val && dst_lat_m && dst_lon_m && out);
    PGW_REQUIRE(nsrc >= 0 && ndst > 0 && nfield >= 1 && nfield <= pgw::kMaxFields && radius > 0.0);
    const double f = sharpness / radius;
    pgw::gauss_interp_kernel<<<(unsigned)((ndst + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        src_lat_m, src_lon_m, src_val, nsrc, nfield, dst_lat_m, dst_lon_m, land_fr, out, ndst, radius, f * f);
    return pgw_check_launch("gauss_interp_kernel");
}
