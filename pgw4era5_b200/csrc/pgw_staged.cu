// Small kernels of the STAGED per-timestep path: the reference's loop run stage by stage
// (step_03_apply_to_era.py:60-343) on float64 device arrays with the stand-alone operators of
// pgw_ops.cu.  It serves the settings the fused column kernel does not cover, i_reinterp = 1
// (:202-216, :330-343) and p_ref_inp = None (:219-251), and is not a performance path.
#include "pgw_column.cuh"

namespace pgw {

__device__ __forceinline__ long long gidx() { return (long long)blockIdx.x * blockDim.x + threadIdx.x; }

// sea ice, skin and soil temperature (step_03:103-146, integrate_tos functions.py:1145-1186);
// the same block as in the prologue of the fused column kernels
__global__ void surface_update_kernel(const __grid_constant__ pgw_timestep_args a) {
    const long long i = gidx();
    if (i >= a.ncol) return;
    const uint32_t c = (uint32_t)i, n = (uint32_t)a.ncol;
    const Pair2 r_sic = load_pair(a.siconc, c), r_ts = load_pair(a.ts, c), r_tos = load_pair(a.tos, c);
    float sic = (float)((double)a.FR_SEA_ICE[c] + blend_f64(a.siconc, r_sic) / 100.0);
    sic = sic < 0.0f ? 0.0f : (sic > 1.0f ? 1.0f : sic);           // np.clip keeps NaN
    const double dts = blend_f64(a.ts, r_ts);
    const double dtos = blend_f64(a.tos, r_tos);
    double comb = dts;
    if (!isnan(sic) && !isnan(dtos)) {
        float fr = sic + a.FR_LAND[c];
        fr = fr < 0.0f ? 0.0f : (fr > 1.0f ? 1.0f : fr);
        comb = (double)fr * dts + (double)(1.0f - fr) * dtos;
    }
    const double clim = (double)a.ts_clim[c];
    a.FR_SEA_ICE_out[c] = sic;
    a.T_SKIN_out[c] = (float)((double)a.T_SKIN[c] + comb);
    for (int s = 0; s < a.nsoil; ++s) {
        const double dso = clim + a.soil_decay[s] * (comb - clim);
        a.T_SO_out[(uint32_t)s * n + c] = (float)((double)a.T_SO[(uint32_t)s * n + c] + dso);
    }
}

// p[l, c] = a[l] + ps[c] * b[l]     (step_03:64-88, :196-199)
__global__ void hybrid_pressure_kernel(const double *__restrict__ ps, const double *__restrict__ a,
                                       const double *__restrict__ b, double *__restrict__ out, int nlev,
                                       long long ncol) {
    const long long c = gidx();
    if (c >= ncol) return;
    const double p = ps[c];
    for (int l = 0; l < nlev; ++l) out[(long long)l * ncol + c] = a[l] + p * b[l];
}

__global__ void axpy_kernel(const double *__restrict__ x, const double *__restrict__ y, double alpha,
                            double *__restrict__ out, long long n) {
    const long long i = gidx();
    if (i < n) out[i] = x[i] + alpha * y[i];
}

// determine_p_ref (functions.py:583-598) per column: the first option p (in the order given) with
// p_min_era > p and p_min_pgw > p, limited by the previous iteration's choice; NaN if none.
__global__ void determine_p_ref_kernel(const double *__restrict__ p_min_era, const double *__restrict__ p_min_pgw,
                                       const double *__restrict__ opts, int nopt,
                                       const double *__restrict__ p_ref_last, double *__restrict__ out,
                                       long long n, uint32_t *err) {
    const long long c = gidx();
    if (c >= n) return;
    const double pe = p_min_era[c], pp = p_min_pgw[c];
    double r = NAN;
    for (int k = 0; k < nopt; ++k) {
        const double p = opts[k];
        if (pe > p && pp > p) {
            r = p_ref_last ? fmin(p, p_ref_last[c]) : p;
            break;
        }
    }
    out[c] = r;
    if (isnan(r)) atomicOr(err, PGW_ERR_NO_PREF);
}

// field.sel(plev = p_ref) with a per-column p_ref (step_03:292-295): exact match of the level value
__global__ void select_plev_kernel(const double *__restrict__ field, const double *__restrict__ plev, int K,
                                   const double *__restrict__ p_ref, double *__restrict__ out, long long n) {
    const long long c = gidx();
    if (c >= n) return;
    const double p = p_ref[c];
    double r = NAN;
    for (int k = 0; k < K; ++k)
        if (plev[k] == p) { r = field[(long long)k * n + c]; break; }
    out[c] = r;
}

// phi_ref_error, adj_ps and max|phi_ref_error| (step_03:286-308)
__global__ void ps_adjust_kernel(const double *__restrict__ phi_pgw, const double *__restrict__ phi_era,
                                 const double *__restrict__ dphi_clim, const double *__restrict__ ps_pgw,
                                 const double *__restrict__ ta_low, double adj_factor, double *__restrict__ adj,
                                 unsigned long long *maxerr, long long n) {
    const long long c = gidx();
    double ae = 0.0;
    if (c < n) {
        const double e = (phi_pgw[c] - phi_era[c]) - dphi_clim[c];
        adj[c] = -adj_factor * ps_pgw[c] / (kRd * ta_low[c]) * e;
        ae = isnan(e) ? 0.0 : fabs(e);                                // np.abs(...).max() skips NaN
    }
    ae = warp_max(ae);
    if ((threadIdx.x & 31) == 0 && ae > 0.0) atomicMax(maxerr, (unsigned long long)__double_as_longlong(ae));
}

}  // namespace pgw

using namespace pgw;
#define PGW_REQUIRE(cond) do { if (!(cond)) return PGW_E_INVALID; } while (0)
static unsigned grid_of(long long n, int bs) { return (unsigned)((n + bs - 1) / bs); }

extern "C" {

int pgw_surface_update(const pgw_timestep_args *a, void *stream) {
    PGW_REQUIRE(a && a->ncol > 0 && a->nsoil >= 0 && a->nsoil <= PGW_MAX_SOIL);
    PGW_REQUIRE(a->FR_SEA_ICE && a->FR_LAND && a->T_SKIN && a->ts_clim && a->siconc.lo && a->siconc.hi &&
                a->ts.lo && a->ts.hi && a->tos.lo && a->tos.hi && a->FR_SEA_ICE_out && a->T_SKIN_out);
    PGW_REQUIRE(a->nsoil == 0 || (a->T_SO && a->T_SO_out));
    surface_update_kernel<<<grid_of(a->ncol, 256), 256, 0, (cudaStream_t)stream>>>(*a);
    return pgw_check_launch("surface_update_kernel");
}

int pgw_hybrid_pressure_f64(const double *ps, const double *a, const double *b, double *out, int nlev,
                            long long ncol, void *stream) {
    PGW_REQUIRE(ps && a && b && out && nlev > 0 && ncol > 0);
    hybrid_pressure_kernel<<<grid_of(ncol, 256), 256, 0, (cudaStream_t)stream>>>(ps, a, b, out, nlev, ncol);
    return pgw_check_launch("hybrid_pressure_kernel");
}

int pgw_axpy_f64(const double *x, const double *y, double alpha, double *out, long long n, void *stream) {
    PGW_REQUIRE(x && y && out && n > 0);
    axpy_kernel<<<grid_of(n, 256), 256, 0, (cudaStream_t)stream>>>(x, y, alpha, out, n);
    return pgw_check_launch("axpy_kernel");
}

int pgw_determine_p_ref_f64(const double *p_min_era, const double *p_min_pgw, const double *opts, int nopt,
                            const double *p_ref_last, double *out, long long n, uint32_t *err, void *stream) {
    PGW_REQUIRE(p_min_era && p_min_pgw && opts && out && err && nopt > 0 && n > 0);
    determine_p_ref_kernel<<<grid_of(n, 256), 256, 0, (cudaStream_t)stream>>>(p_min_era, p_min_pgw, opts, nopt,
                                                                             p_ref_last, out, n, err);
    return pgw_check_launch("determine_p_ref_kernel");
}

int pgw_select_plev_f64(const double *field, const double *plev, int K, const double *p_ref, double *out,
                        long long n, void *stream) {
    PGW_REQUIRE(field && plev && p_ref && out && K > 0 && n > 0);
    select_plev_kernel<<<grid_of(n, 256), 256, 0, (cudaStream_t)stream>>>(field, plev, K, p_ref, out, n);
    return pgw_check_launch("select_plev_kernel");
}

int pgw_ps_adjust_f64(const double *phi_pgw, const double *phi_era, const double *dphi_clim, const double *ps_pgw,
                      const double *ta_low, double adj_factor, double *adj, uint64_t *maxerr, long long n,
                      void *stream) {
    PGW_REQUIRE(phi_pgw && phi_era && dphi_clim && ps_pgw && ta_low && adj && maxerr && n > 0);
    ps_adjust_kernel<<<grid_of(n, 256), 256, 0, (cudaStream_t)stream>>>(
        phi_pgw, phi_era, dphi_clim, ps_pgw, ta_low, adj_factor, adj,
        reinterpret_cast<unsigned long long *>(maxerr), n);
    return pgw_check_launch("ps_adjust_kernel");
}

}  // extern "C"
