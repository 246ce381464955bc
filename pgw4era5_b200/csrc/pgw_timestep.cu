// The fused per-timestep column kernel of libpgw_b200 (sm_100a).
//
// One thread owns one ERA5 column; a warp is 32 adjacent longitudes, so every
// level access of a warp is one coalesced 128-byte line.  The column is swept
// bottom-up so that everything the surface-pressure iteration needs (the
// levels between the surface and p_ref) is seen first and parked in shared
// memory; after the iteration the rest of the column is streamed through
// registers with the adjusted surface pressure already known.  Each input
// element is read once and each output element written once.
//
// Reference semantics restated here (menschj/PGW4ERA5):
//   step_03_apply_to_era.py:64-94    pressures, RELHUM
//   step_03_apply_to_era.py:103-146  sea ice, skin and soil temperature
//   functions.py:288-292             two-point time interpolation of deltas
//   functions.py:343-366, 369-431    surface insertion + vertical interpolation
//   functions.py:511-580             interp_extrap_1d ('constant' mode)
//   step_03_apply_to_era.py:158-173  delta application
//   step_03_apply_to_era.py:182-319  surface-pressure fixed point
//   functions.py:118-125, 128-189    hus from RH, geopotential integration
#include "pgw_common.cuh"

namespace pgw {

__device__ __forceinline__ float blend_f32(const pgw_tslab &s, float w, long long off) {
    const float lo = __ldg(s.lo + off);
    if (w == 0.0f) return lo;
    const float hi = __ldg(s.hi + off);
    return fmaf(w, hi - lo, lo);
}

// scipy interp1d._call_linear with x = [0, x_hi]: slope * x_new + y_lo
__device__ __forceinline__ double blend_f64(const pgw_tslab &s, long long off) {
    const double lo = (double)__ldg(s.lo + off);
    if (s.x_new == 0.0) return lo;
    const double hi = (double)__ldg(s.hi + off);
    return (hi - lo) / s.x_hi * s.x_new + lo;
}

// Downward merge walk over the (pressure-ascending) source nodes of one pair of
// variables.  Invariant after advance(p): lo < 0 (target above the first node),
// or p_lo <= p and (hi is virtual or p < p_hi).
struct Walk2 {
    int lo;
    bool hi_virtual;
    float p_lo, p_hi, inv_w;
    float a_lo, a_hi, a_nx;
    float b_lo, b_hi, b_nx;
};

struct ColumnCtx {
    const pgw_timestep_args *a;
    const float *s_plev;     // ascending
    long long c;
    float wa, wb;            // time weights of the two variables of a pair
};

__device__ __forceinline__ void load_node(const pgw_timestep_args &a, const pgw_tslab &va, float wa,
                                          const pgw_tslab &vb, float wb, int j, long long c,
                                          float &xa, float &xb) {
    const int fj = a.plev_descending ? (a.nplev - 1 - j) : j;
    const long long off = (long long)fj * a.ncol + c;
    xa = blend_f32(va, wa, off);
    xb = blend_f32(vb, wb, off);
}

__device__ __forceinline__ void walk_advance(Walk2 &w, float p, const pgw_timestep_args &a,
                                             const pgw_tslab &va, float wa, const pgw_tslab &vb,
                                             float wb, const float *s_plev, long long c) {
    while (w.lo >= 0 && w.p_lo > p) {
        w.p_hi = w.p_lo; w.a_hi = w.a_lo; w.b_hi = w.b_lo; w.hi_virtual = false;
        --w.lo;
        if (w.lo >= 0) {
            w.p_lo = s_plev[w.lo];
            w.a_lo = w.a_nx; w.b_lo = w.b_nx;
            if (w.lo >= 1) load_node(a, va, wa, vb, wb, w.lo - 1, c, w.a_nx, w.b_nx);
            w.inv_w = __frcp_rn(__log2f(__fdividef(w.p_hi, w.p_lo)));
        }
    }
}

// interp_extrap_1d, 'constant' mode (functions.py:511-580)
__device__ __forceinline__ void walk_eval(const Walk2 &w, float p, float &xa, float &xb) {
    if (w.lo < 0) { xa = w.a_hi; xb = w.b_hi; }                       // below first node
    else if (w.hi_virtual || p == w.p_lo) { xa = w.a_lo; xb = w.b_lo; } // beyond last / exact
    else {
        const float t = __log2f(__fdividef(p, w.p_lo)) * w.inv_w;
        xa = fmaf(t, w.a_hi - w.a_lo, w.a_lo);
        xb = fmaf(t, w.b_hi - w.b_lo, w.b_lo);
    }
}

constexpr int kU = 4;   // levels per register batch (double-buffered)

template <int NT>
__global__ void __launch_bounds__(NT)
pgw_column_kernel(const __grid_constant__ pgw_timestep_args a, const int lst, const int np) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int L = a.nlev, K = a.nplev;
    double *s_ak = reinterpret_cast<double *>(smem);
    double *s_bk = s_ak + (L + 1);
    double *s_akm = s_bk + (L + 1);
    double *s_bkm = s_akm + L;
    double *st_T = s_bkm + L;                                  // [np][NT] T_pgw (float64)
    float *st_e = reinterpret_cast<float *>(st_T + (size_t)np * NT);   // [np][NT] e_pgw
    float *s_akmf = st_e + (size_t)np * NT;
    float *s_bkmf = s_akmf + L;
    float *s_plev = s_bkmf + L;                                // [K] ascending pressure

    const int tid = threadIdx.x;
    for (int i = tid; i <= L; i += NT) { s_ak[i] = a.ak[i]; s_bk[i] = a.bk[i]; }
    for (int i = tid; i < L; i += NT) {
        const double am = a.akm[i], bm = a.bkm[i];
        s_akm[i] = am; s_bkm[i] = bm; s_akmf[i] = (float)am; s_bkmf[i] = (float)bm;
    }
    for (int i = tid; i < K; i += NT)
        s_plev[i] = (float)a.plev[a.plev_descending ? (K - 1 - i) : i];
    __syncthreads();

    const long long n = a.ncol;
    long long c = (long long)blockIdx.x * NT + tid;
    const bool active = c < n;
    if (!active) c = n - 1;          // compute redundantly, never store
    unsigned errbits = 0;

    // ---------------- surface, skin and soil (step_03:103-146) ----------------
    const float ps_f = __ldg(a.PS + c);
    const double PSd = (double)ps_f;
    {
        // FR_SEA_ICE is float32 in the file and updated in place there
        float sic = (float)((double)__ldg(a.FR_SEA_ICE + c) + blend_f64(a.siconc, c) / 100.0);
        sic = sic < 0.0f ? 0.0f : (sic > 1.0f ? 1.0f : sic);           // np.clip keeps NaN
        const double dts = blend_f64(a.ts, c);
        const double dtos = blend_f64(a.tos, c);
        double comb = dts;                                            // integrate_tos
        if (!isnan(sic) && !isnan(dtos)) {
            float fr = sic + __ldg(a.FR_LAND + c);
            fr = fr < 0.0f ? 0.0f : (fr > 1.0f ? 1.0f : fr);
            comb = (double)fr * dts + (double)(1.0f - fr) * dtos;
        }
        const double clim = (double)__ldg(a.ts_clim + c);
        if (active) {
            a.FR_SEA_ICE_out[c] = sic;
            a.T_SKIN_out[c] = (float)((double)__ldg(a.T_SKIN + c) + comb);
            for (int s = 0; s < a.nsoil; ++s) {
                const double dso = clim + a.soil_decay[s] * (comb - clim);
                a.T_SO_out[(long long)s * n + c] =
                    (float)((double)__ldg(a.T_SO + (long long)s * n + c) + dso);
            }
        }
    }

    // ---------------- delta walkers (functions.py:343-431) ----------------
    const float w_ta = (a.ta.x_new == 0.0) ? 0.0f : (float)(a.ta.x_new / a.ta.x_hi);
    const float w_hur = (a.hur.x_new == 0.0) ? 0.0f : (float)(a.hur.x_new / a.hur.x_hi);
    const float w_ua = (a.ua.x_new == 0.0) ? 0.0f : (float)(a.ua.x_new / a.ua.x_hi);
    const float w_va = (a.va.x_new == 0.0) ? 0.0f : (float)(a.va.x_new / a.va.x_hi);

    Walk2 wA, wB;
    {
        // replace_delta_sfc: node s carries (ps_hist, surface delta); nodes above it
        // (in pressure) all hold the surface delta, so they never matter.
        const float psh = (float)blend_f64(a.ps_hist, c);
        int s = K - 1;
        if (!(psh > s_plev[K - 1])) {
            s = -1;
            for (int k = K - 1; k >= 0; --k)
                if (s_plev[k] < psh) { s = k; break; }
        }
        if (s < 0) { errbits |= PGW_ERR_PS_HIST_RANGE; s = 0; }
        wA.lo = s; wA.hi_virtual = true; wA.p_lo = psh; wA.p_hi = psh; wA.inv_w = 0.0f;
        wA.a_lo = (float)blend_f64(a.tas, c);
        wA.b_lo = (float)blend_f64(a.hurs, c);
        wA.a_hi = wA.a_lo; wA.b_hi = wA.b_lo; wA.a_nx = wA.a_lo; wA.b_nx = wA.b_lo;
        if (s >= 1) load_node(a, a.ta, w_ta, a.hur, w_hur, s - 1, c, wA.a_nx, wA.b_nx);

        wB.lo = K - 1; wB.hi_virtual = true; wB.p_lo = s_plev[K - 1]; wB.p_hi = wB.p_lo; wB.inv_w = 0.0f;
        load_node(a, a.ua, w_ua, a.va, w_va, K - 1, c, wB.a_lo, wB.b_lo);
        wB.a_hi = wB.a_lo; wB.b_hi = wB.b_lo; wB.a_nx = wB.a_lo; wB.b_nx = wB.b_lo;
        if (K >= 2) load_node(a, a.ua, w_ua, a.va, w_va, K - 2, c, wB.a_nx, wB.b_nx);
    }
    float min_src_p = (wA.lo == 0) ? wA.p_lo : s_plev[0];

    const double pref = a.p_ref;
    double pb_era = fma(PSd, s_bk[L], s_ak[L]);
    double acc_era = 0.0;
    if (pb_era < pref) errbits |= PGW_ERR_PREF_BELOW_SFC;

    // One model level: RELHUM of the ERA state, interpolated deltas, PGW state.
    // Returns e_pgw (vapour pressure of the PGW state, iteration invariant).
    auto level = [&](int l, float t, float q, float u, float v, double &Tp, float &e_pgw) {
        const float p = fmaf(ps_f, s_bkmf[l], s_akmf[l]);
        walk_advance(wA, p, a, a.ta, w_ta, a.hur, w_hur, s_plev, c);
        walk_advance(wB, p, a, a.ua, w_ua, a.va, w_va, s_plev, c);
        float dta, dhur, dua, dva;
        walk_eval(wA, p, dta, dhur);
        walk_eval(wB, p, dua, dva);
        const float tm273 = t - 273.0f;
        const float e_era = __fdividef(q * p, 0.622f + 0.378f * q);          // functions.py:58-64
        const float rh_pgw = __fdividef(100.0f * e_era, esat_fast(tm273, 0.0f)) + dhur;
        e_pgw = rh_pgw * 0.01f * esat_fast(tm273, dta);                      // functions.py:123
        Tp = (double)t + (double)dta;
        if (active) {
            const long long off = (long long)l * n + c;
            st_stream(a.T_out + off, (float)Tp);
            st_stream(a.U_out + off, u + dua);
            st_stream(a.V_out + off, v + dva);
        }
    };

    auto load_batch = [&](int l0, int l_end, float *t, float *q, float *u, float *v) {
#pragma unroll
        for (int i = 0; i < kU; ++i) {
            const int l = l0 - i;
            if (l >= l_end) {
                const long long off = (long long)l * n + c;
                t[i] = ld_stream(a.T + off); q[i] = ld_stream(a.QV + off);
                u[i] = ld_stream(a.U + off); v[i] = ld_stream(a.V + off);
            }
        }
    };

    // ---------------- phase 1: surface .. p_ref, parked in shared memory ----------------
    {
        float t0[kU], q0[kU], u0[kU], v0[kU], t1[kU], q1[kU], u1[kU], v1[kU];
        load_batch(L - 1, lst, t0, q0, u0, v0);
        for (int l0 = L - 1; l0 >= lst; l0 -= kU) {
            load_batch(l0 - kU, lst, t1, q1, u1, v1);
#pragma unroll
            for (int i = 0; i < kU; ++i) {
                const int l = l0 - i;
                if (l >= lst) {
                    double Tp; float e_pgw;
                    level(l, t0[i], q0[i], u0[i], v0[i], Tp, e_pgw);
                    st_T[(size_t)(l - lst) * NT + tid] = Tp;
                    st_e[(size_t)(l - lst) * NT + tid] = e_pgw;
                    // geopotential of the ERA state (functions.py:128-189)
                    const double pt = fma(PSd, s_bk[l], s_ak[l]);
                    const double tv = (double)t0[i] * (1.0 + 0.61 * (double)q0[i]);
                    const double pte = fmin(fmax(pt, pref), pb_era);
                    acc_era = fma(tv, log_ratio(pb_era, pte), acc_era);
                    pb_era = pt;
                }
            }
#pragma unroll
            for (int i = 0; i < kU; ++i) { t0[i] = t1[i]; q0[i] = q1[i]; u0[i] = u1[i]; v0[i] = v1[i]; }
        }
    }
    const double fis = (double)__ldg(a.FIS + c);
    const double phi_era = fis + kRd * acc_era;
    const double gdzg = blend_f64(a.zg_ref, c) * kG;                 // step_03:292-295
    const double t_low = st_T[(size_t)(L - 1 - lst) * NT + tid];    // ta_pgw on the lowest level

    // ---------------- phase 2: surface-pressure fixed point (step_03:182-319) ----------------
    double dps = 0.0, adj = 0.0, psn = PSd;
    for (int k = 0; k < a.k_spec; ++k) {
        dps += adj;
        psn = PSd + dps;
        if (active) a.dps_traj[(long long)k * n + c] = (float)dps;
        if (psn > a.ps_bound) errbits |= PGW_ERR_PS_BOUND;
        double pb = fma(psn, s_bk[L], s_ak[L]);
        if (pb < pref) errbits |= PGW_ERR_PREF_BELOW_SFC;
        double acc = 0.0;
        for (int l = L - 1; l >= lst; --l) {
            const double pt = fma(psn, s_bk[l], s_ak[l]);
            const float pf = (float)fma(psn, s_bkm[l], s_akm[l]);
            const float e = st_e[(size_t)(l - lst) * NT + tid];
            const double Td = st_T[(size_t)(l - lst) * NT + tid];
            const float hus = __fdividef(0.622f * e, pf - 0.378f * e);      // functions.py:66-72
            const double tv = fma(Td, 0.61 * (double)hus, Td);
            const double pte = fmin(fmax(pt, pref), pb);
            acc = fma(tv, log_ratio(pb, pte), acc);
            pb = pt;
        }
        const double phi_pgw = fis + kRd * acc;
        const double err = (phi_pgw - phi_era) - gdzg;
        adj = -a.adj_factor * psn / (kRd * t_low) * err;
        double ae = (active && !isnan(err)) ? fabs(err) : 0.0;          // max skips NaN (step_03:308)
        ae = warp_max(ae);
        if ((tid & 31) == 0 && ae > 0.0)
            atomicMax(reinterpret_cast<unsigned long long *>(a.maxerr + k),
                      (unsigned long long)__double_as_longlong(ae));
    }

    // ---------------- phase 3: PS, QV of the parked levels, then the upper column ----------------
    const float psn_f = (float)psn;
    if (active) {
        a.PS_out[c] = psn_f;
        a.dps_out[c] = (float)dps;
        for (int l = L - 1; l >= lst; --l) {
            const float pf = (float)fma(psn, s_bkm[l], s_akm[l]);
            const float e = st_e[(size_t)(l - lst) * NT + tid];
            st_stream(a.QV_out + (long long)l * n + c, __fdividef(0.622f * e, pf - 0.378f * e));
        }
    }
    if (lst > 0) {
        float t0[kU], q0[kU], u0[kU], v0[kU], t1[kU], q1[kU], u1[kU], v1[kU];
        load_batch(lst - 1, 0, t0, q0, u0, v0);
        for (int l0 = lst - 1; l0 >= 0; l0 -= kU) {
            load_batch(l0 - kU, 0, t1, q1, u1, v1);
#pragma unroll
            for (int i = 0; i < kU; ++i) {
                const int l = l0 - i;
                if (l >= 0) {
                    double Tp; float e_pgw;
                    level(l, t0[i], q0[i], u0[i], v0[i], Tp, e_pgw);
                    const float pf = (float)fma(psn, s_bkm[l], s_akm[l]);
                    if (active)
                        st_stream(a.QV_out + (long long)l * n + c,
                                  __fdividef(0.622f * e_pgw, pf - 0.378f * e_pgw));
                }
            }
#pragma unroll
            for (int i = 0; i < kU; ++i) { t0[i] = t1[i]; q0[i] = q1[i]; u0[i] = u1[i]; v0[i] = v1[i]; }
        }
    }

    // ---------------- bookkeeping for the host-side checks ----------------
    float p_top = active ? fmaf(ps_f, s_bkmf[0], s_akmf[0]) : INFINITY;   // functions.py:417
    p_top = warp_min(p_top);
    min_src_p = warp_min(active ? min_src_p : INFINITY);
    if ((tid & 31) == 0) {
        atomicMin(reinterpret_cast<unsigned *>(a.stats), __float_as_uint(fmaxf(p_top, 0.0f)));
        atomicMin(reinterpret_cast<unsigned *>(a.stats) + 1, __float_as_uint(fmaxf(min_src_p, 0.0f)));
    }
    if (!active) errbits = 0;
    if (errbits) atomicOr(a.err, errbits);
}

__global__ void pgw_timestep_init_kernel(uint64_t *maxerr, float *stats) {
    const int i = threadIdx.x;
    if (i < PGW_MAX_ITER) maxerr[i] = 0ull;
    if (i < 2) stats[i] = INFINITY;
}

// N = first k with max|err_k| <= thresh (step_03:189,308)
__global__ void pgw_converge_kernel(const uint64_t *maxerr, int k_spec, double thresh,
                                    pgw_timestep_result *res) {
    int n_iter = 0;
    for (int k = 0; k < k_spec; ++k) {
        const double e = __longlong_as_double((long long)maxerr[k]);
        if (!(e > thresh)) { n_iter = k + 1; break; }
    }
    res->n_iter = n_iter;
    res->converged = n_iter > 0;
    res->rewritten = (n_iter > 0 && n_iter < k_spec);
    res->reserved = 0;
}

// If the field converged before k_spec, PS and QV were written for dps_{k_spec};
// rewrite them for dps_N.  e_pgw is recovered from the speculative QV.
__global__ void __launch_bounds__(256)
pgw_rewrite_kernel(const __grid_constant__ pgw_timestep_args a, const pgw_timestep_result *res) {
    if (!res->rewritten) return;
    const long long n = a.ncol;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const int N = res->n_iter;
    const double PSd = (double)a.PS[c];
    const double dps_n = (double)a.dps_traj[(long long)(N - 1) * n + c];
    const double dps_s = (double)a.dps_traj[(long long)(a.k_spec - 1) * n + c];
    const double ps_n = PSd + dps_n, ps_s = PSd + dps_s;
    a.PS_out[c] = (float)ps_n;
    a.dps_out[c] = (float)dps_n;
    for (int l = 0; l < a.nlev; ++l) {
        const double bm = a.bkm[l];
        if (bm == 0.0) continue;                 // pressure independent of ps
        const double am = a.akm[l];
        const float p_s = (float)fma(ps_s, bm, am);
        const float p_n = (float)fma(ps_n, bm, am);
        const long long off = (long long)l * n + c;
        const float q = a.QV_out[off];
        const float e = __fdividef(q * p_s, 0.622f + 0.378f * q);
        a.QV_out[off] = __fdividef(0.622f * e, p_n - 0.378f * e);
    }
}

}  // namespace pgw

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
namespace {

constexpr int kColumnThreads = 128;

// First (topmost) full level whose layer can reach p_ref for any ps <= ps_bound:
// the layer above the first half level with ak + ps_bound*bk >= p_ref.
int stash_top(const double *ak, const double *bk, int nlev, double p_ref, double ps_bound) {
    int h = nlev;
    for (int i = 0; i <= nlev; ++i)
        if (ak[i] + ps_bound * bk[i] >= p_ref) { h = i; break; }
    int lst = h - 1;
    if (lst < 0) lst = 0;
    if (lst > nlev - 1) lst = nlev - 1;
    return lst;
}

size_t column_smem(int nlev, int nplev, int np, int nt) {
    return sizeof(double) * (size_t)(2 * (nlev + 1) + 2 * nlev) +
           (size_t)np * nt * (sizeof(double) + sizeof(float)) +
           sizeof(float) * (size_t)(2 * nlev + nplev) + 16;
}

int validate(const pgw_timestep_args *a) {
    if (!a) return PGW_E_INVALID;
    if (a->ncol <= 0 || a->nlev < 2 || a->nplev < 2 || a->nplev > 64) return PGW_E_INVALID;
    if (a->nsoil < 0 || a->nsoil > PGW_MAX_SOIL) return PGW_E_INVALID;
    if (a->k_spec < 1 || a->k_spec > PGW_MAX_ITER) return PGW_E_INVALID;
    const void *need[] = {a->ak_host, a->bk_host, a->ak, a->bk, a->akm, a->bkm, a->plev, a->PS, a->FIS, a->FR_LAND, a->FR_SEA_ICE,
                          a->T_SKIN, a->T, a->QV, a->U, a->V, a->ta.lo, a->ta.hi, a->hur.lo, a->hur.hi,
                          a->ua.lo, a->ua.hi, a->va.lo, a->va.hi, a->tas.lo, a->tas.hi, a->hurs.lo,
                          a->hurs.hi, a->ps_hist.lo, a->ps_hist.hi, a->ts.lo, a->ts.hi, a->tos.lo, a->tos.hi,
                          a->siconc.lo, a->siconc.hi, a->zg_ref.lo, a->zg_ref.hi, a->ts_clim, a->PS_out,
                          a->T_SKIN_out, a->FR_SEA_ICE_out, a->T_out, a->QV_out, a->U_out, a->V_out,
                          a->dps_out, a->dps_traj, a->maxerr, a->stats, a->err};
    for (const void *p : need) if (!p) return PGW_E_INVALID;
    if (a->nsoil > 0 && (!a->T_SO || !a->T_SO_out)) return PGW_E_INVALID;
    return PGW_OK;
}

}  // namespace

extern "C" {

long long pgw_timestep_smem_bytes(const pgw_timestep_args *a) {
    int rc = validate(a);
    if (rc != PGW_OK) return rc;
    const int lst = stash_top(a->ak_host, a->bk_host, a->nlev, a->p_ref, a->ps_bound);
    return (long long)column_smem(a->nlev, a->nplev, a->nlev - lst, kColumnThreads);
}

int pgw_timestep(const pgw_timestep_args *a, void *stream) {
    int rc = validate(a);
    if (rc != PGW_OK) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const int lst = stash_top(a->ak_host, a->bk_host, a->nlev, a->p_ref, a->ps_bound);
    const int np = a->nlev - lst;
    const size_t smem = column_smem(a->nlev, a->nplev, np, kColumnThreads);
    auto kern = pgw::pgw_column_kernel<kColumnThreads>;
    static thread_local size_t configured = 0;
    if (smem > configured) {
        int dev = 0, max_optin = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        if (smem > (size_t)max_optin) {
            pgw_set_error("column stash needs %zu B of shared memory (%d levels below p_ref), device allows %d",
                          smem, np, max_optin);
            return PGW_E_SMEM;
        }
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return pgw_check_launch("cudaFuncSetAttribute");
        cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                             cudaSharedmemCarveoutMaxShared);
        configured = smem;
    }
    pgw::pgw_timestep_init_kernel<<<1, 64, 0, st>>>(a->maxerr, a->stats);
    const unsigned grid = (unsigned)((a->ncol + kColumnThreads - 1) / kColumnThreads);
    kern<<<grid, kColumnThreads, smem, st>>>(*a, lst, np);
    return pgw_check_launch("pgw_column_kernel");
}

int pgw_timestep_finalize(const pgw_timestep_args *a, pgw_timestep_result *result_dev, void *stream) {
    int rc = validate(a);
    if (rc != PGW_OK || !result_dev) return PGW_E_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    pgw::pgw_converge_kernel<<<1, 1, 0, st>>>(a->maxerr, a->k_spec, a->thresh_phi_ref_max_error, result_dev);
    const unsigned grid = (unsigned)((a->ncol + 255) / 256);
    pgw::pgw_rewrite_kernel<<<grid, 256, 0, st>>>(*a, result_dev);
    return pgw_check_launch("pgw_timestep_finalize");
}

}  // extern "C"
