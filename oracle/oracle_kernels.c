/*
 * TEST INFRASTRUCTURE ONLY -- compiled column loops of the CPU oracle.
 *
 * Plain-C restatement (float64, single thread, same loop structure and the
 * same O(n_targ * n_src) linear scan) of the reference's two numba functions
 *   interp_1d_for_timelatlon   functions.py:479-508
 *   interp_extrap_1d           functions.py:511-580
 * and of the per-column surface insertion
 *   replace_delta_sfc          functions.py:343-366
 * which the reference runs through xr.apply_ufunc(vectorize=True)
 * (functions.py:396-402).  Checked against the unmodified reference functions
 * by tests/test_oracle_golden.py (fixtures: oracle/make_golden.py).
 *
 * Never linked into the product library.
 */
#include <math.h>
#include <stdlib.h>

enum { MODE_OFF = 0, MODE_LINEAR = 1, MODE_CONSTANT = 2, MODE_NAN = 3 };

/* functions.py:511-580; strided column views like numba's array slices.
 * returns 0, or 3 when extrapolation is required but mode is 'off'. */
static int interp_extrap_1d(const double *src_x, const double *src_y, long ns,
                            long sstride, const double *targ_x, double *targ_y,
                            long nt, long tstride, int mode)
{
    for (long ti = 0; ti < nt; ++ti) {
        const double tx = targ_x[ti * tstride];
        long i1 = -1, i2 = -1;
        int require_extrap = 0;
        for (long si = 0; si < ns; ++si) {
            const double sx = src_x[si * sstride];
            if (si == 0 && sx > tx) {                 /* :530-538 */
                if (mode == MODE_LINEAR) { i1 = si; i2 = si + 1; }
                else if (mode == MODE_CONSTANT) { i1 = si; i2 = si; }
                require_extrap = 1;
                break;
            } else if (sx == tx) {                    /* :540-543 */
                i1 = si; i2 = si;
                break;
            } else if (sx > tx) {                     /* :545-548 */
                i1 = si - 1; i2 = si;
                break;
            }
        }
        if (i1 == -1) {                               /* :554-561 */
            if (mode == MODE_LINEAR) { i1 = ns - 2; i2 = ns - 1; }
            else if (mode == MODE_CONSTANT) { i1 = ns - 1; i2 = ns - 1; }
            require_extrap = 1;
        }
        if (require_extrap && mode == MODE_OFF) return 3;   /* :564-566 */
        if (require_extrap && mode == MODE_NAN) {     /* :569-570 */
            targ_y[ti * tstride] = NAN;
        } else if (i1 == i2) {
            targ_y[ti * tstride] = src_y[i1 * sstride];
        } else {
            targ_y[ti * tstride] = src_y[i1 * sstride] +
                (tx - src_x[i1 * sstride]) *
                (src_y[i2 * sstride] - src_y[i1 * sstride]) /
                (src_x[i2 * sstride] - src_x[i1 * sstride]);
        }
    }
    return 0;
}

/* functions.py:479-508.  Arrays are C-order [ntime, K, nlat, nlon]. */
int oracle_interp_1d_for_timelatlon(const double *orig, const double *src_p,
                                    const double *targ_p, double *out,
                                    long ntime, long ks, long kt, long nlat,
                                    long nlon, int mode)
{
    const long plane = nlat * nlon;
    for (long t = 0; t < ntime; ++t)
        for (long j = 0; j < nlat; ++j)
            for (long i = 0; i < nlon; ++i) {
                const long cs = t * ks * plane + j * nlon + i;
                const long ct = t * kt * plane + j * nlon + i;
                if (src_p[cs + (ks - 1) * plane] < src_p[cs]) return 1;   /* :500-501 */
                if (targ_p[ct + (kt - 1) * plane] < targ_p[ct]) return 2; /* :502-503 */
                int rc = interp_extrap_1d(src_p + cs, orig + cs, ks, plane,
                                          targ_p + ct, out + ct, kt, plane, mode);
                if (rc) return rc;
            }
    return 0;
}

/* functions.py:343-366 applied to every column of [1, K, ncol] arrays.
 * returns 0, or 1 for the reference's bare ValueError() (ps_hist < min P, or
 * the empty argwhere when ps_hist == min P / NaN). */
int oracle_replace_delta_sfc(const double *source_P, const double *ps_hist,
                             const double *delta, const double *delta_sfc,
                             double *out_P, double *out_d, long K, long ncol)
{
    for (long c = 0; c < ncol; ++c) {
        double pmax = -INFINITY, pmin = INFINITY;
        for (long k = 0; k < K; ++k) {
            const double p = source_P[k * ncol + c];
            out_P[k * ncol + c] = p;
            out_d[k * ncol + c] = delta[k * ncol + c];
            if (p > pmax) pmax = p;
            if (p < pmin) pmin = p;
        }
        const double ph = ps_hist[c];
        if (ph > pmax) {
            out_P[(K - 1) * ncol + c] = ph;
            out_d[(K - 1) * ncol + c] = delta_sfc[c];
        } else if (ph < pmin) {
            return 1;
        } else {
            long sfc = -1;
            for (long k = 0; k < K; ++k)
                if (ph > source_P[k * ncol + c]) sfc = k;
            if (sfc < 0) return 1;
            for (long k = sfc; k < K; ++k) out_d[k * ncol + c] = delta_sfc[c];
            out_P[sfc * ncol + c] = ph;
        }
    }
    return 0;
}
