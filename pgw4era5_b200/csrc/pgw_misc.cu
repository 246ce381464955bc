// Version string and host-side error text of libpgw_b200.
#include <stdarg.h>
#include <stdio.h>

#include "pgw_common.cuh"

static thread_local char g_last_error[512] = "";

void pgw_set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
    va_end(ap);
}

int pgw_check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        pgw_set_error("%s: %s", what, cudaGetErrorString(e));
        return PGW_E_LAUNCH;
    }
    return PGW_OK;
}

extern "C" {
#define PGW_STR2(x) #x
#define PGW_STR(x) PGW_STR2(x)
const char *pgw_version(void) { return "pgw_b200 0.3.0 (sm_100a, abi " PGW_STR(PGW_B200_ABI_VERSION) ")"; }
int pgw_abi_version(void) { return PGW_B200_ABI_VERSION; }
const char *pgw_last_error(void) { return g_last_error; }
long long pgw_sizeof_timestep_args(void) { return (long long)sizeof(pgw_timestep_args); }
}
