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

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device AND per kernel: remember what each (device, kernel
// variant) has been opted in to.  (A cache keyed by thread alone skipped the opt-in when one thread launched on a
// second GPU, as pool workers of parallel.IterMP and in-process multi-GPU callers do.)
int pgw_ensure_smem(const void *kernel, int slot, size_t smem, int np) {
    constexpr int kMaxDev = 64, kSlots = 16;
    static thread_local size_t configured[kMaxDev][kSlots] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return pgw_check_launch("cudaGetDevice");
    const bool cached = dev >= 0 && dev < kMaxDev && slot >= 0 && slot < kSlots;
    if (cached && smem <= configured[dev][slot]) return PGW_OK;
    int max_optin = 0;
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (smem > (size_t)max_optin) {
        pgw_set_error("column stash needs %zu B of shared memory (%d levels below p_ref), device allows %d",
                      smem, np, max_optin);
        return PGW_E_SMEM;
    }
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return pgw_check_launch("cudaFuncSetAttribute");
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (cached) configured[dev][slot] = smem;
    return PGW_OK;
}

extern "C" {
#define PGW_STR2(x) #x
#define PGW_STR(x) PGW_STR2(x)
const char *pgw_version(void) { return "pgw_b200 0.4.0 (sm_100a, abi " PGW_STR(PGW_B200_ABI_VERSION) ")"; }
int pgw_abi_version(void) { return PGW_B200_ABI_VERSION; }
const char *pgw_last_error(void) { return g_last_error; }
long long pgw_sizeof_timestep_args(void) { return (long long)sizeof(pgw_timestep_args); }
}
