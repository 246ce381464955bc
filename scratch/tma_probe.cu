#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cstdlib>
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "PGW_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra PGW_DONE;\n\t"
        "bra PGW_WAIT;\n\t"
        "PGW_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, int c0, int c1, const void *src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                 ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(src)) : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct P { CUtensorMap in, out; };

template <int MODE>
__global__ void probe(const __grid_constant__ P p, float *dbg) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *buf = reinterpret_cast<float *>(smem);
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 2048);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (MODE >= 1 && threadIdx.x == 0) {
        mbar_arrive_expect_tx(bar, 1024);
        tma_load_2d(buf, &p.in, blockIdx.x * 128, 3, bar);
    }
    if (MODE >= 2) {
        mbar_wait(bar, 0);
        dbg[blockIdx.x * 256 + threadIdx.x] = buf[threadIdx.x];
        dbg[blockIdx.x * 256 + 128 + threadIdx.x] = buf[128 + threadIdx.x];
        buf[threadIdx.x] += 1.0f; buf[128 + threadIdx.x] += 1.0f;
        fence_proxy_async();
    }
    __syncthreads();
    if (MODE >= 3 && threadIdx.x == 0) {
        tma_store_2d(&p.out, blockIdx.x * 128, 3, buf);
        tma_commit();
        tma_wait_all();
    }
}

template <int N>
__device__ __forceinline__ void tma_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(160, 3) pipe(const __grid_constant__ P p, int L) {
    extern __shared__ __align__(1024) unsigned char smem[];
    float *ring = reinterpret_cast<float *>(smem);           // [4][2][128]
    uint64_t *bar_full = reinterpret_cast<uint64_t *>(smem + 4096);
    uint64_t *bar_done = bar_full + 4;
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int i = 0; i < 4; ++i) { mbar_init(bar_full + i, 1); mbar_init(bar_done + i, 128); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int npairs = (L + 1) >> 1;
    if (tid >= 128) {
        if (tid != 128) return;
        const int c0 = blockIdx.x * 128;
        auto load = [&](int j) {
            const int s = j & 3;
            mbar_arrive_expect_tx(bar_full + s, 1024);
            tma_load_2d(ring + s * 256, &p.in, c0, L - 2 - 2 * j, bar_full + s);
        };
        for (int j = 0; j < 4 && j < npairs; ++j) load(j);
        for (int j = 0; j < npairs; ++j) {
            const int s = j & 3;
            mbar_wait(bar_done + s, (j / 4) & 1);
            tma_store_2d(&p.out, c0, L - 2 - 2 * j, ring + s * 256);
            tma_commit();
            tma_wait_read<1>();
            if (j >= 1 && j - 1 + 4 < npairs) load(j - 1 + 4);
        }
        tma_wait_all();
        return;
    }
    for (int j = 0; j < npairs; ++j) {
        float *sl = ring + (j & 3) * 256 + tid;
        mbar_wait(bar_full + (j & 3), (j / 4) & 1);
        sl[0] += 1.0f; sl[128] += 1.0f;
        fence_proxy_async();
        mbar_arrive(bar_done + (j & 3));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
    void *fp = nullptr; cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q);
    printf("entry %d %d %p\n", (int)e, (int)q, fp);
    EncodeTiledFn enc = (EncodeTiledFn)fp;
    const int ncol = 960; const int nlev = argc > 1 ? atoi(argv[1]) : 137;
    float *in, *out, *dbg;
    cudaMalloc(&in, ncol * nlev * 4); cudaMalloc(&out, ncol * nlev * 4); cudaMalloc(&dbg, 8 * 256 * 4);
    std::vector<float> h(ncol * nlev);
    for (int i = 0; i < ncol * nlev; ++i) h[i] = (float)i;
    cudaMemcpy(in, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(out, 0, ncol * nlev * 4);
    P p;
    for (int k = 0; k < 2; ++k) {
        cuuint64_t dims[2] = {(cuuint64_t)ncol, (cuuint64_t)nlev};
        cuuint64_t strides[1] = {(cuuint64_t)ncol * 4};
        cuuint32_t box[2] = {128, 2}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(k ? &p.out : &p.in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, k ? out : in, dims, strides, box, es,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("encode %d -> %d\n", k, (int)r);
    }
    cudaFuncSetAttribute(probe<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4096);
    probe<0><<<8, 128, 4096>>>(p, dbg); printf("mode0: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    probe<1><<<8, 128, 4096>>>(p, dbg); printf("mode1: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    probe<2><<<8, 128, 4096>>>(p, dbg); printf("mode2: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    probe<3><<<8, 128, 4096>>>(p, dbg); printf("mode3: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    cudaMemset(out, 0, ncol * nlev * 4);
    cudaFuncSetAttribute(pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    pipe<<<8, 160, 8192>>>(p, nlev); printf("pipe: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    std::vector<float> o(ncol * nlev), d(8 * 256);
    cudaMemcpy(o.data(), out, o.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(d.data(), dbg, d.size() * 4, cudaMemcpyDeviceToHost);
    printf("dbg[0]=%g (want %d) dbg[128]=%g (want %d) dbg last block col 100: %g\n", d[0], 3 * ncol, d[128], 4 * ncol, d[7 * 256 + 100]);
    printf("out[3*ncol]=%g out[4*ncol+5]=%g out[2*ncol]=%g out[3*ncol+959]=%g\n", o[3 * ncol], o[4 * ncol + 5], o[2 * ncol], o[3 * ncol + 959]);
    int bad = 0;
    for (int i = 0; i < ncol * nlev; ++i) if (o[i] != h[i] + 1.0f) { if (bad < 5) printf("bad %d: %g\n", i, o[i]); ++bad; }
    printf("pipe mismatches: %d\n", bad);
    return 0;
}
