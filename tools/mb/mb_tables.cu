// mb_tables.cu -- microbenchmarks that decide where stage F's hash tables live (tools only, not product):
// random 4-byte loads + atomicMax on (A) own shared memory, (B) distributed shared memory of a cluster,
// (C) an L2-resident global region; (D) random 16-byte loads from an L2-resident window.
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
namespace cg = cooperative_groups;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t rng(uint32_t& s) { s ^= s << 13; s ^= s >> 17; s ^= s << 5; return s; }

// mode 0: loads only, 1: atomicMax only, 2: load + atomicMax (the finder's mix)
template <int CLUSTER>
__global__ void dsmem_kernel(uint32_t wordsPerCta, int iters, int mode, uint32_t* sink) {
    extern __shared__ uint32_t sm[];
    cg::cluster_group cl = cg::this_cluster();
    for (uint32_t i = threadIdx.x; i < wordsPerCta; i += blockDim.x) sm[i] = i;
    cl.sync();
    uint32_t* base[CLUSTER];
#pragma unroll
    for (int r = 0; r < CLUSTER; r++) base[r] = cl.map_shared_rank(sm, r);
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    const uint32_t mask = wordsPerCta - 1;
    for (int it = 0; it < iters; it++) {
        uint32_t a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) a[k] = rng(s);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t* p;
            if (CLUSTER == 1) p = sm + (a[k] & mask);
            else { const uint32_t r = (a[k] >> 20) % CLUSTER; p = base[0]; 
#pragma unroll
                for (int q = 1; q < CLUSTER; q++) if (r == q) p = base[q];
                p += (a[k] & mask); }
            if (mode == 0 || mode == 2) acc += *p;
            if (mode == 1 || mode == 2) atomicMax(p, a[k] >> 3);
        }
    }
    cl.sync();
    if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void l2_kernel(uint32_t* tab, uint32_t mask, int iters, int mode, uint32_t* sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) a[k] = rng(s);
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint32_t* p = tab + (a[k] & mask);
            if (mode == 0 || mode == 2) acc += __ldcg(p);
            if (mode == 1 || mode == 2) atomicMax(p, a[k] >> 3);
        }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

__global__ void l2_vec_kernel(const uint4* tab, uint32_t mask, int iters, uint32_t* sink) {
    uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) * 2654435761u + 12345u, acc = 0;
    for (int it = 0; it < iters; it++) {
        uint32_t a[4];
#pragma unroll
        for (int k = 0; k < 4; k++) a[k] = rng(s);
#pragma unroll
        for (int k = 0; k < 4; k++) { const uint4 v = __ldg(tab + (a[k] & mask)); acc += v.x ^ v.w; }
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

template <int CLUSTER>
static void run_dsmem(int threads, uint32_t wordsPerCta, int iters, int mode, uint32_t* sink, int nSM) {
    auto k = dsmem_kernel<CLUSTER>;
    const size_t smem = (size_t)wordsPerCta * 4;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (CLUSTER > 8) CK(cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    int grid = (nSM / CLUSTER) * CLUSTER;
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = CLUSTER; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int maxClusters = 0; cudaOccupancyMaxActiveClusters(&maxClusters, k, &cfg);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    CK(cudaLaunchKernelEx(&cfg, k, wordsPerCta, 10, mode, sink)); CK(cudaDeviceSynchronize());
    cudaEventRecord(e0);
    CK(cudaLaunchKernelEx(&cfg, k, wordsPerCta, iters, mode, sink));
    cudaEventRecord(e1); CK(cudaDeviceSynchronize());
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double ops = (double)grid * threads * iters * 4 * (mode == 2 ? 2 : 1);
    printf("dsmem cluster=%d threads=%d KB/cta=%zu mode=%d grid=%d maxActiveClusters=%d: %.3f ms  %.1f Gop/s  %.2f op/clk/SM(@1.9GHz, %d SMs)\n", CLUSTER, threads, smem >> 10, mode, grid, maxClusters, ms, ops / ms * 1e-6, ops / (ms * 1e-3) / 1.9e9 / grid, grid);
}

int main() {
    cudaDeviceProp pr; CK(cudaGetDeviceProperties(&pr, 0));
    const int nSM = pr.multiProcessorCount;
    printf("%s, %d SMs, L2 %d MB\n", pr.name, nSM, pr.l2CacheSize >> 20);
    uint32_t* sink; CK(cudaMalloc(&sink, 4));
    const uint32_t W = 49152 / 1 ;   // words per CTA: 192 KB
    for (int mode = 0; mode < 3; mode++) {
        run_dsmem<1>(1024, 32768, 2000, mode, sink, nSM);
        run_dsmem<2>(1024, 32768, 2000, mode, sink, nSM);
        run_dsmem<4>(1024, 32768, 2000, mode, sink, nSM);
        run_dsmem<8>(1024, 32768, 2000, mode, sink, nSM);
    }
    run_dsmem<4>(512, 32768, 2000, 2, sink, nSM);
    run_dsmem<4>(256, 32768, 4000, 2, sink, nSM);
    (void)W;
    // L2-resident tables
    for (int mb : {16, 48, 96, 192}) {
        uint32_t words = (uint32_t)mb << 18; uint32_t pw = 1; while (pw * 2 <= words) pw *= 2;   // power of two <= size
        uint32_t* tab; CK(cudaMalloc(&tab, (size_t)words * 4)); CK(cudaMemset(tab, 0, (size_t)words * 4));
        for (int mode = 0; mode < 3; mode++) {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            const int grid = nSM * 2, threads = 1024, iters = 500;
            l2_kernel<<<grid, threads>>>(tab, pw - 1, 20, mode, sink); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); l2_kernel<<<grid, threads>>>(tab, pw - 1, iters, mode, sink); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double ops = (double)grid * threads * iters * 4 * (mode == 2 ? 2 : 1);
            printf("L2 table %u MB (pow2 %u MB) mode=%d: %.3f ms  %.1f Gop/s\n", mb, pw >> 18, mode, ms, ops / ms * 1e-6);
        }
        {
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            const int grid = nSM * 2, threads = 1024, iters = 500;
            l2_vec_kernel<<<grid, threads>>>((const uint4*)tab, (pw >> 2) - 1, 20, sink); CK(cudaDeviceSynchronize());
            cudaEventRecord(e0); l2_vec_kernel<<<grid, threads>>>((const uint4*)tab, (pw >> 2) - 1, iters, sink); cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            const double ops = (double)grid * threads * iters * 4;
            printf("L2 window %u MB random 16-byte loads: %.3f ms  %.1f Gop/s\n", pw >> 18, ms, ops / ms * 1e-6);
        }
        cudaFree(tab);
    }
    return 0;
}
