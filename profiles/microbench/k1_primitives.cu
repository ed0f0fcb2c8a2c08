// k1_primitives.cu -- design-space microbenchmarks for the scan-context build kernel (K1).
//
// Measures, on the box, the primitive rates that decide where the per-scan ring x sector key grid
// should live and how points should reach it:
//   T1  streaming LDG.128 read bandwidth (upper bound for a read-dominated kernel)
//   T2  shared-memory scattered LDS + atomicMax rates (1 CTA/SM, 115 KB grid band)
//   T3  the same through DSMEM in a 4-CTA cluster (ld/red.shared::cluster to a random peer)
//   T4  global (L2-resident) scattered ld.cg + red.max rates on a per-CTA 460 KB key grid
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o k1_primitives k1_primitives.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
namespace cg = cooperative_groups;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

// ---- T1 ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) t1_stream(const float4 *__restrict__ p, size_t n, float *sink)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 3 * stride < n; i += 4 * stride) {
        float4 a, b, c, d;
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p + i));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + i + stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(c.x), "=f"(c.y), "=f"(c.z), "=f"(c.w) : "l"(p + i + 2 * stride));
        asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(d.x), "=f"(d.y), "=f"(d.z), "=f"(d.w) : "l"(p + i + 3 * stride));
        acc += a.x + b.y + c.z + d.w;
    }
    if (acc == 1234.5f) *sink = acc;
}

// ---- T2: shared memory -----------------------------------------------------------------------------
// mode 0: LDS only; 1: unconditional atomicMax; 2: LDS pre-check + atomicMax when larger (realistic keys)
template <int MODE>
__global__ void __launch_bounds__(1024) t2_smem(int iters, int cells, uint32_t *sink)
{
    extern __shared__ uint32_t grid[];
    for (int i = threadIdx.x; i < cells; i += blockDim.x) grid[i] = 0;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t h = hash32(blockIdx.x * 1024 + threadIdx.x + 1);
    for (int it = 0; it < iters; ++it) {
        h = hash32(h + it);
        const uint32_t cell = (uint32_t)(((uint64_t)h * (uint32_t)cells) >> 32);
        const uint32_t key = hash32(h ^ 0x9e3779b9u) | 0x80000000u;
        if (MODE == 0) acc += grid[cell];
        if (MODE == 1) atomicMax(grid + cell, key);
        if (MODE == 2) { if (key > grid[cell]) atomicMax(grid + cell, key); }
    }
    __syncthreads();
    if (threadIdx.x == 0 && (acc == 77u || grid[5] == 1u)) *sink = acc;
}

// ---- T3: DSMEM -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ld_cluster(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void red_max_cluster(uint32_t addr, uint32_t v)
{
    asm volatile("red.relaxed.cluster.shared::cluster.max.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
// mode 1: unconditional remote red; 2: remote pre-check read + red; 3: pre-check only.  `local_frac256`
// of the accesses (out of 256) go to the CTA's own band.
template <int MODE, int CL>
__global__ void __launch_bounds__(1024) t3_dsmem(int iters, int cells, int local_frac256, uint32_t *sink)
{
    extern __shared__ uint32_t grid[];
    cg::cluster_group cluster = cg::this_cluster();
    const uint32_t my = cluster.block_rank();
    for (int i = threadIdx.x; i < cells; i += blockDim.x) grid[i] = 0;
    cluster.sync();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(grid);
    uint32_t acc = 0;
    uint32_t h = hash32(blockIdx.x * 1024 + threadIdx.x + 1);
    for (int it = 0; it < iters; ++it) {
        h = hash32(h + it);
        const uint32_t cell = (uint32_t)(((uint64_t)h * (uint32_t)cells) >> 32);
        const uint32_t key = hash32(h ^ 0x9e3779b9u) | 0x80000000u;
        uint32_t peer = (h >> 8) % CL;
        if ((int)(h & 255u) < local_frac256) peer = my;
        const uint32_t a = mapa(base + cell * 4, peer);
        if (MODE == 1) red_max_cluster(a, key);
        if (MODE == 2) { if (key > ld_cluster(a)) red_max_cluster(a, key); }
        if (MODE == 3) acc += ld_cluster(a);
    }
    cluster.sync();
    if (threadIdx.x == 0 && (acc == 77u || grid[5] == 1u)) *sink = acc;
}

// ---- T4: global / L2 -------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(1024) t4_global(int iters, int cells, uint32_t *grids, uint32_t *sink)
{
    uint32_t *grid = grids + (size_t)blockIdx.x * cells;
    uint32_t acc = 0;
    uint32_t h = hash32(blockIdx.x * 1024 + threadIdx.x + 1);
    for (int it = 0; it < iters; ++it) {
        h = hash32(h + it);
        const uint32_t cell = (uint32_t)(((uint64_t)h * (uint32_t)cells) >> 32);
        const uint32_t key = hash32(h ^ 0x9e3779b9u) | 0x80000000u;
        if (MODE == 0) acc += __ldcg(grid + cell);
        if (MODE == 1) atomicMax(grid + cell, key);
        if (MODE == 2) { if (key > __ldcg(grid + cell)) atomicMax(grid + cell, key); }
    }
    if (threadIdx.x == 0 && acc == 77u) *sink = acc;
}

template <typename F>
static float time_ms(F launch, int reps = 5)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        launch();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device %s  SMs %d  L2 %d MB  smem/block optin %zu  clock %d MHz\n", prop.name, sms, prop.l2CacheSize >> 20,
           prop.sharedMemPerBlockOptin, prop.clockRate / 1000);
    uint32_t *sink; CK(cudaMalloc(&sink, 4));

    // T1
    {
        const size_t bytes = (size_t)4 << 30;
        float4 *p; CK(cudaMalloc(&p, bytes)); CK(cudaMemset(p, 1, bytes));
        for (int bps : {2, 4, 8}) {
            float ms = time_ms([&] { t1_stream<<<sms * bps, 256>>>(p, bytes / 16, (float *)sink); });
            printf("T1 stream LDG.128  %d CTA/SM x256thr x4 in flight: %.1f GB/s\n", bps, bytes / ms / 1e6);
        }
        CK(cudaFree(p));
    }
    const int cells = 40 * 720;   // one quarter of the 160x720 grid = 115 200 B
    const int iters = 2000;
    const double ops = (double)sms * 1024 * iters;
    // T2
    {
        const size_t smem = (size_t)cells * 4;
        CK(cudaFuncSetAttribute(t2_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(t2_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        CK(cudaFuncSetAttribute(t2_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        float m0 = time_ms([&] { t2_smem<0><<<sms, 1024, smem>>>(iters, cells, sink); });
        float m1 = time_ms([&] { t2_smem<1><<<sms, 1024, smem>>>(iters, cells, sink); });
        float m2 = time_ms([&] { t2_smem<2><<<sms, 1024, smem>>>(iters, cells, sink); });
        printf("T2 smem  LDS scattered            : %.1f Gop/s chip (%.2f cyc/lane/SM @%d MHz)\n", ops / m0 / 1e6, m0 * 1e-3 * prop.clockRate * 1e3 / (1024.0 * iters), prop.clockRate / 1000);
        printf("T2 smem  atomicMax scattered      : %.1f Gop/s chip\n", ops / m1 / 1e6);
        printf("T2 smem  LDS precheck + atomicMax : %.1f Gop/s chip\n", ops / m2 / 1e6);
    }
    // T3
    {
        const size_t smem = (size_t)cells * 4;
        auto run = [&](auto kern, int local256, const char *name) {
            CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            cudaLaunchConfig_t cfg = {};
            const int ctas = (sms / 4) * 4;
            cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int maxc = 0;
            CK(cudaOccupancyMaxActiveClusters(&maxc, kern, &cfg));
            float ms = time_ms([&] { CK(cudaLaunchKernelEx(&cfg, kern, iters, cells, local256, sink)); });
            printf("T3 dsmem cl4 %-28s local %3d/256: %.1f Gop/s chip (grid %d CTAs, max active clusters %d)\n", name, local256,
                   (double)ctas * 1024 * iters / ms / 1e6, ctas, maxc);
        };
        for (int l : {64, 256}) {
            run(t3_dsmem<1, 4>, l, "red.max");
            run(t3_dsmem<2, 4>, l, "ld precheck + red.max");
            run(t3_dsmem<3, 4>, l, "ld only");
        }
    }
    // T4
    {
        const int gcells = 160 * 720;
        uint32_t *grids; CK(cudaMalloc(&grids, (size_t)sms * gcells * 4)); CK(cudaMemset(grids, 0, (size_t)sms * gcells * 4));
        float m0 = time_ms([&] { t4_global<0><<<sms, 1024>>>(iters, gcells, grids, sink); });
        float m1 = time_ms([&] { t4_global<1><<<sms, 1024>>>(iters, gcells, grids, sink); });
        float m2 = time_ms([&] { t4_global<2><<<sms, 1024>>>(iters, gcells, grids, sink); });
        printf("T4 global(L2, %d MB of key grids) ld.cg scattered        : %.1f Gop/s chip\n", (int)((size_t)sms * gcells * 4 >> 20), ops / m0 / 1e6);
        printf("T4 global red.max scattered                              : %.1f Gop/s chip\n", ops / m1 / 1e6);
        printf("T4 global ld.cg precheck + red.max                       : %.1f Gop/s chip\n", ops / m2 / 1e6);
        CK(cudaFree(grids));
    }
    return 0;
}
