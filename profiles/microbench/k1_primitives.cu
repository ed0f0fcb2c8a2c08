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
#include "../../distropointclouds_b200/csrc/dsc_math.cuh"
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


// ---- T5: does the ORDER in which 148 persistent CTAs walk the point stream matter to HBM? ---------------
// Same loads per thread as k1_scan (768 threads, 3 x 128-bit per set, two sets), no compute.
//   mode 0: every CTA owns whole "scans" (region_bytes each, handed out by an atomic cursor) -- 148 private fronts
//   mode 1: CTAs take consecutive 36 KB tiles of one shared front (atomic cursor per tile)
template <int MODE>
__global__ void __launch_bounds__(768, 1) t5_fronts(const float4 *__restrict__ p, size_t n_vec, size_t region_vec,
                                                    unsigned int *cursor, float *sink)
{
    constexpr int NT = 768, U = 3, STEP = NT * U;
    __shared__ unsigned s_item;
    float acc = 0.f;
    const size_t n_items = MODE == 0 ? n_vec / region_vec : n_vec / (2 * STEP);
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(cursor, 1u);
        __syncthreads();
        const size_t item = s_item;
        if (item >= n_items) break;
        const float4 *q = p + (MODE == 0 ? item * region_vec : item * (size_t)(2 * STEP)) + threadIdx.x;
        const int steps = MODE == 0 ? (int)(region_vec / STEP) : 2;
        for (int s = 0; s + 1 < steps + 1 && s < steps; s += 2) {
            float4 a[U], b[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w) : "l"(q + (size_t)s * STEP + u * NT));
            if (s + 1 < steps) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b[u].x), "=f"(b[u].y), "=f"(b[u].z), "=f"(b[u].w) : "l"(q + (size_t)(s + 1) * STEP + u * NT));
#pragma unroll
                for (int u = 0; u < U; ++u) acc += b[u].y;
            }
#pragma unroll
            for (int u = 0; u < U; ++u) acc += a[u].x;
        }
    }
    if (acc == 1234.5f) *sink = acc;
}

// ---- T6: what does each stage of the per-point work cost at k1_scan's launch shape? -----------------------
// Same loads as T5 mode 0 over plausible points (x, y uniform in +-25 m), then
//   mode 0: the screened binning only (cell index folded into a dummy accumulator)
//   mode 1: + LDS pre-check on a 62-row shared-memory key grid (cells outside it are skipped)
//   mode 2: + ATOMS.MAX when the key is larger (the resident-row path of k1_scan, no L2 rows, no decode)
__global__ void t6_fill(float4 *p, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint32_t h = hash32((uint32_t)i * 2654435761u + 17u), g = hash32(h ^ 0x1234567u), k = hash32(g + 99u);
        p[i] = make_float4((h * (1.0f / 4294967296.0f) - 0.5f) * 50.0f, (g * (1.0f / 4294967296.0f) - 0.5f) * 50.0f,
                           (k * (1.0f / 4294967296.0f)) * 3.0f - 2.0f, 0.5f);
    }
}

//   mode 3: + RED.MAX into a private L2 slot for the cells outside the resident rows
//   mode 4: + the warp vote per point that k1_scan uses to park failed-screen points (none are processed here)
//   mode 5: + per-scan decode: resident rows and L2 slot -> output grid (1.89 GB more HBM writes), re-zero, barriers
template <int MODE>
__global__ void __launch_bounds__(768, 1) t6_stages(const float4 *__restrict__ p, size_t n_vec, size_t region_vec,
                                                    unsigned int *cursor, float *sink, const dsc::BinCfg cfg,
                                                    uint32_t *__restrict__ slots, float *__restrict__ out)
{
    constexpr int NT = 768, U = 3, STEP = NT * U, NS = 62, S = 720;
    extern __shared__ uint32_t rows[];
    __shared__ unsigned s_item;
    for (int i = threadIdx.x; i < NS * S; i += NT) rows[i] = 0;
    float acc = 0.f;
    unsigned qcount = 0;
    uint32_t *slot = slots + (size_t)blockIdx.x * (size_t)((160 - NS) * S);
    const size_t n_items = n_vec / region_vec;
    auto point = [&](const float4 v) {
        int ring, sector; bool beyond;
        const bool ok = dsc::screened_core_t<true>(cfg, v.x, v.y, 0.0f, 0.0f, &ring, &sector, &beyond);
        const int cell = ring * S + sector;
        if (MODE == 0) { acc += ok ? (float)cell : 0.f; return; }
        const uint32_t u = __float_as_uint(v.z);
        const uint32_t key = u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);
        if (ok && cell < NS * S) {
            const uint32_t old = rows[cell];
            if (MODE == 1) acc += (float)(old & 1u);
            if (MODE >= 2) { if (key > old) atomicMax(rows + cell, key); }
        } else if (MODE >= 3 && ok) {
            asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(slot + (cell - NS * S)), "r"(key) : "memory");
        }
        if (MODE >= 4) {
            const unsigned m = __ballot_sync(0xffffffffu, !ok);
            if (m) qcount += __popc(m);
        }
    };
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_item = atomicAdd(cursor, 1u);
        __syncthreads();
        const size_t item = s_item;
        if (item >= n_items) break;
        const float4 *q = p + item * region_vec + threadIdx.x;
        const int steps = (int)(region_vec / STEP);
        float4 a[U], b[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w) : "l"(q + u * NT));
        int s = 0;
        while (s + 1 < steps) {
#pragma unroll
            for (int u = 0; u < U; ++u)
                asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b[u].x), "=f"(b[u].y), "=f"(b[u].z), "=f"(b[u].w) : "l"(q + (size_t)(s + 1) * STEP + u * NT));
#pragma unroll
            for (int u = 0; u < U; ++u) point(a[u]);
            s += 2;
            if (s < steps) {
#pragma unroll
                for (int u = 0; u < U; ++u)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a[u].x), "=f"(a[u].y), "=f"(a[u].z), "=f"(a[u].w) : "l"(q + (size_t)s * STEP + u * NT));
            }
#pragma unroll
            for (int u = 0; u < U; ++u) point(b[u]);
        }
        if (s < steps) {
#pragma unroll
            for (int u = 0; u < U; ++u) point(a[u]);
        }
        if (MODE >= 5) {
            __threadfence();
            __syncthreads();
            float4 *oq = reinterpret_cast<float4 *>(out + item * (size_t)(160 * S));
            for (int i = threadIdx.x; i < NS * S / 4; i += NT) {
                const uint4 k4 = reinterpret_cast<uint4 *>(rows)[i];
                reinterpret_cast<uint4 *>(rows)[i] = make_uint4(0, 0, 0, 0);
                __stcs(oq + i, make_float4(dsc::key_to_float(k4.x), dsc::key_to_float(k4.y), dsc::key_to_float(k4.z), dsc::key_to_float(k4.w)));
            }
            const int nv = (160 - NS) * S / 4;
            uint4 *kq = reinterpret_cast<uint4 *>(slot);
            for (int i0 = threadIdx.x; i0 < nv; i0 += NT * 4) {
                uint4 k4[4];
#pragma unroll
                for (int d = 0; d < 4; ++d) if (i0 + d * NT < nv) k4[d] = __ldcg(kq + i0 + d * NT);
#pragma unroll
                for (int d = 0; d < 4; ++d) if (i0 + d * NT < nv) {
                    __stcs(oq + NS * S / 4 + i0 + d * NT, make_float4(dsc::key_to_float(k4[d].x), dsc::key_to_float(k4[d].y), dsc::key_to_float(k4[d].z), dsc::key_to_float(k4[d].w)));
                    __stcg(kq + i0 + d * NT, make_uint4(0, 0, 0, 0));
                }
            }
            __threadfence();
        }
    }
    if (acc == 1234.5f || qcount == 0xffffffffu) *sink = acc;
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
    // T5
    {
        const size_t region_vec = 120000;                       // one 120 000-point scan = 1.92 MB
        const size_t n_vec = region_vec * 4096;                 // 7.86 GB
        float4 *buf; unsigned int *cur; float *sink5;
        CK(cudaMalloc(&buf, n_vec * 16)); CK(cudaMalloc(&cur, 4)); CK(cudaMalloc(&sink5, 4));
        CK(cudaMemset(buf, 0, n_vec * 16));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaMemset(cur, 0, 4));
                CK(cudaEventRecord(e0));
                if (mode == 0) t5_fronts<0><<<148, 768>>>(buf, n_vec, region_vec, cur, sink5);
                else t5_fronts<1><<<148, 768>>>(buf, n_vec, region_vec, cur, sink5);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            printf("T5 148 CTAs x 768 thr, 2 x 3 LDG.128 per thread, %s: %.1f GB/s (%.3f ms for 7.86 GB)\n",
                   mode == 0 ? "148 private fronts (one 1.92 MB scan per CTA at a time)" : "one shared front (36 KB tiles in order)     ",
                   n_vec * 16.0 / best / 1e6, best);
        }
        CK(cudaFree(buf)); CK(cudaFree(cur)); CK(cudaFree(sink5));
    }
    // T6
    {
        const size_t region_vec = 120000 / 2304 * 2304;         // whole steps only
        const size_t n_vec = region_vec * 4096;
        float4 *buf; unsigned int *cur; float *sink6;
        CK(cudaMalloc(&buf, n_vec * 16)); CK(cudaMalloc(&cur, 4)); CK(cudaMalloc(&sink6, 4));
        t6_fill<<<148 * 8, 256>>>(buf, n_vec);
        CK(cudaDeviceSynchronize());
        const dsc::BinCfg cfg = dsc::make_bin_cfg(160, 720, 20.0, 0, 1, 2.0f, 0, 0);
        const int smem = 62 * 720 * 4;
        CK(cudaFuncSetAttribute(t6_stages<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(t6_stages<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(t6_stages<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        uint32_t *slots; float *outg;
        CK(cudaMalloc(&slots, (size_t)148 * 98 * 720 * 4)); CK(cudaMemset(slots, 0, (size_t)148 * 98 * 720 * 4));
        CK(cudaMalloc(&outg, (size_t)4096 * 160 * 720 * 4));
        CK(cudaFuncSetAttribute(t6_stages<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(t6_stages<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        CK(cudaFuncSetAttribute(t6_stages<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const char *names[6] = {"loads + screened binning", "+ LDS pre-check (62 resident rows)", "+ ATOMS.MAX when larger",
                                "+ RED.MAX to a private L2 slot", "+ warp vote per point", "+ per-scan decode and grid write-out"};
        for (int mode = 0; mode < 6; ++mode) {
            float best = 1e9f;
            for (int rep = 0; rep < 5; ++rep) {
                CK(cudaMemset(cur, 0, 4));
                CK(cudaEventRecord(e0));
                if (mode == 0) t6_stages<0><<<148, 768, smem>>>(buf, n_vec, region_vec, cur, sink6, cfg, slots, outg);
                else if (mode == 1) t6_stages<1><<<148, 768, smem>>>(buf, n_vec, region_vec, cur, sink6, cfg, slots, outg);
                else if (mode == 2) t6_stages<2><<<148, 768, smem>>>(buf, n_vec, region_vec, cur, sink6, cfg, slots, outg);
                else if (mode == 3) t6_stages<3><<<148, 768, smem>>>(buf, n_vec, region_vec, cur, sink6, cfg, slots, outg);
                else if (mode == 4) t6_stages<4><<<148, 768, smem>>>(buf, n_vec, region_vec, cur, sink6, cfg, slots, outg);
                else t6_stages<5><<<148, 768, smem>>>(buf, n_vec, region_vec, cur, sink6, cfg, slots, outg);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
                if (ms < best) best = ms;
            }
            printf("T6 %-36s: %.3f ms for %.2f GB = %.1f GB/s, %.1f Gpts/s\n", names[mode], best, n_vec * 16.0 / 1e9,
                   n_vec * 16.0 / best / 1e6, n_vec / best / 1e6);
        }
        CK(cudaFree(buf)); CK(cudaFree(cur)); CK(cudaFree(sink6));
    }
    return 0;
}
