// dsc_k4.cuh -- K4: per-grid min/max -> [0,1] (to_tensor01, tiny_sc_ped.py:54-60).
// Included inside the anonymous namespace of dsc_k1_build.cu.
#pragma once

// ---- K4: to_tensor01 (tiny_sc_ped.py:54-60) ---------------------------------------------------
__global__ void k4_init_minmax(uint32_t *mm, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { mm[2 * i] = 0xffffffffu; mm[2 * i + 1] = 0u; }
}

// key order that also orders -0 < +0 consistently; NaNs are not expected in grids
__global__ void __launch_bounds__(256)
k4_minmax(const float *__restrict__ grids, int64_t cells, uint32_t *__restrict__ mm)
{
    const int g = blockIdx.y;
    const float *p = grids + (int64_t)g * cells;
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells;
         i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t k = dsc::float_to_key(p[i]);
        lo = min(lo, k);
        hi = max(hi, k);
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    hi = __reduce_max_sync(0xffffffffu, hi);
    __shared__ uint32_t s_lo[8], s_hi[8];
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_lo[w] = lo; s_hi[w] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < 8; ++i) { lo = min(lo, s_lo[i]); hi = max(hi, s_hi[i]); }
        atomicMin(mm + 2 * g, lo);
        atomicMax(mm + 2 * g + 1, hi);
    }
}

__device__ __forceinline__ float key_to_float_raw(uint32_t k)
{
    return dsc::u2f((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(256)
k4_apply(const float *__restrict__ grids, int64_t cells, const uint32_t *__restrict__ mm,
         float *__restrict__ out)
{
    const int g = blockIdx.y;
    const float mn = key_to_float_raw(mm[2 * g]), mx = key_to_float_raw(mm[2 * g + 1]);
    const float *p = grids + (int64_t)g * cells;
    float *o = out + (int64_t)g * cells;
    const bool scale = mx > mn;
    // (mx - mn) is a Python-float (f64) subtraction in the reference, then a weak f32 scalar
    const float d = (float)__dsub_rn((double)mx, (double)mn);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cells;
         i += (int64_t)gridDim.x * blockDim.x) {
        const float x = p[i];
        o[i] = scale ? __fdiv_rn(__fsub_rn(x, mn), d) : x;
    }
}

// K4, one pass over HBM: a persistent CTA owns one grid at a time -- pass 1 reads it (128-bit loads) and reduces
// min / max, pass 2 writes the normalised grid with streaming stores.  The head of the grid -- as much as fits the
// CTA's shared memory, about half of a default grid -- is kept there between the passes; the tail, read last and so
// the part most likely to be still in L2, is re-read from L2 (evict-last on the way in).  Round 1 re-read the whole
// grid from L2 and missed on 40 % of the lines (1.40x the algorithmic reads).
constexpr int K4_THREADS = 1024;

__device__ __forceinline__ float4 k4_ld(const float4 *p, uint64_t pol)
{
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}

__global__ void __launch_bounds__(K4_THREADS, 1)
k4_fused(const float *__restrict__ grids, int n, int nvec, float *__restrict__ out, const int keep_vec)
{
    extern __shared__ __align__(16) float4 k4_keep[];     // the first keep_vec 128-bit groups of the grid in hand
    __shared__ uint32_t s_lo[32], s_hi[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // pass 1 parks the grid in L2 (evict last), pass 2 lets go of it (evict first)
    uint64_t pol_keep, pol_drop;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
    for (int g = blockIdx.x; g < n; g += gridDim.x) {
        const float4 *p = reinterpret_cast<const float4 *>(grids) + (size_t)g * nvec;
        float4 *o = reinterpret_cast<float4 *>(out) + (size_t)g * nvec;
        uint32_t lo = 0xffffffffu, hi = 0u;
        constexpr int UN = 4;
        for (int i0 = tid; i0 < nvec; i0 += K4_THREADS * UN) {
            float4 v[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u)
                if (i0 + u * K4_THREADS < nvec) v[u] = k4_ld(p + i0 + u * K4_THREADS, pol_keep);
#pragma unroll
            for (int u = 0; u < UN; ++u)
                if (i0 + u * K4_THREADS < nvec) {
                    if (i0 + u * K4_THREADS < keep_vec) k4_keep[i0 + u * K4_THREADS] = v[u];
                    const uint32_t a = dsc::float_to_key(v[u].x), b = dsc::float_to_key(v[u].y);
                    const uint32_t c = dsc::float_to_key(v[u].z), d = dsc::float_to_key(v[u].w);
                    lo = min(min(lo, min(a, b)), min(c, d));
                    hi = max(max(hi, max(a, b)), max(c, d));
                }
        }
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        if (lane == 0) { s_lo[w] = lo; s_hi[w] = hi; }
        __syncthreads();
        lo = s_lo[lane];
        hi = s_hi[lane];
        lo = __reduce_min_sync(0xffffffffu, lo);
        hi = __reduce_max_sync(0xffffffffu, hi);
        __syncthreads();                         // s_lo / s_hi are rewritten by the next grid
        const float mn = key_to_float_raw(lo), mx = key_to_float_raw(hi);
        const bool scale = mx > mn;
        // (mx - mn) is a Python-float (f64) subtraction in the reference, then a weak f32 scalar
        const float dd = (float)__dsub_rn((double)mx, (double)mn);
        for (int i0 = tid; i0 < nvec; i0 += K4_THREADS * UN) {
            float4 v[UN];
#pragma unroll
            for (int u = 0; u < UN; ++u) {
                const int i = i0 + u * K4_THREADS;
                if (i < keep_vec) v[u] = k4_keep[i];     // (written by this very thread in pass 1)
                else if (i < nvec) v[u] = k4_ld(p + i, pol_drop);
            }
#pragma unroll
            for (int u = 0; u < UN; ++u)
                if (i0 + u * K4_THREADS < nvec) {
                    float4 r = v[u];
                    if (scale) {
                        r.x = __fdiv_rn(__fsub_rn(r.x, mn), dd);
                        r.y = __fdiv_rn(__fsub_rn(r.y, mn), dd);
                        r.z = __fdiv_rn(__fsub_rn(r.z, mn), dd);
                        r.w = __fdiv_rn(__fsub_rn(r.w, mn), dd);
                    }
                    __stcs(o + i0 + u * K4_THREADS, r);
                }
        }
    }
}
