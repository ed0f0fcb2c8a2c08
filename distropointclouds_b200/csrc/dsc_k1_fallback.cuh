// dsc_k1_fallback.cuh -- the generic K1 form: k1_build_keys + k1_decode (RED.MAX on an L2-resident key grid, then a decode
// pass).  Used for what k1_scan does not take: 12-byte point stride, sector counts that are not a multiple of 4,
// unaligned buffers, grids whose rows do not fit shared memory.  The round-1 A/B kernels (CTA-pair half grids, the tile
// scheduler with a shared ring of L2 slots, the TMA-ring point loader) were measured slower than k1_scan and have
// been removed from the library; their numbers are in profiles/r1_notes.md.
// Included inside the anonymous namespace of dsc_k1_build.cu, after dsc_k1_scan.cuh.
#pragma once

constexpr int K1_THREADS = 256;
constexpr int K1_PPT = 8;                       // points per thread per tile
constexpr int K1_TILE = K1_THREADS * K1_PPT;    // 2048 points = 32 KB of stream per CTA

struct StatAcc {
    unsigned nonfinite = 0, overflow = 0, clamped = 0, oor = 0, slow = 0;
};

__device__ __forceinline__ void flush_stats(const StatAcc &a, unsigned long long *st, int64_t n_in_tile,
                                            bool first_tile_thread)
{
    // st layout == dsc_scan_stats: n_points, n_nonfinite, n_sector_overflow, n_clamped_ring,
    // n_out_of_range, n_slow_path
    unsigned v0 = __reduce_add_sync(0xffffffffu, a.nonfinite);
    unsigned v1 = __reduce_add_sync(0xffffffffu, a.overflow);
    unsigned v2 = __reduce_add_sync(0xffffffffu, a.clamped);
    unsigned v3 = __reduce_add_sync(0xffffffffu, a.oor);
    unsigned v4 = __reduce_add_sync(0xffffffffu, a.slow);
    if ((threadIdx.x & 31) == 0) {
        if (v4) atomicAdd(st + 5, (unsigned long long)v4);
        if (v0) atomicAdd(st + 1, (unsigned long long)v0);
        if (v1) atomicAdd(st + 2, (unsigned long long)v1);
        if (v2) atomicAdd(st + 3, (unsigned long long)v2);
        if (v3) atomicAdd(st + 4, (unsigned long long)v3);
    }
    if (first_tile_thread) atomicAdd(st + 0, (unsigned long long)n_in_tile);
}

template <bool FAST>
__device__ __forceinline__ void accumulate_point(const BinCfg &cfg, float x, float y, float z,
                                                 uint32_t *grid, StatAcc &acc)
{
    BinResult b;
    if constexpr (FAST) {
        bool slow;
        b = dsc::bin_screened(cfg, x, y, z, 0.0f, 0.0f, &slow);
        if (slow) {            // too close to a bin boundary (or not finite): redo with the fdlibm chain
            b = dsc::bin_exact(cfg, x, y, z);
            acc.slow++;
        }
    } else {
        b = dsc::bin_exact(cfg, x, y, z);
    }
    if (b.status > dsc::PT_CLAMPED) {
        acc.nonfinite += (b.status == dsc::PT_NONFINITE);
        acc.overflow += (b.status == dsc::PT_SECTOR_OVERFLOW);
        acc.oor += (b.status == dsc::PT_OUT_OF_RANGE);
        return;
    }
    acc.clamped += (b.status == dsc::PT_CLAMPED);
    const float zv = b.z;
    if (zv != zv) return;                          // NaN never wins (scaner_version2.py:107)
    if (!cfg.full_height && !(zv > 0.0f)) return;  // offset mode: cells start at 0, strict '>'
    const uint32_t key = dsc::float_to_key(zv);
    if (key <= 0x007fffffu) return;                // -inf can never beat the -inf init
    atomicMax(grid + b.cell, key);
}

// One CTA per (tile, scan): 2048 points -> RED.MAX on the scan's key grid.
template <bool STRIDE16, bool FAST>
__global__ void __launch_bounds__(K1_THREADS)
k1_build_keys(const uint8_t *__restrict__ pts, const int64_t *__restrict__ offsets, int scan_base,
              uint32_t *__restrict__ keys, unsigned long long *__restrict__ stats, int stride_bytes,
              BinCfg cfg)
{
    const int scan_local = blockIdx.y;
    const int scan = scan_base + scan_local;
    const int64_t p0 = offsets[scan];
    const int64_t n = offsets[scan + 1] - p0;
    const int64_t start = (int64_t)blockIdx.x * K1_TILE;
    if (start >= n) return;
    const int64_t cells = (int64_t)cfg.R * cfg.S;
    uint32_t *grid = keys + (int64_t)scan_local * cells;
    const int64_t remain = n - start;
    const int in_tile = remain < K1_TILE ? (int)remain : K1_TILE;
    StatAcc acc;
    if constexpr (STRIDE16) {
        const uint8_t *base = pts + (p0 + start) * 16;
        float4 v[K1_PPT];
#pragma unroll
        for (int k = 0; k < K1_PPT; ++k) {
            const int i = k * K1_THREADS + threadIdx.x;
            if (i < in_tile) v[k] = ld_stream_f4(base + (int64_t)i * 16);
        }
#pragma unroll
        for (int k = 0; k < K1_PPT; ++k) {
            const int i = k * K1_THREADS + threadIdx.x;
            if (i < in_tile) accumulate_point<FAST>(cfg, v[k].x, v[k].y, v[k].z, grid, acc);
        }
    } else {
        const uint8_t *base = pts + (p0 + start) * (int64_t)stride_bytes;
        for (int k = 0; k < K1_PPT; ++k) {
            const int i = k * K1_THREADS + threadIdx.x;
            if (i < in_tile) {
                const uint8_t *q = base + (int64_t)i * stride_bytes;
                accumulate_point<FAST>(cfg, ld_stream_f1(q), ld_stream_f1(q + 4), ld_stream_f1(q + 8), grid, acc);
            }
        }
    }
    if (stats) flush_stats(acc, stats + (int64_t)scan * DSC_STATS_WORDS, in_tile, threadIdx.x == 0);
}

// keys -> float grid (coalesced 128-bit), and re-zero the scratch for the next chunk.
__global__ void __launch_bounds__(256)
k1_decode(uint32_t *__restrict__ keys, float *__restrict__ out, int64_t n_vec4)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec4; i += stride) {
        uint4 k = reinterpret_cast<uint4 *>(keys)[i];
        float4 f;
        f.x = dsc::key_to_float(k.x);
        f.y = dsc::key_to_float(k.y);
        f.z = dsc::key_to_float(k.z);
        f.w = dsc::key_to_float(k.w);
        __stcs(reinterpret_cast<float4 *>(out) + i, f);
        reinterpret_cast<uint4 *>(keys)[i] = make_uint4(0, 0, 0, 0);
    }
}
__global__ void __launch_bounds__(256)
k1_decode_scalar(uint32_t *__restrict__ keys, float *__restrict__ out, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        out[i] = dsc::key_to_float(keys[i]);
        keys[i] = 0;
    }
}
