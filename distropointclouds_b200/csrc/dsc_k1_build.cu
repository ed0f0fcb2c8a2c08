// dsc_k1_build.cu -- K1: packed point stream -> scan-context grids (+ K4 normalise).
//
// Replaces ScanContext.load_velo_scan + ptcloud2sc_vectorized (scaner_version2.py:57-115) and
// ScanContextComputer._scan_context (scan_context_node.py:102-157).
//
// Data layout in HBM
//   pts      packed float32 [x,y,z,i] rows, 16 B each (the .bin layout), all scans back to back;
//            read exactly once with 128-bit streaming loads
//   keys     scratch uint32 [slots][R*S]: per-cell max of the order-preserving key of z
//            (0 == empty).  Lives in L2 between the build and decode kernels of a chunk.
//   out      float32 [n][R][S]: written exactly once, 128-bit coalesced stores
//
// Algorithmic bytes per scan: 16*N + 4*R*S  (SURVEY 8d).
#include "dsc_internal.h"

namespace {

using dsc::BinCfg;
using dsc::BinResult;

constexpr int K1_THREADS = 256;
constexpr int K1_PPT = 8;                       // points per thread per tile
constexpr int K1_TILE = K1_THREADS * K1_PPT;    // 2048 points = 32 KB of stream per CTA

__device__ __forceinline__ float4 ld_stream_f4(const void *p)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ float ld_stream_f1(const void *p)
{
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}

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
    if (stats) flush_stats(acc, stats + (int64_t)scan * 6, in_tile, threadIdx.x == 0);
}


// ---- K1 main path: shared-memory-resident half grids -----------------------------------------------
// The 160x720 float32 grid (460 800 B) does not fit one SM's shared memory, DSMEM atomics are ~15x
// slower than local ones on B200 (profiles/microbench: 77 vs 1120 Gop/s) and L2 atomics ~6x slower.
// So a scan is owned by a PAIR of CTAs that split the grid by sector: CTA `half` keeps the keys of
// sectors [half*S/2, (half+1)*S/2) for all rings in shared memory (160*360*4 = 230 400 B).  Both CTAs
// stream all points of the scan (the second read hits L2, HBM sees each point once) and a point is
// claimed by the CTA whose half contains its sector.  Because S/2 sectors are exactly 180 degrees,
// "which half" is sign(y) except within a hair of the seam; both CTAs evaluate seam points and the
// computed sector decides.  There is no traffic between the CTAs at all: reduction = local ATOMS.MAX,
// write-out = one coalesced pass per scan that also re-zeroes the grid for the next scan.

// L2 eviction policies: the point stream is read once (evict first); the key slots are the hot working
// set that must stay resident between a scan's REDs and its decode (evict last).
__device__ __forceinline__ uint64_t l2_policy_evict_first()
{
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_stream_f4_hint(const void *p, uint64_t pol)
{
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void red_max_global_hint(uint32_t *p, uint32_t v, uint64_t pol)
{
    asm volatile("red.relaxed.gpu.global.max.L2::cache_hint.u32 [%0], %1, %2;" ::"l"(p), "r"(v), "l"(pol) : "memory");
}
// fire-and-forget reduction at L2: no return value, nothing for the warp to wait on
__device__ __forceinline__ void red_max_global(uint32_t *p, uint32_t v)
{
    asm volatile("red.relaxed.gpu.global.max.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void reds_max_u32(uint32_t addr, uint32_t v)
{
    asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

// The fdlibm chain is ~150 instructions and runs for < 1 % of the points: keep it out of line so the
// streaming loop stays small (instruction cache) and its registers do not count against the hot path.
// Its configuration is parked in shared memory once per CTA (a BinCfg reference would force the
// caller's copy into local memory, by-value arguments cost registers at every call site); the result
// is packed: x = status << 27 | cell, y = bits of z.
struct ExactArgs {
    int R, S, pow2, gap_f32;
    float inv_gap_ring_f, inv_gap_sector_f, gap_ring_f, gap_sector_f, max_length_f, offset;
    double gap_ring_d, gap_sector_d;
};
template <int VARIANT, bool FULL>
__device__ __noinline__ uint2 hg_bin_exact(float x, float y, float z, const ExactArgs *__restrict__ ap)
{
    const ExactArgs a = *ap;                     // lives in shared memory, filled once per CTA
    BinCfg c;
    c.R = a.R; c.S = a.S; c.variant = VARIANT; c.full_height = FULL; c.offset = a.offset;
    c.pow2 = a.pow2; c.gap_f32 = a.gap_f32;
    c.inv_gap_ring_f = a.inv_gap_ring_f; c.inv_gap_sector_f = a.inv_gap_sector_f;
    c.gap_ring_f = a.gap_ring_f; c.gap_sector_f = a.gap_sector_f;
    c.gap_ring_d = a.gap_ring_d; c.gap_sector_d = a.gap_sector_d;
    c.max_length_f = a.max_length_f;
    const BinResult b = dsc::bin_exact(c, x, y, z);
    return make_uint2(((uint32_t)b.status << 27) | (uint32_t)b.cell, __float_as_uint(b.z));
}

__device__ __forceinline__ void fill_exact_args(ExactArgs *ea, const BinCfg &cfg)
{
    ea->R = cfg.R; ea->S = cfg.S; ea->pow2 = cfg.pow2; ea->gap_f32 = cfg.gap_f32;
    ea->inv_gap_ring_f = cfg.inv_gap_ring_f; ea->inv_gap_sector_f = cfg.inv_gap_sector_f;
    ea->gap_ring_f = cfg.gap_ring_f; ea->gap_sector_f = cfg.gap_sector_f;
    ea->max_length_f = cfg.max_length_f; ea->offset = cfg.offset;
    ea->gap_ring_d = cfg.gap_ring_d; ea->gap_sector_d = cfg.gap_sector_d;
}

// Bulk L2 prefetch (UBLKPF.L2, the TMA engine): shared memory is entirely taken by the key grid, so
// there is no room to stage tiles with cp.async.bulk; instead one thread asks the copy engine to pull
// the chunk `dist` steps ahead into L2, which turns the 128-bit register loads into L2 hits.
__device__ __forceinline__ void bulk_prefetch_l2(const void *p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
// chunk index counted from the start of `scan`; runs over into this CTA's next scan
__device__ __forceinline__ void hg_prefetch(const uint8_t *pts, const int64_t *offsets, int scan, int npairs,
                                            int n_scans, int64_t p0, int n, int chunk, int chunk_bytes)
{
    int nch = (int)(((int64_t)n * 16 + chunk_bytes - 1) / chunk_bytes);
    if (chunk >= nch) {
        scan += npairs;
        if (scan >= n_scans) return;
        chunk -= nch;
        p0 = offsets[scan];
        n = (int)(offsets[scan + 1] - p0);
        nch = (int)(((int64_t)n * 16 + chunk_bytes - 1) / chunk_bytes);
        if (chunk >= nch) return;
    }
    const int64_t off = (int64_t)chunk * chunk_bytes;
    const int64_t left = (int64_t)n * 16 - off;
    bulk_prefetch_l2(pts + p0 * 16 + off, (uint32_t)(left < chunk_bytes ? left : chunk_bytes));
}

struct HgConst {
    uint32_t sgrid;      // shared-window address of the key grid
    int Sh;              // sectors per half
    int sector0;         // first sector of this CTA's half
    float sgn;           // +1 for half 0 (y > 0 side), -1 for half 1
    int half;
};

template <bool FAST, int VARIANT, bool FULL>
__device__ __forceinline__ void hg_point(const BinCfg &cfg, const HgConst &k, const ExactArgs *ea, const float4 v,
                                         StatAcc &acc)
{
    // ---- whose point is it?  S/2 sectors are exactly 180 degrees, so it is sign(y) except within a
    // hair of the seam, where both CTAs evaluate the point and the computed sector decides.  Variant A
    // substitutes y == 0 by +0.001 (scaner_version2.py:86).  NaNs compare false => claimed by both.
    float ys = v.y;
    if (VARIANT == 0 && ys == 0.0f) ys = 0.001f;
    const float c = fabsf(v.x) * 0.001953125f;   // seam band: |angle to the 0/180 axis| < 0.11 deg
    if (ys * k.sgn < -c) return;
    int ring, sector;
    bool beyond, ok = false;
    if constexpr (FAST) {
        ok = dsc::screened_core(cfg, v.x, v.y, 0.0f, 0.0f, &ring, &sector, &beyond);
        if (VARIANT == 1) ok = ok && (fabsf(v.z) < INFINITY);    // finite mask (scan_context_node.py:115)
    }
    float zv = FULL ? v.z : __fadd_rn(v.z, cfg.offset);
    bool dropped = false;
    int drop_status = 0;
    if (!ok) {                                   // near a bin boundary, zero / non-finite coordinate, ...
        const uint2 b = hg_bin_exact<VARIANT, FULL>(v.x, v.y, v.z, ea);
        const int status = (int)(b.x >> 27), cell = (int)(b.x & 0x7ffffffu);
        ring = cell / cfg.S;
        sector = cell - ring * cfg.S;
        zv = __uint_as_float(b.y);
        beyond = status == dsc::PT_CLAMPED;
        dropped = status > dsc::PT_CLAMPED;
        drop_status = status;
        if (FAST) acc.slow++;
    } else if (VARIANT == 1 && beyond) {
        dropped = true;
        drop_status = dsc::PT_OUT_OF_RANGE;
    }
    if (dropped) {                               // no sector: count once -- by half 0 if it claims the point
        const bool claim0 = !(ys < -c);
        if (k.half == 0 || !claim0) {
            acc.nonfinite += (drop_status == dsc::PT_NONFINITE);
            acc.overflow += (drop_status == dsc::PT_SECTOR_OVERFLOW);
            acc.oor += (drop_status == dsc::PT_OUT_OF_RANGE);
        } else if (FAST && !ok) {
            acc.slow--;                          // the partner accounts for this point
        }
        return;
    }
    const unsigned local = (unsigned)(sector - k.sector0);
    if (local >= (unsigned)k.Sh) {               // seam point that belongs to the partner
        if (FAST && !ok) acc.slow--;
        return;
    }
    if (VARIANT == 0 && beyond) acc.clamped++;
    const uint32_t u = __float_as_uint(zv);
    uint32_t key = u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);     // order-preserving key
    if (zv != zv) key = 0;                       // NaN never wins (scaner_version2.py:107)
    if (!FULL && !(zv > 0.0f)) key = 0;          // offset mode: cells start at 0, strict '>'
    const uint32_t addr = k.sgrid + (uint32_t)(ring * k.Sh + (int)local) * 4u;
    if (key > lds_u32(addr)) reds_max_u32(addr, key);   // LDS pre-check: most points are not a new maximum
}

template <bool FAST, int VARIANT, bool FULL, int THREADS, int U>
__global__ void __launch_bounds__(THREADS, 1)
k1_halfgrid(const uint8_t *__restrict__ pts, const int64_t *__restrict__ offsets, int n_scans,
            float *__restrict__ out, unsigned long long *__restrict__ stats, const int pf_dist,
            const BinCfg cfg_in)
{
    constexpr int STEP = THREADS * U;           // points per step and register set
    extern __shared__ __align__(16) uint32_t sgrid[];
    __shared__ ExactArgs s_ea;
    const BinCfg &cfg = cfg_in;                 // stays in the constant bank: its address is never taken
    const int tid = threadIdx.x;
    if (tid == 0) fill_exact_args(&s_ea, cfg);
    const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;
    HgConst k;
    k.half = blockIdx.x & 1;
    k.Sh = cfg.S >> 1;
    k.sector0 = k.half * k.Sh;
    k.sgn = k.half ? -1.0f : 1.0f;
    k.sgrid = (uint32_t)__cvta_generic_to_shared(sgrid);
    const int nvec = (cfg.R * k.Sh) >> 2;       // host guarantees Sh % 4 == 0
    const int rowvec = k.Sh >> 2;
    const int dr = THREADS / rowvec, dj = THREADS - dr * rowvec;   // row/column advance per THREADS vec4
    for (int i = tid; i < nvec; i += THREADS) reinterpret_cast<uint4 *>(sgrid)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
    for (int scan = pair; scan < n_scans; scan += npairs) {
        const int64_t p0 = offsets[scan];
        const int n = (int)(offsets[scan + 1] - p0);          // host guarantees < 2^31 points per scan
        const float4 *s4 = reinterpret_cast<const float4 *>(pts + p0 * 16);
        StatAcc acc;
        constexpr int CHUNK = STEP * 16;
        if (pf_dist > 0 && tid == 0 && scan == pair) {       // cold start of this CTA's first scan
            for (int c = k.half; c < pf_dist; c += 2) hg_prefetch(pts, offsets, scan, npairs, n_scans, p0, n, c, CHUNK);
        }
        // ---- stream the scan: two register sets ping-pong, U x 128-bit loads in flight per set.
        // (A rolling "reload the register right after use" scheme was tried: ptxas copies the freshly
        // loaded value at once, which serialises on the load -- see profiles/r1_notes.md.)
        const int nfull = n / STEP;
        const float4 *p = s4 + tid;
        float4 a[U], b[U];
        if (nfull > 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) a[u] = ld_stream_f4(p + u * THREADS);
        }
        int step = 0;
        while (step + 1 < nfull) {
            // the two CTAs of the pair share the prefetching: even chunks by half 0, odd chunks by half 1
            if (pf_dist > 0 && tid == 0)
                hg_prefetch(pts, offsets, scan, npairs, n_scans, p0, n, step + pf_dist + k.half, CHUNK);
#pragma unroll
            for (int u = 0; u < U; ++u) b[u] = ld_stream_f4(p + STEP + u * THREADS);
#pragma unroll
            for (int u = 0; u < U; ++u) hg_point<FAST, VARIANT, FULL>(cfg, k, &s_ea, a[u], acc);
            p += 2 * STEP;
            step += 2;
            if (step < nfull) {
#pragma unroll
                for (int u = 0; u < U; ++u) a[u] = ld_stream_f4(p + u * THREADS);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) hg_point<FAST, VARIANT, FULL>(cfg, k, &s_ea, b[u], acc);
        }
        if (step < nfull) {
#pragma unroll
            for (int u = 0; u < U; ++u) hg_point<FAST, VARIANT, FULL>(cfg, k, &s_ea, a[u], acc);
        }
        for (int i = nfull * STEP + tid; i < n; i += THREADS)              // ragged tail
            hg_point<FAST, VARIANT, FULL>(cfg, k, &s_ea, ld_stream_f4(s4 + i), acc);
        __syncthreads();
        // ---- one coalesced write-out of this CTA's half of every ring row, re-zeroing the keys ----
        float *o = out + (int64_t)scan * cfg.R * cfg.S + k.sector0;
        int r = tid / rowvec, j4 = tid - r * rowvec;
        for (int i = tid; i < nvec; i += THREADS) {
            const uint4 q = reinterpret_cast<uint4 *>(sgrid)[i];
            reinterpret_cast<uint4 *>(sgrid)[i] = make_uint4(0, 0, 0, 0);
            float4 f;
            f.x = dsc::key_to_float(q.x);
            f.y = dsc::key_to_float(q.y);
            f.z = dsc::key_to_float(q.z);
            f.w = dsc::key_to_float(q.w);
            __stcs(reinterpret_cast<float4 *>(o + r * cfg.S) + j4, f);
            r += dr;
            j4 += dj;
            if (j4 >= rowvec) { j4 -= rowvec; ++r; }
        }
        if (stats) flush_stats(acc, stats + (int64_t)scan * 6, n, k.half == 0 && tid == 0);
        __syncthreads();
    }
}


// ---- K1 streaming path: one visit per point, key grid resident in L2 ----------------------------------
// Persistent CTAs pull 2048-point tiles from a global cursor in scan-major order, bin every point once and
// fold it into the scan's key grid with RED.MAX (fire and forget: no return value, no dependent stall).
// The key grids live in a small ring of slots that stays L2-resident (only a handful of scans are in
// flight at any time), so HBM sees each point once and each output cell once.  The CTA that retires the
// last tile of a scan decodes that scan's keys to floats (coalesced 128-bit stores), re-zeroes the slot
// and publishes it for re-use.
constexpr int ST_THREADS = 256;

struct StreamSync {            // zeroed before every launch
    unsigned int cursor32;      // next global tile index
    unsigned int pad[3];        // pad[0]: decode watermark (every scan below it is decoded)
    // followed by: unsigned tiles_done[n_scans]; unsigned decoded[n_scans];
};

// Hybrid placement of the key grid (scan kernel): `ns` ring rows live in shared memory -- rings
// 0 .. ns-2 plus the clamp ring R-1, which together hold 70-90 % of the points of a KITTI-like scan --
// and take the LDS pre-check + ATOMS.MAX path; the remaining rings go to the CTA's L2 slot with RED.MAX.
struct SmemRows {
    uint32_t base;   // shared-window address of the resident rows
    int lim;         // cells below lim are resident rows 0 .. ns-2            (0 when nothing is resident)
    int last;        // first cell of the clamp ring, kept in resident row ns-1 (INT_MAX when nothing is)
    int shift;       // lim - last: clamp-ring cell -> resident row ns-1 (sc_update only)
};

template <bool FAST, int VARIANT, bool FULL, bool DEF = false, bool SMEM = false>
__device__ __forceinline__ void st_point(const BinCfg &cfg, const ExactArgs *ea, const float4 v,
                                         uint32_t *__restrict__ grid, unsigned &n_clamped,
                                         unsigned long long *__restrict__ stats, int scan, uint64_t key_policy = 0,
                                         const SmemRows sm = SmemRows{0u, 0, 0x7fffffff, 0})
{
    int ring = 0, sector = 0;
    bool beyond = false, ok = false;
    if constexpr (FAST) {
        ok = dsc::screened_core_t<DEF>(cfg, v.x, v.y, 0.0f, 0.0f, &ring, &sector, &beyond);
        if (VARIANT == 1) ok = ok && (fabsf(v.z) < INFINITY);    // finite mask (scan_context_node.py:115)
    }
    float zv = FULL ? v.z : __fadd_rn(v.z, cfg.offset);
    if (!ok) {                                   // near a bin boundary, zero / non-finite coordinate, ...
        const uint2 b = hg_bin_exact<VARIANT, FULL>(v.x, v.y, v.z, ea);
        const int status = (int)(b.x >> 27), cell = (int)(b.x & 0x7ffffffu);
        unsigned long long *st = stats ? stats + (int64_t)scan * 6 : nullptr;
        if (st) {                                // rare events: straight to the per-scan counters
            if (FAST) atomicAdd(st + 5, 1ull);
            if (status == dsc::PT_NONFINITE) atomicAdd(st + 1, 1ull);
            if (status == dsc::PT_SECTOR_OVERFLOW) atomicAdd(st + 2, 1ull);
            if (status == dsc::PT_OUT_OF_RANGE) atomicAdd(st + 4, 1ull);
        }
        if (status > dsc::PT_CLAMPED) return;
        n_clamped += (status == dsc::PT_CLAMPED);
        zv = __uint_as_float(b.y);
        ring = 0;
        sector = cell;                           // ring * S + sector below collapses to `cell`
    } else {
        if (VARIANT == 1 && beyond) {
            if (stats) atomicAdd(stats + (int64_t)scan * 6 + 4, 1ull);
            return;
        }
        if (VARIANT == 0) n_clamped += beyond;
    }
    const uint32_t u = __float_as_uint(zv);
    const uint32_t key = u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);     // order-preserving key
    if (zv != zv) return;                        // NaN never wins (scaner_version2.py:107)
    if (!FULL && !(zv > 0.0f)) return;           // offset mode: cells start at 0, strict '>'
    const int S = DEF ? dsc::defgrid::S : cfg.S;
    const int cell = ring * S + sector;           // slow path: ring == 0 and sector already holds the cell
    if constexpr (SMEM) {
        if (cell < sm.lim || cell >= sm.last) {
            const int local = cell < sm.lim ? cell : cell - sm.last + sm.lim;   // clamp ring -> row ns-1
            const uint32_t addr = sm.base + (uint32_t)local * 4u;
            if (key > lds_u32(addr)) reds_max_u32(addr, key);   // LDS pre-check: most points are not a new maximum
            return;
        }
    }
    if (key_policy) red_max_global_hint(grid + cell, key, key_policy);
    else red_max_global(grid + cell, key);
}

// keys (L2) -> float grid (HBM, coalesced 128-bit streaming stores); re-zeroes the slot.  Out of line so
// that its 8-deep load pipeline does not share the streaming loop's register budget.
__device__ __noinline__ void st_decode(uint32_t *__restrict__ grid, float *__restrict__ out, int nvec, int tid)
{
    const uint4 *kq = reinterpret_cast<const uint4 *>(grid);
    uint4 *kz = reinterpret_cast<uint4 *>(grid);
    float4 *oq = reinterpret_cast<float4 *>(out);
    constexpr int DU = 8;                        // 8 x 128-bit L2 loads in flight per thread
    for (int i0 = tid; i0 < nvec; i0 += ST_THREADS * DU) {
        uint4 q[DU];
#pragma unroll
        for (int d = 0; d < DU; ++d) {
            const int i = i0 + d * ST_THREADS;
            if (i < nvec) q[d] = __ldcg(kq + i);  // L2: where the other CTAs' REDs landed
        }
#pragma unroll
        for (int d = 0; d < DU; ++d) {
            const int i = i0 + d * ST_THREADS;
            if (i < nvec) {
                float4 f;
                f.x = dsc::key_to_float(q[d].x);
                f.y = dsc::key_to_float(q[d].y);
                f.z = dsc::key_to_float(q[d].z);
                f.w = dsc::key_to_float(q[d].w);
                __stcs(oq + i, f);
                __stcg(kz + i, make_uint4(0, 0, 0, 0));
            }
        }
    }
}

// slot re-use guard: scan `need` must be decoded and its slot re-zeroed before the slot is written again.
// The watermark (every scan below it is decoded) is far ahead in practice and cached per CTA; the decoder
// of `need` is a running CTA, so the wait terminates.
__device__ __forceinline__ void st_wait_slot(StreamSync *sync, const unsigned *decoded, int need, unsigned &water)
{
    if (need < 0 || (unsigned)need < water) return;
    for (;;) {
        water = *reinterpret_cast<volatile unsigned *>(&sync->pad[0]);
        if ((unsigned)need < water) break;
        if (*reinterpret_cast<const volatile unsigned *>(decoded + need) != 0u) break;
        __nanosleep(64);
    }
    __threadfence();
}

// tile -> (scan, tile in scan)
__device__ __forceinline__ void st_locate(int g, const int64_t *__restrict__ tile_base, int n_scans,
                                          int tiles_uniform, int &scan, int &tile)
{
    if (tiles_uniform > 0) {
        scan = g / tiles_uniform;
        tile = g - scan * tiles_uniform;
    } else {                                     // ragged batch: tile_base[s] = first global tile of scan s
        int lo = 0, hi = n_scans;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (tile_base[mid] <= g) lo = mid; else hi = mid;
        }
        scan = lo;
        tile = (int)(g - tile_base[lo]);
    }
}

template <bool FAST, int VARIANT, bool FULL, int ST_PPT, int MINB>
__global__ void __launch_bounds__(ST_THREADS, MINB)
k1_stream(const uint8_t *__restrict__ pts, const int64_t *__restrict__ offsets,
          const int64_t *__restrict__ tile_base, int n_scans, int total_tiles, int tiles_uniform,
          uint32_t *__restrict__ keys, int n_slots, float *__restrict__ out,
          unsigned long long *__restrict__ stats, StreamSync *__restrict__ sync, const BinCfg cfg)
{
    constexpr int ST_TILE = ST_THREADS * ST_PPT;
    __shared__ int s_tile, s_scan, s_last;
    __shared__ ExactArgs s_ea;
    const int tid = threadIdx.x;
    const int cells = cfg.R * cfg.S;
    if (tid == 0) fill_exact_args(&s_ea, cfg);
    unsigned *tiles_done = reinterpret_cast<unsigned *>(sync + 1);
    unsigned *decoded = tiles_done + n_scans;
    // thread 0 is the CTA's scheduler; `g` is always fetched one tile ahead so that the round trip of
    // the cursor atomic hides behind the current tile, `water` caches the decode watermark
    int g = 0;
    unsigned water = 0;
    if (tid == 0) {
        g = (int)atomicAdd(&sync->cursor32, 1u);
        s_last = 0;
        int scan = 0, tile = -1;
        if (g < total_tiles) {
            st_locate(g, tile_base, n_scans, tiles_uniform, scan, tile);
            st_wait_slot(sync, decoded, scan - n_slots, water);
        }
        s_tile = tile;
        s_scan = scan;
        g = (int)atomicAdd(&sync->cursor32, 1u);
    }
    __syncthreads();
    for (;;) {
        const int tile = s_tile;
        const int scan = s_scan;
        if (tile < 0) break;
        const int64_t p0 = offsets[scan];
        const int n = (int)(offsets[scan + 1] - p0);
        const int start = tile * ST_TILE;
        const int in_tile = min(n - start, ST_TILE);
        uint32_t *grid = keys + (size_t)(scan % n_slots) * cells;
        const float4 *src = reinterpret_cast<const float4 *>(pts + (p0 + start) * 16) + tid;
        float4 v[ST_PPT];
        unsigned n_clamped = 0;
        if (in_tile == ST_TILE) {
#pragma unroll
            for (int k = 0; k < ST_PPT; ++k) v[k] = ld_stream_f4(src + k * ST_THREADS);
#pragma unroll
            for (int k = 0; k < ST_PPT; ++k) st_point<FAST, VARIANT, FULL>(cfg, &s_ea, v[k], grid, n_clamped, stats, scan);
        } else {
#pragma unroll
            for (int k = 0; k < ST_PPT; ++k)
                if (k * ST_THREADS + tid < in_tile) v[k] = ld_stream_f4(src + k * ST_THREADS);
#pragma unroll
            for (int k = 0; k < ST_PPT; ++k)
                if (k * ST_THREADS + tid < in_tile) st_point<FAST, VARIANT, FULL>(cfg, &s_ea, v[k], grid, n_clamped, stats, scan);
        }
        if (stats) {
            unsigned long long *st = stats + (int64_t)scan * 6;
            const unsigned c = __reduce_add_sync(0xffffffffu, n_clamped);
            if ((tid & 31) == 0 && c) atomicAdd(st + 3, (unsigned long long)c);
            if (tid == 0) atomicAdd(st + 0, (unsigned long long)in_tile);
        }
        // ---- retire the tile and pick the next one (one barrier pair per tile) ----
        __syncthreads();                         // every thread has issued its REDs
        if (tid == 0) {
            __threadfence();                     // release: the CTA's REDs are performed before the count moves
            const unsigned ntiles = (unsigned)((n + ST_TILE - 1) / ST_TILE);
            const unsigned done = atomicAdd(tiles_done + scan, 1u) + 1u;
            s_last = done == ntiles;
            if (done == ntiles) __threadfence(); // acquire: the other CTAs' REDs of this scan are visible
            int nscan = 0, ntile = -1;
            if (done != ntiles) {                // (a decoding CTA picks its next tile after the decode)
                if (g < total_tiles) {
                    st_locate(g, tile_base, n_scans, tiles_uniform, nscan, ntile);
                    st_wait_slot(sync, decoded, nscan - n_slots, water);
                }
                s_tile = ntile;
                s_scan = nscan;
                g = (int)atomicAdd(&sync->cursor32, 1u);
            }
        }
        __syncthreads();
        if (s_last) {
            // ---- this CTA retired the scan's last tile: decode it, re-zero the slot, publish ----
            st_decode(grid, out + (size_t)scan * cells, cells >> 2, tid);
            __syncthreads();                     // all zeroing stores issued
            if (tid == 0) {
                __threadfence();                 // release the re-zeroed slot
                atomicExch(decoded + scan, 1u);
                __threadfence();
                // advance the watermark over every leading decoded scan
                unsigned w = *reinterpret_cast<volatile unsigned *>(&sync->pad[0]);
                while (w < (unsigned)n_scans && *reinterpret_cast<volatile unsigned *>(decoded + w) != 0u) {
                    const unsigned prev = atomicCAS(&sync->pad[0], w, w + 1u);
                    w = prev == w ? w + 1u : prev;
                }
                // now pick this CTA's next tile
                int nscan = 0, ntile = -1;
                if (g < total_tiles) {
                    st_locate(g, tile_base, n_scans, tiles_uniform, nscan, ntile);
                    st_wait_slot(sync, decoded, nscan - n_slots, water);
                }
                s_tile = ntile;
                s_scan = nscan;
                s_last = 0;
                g = (int)atomicAdd(&sync->cursor32, 1u);
            }
            __syncthreads();
        }
    }
}


// ---- K1 main path: one persistent CTA per scan, private L2-resident key slot ------------------------------
// Every CTA owns one key grid in global memory (sm_count x 460 KB = 68 MB, resident in the 126 MB L2) and
// processes whole scans: stream the points once (128-bit loads, two register sets ping-pong), bin each
// point once, fold it into the slot with RED.MAX (fire and forget: no return value, no dependent stall,
// no barrier in the loop).  When the scan is done the same CTA decodes the slot to floats (coalesced
// 128-bit streaming stores), re-zeroes it and moves on.  There is no synchronisation between CTAs at all;
// HBM sees each point once and each output cell once.
constexpr int SC_THREADS = 1024;
#ifndef SC_TOUCH
#define SC_TOUCH 0                    // 1: touched-segment flags for the L2-row decode (measured: slower, see r1_notes.md)
#endif
#ifndef SC_DU
#define SC_DU 4                       // 128-bit L2 loads in flight per thread in the decode
#endif

// SPLIT = true: a scan is cut into `parts` contiguous point ranges (work items) that different CTAs
// process concurrently -- for multi-million-point scans and for batches with fewer scans than SMs, where
// one CTA per scan would leave most of the chip idle.  The scan's own output grid doubles as its key
// grid (zeroed by the host before the launch): L2 rows take RED.MAX straight into it, the resident rows
// are folded into it with RED.MAX when the item ends, and the CTA that retires the scan's last item
// decodes the keys to floats in place.  The grid of a scan in flight is 460 KB and stays in L2.
struct SplitArgs {
    const int64_t *item_base;    // [n_scans + 1] first item of each scan (ragged batches), or nullptr
    unsigned int *done;          // [n_scans] items retired per scan (zeroed before the launch)
    int total_items;
    int parts_uniform;           // > 0: every scan has this many items
};

// ---- deferred exact path ------------------------------------------------------------------------------------
// A point that fails the screen needs the ~200-instruction fdlibm chain.  Only ~0.3 % of the points do, but
// taken inline that is a divergent branch which some lane of a warp takes every ~7th point, so the warp as a
// whole spends a quarter of its issue slots there with one or two lanes active.  Instead every warp parks such
// points in a small shared-memory queue (x[64], y[64], z[64]) and runs the exact chain only when 32 are
// waiting -- all lanes busy -- or when the work item ends.
constexpr int SCQ_CAP = 64;                                   // entries per warp
constexpr int SCQ_WARP_BYTES = SCQ_CAP * 12;                  // 768 B per warp

__device__ __forceinline__ void sts_f32(uint32_t addr, float v)
{
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr)
{
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}

// fold one binned point into the key grid: resident rows -> LDS pre-check + ATOMS.MAX, other rows -> RED.MAX (L2)
// The L2 rows of a scan are sparse (5-25 % of their cells ever see a point), yet the decode has to read every
// key back through a ~1 us round trip.  TOUCH: remember which 16-cell segments received a RED (one byte flag per
// segment in shared memory, plain store), so that the decode reads -- and re-zeroes -- only those.
constexpr int SC_SEG_SHIFT = 4;                               // 16 cells = 64 B = 4 x 128-bit per segment
// The flags sit right below the resident rows, one byte per segment of the WHOLE grid and indexed by the
// absolute cell number, so the hot path needs no extra base register: flag(cell) = rows_base - bytes + (cell >> 4).
__host__ __device__ constexpr uint32_t sc_flag_bytes(int cells) { return ((uint32_t)(cells >> SC_SEG_SHIFT) + 1u + 127u) & ~127u; }

__device__ __forceinline__ void sts_u8(uint32_t addr, uint32_t v)
{
    asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t addr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}

template <bool FULL, bool TOUCH>
__device__ __forceinline__ void sc_update(const float zv, const int cell, uint32_t *__restrict__ grid, const SmemRows sm,
                                          const uint32_t flag_bytes)
{
    const uint32_t u = __float_as_uint(zv);
    const uint32_t key = u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);     // order-preserving key
    if (zv != zv) return;                        // NaN never wins (scaner_version2.py:107)
    if (!FULL && !(zv > 0.0f)) return;           // offset mode: cells start at 0, strict '>'
    if (cell < sm.lim || cell >= sm.last) {
        const int local = cell < sm.lim ? cell : cell + sm.shift;               // clamp ring -> row ns-1
        const uint32_t addr = sm.base + (uint32_t)local * 4u;
        if (key > lds_u32(addr)) reds_max_u32(addr, key);   // most points are not a new maximum
        return;
    }
    red_max_global(grid + cell, key);
    if constexpr (TOUCH) sts_u8(sm.base - flag_bytes + (uint32_t)(cell >> SC_SEG_SHIFT), 1u);
}

// exact chain for one queued point (entry address of this lane); returns 1 if the point was clamped into the last ring
template <int VARIANT, bool FULL, bool TOUCH>
__device__ __noinline__ unsigned sc_exact_queued(uint32_t qaddr, const ExactArgs *__restrict__ ea,
                                                 uint32_t *__restrict__ grid, uint32_t sm_base, int sm_lim, int sm_last,
                                                 uint32_t flag_bytes, unsigned long long *__restrict__ st)
{
    const float x = lds_f32(qaddr), y = lds_f32(qaddr + SCQ_CAP * 4), z = lds_f32(qaddr + SCQ_CAP * 8);
    const uint2 b = hg_bin_exact<VARIANT, FULL>(x, y, z, ea);
    const int status = (int)(b.x >> 27), cell = (int)(b.x & 0x7ffffffu);
    if (st) {                                    // rare events: straight to the per-scan counters
        atomicAdd(st + 5, 1ull);
        if (status == dsc::PT_NONFINITE) atomicAdd(st + 1, 1ull);
        if (status == dsc::PT_SECTOR_OVERFLOW) atomicAdd(st + 2, 1ull);
        if (status == dsc::PT_OUT_OF_RANGE) atomicAdd(st + 4, 1ull);
    }
    if (status > dsc::PT_CLAMPED) return 0u;
    sc_update<FULL, TOUCH>(__uint_as_float(b.y), cell, grid, SmemRows{sm_base, sm_lim, sm_last, sm_lim - sm_last}, flag_bytes);
    return status == dsc::PT_CLAMPED ? 1u : 0u;
}

// One point per lane, called warp-uniformly (`active` masks lanes past the end of the item).
template <int VARIANT, bool FULL, bool DEF, bool TOUCH>
__device__ __forceinline__ void sc_point(const BinCfg &cfg, const ExactArgs *ea, const float4 v, const bool active,
                                         uint32_t *__restrict__ grid, unsigned &n_clamped,
                                         unsigned long long *__restrict__ st, const SmemRows sm,
                                         const uint32_t queue, unsigned &qcount, const uint32_t flag_bytes)
{
    int ring, sector;
    bool beyond;
    bool ok = dsc::screened_core_t<DEF>(cfg, v.x, v.y, 0.0f, 0.0f, &ring, &sector, &beyond);
    if (VARIANT == 1) ok = ok && (fabsf(v.z) < INFINITY);    // finite mask (scan_context_node.py:115)
    if (ok && active) {
        if (VARIANT == 1 && beyond) {
            if (st) atomicAdd(st + 4, 1ull);
        } else {
            if (VARIANT == 0) n_clamped += beyond;
            const int S = DEF ? dsc::defgrid::S : cfg.S;
            sc_update<FULL, TOUCH>(FULL ? v.z : __fadd_rn(v.z, cfg.offset), ring * S + sector, grid, sm, flag_bytes);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, active && !ok);
    if (m) {                                     // warp-uniform, rare
        const unsigned lane = threadIdx.x & 31u;
        if (active && !ok) {
            const uint32_t a = queue + (qcount + __popc(m & ((1u << lane) - 1u))) * 4u;
            sts_f32(a, v.x);
            sts_f32(a + SCQ_CAP * 4, v.y);
            sts_f32(a + SCQ_CAP * 8, v.z);
        }
        qcount += __popc(m);
        if (qcount >= 32u) {
            qcount -= 32u;
            __syncwarp();
            n_clamped += sc_exact_queued<VARIANT, FULL, TOUCH>(queue + (qcount + lane) * 4u, ea, grid, sm.base, sm.lim, sm.last, flag_bytes, st);
            __syncwarp();
        }
    }
}

// end of item: the queue's remaining entries (fewer than 32)
template <int VARIANT, bool FULL, bool TOUCH>
__device__ __forceinline__ void sc_drain(const ExactArgs *ea, uint32_t *__restrict__ grid, unsigned &n_clamped,
                                         unsigned long long *__restrict__ st, const SmemRows sm, const uint32_t queue,
                                         unsigned &qcount, const uint32_t flag_bytes)
{
    if (qcount) {
        const unsigned lane = threadIdx.x & 31u;
        __syncwarp();
        if (lane < qcount)
            n_clamped += sc_exact_queued<VARIANT, FULL, TOUCH>(queue + lane * 4u, ea, grid, sm.base, sm.lim, sm.last, flag_bytes, st);
        __syncwarp();
        qcount = 0;
    }
}

// work item -> (scan, first point, point count, parts of the scan, key grid)
template <bool SPLIT>
__device__ __forceinline__ void sc_item(int item, const int64_t *__restrict__ offsets, int n_scans, const SplitArgs &sp,
                                        int &scan, int &parts, int64_t &p0, int &n)
{
    scan = item;
    parts = 1;
    if constexpr (SPLIT) {
        int part;
        st_locate(item, sp.item_base, n_scans, sp.parts_uniform, scan, part);
        parts = sp.parts_uniform > 0 ? sp.parts_uniform : (int)(sp.item_base[scan + 1] - sp.item_base[scan]);
        const int64_t q0 = offsets[scan];
        const int64_t nn = offsets[scan + 1] - q0;
        const int64_t chunk = (nn + parts - 1) / parts;
        const int64_t start = (int64_t)part * chunk;
        p0 = q0 + (start < nn ? start : nn);
        n = (int)(start < nn ? (nn - start < chunk ? nn - start : chunk) : 0);
    } else {
        p0 = offsets[scan];
        n = (int)(offsets[scan + 1] - p0);       // host guarantees < 2^31 points per scan
    }
}

// End of a work item.  Non-split: this CTA owns the whole scan -- decode the resident rows and the private
// L2 slot to floats (coalesced 128-bit streaming stores) and re-zero both.  Split: fold the resident rows
// into the scan's key grid (== its output grid), count the item as retired, and if it was the scan's last
// item decode the grid in place.  Ends with a CTA barrier in both cases.
template <bool SPLIT, int NT, int DU = 4>
__device__ __forceinline__ void sc_finish_item(uint32_t *__restrict__ srows, uint32_t *__restrict__ grid,
                                               float *__restrict__ out_scan, int ns, int R, int S, int scan, int parts,
                                               unsigned int *__restrict__ done, int *s_last,
                                               unsigned int *__restrict__ reset = nullptr, const uint32_t flags = 0u)
{
    const int tid = threadIdx.x;
    const int rowvec = S >> 2;
    const int cells = R * S;
    const int nres = ns * rowvec;
    const int lim = (ns - 1) * rowvec, shift = (R - ns) * rowvec;    // resident row ns-1 holds ring R-1
    if constexpr (SPLIT) {
        __syncthreads();                         // every shared-memory atomic of the item has landed
        for (int i = tid; i < nres; i += NT) {
            const uint4 q = reinterpret_cast<uint4 *>(srows)[i];
            if ((q.x | q.y | q.z | q.w) == 0u) continue;
            reinterpret_cast<uint4 *>(srows)[i] = make_uint4(0, 0, 0, 0);
            uint32_t *g = grid + 4 * (size_t)(i < lim ? i : i + shift);
            if (q.x) red_max_global(g + 0, q.x);
            if (q.y) red_max_global(g + 1, q.y);
            if (q.z) red_max_global(g + 2, q.z);
            if (q.w) red_max_global(g + 3, q.w);
        }
        __threadfence();                         // release: this CTA's REDs are performed ...
        __syncthreads();
        if (tid == 0) {
            if (reset) *reset = 0u;              // every warp has left the item's streaming loop
            const unsigned prev = atomicAdd(done + scan, 1u);          // ... before the item counts as retired
            *s_last = prev + 1u == (unsigned)parts;
            if (prev + 1u == (unsigned)parts) __threadfence();         // acquire: the other CTAs' REDs are visible
        }
        __syncthreads();
        if (*s_last) {
            uint4 *kq = reinterpret_cast<uint4 *>(grid);
            const int nvec = cells >> 2;
            for (int i0 = tid; i0 < nvec; i0 += NT * DU) {
                uint4 q[DU];
#pragma unroll
                for (int d = 0; d < DU; ++d) {
                    const int i = i0 + d * NT;
                    if (i < nvec) q[d] = __ldcg(kq + i);
                }
#pragma unroll
                for (int d = 0; d < DU; ++d) {
                    const int i = i0 + d * NT;
                    if (i < nvec) {
                        float4 f;
                        f.x = dsc::key_to_float(q[d].x);
                        f.y = dsc::key_to_float(q[d].y);
                        f.z = dsc::key_to_float(q[d].z);
                        f.w = dsc::key_to_float(q[d].w);
                        __stcs(reinterpret_cast<float4 *>(kq) + i, f);
                    }
                }
            }
        }
        __syncthreads();                         // s_last is rewritten by the next item
    } else {
        __threadfence();                         // our REDs are performed ...
        __syncthreads();                         // ... before any thread of the CTA reads the keys back
        if (tid == 0 && reset) *reset = 0u;      // every warp has left the item's streaming loop
        float4 *oq = reinterpret_cast<float4 *>(out_scan);
        for (int i = tid; i < nres; i += NT) {       // resident rows: rings 0 .. ns-2 and the clamp ring
            const uint4 q = reinterpret_cast<uint4 *>(srows)[i];
            reinterpret_cast<uint4 *>(srows)[i] = make_uint4(0, 0, 0, 0);
            float4 f;
            f.x = dsc::key_to_float(q.x);
            f.y = dsc::key_to_float(q.y);
            f.z = dsc::key_to_float(q.z);
            f.w = dsc::key_to_float(q.w);
            __stcs(oq + (i < lim ? i : i + shift), f);
        }
        if (ns < R) {                            // L2 rows: rings ns-1 .. R-2
            const uint4 *kq = reinterpret_cast<const uint4 *>(grid);
            uint4 *kz = reinterpret_cast<uint4 *>(grid);
            const int lo = (ns > 0 ? ns - 1 : 0) * rowvec, hi = (ns > 0 ? R - 1 : R) * rowvec;
            if (flags) {
                // only the segments that received a RED are read back (and re-zeroed); the rest is written as zeros
                for (int i0 = lo + tid; i0 < hi; i0 += NT * DU) {
                    uint4 q[DU];
                    bool t[DU];
#pragma unroll
                    for (int d = 0; d < DU; ++d) {
                        const int i = i0 + d * NT;
                        t[d] = i < hi && lds_u8(flags + (uint32_t)(i >> (SC_SEG_SHIFT - 2))) != 0u;
                        if (t[d]) q[d] = __ldcg(kq + i);
                    }
#pragma unroll
                    for (int d = 0; d < DU; ++d) {
                        const int i = i0 + d * NT;
                        if (i < hi) {
                            float4 f = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
                            if (t[d]) {
                                f.x = dsc::key_to_float(q[d].x);
                                f.y = dsc::key_to_float(q[d].y);
                                f.z = dsc::key_to_float(q[d].z);
                                f.w = dsc::key_to_float(q[d].w);
                                __stcg(kz + i, make_uint4(0, 0, 0, 0));
                            }
                            __stcs(oq + i, f);
                        }
                    }
                }
                __syncthreads();                 // every flag has been read ...
                const int nseg4 = (int)(sc_flag_bytes(R * S) >> 2);                      // flags, in 32-bit words
                for (int i = tid; i < nseg4; i += NT)
                    asm volatile("st.shared.u32 [%0], %1;" ::"r"(flags + (uint32_t)i * 4u), "r"(0u) : "memory");
            } else {
                for (int i0 = lo + tid; i0 < hi; i0 += NT * DU) {
                    uint4 q[DU];
#pragma unroll
                    for (int d = 0; d < DU; ++d) {
                        const int i = i0 + d * NT;
                        if (i < hi) q[d] = __ldcg(kq + i);
                    }
#pragma unroll
                    for (int d = 0; d < DU; ++d) {
                        const int i = i0 + d * NT;
                        if (i < hi) {
                            float4 f;
                            f.x = dsc::key_to_float(q[d].x);
                            f.y = dsc::key_to_float(q[d].y);
                            f.z = dsc::key_to_float(q[d].z);
                            f.w = dsc::key_to_float(q[d].w);
                            __stcs(oq + i, f);
                            __stcg(kz + i, make_uint4(0, 0, 0, 0));
                        }
                    }
                }
            }
        }
        __threadfence();                         // the zeroes are performed ...
        __syncthreads();                         // ... before the next scan's REDs
    }
}

// (A flavour with the resident-row count as a template constant -- row-placement tests and clamp-row shift as
// immediates -- was measured 10 % SLOWER than this one, which takes `ns` at run time: profiles/r1_notes.md.)
template <bool FAST, int VARIANT, bool FULL, int U, int NT, bool DEF, bool SPLIT>
__global__ void __launch_bounds__(NT, 1)
k1_scan(const uint8_t *__restrict__ pts, const int64_t *__restrict__ offsets, int n_scans,
        uint32_t *__restrict__ keys, float *__restrict__ out, unsigned long long *__restrict__ stats,
        unsigned int *__restrict__ cursor, const int ns, const BinCfg cfg, const SplitArgs sp)
{
    // NB: this kernel lives on the edge of its register budget (64 regs at 1024 threads, 80 at 768); measured
    // variants that carried more state across the streaming loop (next-item prefetch before the decode,
    // per-warp chunk counters, 8 decode loads in flight) spilled inside the loop and lost 15-40 %
    // (profiles/r1_notes.md).
    constexpr int STEP = NT * U;
    extern __shared__ __align__(16) uint32_t smem_raw[];
    __shared__ int s_next[2];
    __shared__ int s_last;
    __shared__ ExactArgs s_ea;
    const int tid = threadIdx.x;
    const int R = DEF ? dsc::defgrid::R : cfg.R, S = DEF ? dsc::defgrid::S : cfg.S;
    const int cells = R * S;
    const int rowvec = S >> 2;                   // host guarantees S % 4 == 0
    // private L2 slot of this CTA: only the non-resident rows exist in it; the pointer is biased so that it is still
    // indexed by the absolute cell number
    uint32_t *grid = SPLIT ? nullptr
                           : keys + (size_t)blockIdx.x * (size_t)((ns > 0 ? R - ns : R) * S) - (ns > 0 ? (ns - 1) * S : 0);
    const uint64_t pol_stream = l2_policy_evict_first();
    // dynamic shared memory: [touched-segment flags (TOUCH only)][ns ring rows of keys][per-warp exact-path queues]
    constexpr bool TOUCH = SC_TOUCH && FAST && !SPLIT;       // whole-scan items of the screened kernels
    const uint32_t flag_bytes = TOUCH ? sc_flag_bytes(cells) : 0u;
    uint32_t *const srows = reinterpret_cast<uint32_t *>(reinterpret_cast<uint8_t *>(smem_raw) + flag_bytes);
    SmemRows sm;
    sm.base = (uint32_t)__cvta_generic_to_shared(srows);
    sm.lim = ns > 0 ? (ns - 1) * S : 0;
    sm.last = ns > 0 ? (R - 1) * S : 0x7fffffff;
    sm.shift = sm.lim - sm.last;
    const uint32_t queue = sm.base + (uint32_t)ns * (uint32_t)S * 4u + (uint32_t)(tid >> 5) * SCQ_WARP_BYTES;
    unsigned qcount = 0;
    for (int i = tid; i < (int)(flag_bytes >> 4) + ns * rowvec; i += NT) reinterpret_cast<uint4 *>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) {
        fill_exact_args(&s_ea, cfg);
        s_next[0] = (int)atomicAdd(cursor, 1u);
    }
    __syncthreads();
    const int n_items = SPLIT ? sp.total_items : n_scans;
    int item = s_next[0];
    int par = 0;
    while (item < n_items) {
        // work item after this one: fetched now, consumed after the item (the round trip is hidden)
        if (tid == 0) s_next[par ^ 1] = (int)atomicAdd(cursor, 1u);
        int scan, parts, n;
        int64_t p0;
        sc_item<SPLIT>(item, offsets, n_scans, sp, scan, parts, p0, n);
        if constexpr (SPLIT) grid = reinterpret_cast<uint32_t *>(out + (size_t)scan * cells);
        const float4 *s4 = reinterpret_cast<const float4 *>(pts + p0 * 16);
        unsigned n_clamped = 0;
        unsigned long long *st = stats ? stats + (int64_t)scan * 6 : nullptr;
        const int nfull = n / STEP;
        const float4 *p = s4 + tid;
        float4 a[U], b[U];
#define SC_LD(ptr) ld_stream_f4_hint((ptr), pol_stream)
#define SC_PT(v, act)                                                                                           \
    do {                                                                                                        \
        if constexpr (FAST) sc_point<VARIANT, FULL, DEF, TOUCH>(cfg, &s_ea, (v), (act), grid, n_clamped, st, sm, queue, qcount, flag_bytes); \
        else if (act) st_point<false, VARIANT, FULL, DEF, true>(cfg, &s_ea, (v), grid, n_clamped, stats, scan, 0, sm);   \
    } while (0)
        if (nfull > 0) {
#pragma unroll
            for (int u = 0; u < U; ++u) a[u] = SC_LD(p + u * NT);
        }
        int step = 0;
        while (step + 1 < nfull) {
#pragma unroll
            for (int u = 0; u < U; ++u) b[u] = SC_LD(p + STEP + u * NT);
#pragma unroll
            for (int u = 0; u < U; ++u) SC_PT(a[u], true);
            p += 2 * STEP;
            step += 2;
            if (step < nfull) {
#pragma unroll
                for (int u = 0; u < U; ++u) a[u] = SC_LD(p + u * NT);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) SC_PT(b[u], true);
        }
        if (step < nfull) {
#pragma unroll
            for (int u = 0; u < U; ++u) SC_PT(a[u], true);
        }
        for (int i0 = nfull * STEP; i0 < n; i0 += NT) {       // ragged tail, warp-uniform trip count
            const bool act = i0 + tid < n;
            float4 v = make_float4(1.0f, 1.0f, 0.0f, 0.0f);
            if (act) v = SC_LD(s4 + i0 + tid);
            SC_PT(v, act);
        }
        if constexpr (FAST) sc_drain<VARIANT, FULL, TOUCH>(&s_ea, grid, n_clamped, st, sm, queue, qcount, flag_bytes);
#undef SC_LD
#undef SC_PT
        if (stats) {
            const unsigned c = __reduce_add_sync(0xffffffffu, n_clamped);
            if ((tid & 31) == 0 && c) atomicAdd(st + 3, (unsigned long long)c);
            if (tid == 0 && n) atomicAdd(st + 0, (unsigned long long)n);
        }
        sc_finish_item<SPLIT, NT, SC_DU>(srows, grid, out + (size_t)scan * cells, ns, R, S, scan, parts, sp.done, &s_last,
                                         nullptr, TOUCH ? sm.base - flag_bytes : 0u);
        par ^= 1;
        item = s_next[par];                      // published by the barriers of sc_finish_item
    }
}

// ---- K1 main path, TMA flavour: the point stream is staged by the copy engine -----------------------------
// Same work items, key-grid placement and write-out as k1_scan; the difference is the point loader.  One
// elected thread keeps a ring of NST tiles (1024*U points, 16 KB * U each) in flight with cp.async.bulk
// (global -> shared, completion on an mbarrier, L2 evict-first), so the bytes in flight per SM do not depend
// on how the warps are scheduled and cost no registers; every thread then reads its own point with one
// conflict-free LDS.128.  The ring runs ahead across item boundaries: the first tiles of the next item are
// requested before the current item is decoded, so the decode overlaps the loads.  The loads bypass L1, which
// frees the whole 227 KB of shared memory for grid rows + ring.
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t pol)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 lds_f4(uint32_t addr)
{
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

template <bool FAST, int VARIANT, bool FULL, int U, int NST, bool DEF, bool SPLIT>
__global__ void __launch_bounds__(SC_THREADS, 1)
k1_ring(const uint8_t *__restrict__ pts, const int64_t *__restrict__ offsets, int n_scans,
        uint32_t *__restrict__ keys, float *__restrict__ out, unsigned long long *__restrict__ stats,
        unsigned int *__restrict__ cursor, const int ns, const BinCfg cfg, const SplitArgs sp)
{
    constexpr int TILE = SC_THREADS * U;                       // points per tile
    constexpr uint32_t TILE_BYTES = TILE * 16u;
    extern __shared__ __align__(16) uint32_t srows[];          // [ns ring rows of keys][NST tiles]
    __shared__ __align__(8) unsigned long long s_bar[2 * NST];  // full[NST], empty[NST]
    __shared__ int s_next[2];
    __shared__ int s_last;
    __shared__ ExactArgs s_ea;
    const int tid = threadIdx.x;
    const int R = DEF ? dsc::defgrid::R : cfg.R, S = DEF ? dsc::defgrid::S : cfg.S;
    const int cells = R * S;
    const int rowvec = S >> 2;
    uint32_t *grid = SPLIT ? nullptr : keys + (size_t)blockIdx.x * cells;
    const uint64_t pol_stream = l2_policy_evict_first();
    SmemRows sm;
    sm.base = (uint32_t)__cvta_generic_to_shared(srows);
    sm.lim = ns > 0 ? (ns - 1) * S : 0;
    sm.last = ns > 0 ? (R - 1) * S : 0x7fffffff;
    const uint32_t ring = sm.base + (((uint32_t)ns * (uint32_t)S * 4u + 127u) & ~127u);
    const uint32_t bar_full = (uint32_t)__cvta_generic_to_shared(s_bar), bar_empty = bar_full + 8u * NST;
    for (int i = tid; i < ns * rowvec; i += SC_THREADS) reinterpret_cast<uint4 *>(srows)[i] = make_uint4(0, 0, 0, 0);
    const int n_items = SPLIT ? sp.total_items : n_scans;

    // producer state (thread 0 only): g_prod = tiles requested so far; tile g lives in stage g % NST, its
    // `full` phase is (g / NST) & 1 and it may be requested once tile g - NST has been released
    auto request = [&](unsigned g, const uint8_t *src, uint32_t bytes) {
        const unsigned s = g % NST;
        if (g >= (unsigned)NST) mbar_wait(bar_empty + 8u * s, ((g / NST) - 1u) & 1u);
        mbar_expect_tx(bar_full + 8u * s, bytes);
        bulk_load(ring + s * TILE_BYTES, src, bytes, bar_full + 8u * s, pol_stream);
    };
    // first tiles of an item (at most NST), requested ahead of time; g0 = global index of its tile 0
    auto prologue = [&](int item_, unsigned g0) {
        if (item_ >= n_items) return;
        int scan_, parts_, n_;
        int64_t p0_;
        sc_item<SPLIT>(item_, offsets, n_scans, sp, scan_, parts_, p0_, n_);
        const uint8_t *src = pts + p0_ * 16;
        const int nt_ = (n_ + TILE - 1) / TILE;
        for (int k = 0; k < NST && k < nt_; ++k) {
            const int cnt = min(TILE, n_ - k * TILE);
            request(g0 + (unsigned)k, src + (size_t)k * TILE_BYTES, (uint32_t)cnt * 16u);
        }
    };

    if (tid == 0) {
        fill_exact_args(&s_ea, cfg);
        for (int i = 0; i < NST; ++i) {
            mbar_init(bar_full + 8u * i, 1u);
            mbar_init(bar_empty + 8u * i, SC_THREADS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_next[0] = (int)atomicAdd(cursor, 1u);
    }
    __syncthreads();
    int item = s_next[0];
    int par = 0;
    unsigned g = 0;                              // global tile counter of this CTA (same in every thread)
    if (tid == 0) prologue(item, 0u);
    while (item < n_items) {
        if (tid == 0) s_next[par ^ 1] = (int)atomicAdd(cursor, 1u);
        int scan, parts, n;
        int64_t p0;
        sc_item<SPLIT>(item, offsets, n_scans, sp, scan, parts, p0, n);
        if constexpr (SPLIT) grid = reinterpret_cast<uint32_t *>(out + (size_t)scan * cells);
        const uint8_t *src = pts + p0 * 16;
        const int nt = (n + TILE - 1) / TILE;
        unsigned n_clamped = 0;
#define RG_PT(v) st_point<FAST, VARIANT, FULL, DEF, true>(cfg, &s_ea, (v), grid, n_clamped, stats, scan, 0, sm)
        for (int k = 0; k < nt; ++k, ++g) {
            const unsigned s = g % NST;
            const int cnt = min(TILE, n - k * TILE);
            mbar_wait(bar_full + 8u * s, (g / NST) & 1u);
            const uint32_t base = ring + s * TILE_BYTES + (uint32_t)tid * 16u;
            float4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (u * SC_THREADS + tid < cnt) v[u] = lds_f4(base + (uint32_t)u * (SC_THREADS * 16u));
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(bar_empty + 8u * s);      // this warp has its points in registers
            if (tid == 0 && k + NST < nt) {
                const int cn = min(TILE, n - (k + NST) * TILE);
                request(g + NST, src + (size_t)(k + NST) * TILE_BYTES, (uint32_t)cn * 16u);
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (u * SC_THREADS + tid < cnt) RG_PT(v[u]);
        }
#undef RG_PT
        if (stats) {
            unsigned long long *st = stats + (int64_t)scan * 6;
            const unsigned c = __reduce_add_sync(0xffffffffu, n_clamped);
            if ((tid & 31) == 0 && c) atomicAdd(st + 3, (unsigned long long)c);
            if (tid == 0 && n) atomicAdd(st + 0, (unsigned long long)n);
        }
        // the next item's first tiles go out now, so that they land while this item is decoded
        if (tid == 0) prologue(s_next[par ^ 1], g);
        sc_finish_item<SPLIT, SC_THREADS>(srows, grid, out + (size_t)scan * cells, ns, R, S, scan, parts, sp.done, &s_last);
        par ^= 1;
        item = s_next[par];
    }
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
// min / max, pass 2 re-reads it from L2 (460 KB per CTA, 68 MB chip-wide: resident) and writes the normalised
// grid with streaming stores.  HBM sees each grid once in and once out: the algorithmic 2 * 4 * R * S bytes.
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
k4_fused(const float *__restrict__ grids, int n, int nvec, float *__restrict__ out)
{
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
            for (int u = 0; u < UN; ++u)
                if (i0 + u * K4_THREADS < nvec) v[u] = k4_ld(p + i0 + u * K4_THREADS, pol_drop);
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

}  // namespace

// ---- host side ----------------------------------------------------------------------------------

int dsc_check_launch(dsc_ctx *ctx, const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        ctx->err = std::string(what) + ": " + cudaGetErrorString(e);
        return DSC_ERR_CUDA;
    }
    ctx->launches++;
    return DSC_OK;
}

static int validate_params(dsc_ctx *ctx, const dsc_params *p)
{
    if (!p) DSC_FAIL(ctx, DSC_ERR_INVALID, "params is NULL");
    if (p->num_ring <= 0 || p->num_sector <= 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "num_ring/num_sector must be > 0");
    if ((int64_t)p->num_ring * p->num_sector > (int64_t)1 << 26)
        DSC_FAIL(ctx, DSC_ERR_INVALID, "num_ring*num_sector too large");
    if (!(p->max_length > 0.0) || !isfinite(p->max_length)) DSC_FAIL(ctx, DSC_ERR_INVALID, "max_length must be finite and > 0");
    if (p->variant != DSC_VARIANT_A && p->variant != DSC_VARIANT_B) DSC_FAIL(ctx, DSC_ERR_INVALID, "unknown variant");
    return DSC_OK;
}

static int normalize_device(dsc_ctx *ctx, const float *grids, int n, int64_t cells, float *out,
                            cudaStream_t st)
{
    if (n <= 0) return DSC_OK;
    if ((cells & 3) == 0 && cells / 4 < 0x7fffffff && ((((uintptr_t)grids) | ((uintptr_t)out)) & 15) == 0) {
        const int ctas = n < ctx->sm_count ? n : ctx->sm_count;
        k4_fused<<<ctas, K4_THREADS, 0, st>>>(grids, n, (int)(cells / 4), out);
        return dsc_check_launch(ctx, "k4_fused");
    }
    if ((size_t)n > ctx->minmax_cap) {
        if (ctx->minmax) cudaFree(ctx->minmax);
        ctx->minmax = nullptr;
        ctx->minmax_cap = 0;
        DSC_CUDA(ctx, cudaMalloc(&ctx->minmax, sizeof(uint32_t) * 2 * (size_t)n));
        ctx->minmax_cap = (size_t)n;
    }
    int rc;
    for (int g0 = 0; g0 < n; g0 += 32768) {
        const int gn = n - g0 < 32768 ? n - g0 : 32768;
        uint32_t *mm = ctx->minmax + 2 * (size_t)g0;
        k4_init_minmax<<<(gn + 255) / 256, 256, 0, st>>>(mm, gn);
        if ((rc = dsc_check_launch(ctx, "k4_init_minmax"))) return rc;
        int bx = (int)((cells + 256 * 8 - 1) / (256 * 8));
        if (bx < 1) bx = 1;
        if (bx > 64) bx = 64;
        dim3 grid(bx, gn);
        k4_minmax<<<grid, 256, 0, st>>>(grids + (int64_t)g0 * cells, cells, mm);
        if ((rc = dsc_check_launch(ctx, "k4_minmax"))) return rc;
        k4_apply<<<grid, 256, 0, st>>>(grids + (int64_t)g0 * cells, cells, mm, out + (int64_t)g0 * cells);
        if ((rc = dsc_check_launch(ctx, "k4_apply"))) return rc;
    }
    return DSC_OK;
}

static int dsc_reserve_sync(dsc_ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->sync_cap) return DSC_OK;
    if (ctx->sync_buf) { DSC_CUDA(ctx, cudaDeviceSynchronize()); cudaFree(ctx->sync_buf); }
    ctx->sync_buf = nullptr;
    ctx->sync_cap = 0;
    const size_t cap = bytes * 2 + 4096;
    DSC_CUDA(ctx, cudaMalloc(&ctx->sync_buf, cap));
    ctx->sync_cap = cap;
    return DSC_OK;
}

// key scratch (always left zeroed by the kernels that use it)
static int ensure_keys(dsc_ctx *ctx, size_t need, cudaStream_t st)
{
    if (need <= ctx->keys_bytes) return DSC_OK;
    if (ctx->keys) { DSC_CUDA(ctx, cudaDeviceSynchronize()); cudaFree(ctx->keys); }
    ctx->keys = nullptr;
    ctx->keys_bytes = 0;
    DSC_CUDA(ctx, cudaMalloc(&ctx->keys, need));
    ctx->keys_bytes = need;
    DSC_CUDA(ctx, cudaMemsetAsync(ctx->keys, 0, need, st));
    return DSC_OK;
}

// Device-resident batch: pts/out/stats are device pointers, offsets already on the device.
static int build_device(dsc_ctx *ctx, const dsc_params *p, const uint8_t *pts, const int64_t *d_offsets,
                        const int64_t *h_offsets, int n_scans, int stride, float *out, float *out_norm,
                        dsc_scan_stats *stats, cudaStream_t st)
{
    const int64_t cells = (int64_t)p->num_ring * p->num_sector;
    const BinCfg cfg = dsc::make_bin_cfg(p->num_ring, p->num_sector, p->max_length, p->variant,
                                         p->full_height, p->height_offset,
                                         (p->flags & DSC_FLAG_GAP_F32) != 0,
                                         (p->flags & DSC_FLAG_EXACT_ONLY) != 0);
    const bool vec16 = stride == 16 && (((uintptr_t)pts) & 15) == 0;
    int rc;
    if (stats) DSC_CUDA(ctx, cudaMemsetAsync(stats, 0, sizeof(dsc_scan_stats) * (size_t)n_scans, st));
    // ---- main path: persistent CTAs, one scan (or one slice of a big scan) at a time ----
    // Work items: a scan is cut into `parts` point ranges of about sc_item_points points when one CTA per
    // scan would not fill the chip (few scans, or multi-million-point scans); see k1_scan<SPLIT>.
    int64_t total_pts = 0, max_pts = 0;
    for (int s = 0; s < n_scans; ++s) {
        const int64_t n = h_offsets[s + 1] - h_offsets[s];
        total_pts += n;
        if (n > max_pts) max_pts = n;
    }
    int64_t item_pts = ctx->sc_item_points;
    {   // small batches: shrink the items until every SM gets a few of them (not below 16 Ki points)
        const int64_t want = total_pts / ((int64_t)ctx->sm_count * 4);
        if (want < item_pts) item_pts = want < 16384 ? 16384 : want;
    }
    int64_t split_items = 0;
    for (int s = 0; s < n_scans; ++s) {
        const int64_t n = h_offsets[s + 1] - h_offsets[s];
        split_items += n > item_pts ? (n + item_pts - 1) / item_pts : 1;
    }
    const bool sc_shape_ok = vec16 && (p->num_sector % 4) == 0 && (((uintptr_t)out) & 15) == 0;
    const bool want_split = ctx->sc_split == 1 ||
                            (ctx->sc_split < 0 && (n_scans < 2 * ctx->sm_count || max_pts > 3 * ctx->sc_item_points) &&
                             split_items > n_scans);
    const bool split = sc_shape_ok && want_split && (ctx->k1_impl == 0 || ctx->k1_impl == 3) &&
                       split_items < 0x7fff0000LL;
    const bool sc_ok = sc_shape_ok &&
                       (split || (ctx->k1_impl == 0 ? n_scans >= ctx->sm_count / 2 : ctx->k1_impl == 3));
    if (sc_ok) {
        int ctas = ctx->sm_count;
        const int64_t n_items = split ? split_items : (int64_t)n_scans;
        if (ctas > n_items) ctas = (int)n_items;
        // sync block: [cursor (16 B)] [done[n_scans]] [item_base[n_scans+1] (ragged batches)]
        SplitArgs spa = {nullptr, nullptr, 0, 0};
        size_t sync_bytes = 16;
        if (split) {
            bool uniform = true;
            const int64_t n0 = h_offsets[1] - h_offsets[0];
            for (int s = 1; s < n_scans && uniform; ++s) uniform = (h_offsets[s + 1] - h_offsets[s]) == n0;
            const size_t done_bytes = ((size_t)n_scans * 4 + 15) & ~(size_t)15;
            const size_t tb_bytes = uniform ? 0 : sizeof(int64_t) * ((size_t)n_scans + 1);
            if ((rc = dsc_reserve_sync(ctx, 16 + done_bytes + tb_bytes + 64))) return rc;
            spa.done = (unsigned int *)((uint8_t *)ctx->sync_buf + 16);
            spa.total_items = (int)split_items;
            if (uniform) {
                spa.parts_uniform = (int)(split_items / n_scans);
            } else {
                int64_t *d_tb = (int64_t *)((uint8_t *)ctx->sync_buf + 16 + done_bytes);
                std::vector<int64_t> tb((size_t)n_scans + 1);
                int64_t acc_t = 0;
                for (int s = 0; s < n_scans; ++s) {
                    const int64_t n = h_offsets[s + 1] - h_offsets[s];
                    tb[s] = acc_t;
                    acc_t += n > item_pts ? (n + item_pts - 1) / item_pts : 1;
                }
                tb[n_scans] = acc_t;
                // pageable source: the copy is staged by the driver before the call returns
                DSC_CUDA(ctx, cudaMemcpyAsync(d_tb, tb.data(), tb_bytes, cudaMemcpyHostToDevice, st));
                spa.item_base = d_tb;
            }
            sync_bytes = 16 + done_bytes;
            // the output grids double as the key grids of their scans: start from "empty"
            DSC_CUDA(ctx, cudaMemsetAsync(out, 0, (size_t)n_scans * cells * 4, st));
        } else {
            if ((rc = dsc_reserve_sync(ctx, 256))) return rc;
        }
        unsigned int *cursor = (unsigned int *)ctx->sync_buf;
        DSC_CUDA(ctx, cudaMemsetAsync(cursor, 0, sync_bytes, st));
        unsigned long long *stp = (unsigned long long *)stats;
        const bool defgrid = dsc::is_default_grid(cfg) && ctx->no_defgrid == 0;
        const bool full = p->full_height != 0;
        const size_t row_bytes = (size_t)p->num_sector * 4;
        // ---- TMA-ring loader (k1_ring): default grid + screened path only; everything else uses k1_scan ----
        const int ring_cfg = (cfg.fast_ok && defgrid) ? ctx->sc_loader : 0;
        if (ring_cfg > 0) {
            // 2 points per thread x 3 stages (96 KB); 1 x 4, 1 x 6 and 1 x 3 were measured and are slower (r1_notes.md)
            const size_t ring_bytes = 98304;
            int ns = (int)((ctx->smem_optin - ring_bytes - 1024) / row_bytes);
            if (ctx->sc_smem_rows >= 0 && ctx->sc_smem_rows < ns) ns = ctx->sc_smem_rows;
            if (ns > p->num_ring) ns = p->num_ring;
            if (ns < 2) ns = 0;
            const size_t rg_smem = (((size_t)ns * row_bytes + 127) & ~(size_t)127) + ring_bytes;
            if ((rc = ensure_keys(ctx, split ? 0 : (size_t)ctx->sm_count * cells * 4, st))) return rc;   // full-size slots
#define RG_LAUNCH3(V, H, U, N, SP)                                                                              \
    do {                                                                                                        \
        DSC_CUDA(ctx, cudaFuncSetAttribute(k1_ring<true, V, H, U, N, true, SP>,                                 \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rg_smem));         \
        k1_ring<true, V, H, U, N, true, SP><<<ctas, SC_THREADS, rg_smem, st>>>(pts, d_offsets, n_scans,         \
                                                                               ctx->keys, out, stp, cursor, ns, \
                                                                               cfg, spa);                       \
    } while (0)
#define RG_LAUNCH2(V, H, U, N)                                                                                  \
    do {                                                                                                        \
        if (split) RG_LAUNCH3(V, H, U, N, true); else RG_LAUNCH3(V, H, U, N, false);                            \
    } while (0)
#define RG_LAUNCH(V, H)                                                                                         \
    do {                                                                                                        \
        RG_LAUNCH2(V, H, 2, 3);                                                                            \
    } while (0)
            if (p->variant == 0) { if (full) RG_LAUNCH(0, true); else RG_LAUNCH(0, false); }
            else                 { if (full) RG_LAUNCH(1, true); else RG_LAUNCH(1, false); }
#undef RG_LAUNCH
#undef RG_LAUNCH2
#undef RG_LAUNCH3
            if ((rc = dsc_check_launch(ctx, "k1_ring"))) return rc;
            if (out_norm) return normalize_device(ctx, out, n_scans, cells, out_norm, st);
            return DSC_OK;
        }
        // launch shape (threads x 128-bit loads per thread per register set); shapes other than 0 exist for the
        // screened default-grid kernels only.  Fewer threads = more registers per thread (no spills), fewer warps.
        int shape = (cfg.fast_ok && defgrid) ? ctx->sc_cfg : 2;                 // the generic kernels are 1024 x 2
        if (shape != 1 && shape != 2) shape = 0;
        const int nthreads = shape == 0 ? 768 : 1024;
        const size_t q_bytes = cfg.fast_ok ? (size_t)SCQ_WARP_BYTES * (nthreads / 32) : 0;   // per-warp exact-path queues
        // Resident rows.  Shared memory and L1 share 256 KB per SM and the 128-bit streaming loads stage through
        // L1: beyond the 196 KB carve-out step (164 KB for the 1024-thread shapes) the loader starves
        // (profiles/r1_notes.md: 60 rows 1.99 ms, 64 rows 2.14 ms).  Split scans have no private L2 slot --
        // every non-resident point is a RED.MAX on a grid that 16 CTAs share -- and want every row they can get
        // (dense2m: 62 rows 0.62 ms, 72 rows 0.53 ms).
        const size_t sc_budget = split ? ctx->smem_optin
                                       : (size_t)(nthreads < 1024 ? 196 : 164) * 1024 - 1024 - 256 - q_bytes;
        const size_t sc_max = ctx->smem_optin - 1024 - q_bytes;
        // touched-segment flags (whole-scan items of the screened kernels): one byte per 16 cells of the grid
        const size_t f_bytes = (SC_TOUCH && cfg.fast_ok && !split) ? sc_flag_bytes((int)cells) : 0;
        const size_t lim_bytes = (sc_budget < sc_max ? sc_budget : sc_max) - f_bytes;
        int ns = (int)(lim_bytes / row_bytes);
        if (ctx->sc_smem_rows >= 0) ns = ctx->sc_smem_rows;                      // A/B override
        if ((size_t)ns * row_bytes > sc_max - f_bytes) ns = (int)((sc_max - f_bytes) / row_bytes);
        if (ns > p->num_ring) ns = p->num_ring;
        if (ns < 2) ns = 0;
        const size_t sc_smem = (size_t)ns * row_bytes + q_bytes + f_bytes;
        // private key slots (whole-scan items): the non-resident rows of one grid per CTA, L2-resident
        const size_t slot_bytes = (size_t)(ns > 0 ? p->num_ring - ns : p->num_ring) * row_bytes;
        const size_t need = split ? 0 : (size_t)ctx->sm_count * slot_bytes;
        if ((rc = ensure_keys(ctx, need, st))) return rc;
        // optional: keep the slots in the persisting part of L2 for the duration of the launch
        cudaStreamAttrValue win;
        memset(&win, 0, sizeof win);
        if (ctx->l2_persist && need > 0 && ctx->persist_max > 0) {
            const size_t want = need < ctx->persist_max ? need : ctx->persist_max;
            if (ctx->persist_set < want) {
                DSC_CUDA(ctx, cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want));
                ctx->persist_set = want;
            }
            win.accessPolicyWindow.base_ptr = ctx->keys;
            win.accessPolicyWindow.num_bytes = need < ctx->window_max ? need : ctx->window_max;
            const double ratio = (double)ctx->persist_set / (double)win.accessPolicyWindow.num_bytes;
            win.accessPolicyWindow.hitRatio = ratio > 1.0 ? 1.0f : (float)ratio;
            win.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            win.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
            DSC_CUDA(ctx, cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &win));
        }
#define SC_LAUNCH4(F, V, H, U, NT, D, SP)                                                                       \
    do {                                                                                                        \
        DSC_CUDA(ctx, cudaFuncSetAttribute(k1_scan<F, V, H, U, NT, D, SP>,                                      \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));         \
        k1_scan<F, V, H, U, NT, D, SP><<<ctas, NT, sc_smem, st>>>(pts, d_offsets, n_scans, ctx->keys, out, stp, \
                                                                  cursor, ns, cfg, spa);                        \
    } while (0)
#define SC_LAUNCH3(F, V, H, U, NT, D)                                                                           \
    do {                                                                                                        \
        if (split) SC_LAUNCH4(F, V, H, U, NT, D, true); else SC_LAUNCH4(F, V, H, U, NT, D, false);              \
    } while (0)
#define SC_LAUNCH_DEF(V, H)                                                                                     \
    do {                                                                                                        \
        if (shape == 1) SC_LAUNCH3(true, V, H, 3, 1024, true);                                                  \
        else if (shape == 2) SC_LAUNCH3(true, V, H, 2, 1024, true);                                             \
        else SC_LAUNCH3(true, V, H, 3, 768, true);                                                              \
    } while (0)
        if (cfg.fast_ok && defgrid) {
            if (p->variant == 0) { if (full) SC_LAUNCH_DEF(0, true); else SC_LAUNCH_DEF(0, false); }
            else                 { if (full) SC_LAUNCH_DEF(1, true); else SC_LAUNCH_DEF(1, false); }
        } else if (cfg.fast_ok) {
            if (p->variant == 0) { if (full) SC_LAUNCH3(true, 0, true, 2, 1024, false); else SC_LAUNCH3(true, 0, false, 2, 1024, false); }
            else                 { if (full) SC_LAUNCH3(true, 1, true, 2, 1024, false); else SC_LAUNCH3(true, 1, false, 2, 1024, false); }
        } else {
            if (p->variant == 0) { if (full) SC_LAUNCH3(false, 0, true, 2, 1024, false); else SC_LAUNCH3(false, 0, false, 2, 1024, false); }
            else                 { if (full) SC_LAUNCH3(false, 1, true, 2, 1024, false); else SC_LAUNCH3(false, 1, false, 2, 1024, false); }
        }
#undef SC_LAUNCH_DEF
#undef SC_LAUNCH3
#undef SC_LAUNCH4
        if ((rc = dsc_check_launch(ctx, "k1_scan"))) return rc;
        if (win.accessPolicyWindow.num_bytes) {      // leave the caller's stream as we found it
            win.accessPolicyWindow.num_bytes = 0;
            DSC_CUDA(ctx, cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &win));
        }
        if (out_norm) return normalize_device(ctx, out, n_scans, cells, out_norm, st);
        return DSC_OK;
    }
    // ---- tile path (small batches, huge scans): persistent tile kernel, shared L2-resident slot ring ----
    const bool st_ok = vec16 && (cells % 4) == 0 && (((uintptr_t)out) & 15) == 0 && (ctx->k1_impl == 0 || ctx->k1_impl == 4);
    if (st_ok) {
        int n_slots = ctx->st_slots;
        if (n_slots > n_scans) n_slots = n_scans;
        const size_t need = (size_t)n_slots * cells * 4;
        if (need > ctx->keys_bytes) {
            if (ctx->keys) { DSC_CUDA(ctx, cudaDeviceSynchronize()); cudaFree(ctx->keys); }
            ctx->keys = nullptr;
            ctx->keys_bytes = 0;
            DSC_CUDA(ctx, cudaMalloc(&ctx->keys, need));
            ctx->keys_bytes = need;
            DSC_CUDA(ctx, cudaMemsetAsync(ctx->keys, 0, need, st));
        }
        // tile table
        const int ST_TILE = ST_THREADS * (ctx->st_cfg == 1 ? 6 : 8);
        long long total_tiles = 0;
        int uniform = -1;
        for (int s = 0; s < n_scans; ++s) {
            const long long t = (h_offsets[s + 1] - h_offsets[s] + ST_TILE - 1) / ST_TILE;
            if (s == 0) uniform = (int)t; else if (t != uniform) uniform = 0;
            total_tiles += t;
        }
        const size_t sync_bytes = sizeof(StreamSync) + sizeof(unsigned) * 2 * (size_t)n_scans;
        const size_t tb_bytes = uniform > 0 ? 0 : sizeof(int64_t) * ((size_t)n_scans + 1);
        if ((rc = dsc_reserve_sync(ctx, sync_bytes + tb_bytes + 64))) return rc;
        StreamSync *sync = (StreamSync *)ctx->sync_buf;
        int64_t *d_tb = nullptr;
        if (uniform <= 0) {
            d_tb = (int64_t *)((uint8_t *)ctx->sync_buf + ((sync_bytes + 15) & ~(size_t)15));
            std::vector<int64_t> tb((size_t)n_scans + 1);
            long long acc_t = 0;
            for (int s = 0; s < n_scans; ++s) { tb[s] = acc_t; acc_t += (h_offsets[s + 1] - h_offsets[s] + ST_TILE - 1) / ST_TILE; }
            tb[n_scans] = acc_t;
            // pageable source: the copy is staged by the driver before the call returns
            DSC_CUDA(ctx, cudaMemcpyAsync(d_tb, tb.data(), tb_bytes, cudaMemcpyHostToDevice, st));
        }
        DSC_CUDA(ctx, cudaMemsetAsync(sync, 0, sync_bytes, st));
        if (total_tiles > 0x7fff0000LL) DSC_FAIL(ctx, DSC_ERR_INVALID, "batch too large for one call");
        if (total_tiles > 0) {
            unsigned long long *stp = (unsigned long long *)stats;
            const int minb = ctx->st_cfg == 1 ? 4 : 3;
            int grid_ctas = ctx->sm_count * minb;
            if ((long long)grid_ctas > total_tiles) grid_ctas = (int)total_tiles;
#define ST_LAUNCH2(F, V, H, PPT, MINB)                                                                          \
    k1_stream<F, V, H, PPT, MINB><<<grid_ctas, ST_THREADS, 0, st>>>(pts, d_offsets, d_tb, n_scans,              \
                                                                    (int)total_tiles, uniform, ctx->keys,       \
                                                                    n_slots, out, stp, sync, cfg)
#define ST_LAUNCH(F, V, H)                                                                                      \
    do {                                                                                                        \
        if (ctx->st_cfg == 1) ST_LAUNCH2(F, V, H, 6, 4); else ST_LAUNCH2(F, V, H, 8, 3);                        \
    } while (0)
            const bool full = p->full_height != 0;
            if (cfg.fast_ok) {
                if (p->variant == 0) { if (full) ST_LAUNCH(true, 0, true); else ST_LAUNCH(true, 0, false); }
                else                 { if (full) ST_LAUNCH(true, 1, true); else ST_LAUNCH(true, 1, false); }
            } else {
                if (p->variant == 0) { if (full) ST_LAUNCH(false, 0, true); else ST_LAUNCH(false, 0, false); }
                else                 { if (full) ST_LAUNCH(false, 1, true); else ST_LAUNCH(false, 1, false); }
            }
#undef ST_LAUNCH
#undef ST_LAUNCH2
            if ((rc = dsc_check_launch(ctx, "k1_stream"))) return rc;
        }
        // scans without points never retire a tile: their grids are all zeros
        for (int s = 0; s < n_scans; ++s)
            if (h_offsets[s + 1] == h_offsets[s])
                DSC_CUDA(ctx, cudaMemsetAsync(out + (size_t)s * cells, 0, (size_t)cells * 4, st));
        if (out_norm) return normalize_device(ctx, out, n_scans, cells, out_norm, st);
        return DSC_OK;
    }
    // ---- shared-memory half grids, one CTA pair per scan, persistent over the batch (A/B variant) ----
    const size_t hg_smem = (size_t)cells * 2;   // R * S/2 * 4 bytes
    const bool hg_ok = vec16 && (p->num_sector % 8) == 0 && hg_smem <= ctx->smem_optin &&
                       (((uintptr_t)out) & 15) == 0 && ctx->k1_impl != 1 &&
                       ctx->k1_impl == 2;
    if (hg_ok) {
        int pairs = ctx->sm_count / 2;
        if (pairs > n_scans) pairs = n_scans;
        unsigned long long *stp = (unsigned long long *)stats;
#define HG_LAUNCH2(F, V, H, T, U)                                                                               \
    do {                                                                                                        \
        DSC_CUDA(ctx, cudaFuncSetAttribute(k1_halfgrid<F, V, H, T, U>,                                          \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)hg_smem));         \
        k1_halfgrid<F, V, H, T, U><<<2 * pairs, T, hg_smem, st>>>(pts, d_offsets, n_scans, out, stp,        \
                                                                   ctx->pf_dist, cfg);                          \
    } while (0)
#define HG_LAUNCH(F, V, H)                                                                                      \
    do {                                                                                                        \
        if (ctx->hg_cfg == 1) HG_LAUNCH2(F, V, H, 1024, 3);                                                     \
        else if (ctx->hg_cfg == 2) HG_LAUNCH2(F, V, H, 512, 4);                                                 \
        else HG_LAUNCH2(F, V, H, 1024, 2);                                                                      \
    } while (0)
        const bool full = p->full_height != 0;
        if (cfg.fast_ok) {
            if (p->variant == 0) { if (full) HG_LAUNCH(true, 0, true); else HG_LAUNCH(true, 0, false); }
            else                 { if (full) HG_LAUNCH(true, 1, true); else HG_LAUNCH(true, 1, false); }
        } else {
            if (p->variant == 0) { if (full) HG_LAUNCH(false, 0, true); else HG_LAUNCH(false, 0, false); }
            else                 { if (full) HG_LAUNCH(false, 1, true); else HG_LAUNCH(false, 1, false); }
        }
#undef HG_LAUNCH
#undef HG_LAUNCH2
        if ((rc = dsc_check_launch(ctx, "k1_halfgrid"))) return rc;
        if (out_norm) return normalize_device(ctx, out, n_scans, cells, out_norm, st);
        return DSC_OK;
    }
    // ---- fallback path (small batches, odd grids, 12-byte stride): key grid in L2 + RED.MAX ----
    // key scratch: sized to stay L2-resident between build and decode
    int slots = (int)((size_t)48 * 1024 * 1024 / (size_t)(cells * 4));
    if (slots < 1) slots = 1;
    if (slots > 4096) slots = 4096;
    if (slots > n_scans) slots = n_scans;
    const size_t need = (size_t)slots * cells * 4;
    if (need > ctx->keys_bytes) {
        if (ctx->keys) cudaFree(ctx->keys);
        ctx->keys = nullptr;
        ctx->keys_bytes = 0;
        DSC_CUDA(ctx, cudaMalloc(&ctx->keys, need));
        ctx->keys_bytes = need;
        DSC_CUDA(ctx, cudaMemsetAsync(ctx->keys, 0, need, st));
    }
    for (int s0 = 0; s0 < n_scans; s0 += slots) {
        const int g = n_scans - s0 < slots ? n_scans - s0 : slots;
        int64_t max_n = 0;
        for (int s = s0; s < s0 + g; ++s) {
            const int64_t n = h_offsets[s + 1] - h_offsets[s];
            if (n > max_n) max_n = n;
        }
        if (max_n > 0) {
            const int64_t tiles = (max_n + K1_TILE - 1) / K1_TILE;
            if (tiles > 0x7fffffff) DSC_FAIL(ctx, DSC_ERR_INVALID, "scan too large");
            dim3 grid((unsigned)tiles, (unsigned)g);
            unsigned long long *stp = (unsigned long long *)stats;
            if (vec16 && cfg.fast_ok)
                k1_build_keys<true, true><<<grid, K1_THREADS, 0, st>>>(pts, d_offsets, s0, ctx->keys, stp, stride, cfg);
            else if (vec16)
                k1_build_keys<true, false><<<grid, K1_THREADS, 0, st>>>(pts, d_offsets, s0, ctx->keys, stp, stride, cfg);
            else if (cfg.fast_ok)
                k1_build_keys<false, true><<<grid, K1_THREADS, 0, st>>>(pts, d_offsets, s0, ctx->keys, stp, stride, cfg);
            else
                k1_build_keys<false, false><<<grid, K1_THREADS, 0, st>>>(pts, d_offsets, s0, ctx->keys, stp, stride, cfg);
            if ((rc = dsc_check_launch(ctx, "k1_build_keys"))) return rc;
        }
        const int64_t total = (int64_t)g * cells;
        float *o = out + (int64_t)s0 * cells;
        if ((cells & 3) == 0 && (((uintptr_t)o) & 15) == 0) {
            const int64_t nv = total / 4;
            int blocks = (int)((nv + 255) / 256 < (int64_t)ctx->sm_count * 16 ? (nv + 255) / 256
                                                                              : (int64_t)ctx->sm_count * 16);
            k1_decode<<<blocks, 256, 0, st>>>(ctx->keys, o, nv);
        } else {
            int blocks = (int)((total + 255) / 256 < (int64_t)ctx->sm_count * 16 ? (total + 255) / 256
                                                                                 : (int64_t)ctx->sm_count * 16);
            k1_decode_scalar<<<blocks, 256, 0, st>>>(ctx->keys, o, total);
        }
        if ((rc = dsc_check_launch(ctx, "k1_decode"))) return rc;
    }
    if (out_norm) {
        if ((rc = normalize_device(ctx, out, n_scans, cells, out_norm, st))) return rc;
    }
    return DSC_OK;
}

static int upload_offsets(dsc_ctx *ctx, const int64_t *offsets, int n_scans, cudaStream_t st)
{
    const size_t cnt = (size_t)n_scans + 1;
    if (cnt > ctx->d_offsets_cap) {
        if (ctx->d_offsets) cudaFree(ctx->d_offsets);
        if (ctx->h_offsets) cudaFreeHost(ctx->h_offsets);
        ctx->d_offsets = nullptr;
        ctx->h_offsets = nullptr;
        ctx->d_offsets_cap = ctx->h_offsets_cap = 0;
        const size_t cap = cnt < 4096 ? 4096 : cnt * 2;
        DSC_CUDA(ctx, cudaMalloc(&ctx->d_offsets, cap * sizeof(int64_t)));
        DSC_CUDA(ctx, cudaMallocHost(&ctx->h_offsets, cap * sizeof(int64_t)));
        ctx->d_offsets_cap = ctx->h_offsets_cap = cap;
    } else {
        // the pinned staging buffer may still be in flight from the previous async call
        DSC_CUDA(ctx, cudaStreamSynchronize(st));
    }
    memcpy(ctx->h_offsets, offsets, cnt * sizeof(int64_t));
    DSC_CUDA(ctx, cudaMemcpyAsync(ctx->d_offsets, ctx->h_offsets, cnt * sizeof(int64_t),
                                  cudaMemcpyHostToDevice, st));
    return DSC_OK;
}

static void free_stage(dsc_ctx *ctx)
{
    for (int i = 0; i < DSC_NBUF; ++i) {
        if (ctx->stage_pts[i]) cudaFree(ctx->stage_pts[i]);
        if (ctx->stage_out[i]) cudaFree(ctx->stage_out[i]);
        if (ctx->stage_norm[i]) cudaFree(ctx->stage_norm[i]);
        if (ctx->stage_stats[i]) cudaFree(ctx->stage_stats[i]);
        if (ctx->stage_doff[i]) cudaFree(ctx->stage_doff[i]);
        if (ctx->stage_hoff[i]) cudaFreeHost(ctx->stage_hoff[i]);
        ctx->stage_pts[i] = nullptr;
        ctx->stage_out[i] = nullptr;
        ctx->stage_norm[i] = nullptr;
        ctx->stage_stats[i] = nullptr;
        ctx->stage_doff[i] = nullptr;
        ctx->stage_hoff[i] = nullptr;
    }
    ctx->stage_pts_cap = 0;
    ctx->stage_scans_cap = 0;
    ctx->stage_cells = 0;
}

void dsc_k1_release(dsc_ctx *ctx)
{
    free_stage(ctx);
    if (ctx->keys) cudaFree(ctx->keys);
    if (ctx->d_offsets) cudaFree(ctx->d_offsets);
    if (ctx->h_offsets) cudaFreeHost(ctx->h_offsets);
    if (ctx->minmax) cudaFree(ctx->minmax);
    if (ctx->sync_buf) cudaFree(ctx->sync_buf);
    ctx->sync_buf = nullptr;
    ctx->sync_cap = 0;
    ctx->keys = nullptr;
    ctx->d_offsets = nullptr;
    ctx->h_offsets = nullptr;
    ctx->minmax = nullptr;
}

// Host-resident batch: chunked, double-buffered pipeline over three streams
//   s_h2d: points of chunk i+1 go up   |  s_k: kernels of chunk i  |  s_d2h: grids of chunk i-1 come down
// so the PCIe link runs in both directions while the SMs work.
static int build_host(dsc_ctx *ctx, const dsc_params *p, const uint8_t *pts, const int64_t *offsets,
                      int n_scans, int stride, float *out, float *out_norm, dsc_scan_stats *stats)
{
    const int64_t cells = (int64_t)p->num_ring * p->num_sector;
    const size_t chunk_bytes_target = (size_t)64 << 20;
    size_t max_scan_bytes = 0;
    for (int s = 0; s < n_scans; ++s) {
        const size_t b = (size_t)(offsets[s + 1] - offsets[s]) * stride;
        if (b > max_scan_bytes) max_scan_bytes = b;
    }
    const size_t pts_cap = max_scan_bytes > chunk_bytes_target ? max_scan_bytes : chunk_bytes_target;
    const size_t scans_cap = 128;
    if (pts_cap > ctx->stage_pts_cap || scans_cap > ctx->stage_scans_cap || (size_t)cells != ctx->stage_cells) {
        DSC_CUDA(ctx, cudaDeviceSynchronize());
        free_stage(ctx);
        for (int i = 0; i < DSC_NBUF; ++i) {
            DSC_CUDA(ctx, cudaMalloc(&ctx->stage_pts[i], pts_cap + 16));
            DSC_CUDA(ctx, cudaMalloc(&ctx->stage_out[i], scans_cap * cells * 4));
            DSC_CUDA(ctx, cudaMalloc(&ctx->stage_norm[i], scans_cap * cells * 4));
            DSC_CUDA(ctx, cudaMalloc(&ctx->stage_stats[i], scans_cap * sizeof(dsc_scan_stats)));
            DSC_CUDA(ctx, cudaMalloc(&ctx->stage_doff[i], (scans_cap + 1) * sizeof(int64_t)));
            DSC_CUDA(ctx, cudaMallocHost(&ctx->stage_hoff[i], (scans_cap + 1) * sizeof(int64_t)));
        }
        ctx->stage_pts_cap = pts_cap;
        ctx->stage_scans_cap = scans_cap;
        ctx->stage_cells = (size_t)cells;
    }
    int s0 = 0, buf = 0, rc = DSC_OK;
    int used[DSC_NBUF] = {};
    while (s0 < n_scans) {
        int s1 = s0;
        size_t bytes = 0;
        while (s1 < n_scans && (size_t)(s1 - s0) < scans_cap) {
            const size_t b = (size_t)(offsets[s1 + 1] - offsets[s1]) * stride;
            if (s1 > s0 && bytes + b > pts_cap) break;
            bytes += b;
            ++s1;
        }
        const int g = s1 - s0;
        if (used[buf]) {
            // pinned offsets of this buffer are rewritten below: its previous upload must be done;
            // the previous kernels reading stage_pts[buf] must be done before we overwrite it
            DSC_CUDA(ctx, cudaEventSynchronize(ctx->ev_h2d[buf]));
            DSC_CUDA(ctx, cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_k[buf], 0));
        }
        int64_t *ho = ctx->stage_hoff[buf];
        for (int s = 0; s <= g; ++s) ho[s] = offsets[s0 + s] - offsets[s0];
        DSC_CUDA(ctx, cudaMemcpyAsync(ctx->stage_doff[buf], ho, (size_t)(g + 1) * sizeof(int64_t),
                                      cudaMemcpyHostToDevice, ctx->s_h2d));
        if (bytes)
            DSC_CUDA(ctx, cudaMemcpyAsync(ctx->stage_pts[buf], pts + (size_t)offsets[s0] * stride, bytes,
                                          cudaMemcpyHostToDevice, ctx->s_h2d));
        DSC_CUDA(ctx, cudaEventRecord(ctx->ev_h2d[buf], ctx->s_h2d));
        DSC_CUDA(ctx, cudaStreamWaitEvent(ctx->s_k, ctx->ev_h2d[buf], 0));
        if (used[buf]) DSC_CUDA(ctx, cudaStreamWaitEvent(ctx->s_k, ctx->ev_d2h[buf], 0));
        if ((rc = build_device(ctx, p, (const uint8_t *)ctx->stage_pts[buf], ctx->stage_doff[buf], ho, g, stride,
                               ctx->stage_out[buf], out_norm ? ctx->stage_norm[buf] : nullptr,
                               stats ? ctx->stage_stats[buf] : nullptr, ctx->s_k)))
            return rc;
        DSC_CUDA(ctx, cudaEventRecord(ctx->ev_k[buf], ctx->s_k));
        DSC_CUDA(ctx, cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_k[buf], 0));
        DSC_CUDA(ctx, cudaMemcpyAsync(out + (int64_t)s0 * cells, ctx->stage_out[buf], (size_t)g * cells * 4,
                                      cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (out_norm)
            DSC_CUDA(ctx, cudaMemcpyAsync(out_norm + (int64_t)s0 * cells, ctx->stage_norm[buf],
                                          (size_t)g * cells * 4, cudaMemcpyDeviceToHost, ctx->s_d2h));
        if (stats)
            DSC_CUDA(ctx, cudaMemcpyAsync(stats + s0, ctx->stage_stats[buf], (size_t)g * sizeof(dsc_scan_stats),
                                          cudaMemcpyDeviceToHost, ctx->s_d2h));
        DSC_CUDA(ctx, cudaEventRecord(ctx->ev_d2h[buf], ctx->s_d2h));
        used[buf] = 1;
        s0 = s1;
        buf = buf + 1 == DSC_NBUF ? 0 : buf + 1;
    }
    DSC_CUDA(ctx, cudaStreamSynchronize(ctx->s_d2h));
    DSC_CUDA(ctx, cudaStreamSynchronize(ctx->s_k));
    DSC_CUDA(ctx, cudaStreamSynchronize(ctx->s_h2d));
    return DSC_OK;
}

extern "C" int dsc_sc_build_batch(dsc_ctx *ctx, const dsc_params *params, const void *pts,
                                  const int64_t *offsets, int n_scans, int pts_stride_bytes, int mem,
                                  float *out_grids, float *out_norm, dsc_scan_stats *stats, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    int rc;
    if ((rc = validate_params(ctx, params))) return rc;
    if (n_scans < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "n_scans < 0");
    if (n_scans == 0) return DSC_OK;
    if (!offsets || !out_grids) DSC_FAIL(ctx, DSC_ERR_INVALID, "offsets/out_grids is NULL");
    if (pts_stride_bytes != 16 && pts_stride_bytes != 12) DSC_FAIL(ctx, DSC_ERR_INVALID, "pts_stride_bytes must be 12 or 16");
    if (mem != DSC_MEM_HOST && mem != DSC_MEM_DEVICE) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad mem");
    for (int s = 0; s < n_scans; ++s) {
        if (offsets[s + 1] < offsets[s] || offsets[s] < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "offsets must be non-decreasing and >= 0");
        if (offsets[s + 1] - offsets[s] > 0x7fff0000LL) DSC_FAIL(ctx, DSC_ERR_INVALID, "more than 2^31 points in one scan");
    }
    if (offsets[n_scans] > offsets[0] && !pts) DSC_FAIL(ctx, DSC_ERR_INVALID, "pts is NULL");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (mem == DSC_MEM_HOST) {
        return build_host(ctx, params, (const uint8_t *)pts, offsets, n_scans, pts_stride_bytes, out_grids,
                          out_norm, stats);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if ((rc = upload_offsets(ctx, offsets, n_scans, st))) return rc;
    if (ctx->timing) DSC_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    rc = build_device(ctx, params, (const uint8_t *)pts, ctx->d_offsets, offsets, n_scans, pts_stride_bytes,
                      out_grids, out_norm, stats, st);
    if (rc) return rc;
    if (ctx->timing) {
        DSC_CUDA(ctx, cudaEventRecord(ctx->ev1, st));
        DSC_CUDA(ctx, cudaEventSynchronize(ctx->ev1));
        DSC_CUDA(ctx, cudaEventElapsedTime(&ctx->last_ms, ctx->ev0, ctx->ev1));
    }
    return DSC_OK;
}

extern "C" int dsc_sc_normalize01(dsc_ctx *ctx, const float *grids, int n_grids, int64_t cells_per_grid,
                                  int mem, float *out, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    if (n_grids < 0 || cells_per_grid <= 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad sizes");
    if (n_grids == 0) return DSC_OK;
    if (!grids || !out) DSC_FAIL(ctx, DSC_ERR_INVALID, "NULL buffer");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    if (mem == DSC_MEM_DEVICE) return normalize_device(ctx, grids, n_grids, cells_per_grid, out, (cudaStream_t)stream);
    const size_t bytes = (size_t)n_grids * cells_per_grid * 4;
    int rc;
    if ((rc = dsc_reserve_scratch(ctx, bytes * 2))) return rc;
    float *d_in = (float *)ctx->scratch, *d_out = d_in + (size_t)n_grids * cells_per_grid;
    cudaStream_t st = ctx->s_k;
    DSC_CUDA(ctx, cudaMemcpyAsync(d_in, grids, bytes, cudaMemcpyHostToDevice, st));
    if ((rc = normalize_device(ctx, d_in, n_grids, cells_per_grid, d_out, st))) return rc;
    DSC_CUDA(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, st));
    DSC_CUDA(ctx, cudaStreamSynchronize(st));
    return DSC_OK;
}
