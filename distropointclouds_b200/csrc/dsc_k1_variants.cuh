// dsc_k1_variants.cuh -- the other K1 forms: the generic fallback (k1_build_keys + k1_decode: 12-byte stride, odd
// sector counts, unaligned buffers) and the measured alternatives kept for A/B runs behind
// dsc_set_option("k1_impl" / "sc_loader"): CTA-pair half grids (k1_halfgrid), tile scheduler with a shared ring of
// L2 slots (k1_stream), TMA-ring point loader (k1_ring).  Numbers: profiles/r1_notes.md.
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
