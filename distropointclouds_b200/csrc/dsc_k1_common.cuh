// dsc_k1_common.cuh -- device helpers shared by every K1 kernel: streaming loads and L2 policies, fire-and-forget
// reductions, the out-of-line exact (fdlibm) binning call, the hybrid shared-memory / L2 key-grid placement and
// the point function of the generic kernels.  Included inside the anonymous namespace of dsc_k1_build.cu.
#pragma once

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
