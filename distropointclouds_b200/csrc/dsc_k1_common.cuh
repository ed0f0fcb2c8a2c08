// dsc_k1_common.cuh -- device helpers shared by the K1 kernels: streaming loads and L2 policies, fire-and-forget
// reductions, the out-of-line exact (fdlibm) binning call and the placement of the shared-memory ring rows.
// Included inside the anonymous namespace of dsc_k1_build.cu.
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

// Placement of the key grid (scan kernel): `ns` ring rows live in shared memory -- rings 0 .. ns-2 plus the
// clamp ring R-1, which together hold 70-90 % of the points of a KITTI-like scan -- and take the LDS pre-check +
// ATOMS.MAX path; points of the remaining rings are spilled and reduced band by band in a second pass (whole-scan
// items) or go to the scan's key grid with RED.MAX (split items).
struct SmemRows {
    uint32_t base;   // shared-window address of the resident rows
    int c0;          // first cell of the resident window (ring r0)
    int lim;         // cells c0 .. c0+lim-1 are resident rows 0 .. ns-2
    int last;        // first cell of the clamp ring, kept in resident row ns-1
    int shift;       // lim - last: clamp-ring cell -> resident row ns-1
};

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
