// dsc_k1_scan.cuh -- K1 main path: k1_scan, the persistent scan / work-item kernel.
// Included inside the anonymous namespace of dsc_k1_build.cu, after dsc_k1_common.cuh.
#pragma once

// ---- K1 main path: persistent CTAs, one work item (a scan, or a slice of a big scan) at a time ------------------
// A CTA streams the item's points once (128-bit loads, two register sets ping-pong), bins each point once and
// folds it into the key grid: the hot ring rows live in shared memory (LDS pre-check + ATOMS.MAX), the other rows
// in the CTA's private L2-resident slot (RED.MAX, fire and forget).  When a scan is done the same CTA decodes
// rows and slot to floats (coalesced 128-bit streaming stores) and re-zeroes both.  There is no synchronisation
// between CTAs for whole-scan items; HBM sees each point once and each output cell once.
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
