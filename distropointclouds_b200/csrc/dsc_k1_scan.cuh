// dsc_k1_scan.cuh -- K1 main path: k1_scan, the persistent scan / work-item kernel.
// Included inside the anonymous namespace of dsc_k1_build.cu, after dsc_k1_common.cuh.
#pragma once

// ---- K1 main path: persistent CTAs, one work item (a scan, or a slice of a big scan) at a time ------------------
// A CTA streams the item's points once (128-bit loads, two register sets ping-pong), bins each point once and
// folds it into the key grid.  The 460 KB grid does not fit one SM, so the reduction runs in ring bands:
//   pass 1 (the stream)  the hot ring rows -- rings 0 .. ns-2 and the clamp ring R-1, 75-90 % of the points of a
//                        KITTI-like scan -- live in shared memory (LDS pre-check + ATOMS.MAX).  A point of any other
//                        ring is NOT reduced yet: its (cell, key) pair is appended to the warp's private spill list
//                        (one ballot, one coalesced 8-byte store per lane; the list is a few hundred KB per CTA and
//                        stays in L2).
//   write-out            the resident rows are decoded to floats (coalesced 128-bit streaming stores) and re-zeroed;
//   pass 2 (per band)    the same shared-memory rows now stand for the next band of up to ns rings: every warp
//                        re-reads its own spill list (L2 hits, 128-bit loads) and folds the band's entries in with
//                        the same LDS pre-check + ATOMS.MAX; the band is decoded and written out; repeat until
//                        ring R-2.
// So every max-reduction is a shared-memory atomic, HBM sees each point once and each output cell once, there are
// no global atomics, no key slots to read back and re-zero, and no synchronisation between CTAs for whole-scan items.
// (Round 1 reduced the non-resident rings with RED.MAX on a private L2 slot per CTA: 72 M scattered L2 atomics per
// launch, a 282 KB read-back + re-zero per scan and 1.2x write traffic from evicted slot lines: profiles/r1_notes.md.)
constexpr int SC_THREADS = 1024;
#ifndef SC_DU
#define SC_DU 4                       // 128-bit L2 loads in flight per thread in the split-mode decode
#endif
#ifndef SC_LU
#define SC_LU 4                       // 128-bit spill-list loads in flight per lane in pass 2
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
__device__ __forceinline__ unsigned lanemask_lt()
{
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
// spill-list traffic goes to L2 and no further (st.cg / ld.cg): written once, read back by the same SM
__device__ __forceinline__ void st_spill(uint2 *p, uint32_t cell, uint32_t key)
{
    asm volatile("st.global.cg.v2.u32 [%0], {%1,%2};" ::"l"(p), "r"(cell), "r"(key) : "memory");
}
__device__ __forceinline__ uint4 ld_spill2(const uint2 *p)
{
    uint4 v;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}

// Fold one binned point into the key grid.  Called warp-uniformly; `valid` masks lanes without a point.
//   resident cell      LDS pre-check + ATOMS.MAX (most points are not a new maximum)
//   any other cell     SPILL: appended to the warp's spill list (`gp` is the list's write pointer, advanced for the whole
//                      warp);  !SPILL (split items): RED.MAX straight into the scan's key grid (`gp` is that grid)
template <bool FULL, bool SPILL>
__device__ __forceinline__ void sc_commit(const float zv, const int cell, bool valid, const SmemRows sm, uint32_t *&gp)
{
    const uint32_t u = __float_as_uint(zv);
    const uint32_t key = u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u);     // order-preserving key
    valid = valid && (zv == zv);                 // NaN never wins (scaner_version2.py:107)
    if (!FULL) valid = valid && (zv > 0.0f);     // offset mode: cells start at 0, strict '>'
    const bool res = cell < sm.lim || cell >= sm.last;
    if (valid && res) {
        const int local = cell < sm.lim ? cell : cell + sm.shift;               // clamp ring -> row ns-1
        const uint32_t addr = sm.base + (uint32_t)local * 4u;
        if (key > lds_u32(addr)) reds_max_u32(addr, key);
    }
    if constexpr (SPILL) {
        const unsigned m = __ballot_sync(0xffffffffu, valid && !res);
        if (m) {                                 // warp-uniform
            uint2 *w = reinterpret_cast<uint2 *>(gp);
            if (valid && !res) st_spill(w + __popc(m & lanemask_lt()), (uint32_t)cell, key);
            gp = reinterpret_cast<uint32_t *>(w + __popc(m));
        }
    } else {
        if (valid && !res) red_max_global(gp + cell, key);
    }
}

// exact chain for one queued point (entry address of this lane): packed result of hg_bin_exact, counters updated
template <int VARIANT, bool FULL>
__device__ __noinline__ uint2 sc_exact_queued(uint32_t qaddr, const ExactArgs *__restrict__ ea,
                                              unsigned long long *__restrict__ st)
{
    const float x = lds_f32(qaddr), y = lds_f32(qaddr + SCQ_CAP * 4), z = lds_f32(qaddr + SCQ_CAP * 8);
    const uint2 b = hg_bin_exact<VARIANT, FULL>(x, y, z, ea);
    const int status = (int)(b.x >> 27);
    if (st) {                                    // rare events: straight to the per-scan counters
        atomicAdd(st + 5, 1ull);
        if (status == dsc::PT_NONFINITE) atomicAdd(st + 1, 1ull);
        if (status == dsc::PT_SECTOR_OVERFLOW) atomicAdd(st + 2, 1ull);
        if (status == dsc::PT_OUT_OF_RANGE) atomicAdd(st + 4, 1ull);
    }
    return b;
}

// run the exact chain for the queue entries [first, first + count) of this warp (count <= 32) and commit them
template <int VARIANT, bool FULL, bool SPILL>
__device__ __forceinline__ void sc_exact_batch(const uint32_t queue, const unsigned first, const unsigned count,
                                               const ExactArgs *ea, uint32_t *&gp, unsigned &n_clamped,
                                               unsigned long long *__restrict__ st, const SmemRows sm)
{
    const unsigned lane = threadIdx.x & 31u;
    const bool mine = lane < count;
    uint2 b = make_uint2((uint32_t)dsc::PT_NONFINITE << 27, 0u);
    __syncwarp();
    if (mine) b = sc_exact_queued<VARIANT, FULL>(queue + (first + lane) * 4u, ea, st);
    __syncwarp();
    const int status = (int)(b.x >> 27);
    n_clamped += (mine && status == dsc::PT_CLAMPED) ? 1u : 0u;
    sc_commit<FULL, SPILL>(__uint_as_float(b.y), (int)(b.x & 0x7ffffffu), mine && status <= dsc::PT_CLAMPED, sm, gp);
}

// One point per lane, called warp-uniformly (`active` masks lanes past the end of the item).
template <int VARIANT, bool FULL, bool DEF, bool SPILL>
__device__ __forceinline__ void sc_point(const BinCfg &cfg, const ExactArgs *ea, const float4 v, const bool active,
                                         uint32_t *&gp, unsigned &n_clamped, unsigned long long *__restrict__ st,
                                         const SmemRows sm, const uint32_t queue, unsigned &qcount)
{
    int ring, sector;
    bool beyond;
    bool ok = dsc::screened_core_t<DEF>(cfg, v.x, v.y, 0.0f, 0.0f, &ring, &sector, &beyond);
    if (VARIANT == 1) ok = ok && (fabsf(v.z) < INFINITY);    // finite mask (scan_context_node.py:115)
    bool take = ok && active;
    if (VARIANT == 1) {
        if (take && beyond && st) atomicAdd(st + 4, 1ull);   // r > max_length: dropped (scan_context_node.py:134-137)
        take = take && !beyond;
    } else {
        n_clamped += (take && beyond) ? 1u : 0u;
    }
    const int S = DEF ? dsc::defgrid::S : cfg.S;
    sc_commit<FULL, SPILL>(FULL ? v.z : __fadd_rn(v.z, cfg.offset), ring * S + sector, take, sm, gp);
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
            sc_exact_batch<VARIANT, FULL, SPILL>(queue, qcount, 32u, ea, gp, n_clamped, st, sm);
        }
    }
}

// exact-only flavour (DSC_FLAG_EXACT_ONLY, non-dyadic gaps): every point takes the fdlibm chain
template <int VARIANT, bool FULL, bool SPILL>
__device__ __forceinline__ void sc_point_exact(const ExactArgs *ea, const float4 v, const bool active, uint32_t *&gp,
                                               unsigned &n_clamped, unsigned long long *__restrict__ st, const SmemRows sm)
{
    uint2 b = make_uint2((uint32_t)dsc::PT_NONFINITE << 27, 0u);
    if (active) {
        b = hg_bin_exact<VARIANT, FULL>(v.x, v.y, v.z, ea);
        const int status = (int)(b.x >> 27);
        if (st && status > dsc::PT_CLAMPED) {
            if (status == dsc::PT_NONFINITE) atomicAdd(st + 1, 1ull);
            if (status == dsc::PT_SECTOR_OVERFLOW) atomicAdd(st + 2, 1ull);
            if (status == dsc::PT_OUT_OF_RANGE) atomicAdd(st + 4, 1ull);
        }
    }
    __syncwarp();
    const int status = (int)(b.x >> 27);
    n_clamped += (active && status == dsc::PT_CLAMPED) ? 1u : 0u;
    sc_commit<FULL, SPILL>(__uint_as_float(b.y), (int)(b.x & 0x7ffffffu), active && status <= dsc::PT_CLAMPED, sm, gp);
}

// end of item: the queue's remaining entries (fewer than 32)
template <int VARIANT, bool FULL, bool SPILL>
__device__ __forceinline__ void sc_drain(const ExactArgs *ea, uint32_t *&gp, unsigned &n_clamped,
                                         unsigned long long *__restrict__ st, const SmemRows sm, const uint32_t queue,
                                         unsigned &qcount)
{
    if (qcount) {                                // warp-uniform
        sc_exact_batch<VARIANT, FULL, SPILL>(queue, 0u, qcount, ea, gp, n_clamped, st, sm);
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

__device__ __forceinline__ float4 sc_decode4(const uint4 q)
{
    float4 f;
    f.x = dsc::key_to_float(q.x);
    f.y = dsc::key_to_float(q.y);
    f.z = dsc::key_to_float(q.z);
    f.w = dsc::key_to_float(q.w);
    return f;
}

// decode `nvec` 128-bit groups of keys from shared memory to floats, re-zeroing the rows
template <int NT>
__device__ __forceinline__ void sc_flush_rows(uint32_t *__restrict__ srows, float4 *__restrict__ dst, int nvec)
{
    for (int i = threadIdx.x; i < nvec; i += NT) {
        const uint4 q = reinterpret_cast<uint4 *>(srows)[i];
        reinterpret_cast<uint4 *>(srows)[i] = make_uint4(0, 0, 0, 0);
        __stcs(dst + i, sc_decode4(q));
    }
}

// End of a whole-scan item (this CTA owns the scan): write out the resident rows, then reduce and write out the other
// rings band by band from the spill lists.  `wseg`/`wend`: this warp's list.  Ends with a CTA barrier.
template <int NT>
__device__ __forceinline__ void sc_finish_scan(uint32_t *__restrict__ srows, const uint32_t sbase,
                                               float *__restrict__ out_scan, const int ns, const int R, const int S,
                                               const uint2 *__restrict__ wseg, const uint2 *__restrict__ wend)
{
    const int rowvec = S >> 2;
    float4 *oq = reinterpret_cast<float4 *>(out_scan);
    __syncthreads();                             // every shared-memory atomic and every list store of the item is done
    // resident rows: rings 0 .. ns-2, and the clamp ring R-1 kept in row ns-1
    sc_flush_rows<NT>(srows, oq, (ns - 1) * rowvec);
    for (int i = threadIdx.x; i < rowvec; i += NT) {
        uint4 *src = reinterpret_cast<uint4 *>(srows) + (ns - 1) * rowvec + i;
        const uint4 q = *src;
        *src = make_uint4(0, 0, 0, 0);
        __stcs(oq + (size_t)(R - 1) * rowvec + i, sc_decode4(q));
    }
    const unsigned lane = threadIdx.x & 31u;
    for (int row0 = ns - 1; row0 < R - 1; row0 += ns) {         // bands of up to ns rings
        const int rows = R - 1 - row0 < ns ? R - 1 - row0 : ns;
        const uint32_t c_lo = (uint32_t)row0 * (uint32_t)S, c_n = (uint32_t)rows * (uint32_t)S;
        __syncthreads();                         // the rows are zero again
        // this warp's list, two entries per 128-bit load, SC_LU loads in flight per lane
        for (const uint2 *e0 = wseg + 2u * lane; e0 < wend; e0 += 64 * SC_LU) {
            uint4 q[SC_LU];
#pragma unroll
            for (int d = 0; d < SC_LU; ++d) {
                const uint2 *e = e0 + 64 * d;
                if (e < wend) q[d] = ld_spill2(e);
            }
#pragma unroll
            for (int d = 0; d < SC_LU; ++d) {
                const uint2 *e = e0 + 64 * d;
                if (e < wend) {
                    const uint32_t l0 = q[d].x - c_lo;
                    if (l0 < c_n) {
                        const uint32_t a = sbase + l0 * 4u;
                        if (q[d].y > lds_u32(a)) reds_max_u32(a, q[d].y);
                    }
                    const uint32_t l1 = q[d].z - c_lo;
                    if (e + 1 < wend && l1 < c_n) {
                        const uint32_t a = sbase + l1 * 4u;
                        if (q[d].w > lds_u32(a)) reds_max_u32(a, q[d].w);
                    }
                }
            }
        }
        __syncthreads();
        sc_flush_rows<NT>(srows, oq + (size_t)row0 * rowvec, rows * rowvec);
    }
    __syncthreads();                             // rows are zero before the next scan's atomics
}

// End of a split work item: fold the resident rows into the scan's key grid (== its output grid), count the item as
// retired, and if it was the scan's last item decode the grid in place.  Ends with a CTA barrier.
template <int NT, int DU = 4>
__device__ __forceinline__ void sc_finish_split(uint32_t *__restrict__ srows, uint32_t *__restrict__ grid, int ns, int R,
                                                int S, int scan, int parts, unsigned int *__restrict__ done, int *s_last)
{
    const int tid = threadIdx.x;
    const int rowvec = S >> 2;
    const int cells = R * S;
    const int nres = ns * rowvec;
    const int lim = (ns - 1) * rowvec, shift = (R - ns) * rowvec;    // resident row ns-1 holds ring R-1
    __syncthreads();                             // every shared-memory atomic of the item has landed
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
    __threadfence();                             // release: this CTA's REDs are performed ...
    __syncthreads();
    if (tid == 0) {
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
                if (i < nvec) __stcs(reinterpret_cast<float4 *>(kq) + i, sc_decode4(q[d]));
            }
        }
    }
    __syncthreads();                             // s_last is rewritten by the next item
}

// (A flavour with the resident-row count as a template constant -- row-placement tests and clamp-row shift as
// immediates -- was measured 10 % SLOWER than this one, which takes `ns` at run time: profiles/r1_notes.md.)
//   spill       whole-scan items: [gridDim.x][spill_cap] (cell, key) pairs, one region per CTA, cut into per-warp lists
//   spill_cap   entries per CTA: >= ceil(max points per scan / NT) * NT
template <bool FAST, int VARIANT, bool FULL, int U, int NT, bool DEF, bool SPLIT>
__global__ void __launch_bounds__(NT, 1)
k1_scan(const uint8_t *__restrict__ pts, const int64_t *__restrict__ offsets, int n_scans,
        uint2 *__restrict__ spill, const size_t spill_cap, float *__restrict__ out, unsigned long long *__restrict__ stats,
        unsigned int *__restrict__ cursor, const int ns, const BinCfg cfg, const SplitArgs sp)
{
    // NB: this kernel lives on the edge of its register budget (64 regs at 1024 threads, 80 at 768); measured
    // variants that carried more state across the streaming loop (next-item prefetch before the decode,
    // per-warp chunk counters, 8 decode loads in flight) spilled inside the loop and lost 15-40 %
    // (profiles/r1_notes.md).
    constexpr int STEP = NT * U;
    constexpr bool SPILL = !SPLIT;
    extern __shared__ __align__(16) uint32_t smem_raw[];
    __shared__ int s_next[2];
    __shared__ int s_last;
    __shared__ ExactArgs s_ea;
    const int tid = threadIdx.x;
    const int R = DEF ? dsc::defgrid::R : cfg.R, S = DEF ? dsc::defgrid::S : cfg.S;
    const int cells = R * S;
    const int rowvec = S >> 2;                   // host guarantees S % 4 == 0
    const uint64_t pol_stream = l2_policy_evict_first();
    // dynamic shared memory: [ns ring rows of keys][per-warp exact-path queues]
    uint32_t *const srows = smem_raw;
    SmemRows sm;
    sm.base = (uint32_t)__cvta_generic_to_shared(srows);
    sm.lim = (ns - 1) * S;                       // host guarantees ns >= 2
    sm.last = (R - 1) * S;
    sm.shift = sm.lim - sm.last;
    const uint32_t queue = sm.base + (uint32_t)ns * (uint32_t)S * 4u + (uint32_t)(tid >> 5) * SCQ_WARP_BYTES;
    unsigned qcount = 0;
    for (int i = tid; i < ns * rowvec; i += NT) reinterpret_cast<uint4 *>(smem_raw)[i] = make_uint4(0, 0, 0, 0);
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
        // gp: split items -- the scan's key grid (== its output grid); whole scans -- write pointer of this warp's spill list
        uint32_t *gp;
        const uint2 *wseg = nullptr;
        if constexpr (SPLIT) {
            gp = reinterpret_cast<uint32_t *>(out + (size_t)scan * cells);
        } else {
            const unsigned seg = (unsigned)((n + NT - 1) / NT) * 32u;           // most points one warp can see
            wseg = spill + (size_t)blockIdx.x * spill_cap + (size_t)(tid >> 5) * seg;
            gp = reinterpret_cast<uint32_t *>(const_cast<uint2 *>(wseg));
        }
        const float4 *s4 = reinterpret_cast<const float4 *>(pts + p0 * 16);
        unsigned n_clamped = 0;
        unsigned long long *st = stats ? stats + (int64_t)scan * DSC_STATS_WORDS : nullptr;
        const int nfull = n / STEP;
        const float4 *p = s4 + tid;
        float4 a[U], b[U];
#define SC_LD(ptr) ld_stream_f4_hint((ptr), pol_stream)
#define SC_PT(v, act)                                                                                           \
    do {                                                                                                        \
        if constexpr (FAST) sc_point<VARIANT, FULL, DEF, SPILL>(cfg, &s_ea, (v), (act), gp, n_clamped, st, sm, queue, qcount); \
        else sc_point_exact<VARIANT, FULL, SPILL>(&s_ea, (v), (act), gp, n_clamped, st, sm);                   \
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
        if constexpr (FAST) sc_drain<VARIANT, FULL, SPILL>(&s_ea, gp, n_clamped, st, sm, queue, qcount);
#undef SC_LD
#undef SC_PT
        if (stats) {
            const unsigned c = __reduce_add_sync(0xffffffffu, n_clamped);
            if ((tid & 31) == 0 && c) atomicAdd(st + 3, (unsigned long long)c);
            if (tid == 0 && n) atomicAdd(st + 0, (unsigned long long)n);
        }
        if constexpr (SPLIT)
            sc_finish_split<NT, SC_DU>(srows, gp, ns, R, S, scan, parts, sp.done, &s_last);
        else
            sc_finish_scan<NT>(srows, sm.base, out + (size_t)scan * cells, ns, R, S, wseg, reinterpret_cast<const uint2 *>(gp));
        par ^= 1;
        item = s_next[par];                      // published by the barriers of the finish
    }
}
