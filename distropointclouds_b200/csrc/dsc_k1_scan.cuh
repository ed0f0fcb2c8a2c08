// dsc_k1_scan.cuh -- K1 main path: k1_scan, the persistent scan / work-item kernel.
// Included inside the anonymous namespace of dsc_k1_build.cu, after dsc_k1_common.cuh.
#pragma once

// ---- K1 main path: persistent CTAs, one work item (a scan, or a slice of a big scan) at a time ------------------
// A CTA streams the item's points once (128-bit loads, two register sets ping-pong), bins each point once and
// folds it into the key grid.  The 460 KB grid does not fit one SM, so the reduction runs in ring bands:
//   pass 1 (the stream)  the hot ring rows -- a window of ns-1 rings starting at ring r0, plus the clamp ring R-1 --
//                        live in shared memory (LDS pre-check + ATOMS.MAX).  r0 is chosen per launch by k1_probe
//                        from a ring histogram of 64 Ki sampled points (a LiDAR sees nothing closer than ~3 m, so
//                        rings 0..20 are empty and the dense ground rings start behind them: the best window holds
//                        85 % of the points of the bench batch, the window at ring 0 only 70 %).  A point of any
//                        other ring is NOT reduced yet: its (cell, key) pair is appended to the warp's private
//                        spill list (one ballot, one coalesced 8-byte store per lane; the lists are ~100 KB per CTA
//                        and stay in L2).
//   write-out            the resident rows are decoded to floats (coalesced 128-bit streaming stores) and re-zeroed;
//   pass 2 (per band)    the same shared-memory rows now stand for the next band of up to ns of the remaining rings:
//                        every warp folds the band's entries of its own spill list in with the same LDS pre-check +
//                        ATOMS.MAX (L2 hits, 128-bit loads); the band is decoded and written out; repeat until
//                        every ring is done (two bands on the default grid).
// So every max-reduction is a shared-memory atomic, HBM sees each point once and each output cell once, there are
// no global atomics, no key slots to read back and re-zero, and no synchronisation between CTAs for whole-scan items.
// (Round 1 reduced the non-resident rings with RED.MAX on a private L2 slot per CTA: 72 M scattered L2 atomics per
// launch, a 282 KB read-back + re-zero per scan and 1.2x write traffic from evicted slot lines: profiles/r1_notes.md.)
#ifndef SC_DU
#define SC_DU 4                       // 128-bit L2 loads in flight per thread in the split-mode decode
#endif
#ifndef SC_UT
#define SC_UT 2                       // loads per thread in the masked register sets of an item's ragged tail (3 spills in the loop)
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
#ifndef SC_QCAP
#define SC_QCAP 64
#endif
constexpr int SCQ_CAP = SC_QCAP;                              // entries per warp (>= 32; below 64 a batch may run part-filled)
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

// max-reduce `key` into the shared-memory cell at `addr` when `doit`: LDS pre-check (most points are not a new maximum)
// + ATOMS.MAX, predicated instead of branched.
// `w` is the point's intensity word -- dead data, but it doubles as the destination of the LDS: that keeps its register
// occupied from the 128-bit point load until here.  Otherwise ptxas hands the (dead) fourth register of the in-flight
// load to some early instruction of the PREVIOUS point, which then waits for the whole load (a write-after-write on
// the long scoreboard: one POPC held 20 % of all stall samples in profiles/r2d) and defeats the register ping-pong.
__device__ __forceinline__ void smem_max_if(const uint32_t addr, const uint32_t key, const bool doit, uint32_t w)
{
    asm volatile("{\n"
                 ".reg .pred p, q;\n"
                 "setp.ne.u32 p, %3, 0;\n"
                 "@p ld.shared.u32 %0, [%1];\n"
                 "setp.gt.and.u32 q, %2, %0, p;\n"
                 "@q red.shared.max.u32 [%1], %2;\n"
                 "}" : "+r"(w) : "r"(addr), "r"(key), "r"((uint32_t)doit) : "memory");
}

__device__ __forceinline__ uint32_t sc_tok3(const float a, const float b, const float c)
{
    uint32_t t;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(t) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)), "r"(__float_as_uint(c)));
    return t;
}
__device__ __forceinline__ uint32_t sc_tok2(const float a, const float b)
{
    uint32_t t;
    asm volatile("xor.b32 %0, %1, %2;" : "=r"(t) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b)));
    return t;
}

// What the key grid holds for a height.  Full-height mode: the order-preserving key (negative heights count, 0 == empty
// == below everything).  Offset mode: only heights > 0 ever count (strict '>' against cells that start at 0), and for
// positive floats the raw bit pattern already orders like the value -- the "key" is the float itself, 0 == empty ==
// 0.0f, and neither the points nor the 115 200 cells written out per scan need a conversion.
template <bool FULL>
__device__ __forceinline__ uint32_t sc_key(const float zv)
{
    const uint32_t u = __float_as_uint(zv);
    return FULL ? (u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u)) : u;
}
// a height takes part in the max: never a NaN (scaner_version2.py:107); offset mode: cells start at 0, strict '>'
template <bool FULL>
__device__ __forceinline__ bool sc_counts(const float zv)
{
    return FULL ? (zv == zv) : (zv > 0.0f);
}

// one real use of every intensity word of a register set (see SC_TOK in k1_scan): lop3 over three words at a time
template <int U>
__device__ __forceinline__ uint32_t sc_tok_set(const float4 (&s)[U])
{
    uint32_t t = 0u;
#pragma unroll
    for (int u = 0; u + 2 < U; u += 3) t ^= sc_tok3(s[u].w, s[u + 1].w, s[u + 2].w);
    if (U % 3 == 2) t ^= sc_tok2(s[U - 2].w, s[U - 1].w);
    if (U % 3 == 1) t ^= __float_as_uint(s[U - 1].w);
    return t;
}

// Fold one binned point into the key grid.  Called warp-uniformly; `valid` masks lanes without a (counting) point.
//   resident cell      LDS pre-check + ATOMS.MAX
//   any other cell     SPILL: appended to the warp's spill list (`gbase` = the launch's spill buffer, `goff` = the warp's
//                      write position in it, in entries, advanced for the whole warp; `lt` = %lanemask_lt);
//                      !SPILL (split items): RED.MAX straight into the scan's key grid (`gbase`)
// shared-memory part of a commit; returns "this lane's point belongs to a ring that is not resident"
__device__ __forceinline__ bool sc_commit_rows(const uint32_t key, const int cell, const bool valid, const SmemRows sm, const uint32_t w)
{
    const int rel = cell - sm.c0;
    const bool win = (uint32_t)rel < (uint32_t)sm.lim;                          // rings r0 .. r0+ns-2
    const bool res = win || cell >= sm.last;
    const int local = win ? rel : cell + sm.shift;                              // clamp ring -> row ns-1
    smem_max_if(sm.base + (uint32_t)local * 4u, key, valid && res, w);
    return valid && !res;
}
// append the non-resident points of the warp (`sp` lanes) to its spill list; called warp-uniformly
__device__ __forceinline__ void sc_spill(const bool sp, const uint32_t key, const int cell, uint32_t *const gbase, uint32_t &goff,
                                         const unsigned lt)
{
    const unsigned m = __ballot_sync(0xffffffffu, sp);
    if (sp) st_spill(reinterpret_cast<uint2 *>(gbase) + (goff + __popc(m & lt)), (uint32_t)cell, key);
    goff += __popc(m);
}
template <bool SPILL>
__device__ __forceinline__ void sc_commit(const uint32_t key, const int cell, const bool valid, const SmemRows sm, uint32_t *const gbase,
                                          uint32_t &goff, const unsigned lt, const uint32_t w = 0u)
{
    const bool sp = sc_commit_rows(key, cell, valid, sm, w);
    if constexpr (SPILL) {
        // most warps have no spilling lane (consecutive points of a beam share their rings): one warp-uniform branch
        // around rank, address and store instead of six predicated instructions
        if (__any_sync(0xffffffffu, sp)) sc_spill(sp, key, cell, gbase, goff, lt);
    } else {
        if (sp) red_max_global(gbase + cell, key);
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

// run the exact chain for one queue entry per lane (`qaddr`, lanes with `mine`) and commit the results; warp-uniform
template <int VARIANT, bool FULL, bool SPILL>
__device__ __forceinline__ void sc_exact_lanes(const uint32_t qaddr, const bool mine, const ExactArgs *ea, uint32_t *const gbase,
                                               uint32_t &goff, unsigned &n_clamped, unsigned long long *__restrict__ st,
                                               const SmemRows sm, const unsigned lt)
{
    uint2 b = make_uint2((uint32_t)dsc::PT_NONFINITE << 27, 0u);
    __syncwarp();
    if (mine) b = sc_exact_queued<VARIANT, FULL>(qaddr, ea, st);
    __syncwarp();
    const int status = (int)(b.x >> 27);
    n_clamped += (mine && status == dsc::PT_CLAMPED) ? 1u : 0u;
    const float zv = __uint_as_float(b.y);
    sc_commit<SPILL>(sc_key<FULL>(zv), (int)(b.x & 0x7ffffffu), mine && status <= dsc::PT_CLAMPED && sc_counts<FULL>(zv), sm, gbase, goff, lt);
}
// ... for the queue entries [first, first + count) of this warp (count <= 32)
template <int VARIANT, bool FULL, bool SPILL>
__device__ __forceinline__ void sc_exact_batch(const uint32_t queue, const unsigned first, const unsigned count,
                                               const ExactArgs *ea, uint32_t *const gbase, uint32_t &goff, unsigned &n_clamped,
                                               unsigned long long *__restrict__ st, const SmemRows sm, const unsigned lt)
{
    const unsigned lane = threadIdx.x & 31u;
    sc_exact_lanes<VARIANT, FULL, SPILL>(queue + (first + lane) * 4u, lane < count, ea, gbase, goff, n_clamped, st, sm, lt);
}

// Park the points that failed the screen (`slow` lanes) in the warp's exact-path queue; run the exact chain when 32 wait.
// Called warp-uniformly, on the rare path only.
template <int VARIANT, bool FULL, bool SPILL, bool STATS>
__device__ __forceinline__ void sc_park(const ExactArgs *ea, const float4 v, const bool slow, uint32_t *const gbase, uint32_t &goff,
                                        unsigned &n_clamped, unsigned long long *__restrict__ stats, const int scan,
                                        const SmemRows sm, const uint32_t qoff, unsigned &qcount, const unsigned lt)
{
    // (the thread index is read here, opaquely, so that the queue address is formed on this path instead of being
    // hoisted out of the streaming loop into a register of its own)
    unsigned t;
    asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
    const uint32_t queue = sm.base + qoff + (t >> 5) * SCQ_WARP_BYTES;
    const unsigned m = __ballot_sync(0xffffffffu, slow);
    if (SCQ_CAP < 64 && qcount + __popc(m) > (unsigned)SCQ_CAP) {               // no room: run what is waiting, part-filled
        unsigned long long *st = STATS ? stats + (int64_t)scan * DSC_STATS_WORDS : nullptr;
        sc_exact_batch<VARIANT, FULL, SPILL>(queue, 0u, qcount, ea, gbase, goff, n_clamped, st, sm, lt);
        qcount = 0;
    }
    if (slow) {
        const uint32_t a = queue + (qcount + __popc(m & lt)) * 4u;
        sts_f32(a, v.x);
        sts_f32(a + SCQ_CAP * 4, v.y);
        sts_f32(a + SCQ_CAP * 8, v.z);
    }
    qcount += __popc(m);
    if (qcount >= 32u) {
        qcount -= 32u;
        unsigned long long *st = STATS ? stats + (int64_t)scan * DSC_STATS_WORDS : nullptr;
        sc_exact_batch<VARIANT, FULL, SPILL>(queue, qcount, 32u, ea, gbase, goff, n_clamped, st, sm, lt);
    }
}

// One register set -- U points per lane -- called warp-uniformly; the lane's first `mine` points exist (all U in the
// streaming loop, fewer in the item's last, ragged set).
// Point by point: screen, commit; then ONE vote per set for the points that failed the screen.
// (Screening the U points first as one branch-free block, so that ptxas interleaves the U dependent chains, and
// committing afterwards was measured 6 % slower -- the (cell, key, valid) triples stay live, predicates get packed into
// registers and the next set's loads issue later -- and so was ONE screen vote per set, which removes the branch between
// the points and lets ptxas interleave them by itself: profiles/r2_notes.md, r2k/r2l.)
// `qoff`: byte offset of the exact-path queues from the rows; the warp's own queue address and the scan's counter block
// are only formed on the rare paths.  STATS = false: no per-scan counters are kept at all (stats == NULL).
template <int VARIANT, bool FULL, bool DEF, bool SPILL, bool STATS, int U>
__device__ __forceinline__ void sc_set(const BinCfg &cfg, const ExactArgs *ea, const float4 (&v)[U], const int mine,
                                       uint32_t *const gbase, uint32_t &goff, unsigned &n_clamped, unsigned long long *__restrict__ stats,
                                       const int scan, const SmemRows sm, const uint32_t qoff, unsigned &qcount,
                                       const unsigned lt, const uint32_t wtok = 0u)
{
    const int S = DEF ? dsc::defgrid::S : cfg.S;
    bool slow[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
        int ring, sector;
        bool beyond;
        const bool active = u < mine;
        bool ok = dsc::screened_core_t<DEF>(cfg, v[u].x, v[u].y, 0.0f, 0.0f, &ring, &sector, &beyond);
        if (VARIANT == 1) ok = ok && (fabsf(v[u].z) < INFINITY);    // finite mask (scan_context_node.py:115)
        bool take = ok && active;
        if (VARIANT == 1) {
            if (STATS && take && beyond)                         // r > max_length: dropped (scan_context_node.py:134-137)
                atomicAdd(stats + (int64_t)scan * DSC_STATS_WORDS + 4, 1ull);
            take = take && !beyond;
        } else if (STATS) {
            n_clamped += (take && beyond) ? 1u : 0u;
        }
        const float zv = FULL ? v[u].z : __fadd_rn(v[u].z, cfg.offset);
        slow[u] = active && !ok;
        {   // ONE vote per point for "some lane spills or failed the screen": most warps have neither (consecutive points
            // of a beam share their rings; 0.3 % of the points fail the screen)
            const uint32_t key = sc_key<FULL>(zv);
            const int cell = ring * S + sector;
            const bool sp = sc_commit_rows(key, cell, take && sc_counts<FULL>(zv), sm, u == 0 ? wtok : 0u);
            if constexpr (!SPILL) {
                if (sp) red_max_global(gbase + cell, key);
            }
            if (__any_sync(0xffffffffu, (SPILL && sp) || slow[u])) {
                if constexpr (SPILL) sc_spill(sp, key, cell, gbase, goff, lt);
                if (__any_sync(0xffffffffu, slow[u]))
                    sc_park<VARIANT, FULL, SPILL, STATS>(ea, v[u], slow[u], gbase, goff, n_clamped, stats, scan, sm, qoff, qcount, lt);
            }
        }
    }
}

// exact-only flavour (DSC_FLAG_EXACT_ONLY, non-dyadic gaps): every point takes the fdlibm chain
template <int VARIANT, bool FULL, bool SPILL>
__device__ __forceinline__ void sc_point_exact(const ExactArgs *ea, const float4 v, const bool active, uint32_t *const gbase, uint32_t &goff,
                                               unsigned &n_clamped, unsigned long long *__restrict__ stats, const int scan,
                                               const SmemRows sm, const unsigned lt)
{
    unsigned long long *st = stats ? stats + (int64_t)scan * DSC_STATS_WORDS : nullptr;
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
    const float zv = __uint_as_float(b.y);
    sc_commit<SPILL>(sc_key<FULL>(zv), (int)(b.x & 0x7ffffffu), active && status <= dsc::PT_CLAMPED && sc_counts<FULL>(zv), sm, gbase, goff, lt);
}

// End of item: the queues' remaining entries (fewer than 32 per warp).  Each warp draining its own ~13 entries would run
// the ~300-instruction exact chain 24 times per scan, half empty -- 70 % of all its executions.  Groups of G warps pool
// what they have left instead (a named barrier over the group, counts through shared memory) and deal the pooled
// entries out in batches of 32, round-robin: ceil(total / 32) executions per group, still in parallel.
// A warp may so commit up to 32 points that it did not stream itself: its spill list has one more row of 32 entries
// than its own points need (see `seg` in k1_scan).
template <int VARIANT, bool FULL, bool SPILL, int NT>
__device__ __forceinline__ void sc_drain(const ExactArgs *ea, uint32_t *const gbase, uint32_t &goff, unsigned &n_clamped,
                                         unsigned long long *__restrict__ st, const SmemRows sm, const uint32_t qoff,
                                         unsigned &qcount, const unsigned lt, int *__restrict__ s_qcnt)
{
    constexpr int W = NT / 32;
    constexpr int G = (W % 4 == 0) ? 4 : (W % 2 == 0) ? 2 : 1;
    const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    if constexpr (G == 1) {
        if (qcount) {                            // warp-uniform
            const uint32_t queue = sm.base + qoff + warp * SCQ_WARP_BYTES;
            sc_exact_batch<VARIANT, FULL, SPILL>(queue, 0u, qcount, ea, gbase, goff, n_clamped, st, sm, lt);
        }
    } else {
        const unsigned g = warp / G, wi = warp % G;
        if (lane == 0) s_qcnt[warp] = (int)qcount;
        asm volatile("bar.sync %0, %1;" ::"r"(1u + g), "r"((unsigned)(G * 32)) : "memory");    // barrier 0 is __syncthreads'
        int c[G], total = 0;
#pragma unroll
        for (int k = 0; k < G; ++k) {
            c[k] = s_qcnt[g * G + k];
            total += c[k];
        }
        const uint32_t gq = sm.base + qoff + g * G * SCQ_WARP_BYTES;            // the group's first queue
        for (int b = (int)wi; b * 32 < total; b += G) {                         // group-uniform bounds, warp-uniform trip count
            int j = b * 32 + (int)lane;
            const bool mine = j < total;
            int q = 0;
#pragma unroll
            for (int k = 0; k + 1 < G; ++k)
                if (q == k && j >= c[k]) { j -= c[k]; q = k + 1; }
            sc_exact_lanes<VARIANT, FULL, SPILL>(gq + (uint32_t)q * SCQ_WARP_BYTES + (uint32_t)j * 4u, mine, ea, gbase, goff, n_clamped,
                                                 st, sm, lt);
        }
    }
    qcount = 0;
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

template <bool FULL>
__device__ __forceinline__ float4 sc_decode4(const uint4 q)
{
    float4 f;
    if constexpr (FULL) {
        f.x = dsc::key_to_float(q.x);
        f.y = dsc::key_to_float(q.y);
        f.z = dsc::key_to_float(q.z);
        f.w = dsc::key_to_float(q.w);
    } else {                                     // offset mode: the keys are the floats (sc_key)
        f = make_float4(__uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z), __uint_as_float(q.w));
    }
    return f;
}

// decode `nvec` 128-bit groups of keys from shared memory to floats, re-zeroing the rows.  (Reading 2 or 4 groups per
// thread before the first is decoded was measured 0.4 % / 1.4 % slower than this plain loop, which ptxas unrolls itself.)
template <int NT, bool FULL>
__device__ __forceinline__ void sc_flush_rows(uint32_t *__restrict__ srows, float4 *__restrict__ dst, int nvec)
{
    uint4 *const s4 = reinterpret_cast<uint4 *>(srows);
    for (int i = threadIdx.x; i < nvec; i += NT) {
        const uint4 q = s4[i];
        s4[i] = make_uint4(0, 0, 0, 0);
        __stcs(dst + i, sc_decode4<FULL>(q));
    }
}

// fold one spill entry into the band that the shared-memory rows currently stand for: the band is the cell range
// [lo1, lo1 + n1) followed by [lo2, lo2 + n2) (the remaining rings are the two intervals either side of the window)
__device__ __forceinline__ void sc_band_entry(const uint32_t cell, const uint32_t key, const uint32_t sbase,
                                              const uint32_t lo1, const uint32_t n1, const uint32_t lo2, const uint32_t n2,
                                              const bool present)
{
    // branch-free: both interval tests, one predicated LDS pre-check + ATOMS.MAX
    const uint32_t l1 = cell - lo1, l2 = cell - lo2;
    const bool in1 = l1 < n1, in2 = l2 < n2;
    const uint32_t l = in1 ? l1 : l2 + n1;
    smem_max_if(sbase + l * 4u, key, present && (in1 || in2), 0u);
}

// End of a whole-scan item (this CTA owns the scan): write out the resident rows, then reduce and write out the other
// rings band by band from the spill lists.  `wseg`: this warp's list of `len` entries.  Ends with a CTA barrier.
template <int NT, bool FULL>
__device__ __forceinline__ void sc_finish_scan(uint32_t *__restrict__ srows, const uint32_t sbase,
                                               float *__restrict__ out_scan, const int ns, const int r0, const int R,
                                               const int S, const uint2 *__restrict__ wseg, const int len)
{
    const int rowvec = S >> 2;
    float4 *oq = reinterpret_cast<float4 *>(out_scan);
    const int lane2 = 2 * (int)(threadIdx.x & 31u);
    // this warp's list is read two entries per 128-bit load (L2 hits), SC_LU loads per lane and batch
    auto load_batch = [&](uint4(&q)[SC_LU], const int i0) {
#pragma unroll
        for (int d = 0; d < SC_LU; ++d) {
            q[d] = make_uint4(0u, 0u, 0u, 0u);
            if (i0 + 64 * d < len) q[d] = ld_spill2(wseg + i0 + 64 * d);
        }
    };
    __syncthreads();                             // every shared-memory atomic and every list store of the item is done
    // resident rows: rings r0 .. r0+ns-2, and the clamp ring R-1 kept in row ns-1
    sc_flush_rows<NT, FULL>(srows, oq + (size_t)r0 * rowvec, (ns - 1) * rowvec);
    sc_flush_rows<NT, FULL>(srows + (size_t)(ns - 1) * S, oq + (size_t)(R - 1) * rowvec, rowvec);
    // remaining rings, numbered v = 0 .. R-ns-1: ring v below the window, ring v + ns - 1 above it
    const int nv = R - ns;
    for (int b0 = 0; b0 < nv; b0 += ns) {
        const int b1 = b0 + ns < nv ? b0 + ns : nv;
        const int rows1 = b0 < r0 ? (b1 < r0 ? b1 : r0) - b0 : 0;               // rings b0 .. b0+rows1-1
        const int rows2 = (b1 - b0) - rows1;                                     // rings ring2 .. ring2+rows2-1
        const int ring2 = (b0 > r0 ? b0 : r0) + ns - 1;
        const uint32_t lo1 = (uint32_t)b0 * (uint32_t)S, n1 = (uint32_t)rows1 * (uint32_t)S;
        const uint32_t lo2 = (uint32_t)ring2 * (uint32_t)S, n2 = (uint32_t)rows2 * (uint32_t)S;
        __syncthreads();                         // the rows are zero again
        for (int i0 = lane2; i0 - lane2 < len; i0 += 64 * SC_LU) {              // warp-uniform trip count
            uint4 q[SC_LU];
            load_batch(q, i0);
#pragma unroll
            for (int d = 0; d < SC_LU; ++d) {
                const int i = i0 + 64 * d;
                sc_band_entry(q[d].x, q[d].y, sbase, lo1, n1, lo2, n2, i < len);
                sc_band_entry(q[d].z, q[d].w, sbase, lo1, n1, lo2, n2, i + 1 < len);
            }
        }
        __syncthreads();
        sc_flush_rows<NT, FULL>(srows, oq + (size_t)b0 * rowvec, rows1 * rowvec);
        sc_flush_rows<NT, FULL>(srows + (size_t)rows1 * S, oq + (size_t)ring2 * rowvec, rows2 * rowvec);
    }
    __syncthreads();                             // rows are zero before the next scan's atomics
}

// End of a split work item: fold the resident rows into the scan's key grid (== its output grid), count the item as
// retired, and if it was the scan's last item decode the grid in place.  Ends with a CTA barrier.
template <int NT, bool FULL, int DU = 4>
__device__ __forceinline__ void sc_finish_split(uint32_t *__restrict__ srows, uint32_t *__restrict__ grid, int ns, int r0,
                                                int R, int S, int scan, int parts, unsigned int *__restrict__ done,
                                                int *s_last)
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
        uint32_t *g = grid + 4 * (size_t)(i < lim ? i + r0 * rowvec : i + shift);
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
    if (FULL && *s_last) {                       // (offset mode: the keys are the floats already)
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
                if (i < nvec) __stcs(reinterpret_cast<float4 *>(kq) + i, sc_decode4<true>(q[d]));
            }
        }
    }
    __syncthreads();                             // s_last is rewritten by the next item
}

// (A flavour with the resident-row count as a template constant -- row-placement tests and clamp-row shift as
// immediates -- was measured 10 % SLOWER than this one, which takes `ns` at run time: profiles/r1_notes.md.)
//   spill       whole-scan items: [gridDim.x][spill_cap] (cell, key) pairs, one region per CTA, cut into per-warp lists
//   spill_cap   entries per CTA: >= ceil(max points per scan / NT) * NT
//   window      device int: first ring of the resident window, written by k1_probe earlier in the stream
template <bool FAST, int VARIANT, bool FULL, int U, int NT, bool DEF, bool SPLIT, bool STATS>
__global__ void __launch_bounds__(NT, 1)
k1_scan(const uint8_t *__restrict__ pts, const int64_t *__restrict__ offsets, int n_scans,
        uint2 *__restrict__ spill, const size_t spill_cap, float *__restrict__ out, unsigned long long *__restrict__ stats,
        unsigned int *__restrict__ cursor, const int ns, const int *__restrict__ window, const BinCfg cfg,
        const SplitArgs sp)
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
    __shared__ int s_qcnt[NT / 32];
    __shared__ ExactArgs s_ea;
    const int tid = threadIdx.x;
    const int R = DEF ? dsc::defgrid::R : cfg.R, S = DEF ? dsc::defgrid::S : cfg.S;
    const int cells = R * S;
    const int rowvec = S >> 2;                   // host guarantees S % 4 == 0
    const uint64_t pol_stream = l2_policy_evict_first();
    // dynamic shared memory: [ns ring rows of keys][per-warp exact-path queues]
    uint32_t *const srows = smem_raw;
    const int r0 = *window;                      // first ring of the resident window (k1_probe): 0 <= r0 <= R - ns
    SmemRows sm;
    sm.base = (uint32_t)__cvta_generic_to_shared(srows);
    // keep both in registers: passed through a shuffle, which ptxas cannot re-derive in the loop (re-reading
    // SR_LTMASK or rebuilding the shared-window address costs 1-5 instructions per point in an issue-bound loop)
    sm.base = __shfl_sync(0xffffffffu, sm.base, 0);
    const unsigned lt = __shfl_sync(0xffffffffu, lanemask_lt(), (int)(tid & 31));
    sm.c0 = __shfl_sync(0xffffffffu, r0 * S, 0);     // (same: otherwise re-derived from r0 with two instructions per point)
    sm.lim = (ns - 1) * S;                       // host guarantees ns >= 2
    sm.last = (R - 1) * S;
    sm.shift = sm.lim - sm.last;
    const uint32_t qoff = (uint32_t)ns * (uint32_t)S * 4u;
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
        // gbase / goff: split items -- the scan's key grid (== its output grid); whole scans -- this CTA's spill region
        // (CTA-uniform: lives in uniform registers) and this warp's write position in it (one register)
        uint32_t *gbase;
        uint32_t goff = 0;
        {   // ---- stream phase.  Nothing but the loop state may stay live here (80 registers): the item's facts
            // (scan, first point, count) are recomputed for the finish instead of being carried through the loop.
            int scan, parts, n;
            int64_t p0;
            sc_item<SPLIT>(item, offsets, n_scans, sp, scan, parts, p0, n);
            if constexpr (SPLIT) {
                gbase = reinterpret_cast<uint32_t *>(out + (size_t)scan * cells);
            } else {
                // the write position counts entries from the start of the launch's spill buffer (host guarantees
                // gridDim.x * spill_cap < 2^32): the store address is one multiply-add away from it
                const unsigned seg = (unsigned)((n + NT - 1) / NT + 1) * 32u;   // most entries one warp can append (see sc_drain)
                gbase = reinterpret_cast<uint32_t *>(spill);
                goff = blockIdx.x * (uint32_t)spill_cap + (uint32_t)(tid >> 5) * seg;
            }
            // the stream position is a 32-bit index of 16-byte rows from `pts` (host guarantees < 2^32 rows per call):
            // one register across the loop instead of a 64-bit pointer
            const float4 *const pts4 = reinterpret_cast<const float4 *>(pts);
            unsigned n_clamped = 0;
            const int nfull = n / STEP;
            uint32_t p = (uint32_t)p0 + (uint32_t)tid;
            float4 a[U], b[U];
#define SC_LD(idx) ld_stream_f4_hint(pts4 + (idx), pol_stream)
#define SC_SET(v, tok)                                                                                          \
    do {                                                                                                        \
        if constexpr (FAST) {                                                                                   \
            sc_set<VARIANT, FULL, DEF, SPILL, STATS, U>(cfg, &s_ea, (v), U, gbase, goff, n_clamped, stats, scan, sm, qoff, qcount, lt, (tok)); \
        } else {                                                                                                \
            _Pragma("unroll") for (int u = 0; u < U; ++u)                                                       \
                sc_point_exact<VARIANT, FULL, SPILL>(&s_ea, (v)[u], true, gbase, goff, n_clamped, stats, scan, sm, lt); \
        }                                                                                                       \
    } while (0)
            // SC_TOK: one real use of the (dead) intensity words of a register set when its processing starts.  Without it
            // ptxas hands those registers -- the fourth destination registers of 128-bit loads that are still in
            // flight -- to instructions of the set processed BEFORE, which then wait for the whole load (write-after-
            // write on the long scoreboard: one POPC held 20 % of all stall samples in profiles/r2d_*) and the register
            // ping-pong is defeated.  The XOR of the words becomes the scratch destination of the first point's LDS.
#define SC_TOK(s) sc_tok_set<U>(s)
            if (nfull > 0) {
#pragma unroll
                for (int u = 0; u < U; ++u) a[u] = SC_LD(p + u * NT);
            }
            int left = nfull;                    // register sets still to process (one counter: the loop lives on the
            while (left >= 2) {                  // edge of its register budget)
#pragma unroll
                for (int u = 0; u < U; ++u) b[u] = SC_LD(p + STEP + u * NT);
                const uint32_t ta = SC_TOK(a);
                SC_SET(a, ta);
                p += 2 * STEP;
                left -= 2;
                if (left > 0) {
#pragma unroll
                    for (int u = 0; u < U; ++u) a[u] = SC_LD(p + u * NT);
                }
                const uint32_t tb = SC_TOK(b);
                SC_SET(b, tb);
            }
            if (left > 0) {
                const uint32_t ta = SC_TOK(a);
                SC_SET(a, ta);
            }
#undef SC_TOK
#undef SC_SET
            {   // ragged tail (fewer than STEP points): masked register sets of UT loads, all of a set's loads in flight at once
                constexpr int UT = U < SC_UT ? U : SC_UT;
                for (int i0 = nfull * STEP; i0 < n; i0 += UT * NT) {                    // CTA-uniform trip count
                    const int rem = n - i0;
                    const int mine = rem > tid ? (rem - tid + NT - 1) / NT : 0;         // >= UT: all of this lane's points exist
                    const uint32_t pt = (uint32_t)p0 + (uint32_t)(i0 + tid);
                    float4 t[UT];
#pragma unroll
                    for (int u = 0; u < UT; ++u) {
                        t[u] = make_float4(1.0f, 1.0f, 0.0f, 0.0f);
                        if (u < mine) t[u] = SC_LD(pt + u * NT);
                    }
                    if constexpr (FAST) {
                        sc_set<VARIANT, FULL, DEF, SPILL, STATS, UT>(cfg, &s_ea, t, mine, gbase, goff, n_clamped, stats, scan, sm, qoff, qcount, lt);
                    } else {
#pragma unroll
                        for (int u = 0; u < UT; ++u)
                            sc_point_exact<VARIANT, FULL, SPILL>(&s_ea, t[u], u < mine, gbase, goff, n_clamped, stats, scan, sm, lt);
                    }
                }
            }
#undef SC_LD
            unsigned long long *st = STATS ? stats + (int64_t)scan * DSC_STATS_WORDS : nullptr;
            if constexpr (FAST) sc_drain<VARIANT, FULL, SPILL, NT>(&s_ea, gbase, goff, n_clamped, st, sm, qoff, qcount, lt, s_qcnt);
            if (STATS) {
                const unsigned c = __reduce_add_sync(0xffffffffu, n_clamped);
                if ((tid & 31) == 0 && c) atomicAdd(st + 3, (unsigned long long)c);
                if (tid == 0 && n) atomicAdd(st + 0, (unsigned long long)n);
            }
        }
        {   // ---- finish phase
            int item2 = item;
            asm volatile("" : "+r"(item2));      // opaque: do not carry scan / n / p0 through the stream phase in registers
            int scan, parts, n;
            int64_t p0;
            sc_item<SPLIT>(item2, offsets, n_scans, sp, scan, parts, p0, n);
            if constexpr (SPLIT) {
                sc_finish_split<NT, FULL, SC_DU>(srows, gbase, ns, r0, R, S, scan, parts, sp.done, &s_last);
            } else {
                const unsigned seg = (unsigned)((n + NT - 1) / NT + 1) * 32u;
                const uint32_t w0 = blockIdx.x * (uint32_t)spill_cap + (uint32_t)(tid >> 5) * seg;
                sc_finish_scan<NT, FULL>(srows, sm.base, out + (size_t)scan * cells, ns, r0, R, S, spill + w0, (int)(goff - w0));
            }
        }
        par ^= 1;
        item = s_next[par];                      // published by the barriers of the finish
    }
}

// ---- k1_probe: where to put the resident window -------------------------------------------------------------
// Ring histogram of PROBE_BLOCKS * PROBE_THREADS * PROBE_PPT pseudo-randomly sampled points of the batch; the last
// block to finish picks the start ring r0 of the window of `nwin` rings that holds the most samples (ring R-1, the
// clamp ring, is resident anyway and not part of the search) and writes it to *window.  A heuristic only: the grids
// do not depend on r0.  probe[] = ring histogram [R], then the finished-blocks counter: zero when the kernel starts
// (zeroed once when the buffer is allocated) and zero again when it ends -- the last block cleans up after itself, and
// block 0 resets the work cursor of the k1_scan launch that follows, so a call needs no memset of its own.
constexpr int PROBE_BLOCKS = 32, PROBE_THREADS = 256, PROBE_PPT = 8, PROBE_MAXR = 4096;

__global__ void __launch_bounds__(PROBE_THREADS)
k1_probe(const uint8_t *__restrict__ pts, const int64_t first, const int64_t total, const float inv_gap_ring, const int R,
         const int nwin, unsigned int *__restrict__ probe, int *__restrict__ window, unsigned int *__restrict__ cursor)
{
    __shared__ unsigned int sh[PROBE_MAXR];
    __shared__ unsigned long long s_best[PROBE_THREADS / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    if (blockIdx.x == 0 && tid == 0 && cursor) *cursor = 0u;
    for (int i = tid; i < R; i += PROBE_THREADS) sh[i] = 0u;
    __syncthreads();
    const float4 *p4 = reinterpret_cast<const float4 *>(pts) + first;
#pragma unroll
    for (int k = 0; k < PROBE_PPT; ++k) {
        const uint32_t g = (uint32_t)((blockIdx.x * PROBE_THREADS + tid) * PROBE_PPT + k);
        const uint64_t idx = ((uint64_t)(g * 2654435761u) * (uint64_t)total) >> 32;       // uniform over the batch
        const float4 v = ld_stream_f4(p4 + idx);
        const float r = sqrtf(v.x * v.x + v.y * v.y) * inv_gap_ring;
        if (r == r) atomicAdd(&sh[r >= (float)(R - 1) ? R - 1 : (int)r], 1u);
    }
    __syncthreads();
    for (int i = tid; i < R; i += PROBE_THREADS)
        if (sh[i]) atomicAdd(probe + i, sh[i]);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        s_last = atomicAdd(probe + R, 1u) == gridDim.x - 1;
        if (s_last) __threadfence();
    }
    __syncthreads();
    if (!s_last) return;
    for (int i = tid; i < R; i += PROBE_THREADS) {
        sh[i] = __ldcg(probe + i);
        probe[i] = 0u;                                            // leave the histogram clean for the next call
    }
    if (tid == 0) probe[R] = 0u;
    __syncthreads();
    // windows [s, s + nwin) below the clamp ring: every thread sums its candidates, the best (count, lowest start) wins
    unsigned long long best = 0ull;                               // (count << 32) | ~start: max = most samples, lowest start
    for (int s0 = tid; s0 + nwin <= R - 1; s0 += PROBE_THREADS) {   // (host guarantees nwin <= R - 2: s0 = 0 is a candidate)
        unsigned int c = 0;
        for (int i = s0; i < s0 + nwin; ++i) c += sh[i];
        const unsigned long long cand = ((unsigned long long)c << 32) | (unsigned int)~s0;
        best = cand > best ? cand : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        const unsigned long long o = __shfl_xor_sync(0xffffffffu, best, d);
        best = o > best ? o : best;
    }
    if ((tid & 31) == 0) s_best[tid >> 5] = best;
    __syncthreads();
    if (tid == 0) {
        for (int w = 1; w < PROBE_THREADS / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
        *window = (int)~(unsigned int)(best & 0xffffffffull);
    }
}
