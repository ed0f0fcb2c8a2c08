// dsc_k23_cells.cu -- K2 box rasteriser, K3 grid->points reconstructor, K5 covered-cell helpers.
//
// All three are ordered stream compactions over the (ring, sector) grid with float64 payloads:
//   K2  cells_covered_by_box          kitti_dataset.py:204-253
//   K3  scan_context_to_pointcloud    reconstruct_pointcloud.py:47-77 (loop twin scaner_version2.py:216-262)
//   K5  add_covered_cells_to_pointcloud / np.isin recolour   reconstruct_pointcloud.py:104-130, :153-164
//
// The trigonometry (cos/sin of the cell-centre angles and of the box yaw) is done by the caller
// with NumPy, exactly as the reference does it, and handed over as tables; the kernels only use
// individually rounded f64 mul/add/sub/compare, so outputs are bit-identical to the reference on
// any host.  One warp owns one (item, ring) row: pass A counts, a two-level exclusive scan turns
// counts into CSR offsets, pass C re-evaluates the predicate and writes in row-major order.
#include "dsc_internal.h"

namespace {

constexpr int ROW_WARPS = 8;  // warps per CTA in the row kernels

// ---- K2 predicate -----------------------------------------------------------------------------
struct Box { double x, y, l, w, c, s; };

__device__ __forceinline__ bool box_covers(const Box &b, double rc, double ct, double st)
{
    const double Xc = __dmul_rn(rc, ct);                    // :228
    const double Yc = __dmul_rn(rc, st);                    // :229
    const double X = __dsub_rn(Xc, b.x);                    // :232
    const double Y = __dsub_rn(Yc, b.y);
    const double Xl = __dadd_rn(__dmul_rn(b.c, X), __dmul_rn(b.s, Y));    // :239
    const double Yl = __dadd_rn(__dmul_rn(-b.s, X), __dmul_rn(b.c, Y));   // :240
    return fabs(Xl) <= __dmul_rn(0.5, b.l) && fabs(Yl) <= __dmul_rn(0.5, b.w);   // :242-245
}

// conservative ring cull: a covered cell centre P satisfies | |P| - |C| | <= half diagonal
__device__ __forceinline__ bool ring_may_touch(const Box &b, double rc)
{
    const double d = sqrt(b.x * b.x + b.y * b.y);
    const double h = 0.5 * sqrt(b.l * b.l + b.w * b.w);
    const double m = 1e-6 * (1.0 + d + h);
    return !(rc < d - h - m) && !(rc > d + h + m);          // NaNs fall through to the full test
}

template <bool WRITE>
__global__ void __launch_bounds__(ROW_WARPS * 32)
k2_rows(const double *__restrict__ boxes, int n_boxes, int R, int S, const double *__restrict__ rcs,
        const double *__restrict__ cosT, const double *__restrict__ sinT, int32_t *__restrict__ row_counts,
        const int64_t *__restrict__ row_offsets, int32_t *__restrict__ cells_out, int64_t cap,
        uint8_t *__restrict__ dense)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= (int64_t)n_boxes * R) return;
    const int bi = (int)(row / R), ring = (int)(row % R);
    if (WRITE && row_counts[row] == 0) return;
    const double *bp = boxes + (int64_t)bi * 6;
    const Box b = {bp[0], bp[1], bp[2], bp[3], bp[4], bp[5]};
    const double rc = rcs[ring];
    int count = 0;
    if (ring_may_touch(b, rc)) {
        int64_t base = WRITE ? row_offsets[row] : 0;
        for (int j0 = 0; j0 < S; j0 += 32) {
            const int j = j0 + lane;
            const bool in = j < S && box_covers(b, rc, cosT[j], sinT[j]);
            const unsigned m = __ballot_sync(0xffffffffu, in);
            if (WRITE) {
                if (in) {
                    const int64_t o = base + __popc(m & ((1u << lane) - 1u));
                    if (o < cap) { cells_out[2 * o] = ring; cells_out[2 * o + 1] = j; }
                    if (dense) dense[(int64_t)ring * S + j] = 1;
                }
                base += __popc(m);
            } else {
                count += __popc(m);
            }
        }
    }
    if (!WRITE && lane == 0) row_counts[row] = count;
}

// ---- K3 predicate: Z != 0 (full mode) or Z > 0 (offset mode)  :61-66 ----------------------------
template <bool WRITE>
__global__ void __launch_bounds__(ROW_WARPS * 32)
k3_rows(const float *__restrict__ grids, int n_grids, int R, int S, int full_height, float offset_f,
        const double *__restrict__ rcs, const double *__restrict__ cosT, const double *__restrict__ sinT,
        int32_t *__restrict__ row_counts, const int64_t *__restrict__ row_offsets,
        double *__restrict__ pts_out, int32_t *__restrict__ ids_out, int64_t cap)
{
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (row >= (int64_t)n_grids * R) return;
    const int ring = (int)(row % R);
    if (WRITE && row_counts[row] == 0) return;
    const float *g = grids + row * S;
    const double rc = rcs[ring];
    int count = 0;
    int64_t base = WRITE ? row_offsets[row] : 0;
    for (int j0 = 0; j0 < S; j0 += 32) {
        const int j = j0 + lane;
        const float z = j < S ? g[j] : 0.0f;
        const bool keep = full_height ? (z != 0.0f) : (z > 0.0f);
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (WRITE) {
            if (keep) {
                const int64_t o = base + __popc(m & ((1u << lane) - 1u));
                if (o < cap) {
                    pts_out[3 * o + 0] = __dmul_rn(rc, cosT[j]);   // :58  X = r * cos(theta)
                    pts_out[3 * o + 1] = __dmul_rn(rc, sinT[j]);   // :59
                    // :66 the offset is removed in the grid's dtype (float32 .npy grids)
                    pts_out[3 * o + 2] = full_height ? (double)z : (double)__fsub_rn(z, offset_f);
                    ids_out[o] = ring * S + j;                     // :74
                }
            }
            base += __popc(m);
        } else {
            count += __popc(m);
        }
    }
    if (!WRITE && lane == 0) row_counts[row] = count;
}

// ---- K3, single pass over HBM ---------------------------------------------------------------------
// Persistent CTAs take grids in ticket order.  Pass A reads the grid once (128-bit loads, every load of a ring row in
// flight at once), counts the kept cells of each row and COMPACTS them -- height and sector index -- into shared memory
// (a grid keeps 15-30 % of its 115 200 cells: ~30 k entries of 6 bytes).  A CTA-wide scan turns the counts into row
// offsets, the grid's base offset comes from a decoupled look-back over the totals of the grids before it (they are
// owned by CTAs that drew earlier tickets and are running, so the wait terminates), and pass C expands the compacted
// entries straight from shared memory into points, in row-major order, staged per warp so that they leave as
// contiguous stores.  HBM sees each grid once and each output byte once, and nothing is re-read from L2 (round 1
// re-read every row in pass C: 1.5x the algorithmic read traffic).  Rows that do not fit the compaction buffer
// (unusually dense grids) are compacted and emitted in further rounds from L2.
constexpr int K3F_THREADS = 1024;
constexpr int K3F_CTAS_PER_SM = 1;
constexpr int K3F_WARPS = K3F_THREADS / 32;
constexpr int K3F_MAXR = 2048;                                 // rows per grid the shared-memory scan can hold
constexpr int K3F_MAXROUNDS = 8;                               // 128 cells per warp round: S <= 1024 (sector fits 16 bits)
constexpr int K3F_STAGE_BYTES = 32 * 24;                       // per-warp output staging: 32 points x 3 doubles
constexpr unsigned long long K3F_AGG = 1ull << 62, K3F_PFX = 2ull << 62, K3F_VAL = (1ull << 62) - 1;

// One ring row per warp at a time, read as 128-bit vectors with every load of the row in flight at once
// (ROUNDS x 512 B per warp); lane l of round k holds cells 4 * (32 k + l) .. + 3.
__device__ __forceinline__ float4 k3_ld(const float4 *p, uint64_t pol)
{
    float4 v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}

// pass A for one row: count the kept cells, claim room in the compaction buffer, write (height, sector) in row-major
// order.  Returns the count; *pos_out = first entry, or -1 when the buffer is full (the row is done in a later round).
template <int ROUNDS>
__device__ __forceinline__ int k3_compact_row(const float4 *row, int rowvec, int full_height, uint64_t pol, int *s_alloc,
                                              int zcap, float *zbuf, unsigned short *jbuf, int *pos_out)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt = (1u << lane) - 1u;
    float4 v[ROUNDS];
#pragma unroll
    for (int k = 0; k < ROUNDS; ++k) {
        const int i = k * 32 + lane;
        v[k] = i < rowvec ? k3_ld(row + i, pol) : make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    unsigned bal[ROUNDS][4];
    int count = 0;
#pragma unroll
    for (int k = 0; k < ROUNDS; ++k) {
        const float zc[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            bal[k][c] = __ballot_sync(0xffffffffu, full_height ? (zc[c] != 0.0f) : (zc[c] > 0.0f));
            count += __popc(bal[k][c]);
        }
    }
    int pos = 0;
    if (lane == 0) {
        pos = count ? atomicAdd(s_alloc, count) : 0;
        if (pos + count > zcap) pos = -1;                  // (the claim stays: later rows of this round overflow too)
    }
    pos = __shfl_sync(0xffffffffu, pos, 0);
    *pos_out = pos;
    if (pos >= 0 && count) {
        int before = 0;                                    // kept cells of the earlier rounds
#pragma unroll
        for (int k = 0; k < ROUNDS; ++k) {
            const float zc[4] = {v[k].x, v[k].y, v[k].z, v[k].w};
            int rank = pos + before + __popc(bal[k][0] & lt) + __popc(bal[k][1] & lt) + __popc(bal[k][2] & lt) +
                       __popc(bal[k][3] & lt);             // kept cells of lower lanes come first, then lower components
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if ((bal[k][c] >> lane) & 1u) {
                    zbuf[rank] = zc[c];
                    jbuf[rank] = (unsigned short)(4 * (k * 32 + lane) + c);
                    ++rank;
                }
            }
            before += __popc(bal[k][0]) + __popc(bal[k][1]) + __popc(bal[k][2]) + __popc(bal[k][3]);
        }
    }
    return count;
}

template <int ROUNDS>
__global__ void __launch_bounds__(K3F_THREADS, K3F_CTAS_PER_SM)
k3_fused(const float *__restrict__ grids, int n_grids, int R, int S, int full_height, float offset_f,
         const double *__restrict__ rcs, const double *__restrict__ cosT, const double *__restrict__ sinT,
         unsigned int *__restrict__ ticket, unsigned long long *__restrict__ status,
         int64_t *__restrict__ item_offsets, double *__restrict__ pts_out, int32_t *__restrict__ ids_out, int64_t cap,
         const int zcap)
{
    __shared__ int s_g, s_alloc, s_left, s_rowA, s_rowC;
    __shared__ volatile int s_ready;             // the grid's base offset has been published (look-back done)
    __shared__ long long s_base;
    // dynamic: [per-warp staging][cos S][sin S][ring centres R][row offsets R+1][row buffer positions R][heights zcap][sectors zcap]
    extern __shared__ __align__(16) double k3_smem[];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    double *dbuf = k3_smem + (size_t)w * (K3F_STAGE_BYTES / 8);
    // the polar tables live in shared memory for the whole launch: pass C reads two of their entries per point, and from
    // global memory that was an L1/L2 round trip in front of every DMUL (5 % of the stall samples, profiles/r2j_k3_fused.*)
    double *s_cos = k3_smem + (size_t)K3F_WARPS * (K3F_STAGE_BYTES / 8);
    double *s_sin = s_cos + S;
    double *s_rc = s_sin + S;
    int *s_cnt = reinterpret_cast<int *>(s_rc + R);
    int *s_pos = s_cnt + (R + 1);
    float *zbuf = reinterpret_cast<float *>(s_pos + R + ((R + 1 + R) & 1));
    unsigned short *jbuf = reinterpret_cast<unsigned short *>(zbuf + zcap);
    const int rowvec = S >> 2;
    // pass A reads the grid once and lets go of it; rows that have to come back in a later round are parked in L2
    uint64_t pol_keep, pol_drop;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_drop));
    for (int i = tid; i < S; i += K3F_THREADS) {
        s_cos[i] = cosT[i];
        s_sin[i] = sinT[i];
    }
    for (int i = tid; i < R; i += K3F_THREADS) s_rc[i] = rcs[i];
    for (;;) {
        if (tid == 0) {
            s_g = (int)atomicAdd(ticket, 1u);
            s_alloc = 0;
            s_left = 0;
            s_rowA = 0;
            s_rowC = 0;
            s_ready = 0;
        }
        __syncthreads();
        const int g = s_g;
        if (g >= n_grids) break;
        const float4 *gp = reinterpret_cast<const float4 *>(grids + (size_t)g * R * S);
        // ---- pass A: count + compact every row; rows are handed out dynamically (dense rows take longer) ----
        for (;;) {
            int r = 0;
            if (lane == 0) r = atomicAdd(&s_rowA, 1);
            r = __shfl_sync(0xffffffffu, r, 0);
            if (r >= R) break;
            int pos;
            const int count = k3_compact_row<ROUNDS>(gp + (size_t)r * rowvec, rowvec, full_height, pol_keep, &s_alloc, zcap,
                                                     zbuf, jbuf, &pos);
            if (lane == 0) {
                s_cnt[r] = count;
                s_pos[r] = pos;
                if (pos < 0) atomicAdd(&s_left, 1);
            }
        }
        __syncthreads();
        // ---- exclusive scan of the row counts (warp 0) ----
        int run = 0;
        if (w == 0) {
            for (int r0 = 0; r0 < R; r0 += 32) {
                const int r = r0 + lane;
                const int c = r < R ? s_cnt[r] : 0;
                int inc = c;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) {
                    const int t = __shfl_up_sync(0xffffffffu, inc, d);
                    if (lane >= d) inc += t;
                }
                if (r < R) s_cnt[r] = run + inc - c;
                run += __shfl_sync(0xffffffffu, inc, 31);
            }
            if (lane == 0) s_cnt[R] = run;
        }
        __syncthreads();
        // ---- warp 0: decoupled look-back for the grid's base offset, 32 predecessors per step (lane l looks at grid
        //      q - l).  The other warps already expand rows (pass C) and only wait for the base before their first store.
        if (w == 0) {
            const unsigned long long total = (unsigned long long)run;
            volatile unsigned long long *stv = status;
            long long base = 0;
            if (g == 0) {
                if (lane == 0) stv[0] = K3F_PFX | total;
            } else {
                if (lane == 0) {
                    stv[g] = K3F_AGG | total;
                    __threadfence();
                }
                int q = g - 1;
                for (;;) {
                    const int idx = q - lane;
                    unsigned long long sv = 2ull << 62;                                   // before grid 0: prefix 0
                    if (idx >= 0) sv = stv[idx];
                    const unsigned not_ready = __ballot_sync(0xffffffffu, sv == 0ull);
                    const unsigned is_pfx = __ballot_sync(0xffffffffu, (sv & K3F_PFX) != 0ull);
                    const int first_pfx = is_pfx ? __ffs(is_pfx) - 1 : 32;
                    const int first_nr = not_ready ? __ffs(not_ready) - 1 : 32;
                    if (first_nr < first_pfx) {              // a predecessor on the way has not published yet
                        __nanosleep(40);
                        continue;
                    }
                    long long v = (lane <= first_pfx) ? (long long)(sv & K3F_VAL) : 0ll;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
                    base += v;
                    if (first_pfx < 32) break;
                    q -= 32;
                }
                if (lane == 0) stv[g] = K3F_PFX | ((unsigned long long)base + total);
            }
            if (lane == 0) {
                item_offsets[g] = base;
                if (g == n_grids - 1) item_offsets[n_grids] = base + (long long)total;
                s_base = base;
                __threadfence_block();
                s_ready = 1;
            }
        }
        bool have_base = false;
        long long gbase = 0;
        for (;;) {
            // ---- pass C: expand the compacted rows into points, row-major, 32 points per warp step ----
            for (;;) {
                int r = 0;
                if (lane == 0) r = atomicAdd(&s_rowC, 1);
                r = __shfl_sync(0xffffffffu, r, 0);
                if (r >= R) break;
                const int pos = s_pos[r];
                if (pos < 0) continue;                       // -1: waits for a later round, -2: already written
                const int rowoff = s_cnt[r];
                const int count = s_cnt[r + 1] - rowoff;
                if (count == 0) continue;
                const double rc = s_rc[r];
                const int idbase = r * S;
                if (!have_base) {                            // first store of this warp for this grid
                    while (!s_ready) __nanosleep(20);
                    __threadfence_block();
                    gbase = s_base;
                    have_base = true;
                }
                // 32 points per step: each lane expands one compacted entry into its staging slot, then the warp
                // copies the 768 contiguous bytes out with three fully coalesced 8-byte stores per lane
                const int64_t obase = gbase + rowoff;
                double *dst = pts_out + 3 * obase + lane;
                int32_t *idst = ids_out + obase + lane;
                const bool fits = obase + count <= cap;      // (capacity overflow is reported by the host wrapper)
                const float *zp = zbuf + pos + lane;
                const unsigned short *jp = jbuf + pos + lane;
                for (int e0 = 0; e0 < count; e0 += 32, dst += 96, idst += 32) {
                    const int m = count - e0;                // >= 32 except in the row's last step
                    int id = 0;
                    if (lane < m) {
                        const float z = zp[e0];
                        const int j = jp[e0];
                        dbuf[3 * lane + 0] = __dmul_rn(rc, s_cos[j]);  // :58  X = r * cos(theta)
                        dbuf[3 * lane + 1] = __dmul_rn(rc, s_sin[j]);  // :59
                        // :66 the offset is removed in the grid's dtype (float32 .npy grids)
                        dbuf[3 * lane + 2] = full_height ? (double)z : (double)__fsub_rn(z, offset_f);
                        id = idbase + j;                               // :74
                    }
                    __syncwarp();
                    if (m >= 32 && fits) {
                        __stcs(dst, dbuf[lane]);
                        __stcs(dst + 32, dbuf[lane + 32]);
                        __stcs(dst + 64, dbuf[lane + 64]);
                        __stcs(idst, id);
                    } else {
                        for (int t = lane; t < 3 * m && t < 96; t += 32)
                            if (3 * (obase + e0) + t < 3 * cap) __stcs(dst + (t - lane), dbuf[t]);
                        if (lane < m && obase + e0 + lane < cap) __stcs(idst, id);
                    }
                    __syncwarp();
                }
                if (lane == 0) s_pos[r] = -2;
            }
            __syncthreads();
            if (s_left == 0) break;                          // CTA-uniform
            // ---- a further round for the rows that did not fit: compact them (from L2) into the freed buffer ----
            __syncthreads();
            if (tid == 0) {
                s_alloc = 0;
                s_left = 0;
                s_rowC = 0;
            }
            __syncthreads();
            for (int r = w; r < R; r += K3F_WARPS) {
                if (s_pos[r] != -1) continue;
                int pos;
                k3_compact_row<ROUNDS>(gp + (size_t)r * rowvec, rowvec, full_height, pol_drop, &s_alloc, zcap, zbuf, jbuf, &pos);
                if (lane == 0) {
                    s_pos[r] = pos;
                    if (pos < 0) atomicAdd(&s_left, 1);
                }
            }
            __syncthreads();
        }
        __syncthreads();                         // the row tables and the buffer are rewritten by the next grid
    }
}

// ---- two-level exclusive scan: rows within an item, then items ---------------------------------
// level 1: one warp per item: row_offsets[item][r] = sum_{q<r} counts, item_totals[item] = sum
__global__ void __launch_bounds__(ROW_WARPS * 32)
scan_rows_in_item(const int32_t *__restrict__ row_counts, int n_items, int R, int64_t *__restrict__ row_offsets,
                  int64_t *__restrict__ item_totals)
{
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
    if (item >= n_items) return;
    int64_t run = 0;
    for (int r0 = 0; r0 < R; r0 += 32) {
        const int r = r0 + lane;
        int v = r < R ? row_counts[(int64_t)item * R + r] : 0;
        int inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (r < R) row_offsets[(int64_t)item * R + r] = run + (inc - v);
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) item_totals[item] = run;
}
// level 2: single CTA: item_offsets[i] = exclusive scan of item_totals, item_offsets[n] = total
__global__ void __launch_bounds__(1024)
scan_items(const int64_t *__restrict__ item_totals, int n_items, int64_t *__restrict__ item_offsets)
{
    __shared__ int64_t warp_sum[32];
    __shared__ int64_t carry;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int i0 = 0; i0 < n_items; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        const int64_t v = i < n_items ? item_totals[i] : 0;
        int64_t inc = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += t;
        }
        if (lane == 31) warp_sum[w] = inc;
        __syncthreads();
        if (w == 0) {
            int64_t s = warp_sum[lane];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t t = __shfl_up_sync(0xffffffffu, s, d);
                if (lane >= d) s += t;
            }
            warp_sum[lane] = s;  // inclusive
        }
        __syncthreads();
        const int64_t before = carry + (w ? warp_sum[w - 1] : 0);
        if (i < n_items) item_offsets[i] = before + inc - v;
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) item_offsets[n_items] = carry;
}
// level 3: make the row offsets absolute
__global__ void add_item_base(int64_t *__restrict__ row_offsets, const int64_t *__restrict__ item_offsets,
                              int64_t n_rows, int R)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_rows) row_offsets[i] += item_offsets[i / R];
}

// ---- K5 ---------------------------------------------------------------------------------------
__device__ __forceinline__ int clipi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

__global__ void k5_set_bits(const int32_t *__restrict__ cells, int64_t n_cells, int R, int S, uint32_t *__restrict__ bits)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const int r = clipi(cells[2 * i], 0, R - 1), s = clipi(cells[2 * i + 1], 0, S - 1);   // :154-155
    const int id = r * S + s;                                                               // :158
    atomicOr(bits + (id >> 5), 1u << (id & 31));
}
__global__ void k5_lookup(const int32_t *__restrict__ ids, int64_t n, int64_t n_cells_total,
                          const uint32_t *__restrict__ bits, uint8_t *__restrict__ mask)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids[i];
    mask[i] = (id >= 0 && id < n_cells_total) ? (uint8_t)((bits[id >> 5] >> (id & 31)) & 1u) : 0;   // :161
}
__global__ void k5_gather(const float *__restrict__ grid, const int32_t *__restrict__ cells, int64_t n_cells,
                          int R, int S, const double *__restrict__ rcs, const double *__restrict__ cosT,
                          const double *__restrict__ sinT, double *__restrict__ pts)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_cells) return;
    const int r = clipi(cells[2 * i], 0, R - 1), s = clipi(cells[2 * i + 1], 0, S - 1);   // :109-110
    pts[3 * i + 0] = __dmul_rn(rcs[r], cosT[s]);                                           // :118
    pts[3 * i + 1] = __dmul_rn(rcs[r], sinT[s]);                                           // :119
    pts[3 * i + 2] = (double)grid[(int64_t)r * S + s];                                     // :120
}

}  // namespace

// ---- host side ----------------------------------------------------------------------------------

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int run_scan(dsc_ctx *ctx, int32_t *row_counts, int n_items, int R, int64_t *row_offsets,
                    int64_t *item_totals, int64_t *item_offsets, cudaStream_t st)
{
    int rc;
    scan_rows_in_item<<<(n_items + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, st>>>(row_counts, n_items, R,
                                                                                        row_offsets, item_totals);
    if ((rc = dsc_check_launch(ctx, "scan_rows_in_item"))) return rc;
    scan_items<<<1, 1024, 0, st>>>(item_totals, n_items, item_offsets);
    if ((rc = dsc_check_launch(ctx, "scan_items"))) return rc;
    const int64_t n_rows = (int64_t)n_items * R;
    add_item_base<<<(unsigned)((n_rows + 255) / 256), 256, 0, st>>>(row_offsets, item_offsets, n_rows, R);
    return dsc_check_launch(ctx, "add_item_base");
}

static int need_tables(dsc_ctx *ctx)
{
    if (!ctx->d_rc) DSC_FAIL(ctx, DSC_ERR_INVALID, "dsc_set_polar_tables has not been called");
    return DSC_OK;
}

extern "C" int dsc_raster_boxes(dsc_ctx *ctx, const double *boxes, int n_boxes, int mem, int64_t *cell_offsets,
                                int32_t *cells_out, int64_t cap, uint8_t *dense_mask, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    int rc;
    if ((rc = need_tables(ctx))) return rc;
    if (n_boxes < 0 || cap < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad sizes");
    if (!cell_offsets) DSC_FAIL(ctx, DSC_ERR_INVALID, "cell_offsets is NULL");
    if (n_boxes > 0 && !boxes) DSC_FAIL(ctx, DSC_ERR_INVALID, "boxes is NULL");
    if (cap > 0 && !cells_out) DSC_FAIL(ctx, DSC_ERR_INVALID, "cells_out is NULL");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    const int R = ctx->tab_R, S = ctx->tab_S;
    const bool host = mem == DSC_MEM_HOST;
    cudaStream_t st = host ? ctx->s_k : (cudaStream_t)stream;
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    const int64_t n_rows = (int64_t)n_boxes * R;
    // scratch layout
    size_t o = 0;
    const size_t o_counts = o;   o = align_up(o + (size_t)(n_rows + 1) * 4, 256);
    const size_t o_rowoff = o;   o = align_up(o + (size_t)(n_rows + 1) * 8, 256);
    const size_t o_tot = o;      o = align_up(o + (size_t)(n_boxes + 1) * 8, 256);
    const size_t o_itemoff = o;  o = align_up(o + (size_t)(n_boxes + 1) * 8, 256);
    const size_t o_boxes = o;    o = align_up(o + (host ? (size_t)n_boxes * 48 : 0), 256);
    const size_t o_cells = o;    o = align_up(o + (host ? (size_t)cap * 8 : 0), 256);
    const size_t o_dense = o;    o = align_up(o + (host && dense_mask ? (size_t)R * S : 0), 256);
    if ((rc = dsc_reserve_scratch(ctx, o))) return rc;
    uint8_t *sc = (uint8_t *)ctx->scratch;
    int32_t *row_counts = (int32_t *)(sc + o_counts);
    int64_t *row_offsets = (int64_t *)(sc + o_rowoff);
    int64_t *item_totals = (int64_t *)(sc + o_tot);
    int64_t *item_offsets = host ? (int64_t *)(sc + o_itemoff) : cell_offsets;
    const double *d_boxes = host ? (const double *)(sc + o_boxes) : boxes;
    int32_t *d_cells = host ? (int32_t *)(sc + o_cells) : cells_out;
    uint8_t *d_dense = dense_mask ? (host ? sc + o_dense : dense_mask) : nullptr;
    if (host && n_boxes)
        DSC_CUDA(ctx, cudaMemcpyAsync((void *)d_boxes, boxes, (size_t)n_boxes * 48, cudaMemcpyHostToDevice, st));
    if (d_dense) DSC_CUDA(ctx, cudaMemsetAsync(d_dense, 0, (size_t)R * S, st));
    if (n_boxes == 0) {
        DSC_CUDA(ctx, cudaMemsetAsync(item_offsets, 0, 8, st));
    } else {
        const unsigned blocks = (unsigned)((n_rows + ROW_WARPS - 1) / ROW_WARPS);
        k2_rows<false><<<blocks, ROW_WARPS * 32, 0, st>>>(d_boxes, n_boxes, R, S, ctx->d_rc, ctx->d_cos, ctx->d_sin,
                                                          row_counts, nullptr, nullptr, 0, nullptr);
        if ((rc = dsc_check_launch(ctx, "k2_rows<count>"))) return rc;
        if ((rc = run_scan(ctx, row_counts, n_boxes, R, row_offsets, item_totals, item_offsets, st))) return rc;
        k2_rows<true><<<blocks, ROW_WARPS * 32, 0, st>>>(d_boxes, n_boxes, R, S, ctx->d_rc, ctx->d_cos, ctx->d_sin,
                                                         row_counts, row_offsets, d_cells, cap, d_dense);
        if ((rc = dsc_check_launch(ctx, "k2_rows<write>"))) return rc;
    }
    if (!host) return dsc_leave_stream(ctx, st);
    DSC_CUDA(ctx, cudaMemcpyAsync(cell_offsets, item_offsets, (size_t)(n_boxes + 1) * 8, cudaMemcpyDeviceToHost, st));
    DSC_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t total = cell_offsets[n_boxes];
    const int64_t ncopy = total < cap ? total : cap;
    if (ncopy > 0) DSC_CUDA(ctx, cudaMemcpy(cells_out, d_cells, (size_t)ncopy * 8, cudaMemcpyDeviceToHost));
    if (dense_mask) DSC_CUDA(ctx, cudaMemcpy(dense_mask, d_dense, (size_t)R * S, cudaMemcpyDeviceToHost));
    if (total > cap) {
        char buf[128];
        snprintf(buf, sizeof buf, "cells_out capacity %lld < required %lld", (long long)cap, (long long)total);
        DSC_FAIL(ctx, DSC_ERR_CAPACITY, buf);
    }
    return DSC_OK;
}

extern "C" int dsc_reconstruct_batch(dsc_ctx *ctx, const float *grids, int n_grids, int full_height,
                                     double height_offset, int mem, int64_t *pt_offsets, double *pts_out,
                                     int32_t *bin_ids_out, int64_t cap, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    int rc;
    if ((rc = need_tables(ctx))) return rc;
    if (n_grids < 0 || cap < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad sizes");
    if (!pt_offsets) DSC_FAIL(ctx, DSC_ERR_INVALID, "pt_offsets is NULL");
    if (n_grids > 0 && !grids) DSC_FAIL(ctx, DSC_ERR_INVALID, "grids is NULL");
    if (cap > 0 && (!pts_out || !bin_ids_out)) DSC_FAIL(ctx, DSC_ERR_INVALID, "output buffer is NULL");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    const int R = ctx->tab_R, S = ctx->tab_S;
    const int64_t cells = (int64_t)R * S;
    const bool host = mem == DSC_MEM_HOST;
    cudaStream_t st = host ? ctx->s_k : (cudaStream_t)stream;
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    const int64_t n_rows = (int64_t)n_grids * R;
    size_t o = 0;
    const size_t o_counts = o;   o = align_up(o + (size_t)(n_rows + 1) * 4, 256);
    const size_t o_rowoff = o;   o = align_up(o + (size_t)(n_rows + 1) * 8, 256);
    const size_t o_tot = o;      o = align_up(o + (size_t)(n_grids + 1) * 8, 256);
    const size_t o_itemoff = o;  o = align_up(o + (size_t)(n_grids + 1) * 8, 256);
    const size_t o_grids = o;    o = align_up(o + (host ? (size_t)n_grids * cells * 4 : 0), 256);
    const size_t o_pts = o;      o = align_up(o + (host ? (size_t)cap * 24 : 0), 256);
    const size_t o_ids = o;      o = align_up(o + (host ? (size_t)cap * 4 : 0), 256);
    if ((rc = dsc_reserve_scratch(ctx, o))) return rc;
    uint8_t *sc = (uint8_t *)ctx->scratch;
    int32_t *row_counts = (int32_t *)(sc + o_counts);
    int64_t *row_offsets = (int64_t *)(sc + o_rowoff);
    int64_t *item_totals = (int64_t *)(sc + o_tot);
    int64_t *item_offsets = host ? (int64_t *)(sc + o_itemoff) : pt_offsets;
    const float *d_grids = host ? (const float *)(sc + o_grids) : grids;
    double *d_pts = host ? (double *)(sc + o_pts) : pts_out;
    int32_t *d_ids = host ? (int32_t *)(sc + o_ids) : bin_ids_out;
    if (host && n_grids)
        DSC_CUDA(ctx, cudaMemcpyAsync((void *)d_grids, grids, (size_t)n_grids * cells * 4, cudaMemcpyHostToDevice, st));
    if (n_grids == 0) {
        DSC_CUDA(ctx, cudaMemsetAsync(item_offsets, 0, 8, st));
    } else {
        const int rounds = (S / 4 + 31) / 32;
        if (R <= K3F_MAXR && (S & 3) == 0 && rounds <= K3F_MAXROUNDS && (((uintptr_t)d_grids) & 15) == 0) {
            // single pass: ticket counter + look-back status words live in the (otherwise unused) counts area
            unsigned int *ticket = (unsigned int *)row_counts;
            unsigned long long *status = (unsigned long long *)row_offsets;
            DSC_CUDA(ctx, cudaMemsetAsync(ticket, 0, 16, st));
            DSC_CUDA(ctx, cudaMemsetAsync(status, 0, (size_t)n_grids * 8, st));
            const int max_ctas = ctx->sm_count * K3F_CTAS_PER_SM;
            const int ctas = n_grids < max_ctas ? n_grids : max_ctas;   // tickets are drawn by running CTAs only: look-back cannot stall
            // shared memory: per-warp staging, the two row tables, and as many compaction entries (6 B) as fit
            const size_t k3_fixed = (size_t)K3F_WARPS * K3F_STAGE_BYTES + (size_t)(2 * S + R) * 8 + (size_t)(2 * R + 2) * 4 + 16;
            int zcap = (int)((ctx->smem_optin - 1024 - k3_fixed) / 6) & ~7;
            if (zcap > R * S) zcap = (R * S + 7) & ~7;
            const int k3_smem = (int)(k3_fixed + (size_t)zcap * 6);
#define K3F_LAUNCH(RN)                                                                                          \
    DSC_CUDA(ctx, cudaFuncSetAttribute(k3_fused<RN>, cudaFuncAttributeMaxDynamicSharedMemorySize, k3_smem));    \
    k3_fused<RN><<<ctas, K3F_THREADS, k3_smem, st>>>(d_grids, n_grids, R, S, full_height, (float)height_offset, ctx->d_rc, \
                                               ctx->d_cos, ctx->d_sin, ticket, status, item_offsets, d_pts, d_ids, cap, zcap)
            if (rounds <= 2) { K3F_LAUNCH(2); }
            else if (rounds <= 4) { K3F_LAUNCH(4); }
            else if (rounds <= 6) { K3F_LAUNCH(6); }
            else { K3F_LAUNCH(8); }
#undef K3F_LAUNCH
            if ((rc = dsc_check_launch(ctx, "k3_fused"))) return rc;
        } else {
        const unsigned blocks = (unsigned)((n_rows + ROW_WARPS - 1) / ROW_WARPS);
        k3_rows<false><<<blocks, ROW_WARPS * 32, 0, st>>>(d_grids, n_grids, R, S, full_height, (float)height_offset,
                                                          ctx->d_rc, ctx->d_cos, ctx->d_sin, row_counts, nullptr,
                                                          nullptr, nullptr, 0);
        if ((rc = dsc_check_launch(ctx, "k3_rows<count>"))) return rc;
        if ((rc = run_scan(ctx, row_counts, n_grids, R, row_offsets, item_totals, item_offsets, st))) return rc;
        k3_rows<true><<<blocks, ROW_WARPS * 32, 0, st>>>(d_grids, n_grids, R, S, full_height, (float)height_offset,
                                                         ctx->d_rc, ctx->d_cos, ctx->d_sin, row_counts, row_offsets,
                                                         d_pts, d_ids, cap);
        if ((rc = dsc_check_launch(ctx, "k3_rows<write>"))) return rc;
        }
    }
    if (!host) return dsc_leave_stream(ctx, st);
    DSC_CUDA(ctx, cudaMemcpyAsync(pt_offsets, item_offsets, (size_t)(n_grids + 1) * 8, cudaMemcpyDeviceToHost, st));
    DSC_CUDA(ctx, cudaStreamSynchronize(st));
    const int64_t total = pt_offsets[n_grids];
    const int64_t ncopy = total < cap ? total : cap;
    if (ncopy > 0) {
        DSC_CUDA(ctx, cudaMemcpy(pts_out, d_pts, (size_t)ncopy * 24, cudaMemcpyDeviceToHost));
        DSC_CUDA(ctx, cudaMemcpy(bin_ids_out, d_ids, (size_t)ncopy * 4, cudaMemcpyDeviceToHost));
    }
    if (total > cap) {
        char buf[128];
        snprintf(buf, sizeof buf, "pts_out capacity %lld < required %lld", (long long)cap, (long long)total);
        DSC_FAIL(ctx, DSC_ERR_CAPACITY, buf);
    }
    return DSC_OK;
}

extern "C" int dsc_covered_mask(dsc_ctx *ctx, const int32_t *bin_ids, int64_t n_ids, const int32_t *cells,
                                int64_t n_cells, int mem, uint8_t *mask_out, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    int rc;
    if ((rc = need_tables(ctx))) return rc;
    if (n_ids < 0 || n_cells < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad sizes");
    if (n_ids == 0) return DSC_OK;
    if (!bin_ids || !mask_out || (n_cells && !cells)) DSC_FAIL(ctx, DSC_ERR_INVALID, "NULL buffer");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    const int R = ctx->tab_R, S = ctx->tab_S;
    const int64_t total_cells = (int64_t)R * S;
    const size_t bits_bytes = align_up((size_t)(total_cells + 31) / 32 * 4, 256);
    const bool host = mem == DSC_MEM_HOST;
    cudaStream_t st = host ? ctx->s_k : (cudaStream_t)stream;
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    size_t o = bits_bytes;
    const size_t o_ids = o;    o = align_up(o + (host ? (size_t)n_ids * 4 : 0), 256);
    const size_t o_cells = o;  o = align_up(o + (host ? (size_t)n_cells * 8 : 0), 256);
    const size_t o_mask = o;   o = align_up(o + (host ? (size_t)n_ids : 0), 256);
    if ((rc = dsc_reserve_scratch(ctx, o))) return rc;
    uint8_t *sc = (uint8_t *)ctx->scratch;
    uint32_t *bits = (uint32_t *)sc;
    const int32_t *d_ids = host ? (const int32_t *)(sc + o_ids) : bin_ids;
    const int32_t *d_cells = host ? (const int32_t *)(sc + o_cells) : cells;
    uint8_t *d_mask = host ? sc + o_mask : mask_out;
    DSC_CUDA(ctx, cudaMemsetAsync(bits, 0, bits_bytes, st));
    if (host) {
        DSC_CUDA(ctx, cudaMemcpyAsync((void *)d_ids, bin_ids, (size_t)n_ids * 4, cudaMemcpyHostToDevice, st));
        if (n_cells) DSC_CUDA(ctx, cudaMemcpyAsync((void *)d_cells, cells, (size_t)n_cells * 8, cudaMemcpyHostToDevice, st));
    }
    if (n_cells) {
        k5_set_bits<<<(unsigned)((n_cells + 255) / 256), 256, 0, st>>>(d_cells, n_cells, R, S, bits);
        if ((rc = dsc_check_launch(ctx, "k5_set_bits"))) return rc;
    }
    k5_lookup<<<(unsigned)((n_ids + 255) / 256), 256, 0, st>>>(d_ids, n_ids, total_cells, bits, d_mask);
    if ((rc = dsc_check_launch(ctx, "k5_lookup"))) return rc;
    if (host) {
        DSC_CUDA(ctx, cudaMemcpyAsync(mask_out, d_mask, (size_t)n_ids, cudaMemcpyDeviceToHost, st));
        DSC_CUDA(ctx, cudaStreamSynchronize(st));
        return DSC_OK;
    }
    return dsc_leave_stream(ctx, st);
}

extern "C" int dsc_gather_cells(dsc_ctx *ctx, const float *grid, const int32_t *cells, int64_t n_cells, int mem,
                                double *pts_out, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    int rc;
    if ((rc = need_tables(ctx))) return rc;
    if (n_cells < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad sizes");
    if (n_cells == 0) return DSC_OK;
    if (!grid || !cells || !pts_out) DSC_FAIL(ctx, DSC_ERR_INVALID, "NULL buffer");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    const int R = ctx->tab_R, S = ctx->tab_S;
    const bool host = mem == DSC_MEM_HOST;
    cudaStream_t st = host ? ctx->s_k : (cudaStream_t)stream;
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    size_t o = 0;
    const size_t o_grid = o;   o = align_up(o + (host ? (size_t)R * S * 4 : 0), 256);
    const size_t o_cells = o;  o = align_up(o + (host ? (size_t)n_cells * 8 : 0), 256);
    const size_t o_pts = o;    o = align_up(o + (host ? (size_t)n_cells * 24 : 0), 256);
    if ((rc = dsc_reserve_scratch(ctx, o))) return rc;
    uint8_t *sc = (uint8_t *)ctx->scratch;
    const float *d_grid = host ? (const float *)(sc + o_grid) : grid;
    const int32_t *d_cells = host ? (const int32_t *)(sc + o_cells) : cells;
    double *d_pts = host ? (double *)(sc + o_pts) : pts_out;
    if (host) {
        DSC_CUDA(ctx, cudaMemcpyAsync((void *)d_grid, grid, (size_t)R * S * 4, cudaMemcpyHostToDevice, st));
        DSC_CUDA(ctx, cudaMemcpyAsync((void *)d_cells, cells, (size_t)n_cells * 8, cudaMemcpyHostToDevice, st));
    }
    k5_gather<<<(unsigned)((n_cells + 255) / 256), 256, 0, st>>>(d_grid, d_cells, n_cells, R, S, ctx->d_rc, ctx->d_cos,
                                                                 ctx->d_sin, d_pts);
    if ((rc = dsc_check_launch(ctx, "k5_gather"))) return rc;
    if (host) {
        DSC_CUDA(ctx, cudaMemcpyAsync(pts_out, d_pts, (size_t)n_cells * 24, cudaMemcpyDeviceToHost, st));
        DSC_CUDA(ctx, cudaStreamSynchronize(st));
        return DSC_OK;
    }
    return dsc_leave_stream(ctx, st);
}
