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

#include "dsc_k1_common.cuh"
#include "dsc_k1_scan.cuh"
#include "dsc_k1_fallback.cuh"
#include "dsc_k4.cuh"

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
        // the head of each grid stays in shared memory between the two passes (everything the SM offers)
        const size_t k4_smem = ((ctx->smem_optin - 1024) / 16) * 16;
        int keep_vec = (int)(k4_smem / 16);
        if (keep_vec > (int)(cells / 4)) keep_vec = (int)(cells / 4);
        DSC_CUDA(ctx, cudaFuncSetAttribute(k4_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4_smem));
        k4_fused<<<ctas, K4_THREADS, k4_smem, st>>>(grids, n, (int)(cells / 4), out, keep_vec);
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
                                         ((p->flags & DSC_FLAG_GAP_F32) ? 3 : 0) | ((p->flags & DSC_FLAG_GAP_RING_F32) ? 1 : 0) |
                                             ((p->flags & DSC_FLAG_GAP_SECTOR_F32) ? 2 : 0),
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
    // (k1_scan addresses the stream with 32-bit row indices)
    const bool sc_shape_ok = vec16 && (p->num_sector % 4) == 0 && (((uintptr_t)out) & 15) == 0 && ctx->k1_impl != 1 &&
                             h_offsets[n_scans] < 0xffff0000LL;
    const bool want_split = ctx->sc_split == 1 ||
                            (ctx->sc_split < 0 && (n_scans < 2 * ctx->sm_count || max_pts > 3 * ctx->sc_item_points) &&
                             split_items > n_scans);
    const bool split = sc_shape_ok && want_split && split_items < 0x7fff0000LL;
    const bool defgrid = dsc::is_default_grid(cfg) && ctx->no_defgrid == 0;
    // launch shape (threads x 128-bit loads per thread per register set); shapes other than 2 exist for the screened
    // default-grid kernels only.  ptxas keeps the loads of both register sets on one scoreboard, so one set is in flight
    // while the other is processed: the bytes in flight are the size of a set, and the biggest set that still leaves
    // 24 warps and spills nothing wins -- 768 x 6 (80 registers, 72 KB per set) 1.770 ms, 768 x 5 1.778, 896 x 4 1.789,
    // 768 x 4 1.794, 896 x 3 1.820, 768 x 3 1.838, 640 x 6 1.808, 896 x 5 (spills) 1.83 (profiles/r2_notes.md, r2q-r2s).
    // Split items reduce with RED.MAX on a grid shared by several CTAs and prefer 768 x 3 (dense2m: 0.575 ms vs 0.627 ms
    // for 768 x 4 and 0.71 ms for 896 x 3).
    int shape = (cfg.fast_ok && defgrid) ? (ctx->sc_cfg >= 0 ? ctx->sc_cfg : (split ? 1 : 0)) : 2;   // the generic kernels are 1024 x 2
#ifdef SC_EXTRA_SHAPES                            // A/B builds: 896 x 3, 768 x 5, 896 x 4, 1024 x 3, 640 x 6
    if (shape < 0 || shape > 7) shape = 0;
    const int nthreads = (shape == 3 || shape == 5) ? 896 : shape == 7 ? 640 : (shape == 2 || shape == 6) ? 1024 : 768;
#else
    if (shape != 1 && shape != 2) shape = 0;
    const int nthreads = shape == 2 ? 1024 : 768;
#endif
    const size_t row_bytes = (size_t)p->num_sector * 4;
    const size_t q_bytes = cfg.fast_ok ? (size_t)SCQ_WARP_BYTES * (nthreads / 32) : 0;       // per-warp exact-path queues
    // Resident rows.  Shared memory and L1 share 256 KB per SM; beyond the 196 KiB carve-out step the streaming loads
    // starve (profiles/r1_notes.md: 60 rows 1.99 ms, 64 rows 2.14 ms; the 1024-thread shapes too: 49 rows 1.90 ms,
    // 58 rows 1.86 ms, r2p).
    // Split items have no second pass -- every non-resident point is a RED.MAX on a grid that 16 CTAs share --
    // and want every row they can get (dense2m: 62 rows 0.62 ms, 72 rows 0.53 ms).
    const size_t sc_budget = split ? ctx->smem_optin : (size_t)196 * 1024 - 1024 - 256 - q_bytes;
    const size_t sc_max = ctx->smem_optin - 1024 - q_bytes;
    int ns = (int)((sc_budget < sc_max ? sc_budget : sc_max) / row_bytes);
    if (ctx->sc_smem_rows >= 0) ns = ctx->sc_smem_rows;                          // A/B override
    if ((size_t)ns * row_bytes > sc_max) ns = (int)(sc_max / row_bytes);
    if (ns > p->num_ring) ns = p->num_ring;
    // whole scans reduce the rings beyond the resident ones in bands of ns rows, re-reading the spill list once
    // per band: grids with very many rings for their row size go to the generic path instead
    const int bands = ns >= 2 ? (p->num_ring - ns + ns - 1) / ns : 0x7fffffff;
    // (whole-scan items address their spill lists with 32-bit entry indices: sm_count regions of max_pts entries)
    const bool spill_ok = (int64_t)ctx->sm_count * (max_pts + 2 * nthreads) < 0xffff0000LL;
    const bool sc_ok = sc_shape_ok && ns >= 2 && (split || (bands <= 8 && spill_ok)) &&
                       (split || ctx->k1_impl == 3 || n_scans >= ctx->sm_count / 2);
    if (sc_ok) {
        int ctas = ctx->sm_count;
        const int64_t n_items = split ? split_items : (int64_t)n_scans;
        if (ctas > n_items) ctas = (int)n_items;
        // sync block: [cursor (16 B)] [done[n_scans]] [item_base[n_scans+1] (ragged batches)]
        SplitArgs spa = {nullptr, nullptr, 0, 0};
        size_t sync_bytes = 16;
        size_t spill_cap = 0;                    // entries per CTA
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
            // spill lists: one region per CTA, big enough for every point of the largest scan (only the part that
            // is used -- the points outside the resident rings, ~15 % on KITTI-like scans -- is ever touched)
            spill_cap = (size_t)((max_pts + nthreads - 1) / nthreads + 1) * (size_t)nthreads;     // (+1: see sc_drain)
            if (spill_cap < (size_t)nthreads) spill_cap = (size_t)nthreads;
            const size_t need = (size_t)ctas * spill_cap * sizeof(uint2);
            if (need > ctx->spill_bytes) {
                if (ctx->spill) { DSC_CUDA(ctx, cudaDeviceSynchronize()); cudaFree(ctx->spill); }
                ctx->spill = nullptr;
                ctx->spill_bytes = 0;
                DSC_CUDA(ctx, cudaMalloc(&ctx->spill, need));
                ctx->spill_bytes = need;
            }
        }
        unsigned int *cursor = (unsigned int *)ctx->sync_buf;
        // resident window: ring histogram of sampled points -> first ring r0 (device-side; no host round trip).  The
        // probe leaves its histogram zeroed for the next call and resets the work cursor of the launch that follows,
        // so the common whole-scan call is two launches and no memset.
        if (!ctx->probe) {
            DSC_CUDA(ctx, cudaMalloc(&ctx->probe, (PROBE_MAXR + 4) * sizeof(unsigned int)));
            DSC_CUDA(ctx, cudaMemsetAsync(ctx->probe, 0, (PROBE_MAXR + 4) * sizeof(unsigned int), st));
        }
        unsigned int *probe = (unsigned int *)ctx->probe;
        int *window = (int *)(probe + PROBE_MAXR + 2);
        const bool run_probe = total_pts > 0 && p->num_ring <= PROBE_MAXR && ns < p->num_ring && ctx->sc_window < 0;
        if (!run_probe || sync_bytes > 16) DSC_CUDA(ctx, cudaMemsetAsync(cursor, 0, sync_bytes, st));
        if (run_probe) {
            k1_probe<<<PROBE_BLOCKS, PROBE_THREADS, 0, st>>>(pts, h_offsets[0], total_pts, (float)(p->num_ring / p->max_length),
                                                             p->num_ring, ns - 1, probe, window, cursor);
            if ((rc = dsc_check_launch(ctx, "k1_probe"))) return rc;
        } else {                                         // no probe: window at ring 0, or the A/B override
            const int w = ctx->sc_window <= 0 ? 0
                        : ctx->sc_window <= p->num_ring - ns ? ctx->sc_window : (p->num_ring - ns > 0 ? p->num_ring - ns : 0);
            DSC_CUDA(ctx, cudaMemcpyAsync(window, &w, sizeof(int), cudaMemcpyHostToDevice, st));
        }
        unsigned long long *stp = (unsigned long long *)stats;
        const bool full = p->full_height != 0;
        const size_t sc_smem = (size_t)ns * row_bytes + q_bytes;
#define SC_LAUNCH5(F, V, H, U, NT, D, SP, ST)                                                                   \
    do {                                                                                                        \
        DSC_CUDA(ctx, cudaFuncSetAttribute(k1_scan<F, V, H, U, NT, D, SP, ST>,                                  \
                                           cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sc_smem));         \
        k1_scan<F, V, H, U, NT, D, SP, ST><<<ctas, NT, sc_smem, st>>>(pts, d_offsets, n_scans,                  \
                                                                      (uint2 *)ctx->spill, spill_cap, out, stp, \
                                                                      cursor, ns, window, cfg, spa);            \
    } while (0)
#define SC_LAUNCH4(F, V, H, U, NT, D, SP)                                                                       \
    do {                                                                                                        \
        if (stp) SC_LAUNCH5(F, V, H, U, NT, D, SP, true); else SC_LAUNCH5(F, V, H, U, NT, D, SP, false);        \
    } while (0)
#define SC_LAUNCH3(F, V, H, U, NT, D)                                                                           \
    do {                                                                                                        \
        if (split) SC_LAUNCH4(F, V, H, U, NT, D, true); else SC_LAUNCH4(F, V, H, U, NT, D, false);              \
    } while (0)
#ifdef SC_EXTRA_SHAPES
#define SC_EXTRA_LAUNCH(V, H)                                                                                   \
        else if (shape == 3) SC_LAUNCH3(true, V, H, 3, 896, true);                                              \
        else if (shape == 4) SC_LAUNCH3(true, V, H, 5, 768, true);                                              \
        else if (shape == 5) SC_LAUNCH3(true, V, H, 4, 896, true);                                              \
        else if (shape == 6) SC_LAUNCH3(true, V, H, 3, 1024, true);                                             \
        else if (shape == 7) SC_LAUNCH3(true, V, H, 6, 640, true);
#else
#define SC_EXTRA_LAUNCH(V, H)
#endif
#define SC_LAUNCH_DEF(V, H)                                                                                     \
    do {                                                                                                        \
        if (shape == 1) SC_LAUNCH3(true, V, H, 3, 768, true);                                                   \
        else if (shape == 2) SC_LAUNCH3(true, V, H, 2, 1024, true);                                             \
        SC_EXTRA_LAUNCH(V, H)                                                                                   \
        else SC_LAUNCH3(true, V, H, 6, 768, true);                                                              \
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
#undef SC_LAUNCH5
        if ((rc = dsc_check_launch(ctx, "k1_scan"))) return rc;
        if (out_norm) return normalize_device(ctx, out, n_scans, cells, out_norm, st);
        return DSC_OK;
    }
    // ---- generic path (12-byte stride, odd grids, unaligned buffers): key grid in L2 + RED.MAX, then a decode pass ----
    // key scratch: sized to stay L2-resident between build and decode
    int slots = (int)((size_t)48 * 1024 * 1024 / (size_t)(cells * 4));
    if (slots < 1) slots = 1;
    if (slots > 4096) slots = 4096;
    if (slots > n_scans) slots = n_scans;
    if ((rc = ensure_keys(ctx, (size_t)slots * cells * 4, st))) return rc;
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

// Offsets of a device-mode call: a small ring of pinned + device buffer pairs, each guarded by the event of the call
// that used it last, so that back-to-back asynchronous calls never wait for one another on the host.
int dsc_upload_offsets(dsc_ctx *ctx, const int64_t *offsets, int n_scans, cudaStream_t st, int *slot_out)
{
    const size_t cnt = (size_t)n_scans + 1;
    const int slot = ctx->off_next;
    ctx->off_next = (slot + 1) % DSC_OFF_RING;
    dsc_ctx::OffSlot &o = ctx->off[slot];
    if (o.used) DSC_CUDA(ctx, cudaEventSynchronize(o.ev));       // DSC_OFF_RING calls ago: long since complete
    if (cnt > o.cap) {
        if (o.d) cudaFree(o.d);
        if (o.h) cudaFreeHost(o.h);
        o.d = o.h = nullptr;
        o.cap = 0;
        const size_t cap = cnt < 4096 ? 4096 : cnt * 2;
        DSC_CUDA(ctx, cudaMalloc(&o.d, cap * sizeof(int64_t)));
        DSC_CUDA(ctx, cudaMallocHost(&o.h, cap * sizeof(int64_t)));
        o.cap = cap;
    }
    if (!o.ev) DSC_CUDA(ctx, cudaEventCreateWithFlags(&o.ev, cudaEventDisableTiming));
    memcpy(o.h, offsets, cnt * sizeof(int64_t));
    DSC_CUDA(ctx, cudaMemcpyAsync(o.d, o.h, cnt * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    *slot_out = slot;
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
    if (ctx->spill) cudaFree(ctx->spill);
    if (ctx->probe) cudaFree(ctx->probe);
    ctx->probe = nullptr;
    for (int i = 0; i < DSC_OFF_RING; ++i) {
        if (ctx->off[i].d) cudaFree(ctx->off[i].d);
        if (ctx->off[i].h) cudaFreeHost(ctx->off[i].h);
        if (ctx->off[i].ev) cudaEventDestroy(ctx->off[i].ev);
        ctx->off[i] = dsc_ctx::OffSlot();
    }
    if (ctx->minmax) cudaFree(ctx->minmax);
    if (ctx->sync_buf) cudaFree(ctx->sync_buf);
    ctx->sync_buf = nullptr;
    ctx->sync_cap = 0;
    ctx->keys = nullptr;
    ctx->keys_bytes = 0;
    ctx->spill = nullptr;
    ctx->spill_bytes = 0;
    ctx->minmax = nullptr;
    ctx->minmax_cap = 0;
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
    if ((rc = dsc_enter_stream(ctx, ctx->s_k))) return rc;     // a device-mode call may still be using the engine scratch
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
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    int slot = 0;
    if ((rc = dsc_upload_offsets(ctx, offsets, n_scans, st, &slot))) return rc;
    if (ctx->timing) DSC_CUDA(ctx, cudaEventRecord(ctx->ev0, st));
    rc = build_device(ctx, params, (const uint8_t *)pts, ctx->off[slot].d, offsets, n_scans, pts_stride_bytes,
                      out_grids, out_norm, stats, st);
    if (rc) return rc;
    DSC_CUDA(ctx, cudaEventRecord(ctx->off[slot].ev, st));
    ctx->off[slot].used = 1;
    if ((rc = dsc_leave_stream(ctx, st))) return rc;
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
    int rc;
    if (mem == DSC_MEM_DEVICE) {
        if ((rc = dsc_enter_stream(ctx, (cudaStream_t)stream))) return rc;
        if ((rc = normalize_device(ctx, grids, n_grids, cells_per_grid, out, (cudaStream_t)stream))) return rc;
        return dsc_leave_stream(ctx, (cudaStream_t)stream);
    }
    const size_t bytes = (size_t)n_grids * cells_per_grid * 4;
    if ((rc = dsc_reserve_scratch(ctx, bytes * 2))) return rc;
    float *d_in = (float *)ctx->scratch, *d_out = d_in + (size_t)n_grids * cells_per_grid;
    cudaStream_t st = ctx->s_k;
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    DSC_CUDA(ctx, cudaMemcpyAsync(d_in, grids, bytes, cudaMemcpyHostToDevice, st));
    if ((rc = normalize_device(ctx, d_in, n_grids, cells_per_grid, d_out, st))) return rc;
    DSC_CUDA(ctx, cudaMemcpyAsync(out, d_out, bytes, cudaMemcpyDeviceToHost, st));
    DSC_CUDA(ctx, cudaStreamSynchronize(st));
    return DSC_OK;
}
