// dsc_internal.h -- context and helpers shared by the engine's translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/dsc.h"
#include "dsc_math.cuh"

// staging buffers of the host-buffer pipeline (H2D of chunk i+2 / i+1, kernels of chunk i, D2H of chunk i-1)
#ifndef DSC_NBUF
#define DSC_NBUF 3
#endif
#define DSC_OFF_RING 4
#define DSC_STATS_WORDS ((int)(sizeof(dsc_scan_stats) / 8))     // counters per scan, as 64-bit words

struct dsc_ctx {
    int device = 0;
    int sm_count = 148;
    int l2_bytes = 0;
    size_t smem_optin = 0;
    std::string err;
    int64_t launches = 0;
    int timing = 0;
    int sc_smem_rows = -1;           // scan kernel: ring rows kept in shared memory (-1 = as many as fit, 0 = none)
    int no_defgrid = 0;              // 1: do not use the default-grid specialisation (A/B runs)
    int sc_cfg = -1;                 // scan kernel launch shape (screened default-grid kernels): -1 = auto, 0 = 768 thr x 6 loads, 1 = 768x3, 2 = 1024x2
    int sc_split = -1;               // scan kernel: -1 auto, 0 never split scans into items, 1 always
    int sc_item_points = 131072;     // scan kernel: target points per work item when scans are split
    void *sync_buf = nullptr;        // scan kernel: work-item cursor, per-scan retire counters, item table
    size_t sync_cap = 0;
    int k1_impl = 0;                 // 0 auto, 1 generic L2-key kernels only, 3 scan kernel whenever the shape allows
    float last_ms = -1.0f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // K1 scratch
    uint32_t *keys = nullptr;        // generic path: [slots][cells] order-preserving keys (zero == empty; left zeroed)
    size_t keys_bytes = 0;
    void *spill = nullptr;           // scan kernel: per-CTA spill lists of (cell, key) pairs
    size_t spill_bytes = 0;
    void *probe = nullptr;           // scan kernel: ring histogram of k1_probe + the resident window's first ring
    int sc_window = -1;              // -1: k1_probe picks the window; >= 0: fixed first ring (A/B runs)
    void *keys64 = nullptr;          // float64-input path: [slots][cells] 64-bit (variant A) or 32-bit (variant B) keys
    size_t keys64_bytes = 0;
    struct OffSlot {                 // offsets of a device-mode call: pinned staging + device copy + "consumed" event
        int64_t *h = nullptr, *d = nullptr;
        size_t cap = 0;              // in elements
        cudaEvent_t ev = nullptr;
        int used = 0;
    };
    OffSlot off[DSC_OFF_RING];
    int off_next = 0;
    cudaEvent_t ev_call = nullptr;   // end of the last device-mode call (the engine scratch is shared between calls)
    int ev_call_set = 0;
    uint32_t *minmax = nullptr;      // [n][2] key-encoded per-grid min / max for normalise
    size_t minmax_cap = 0;           // in grids
    int *work_counter = nullptr;     // persistent-kernel tile cursor

    // host-path pipeline: H2D stream -> kernel stream -> D2H stream, double buffered
    cudaStream_t s_h2d = nullptr, s_k = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_h2d[DSC_NBUF] = {}, ev_k[DSC_NBUF] = {}, ev_d2h[DSC_NBUF] = {};
    void *stage_pts[DSC_NBUF] = {};
    size_t stage_pts_cap = 0;
    float *stage_out[DSC_NBUF] = {};
    float *stage_norm[DSC_NBUF] = {};
    dsc_scan_stats *stage_stats[DSC_NBUF] = {};
    int64_t *stage_doff[DSC_NBUF] = {};   // device offsets per buffer
    int64_t *stage_hoff[DSC_NBUF] = {};   // pinned offsets per buffer
    size_t stage_scans_cap = 0;
    size_t stage_cells = 0;

    // polar tables (K2/K3/K5)
    int tab_R = 0, tab_S = 0;
    double *d_rc = nullptr, *d_cos = nullptr, *d_sin = nullptr;

    // generic device scratch for K2/K3 (counts, offsets, temp copies of host inputs)
    void *scratch = nullptr;
    size_t scratch_cap = 0;
};

#define DSC_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
            return DSC_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define DSC_FAIL(ctx, code, msg)                                                              \
    do {                                                                                      \
        (ctx)->err = (msg);                                                                   \
        return (code);                                                                        \
    } while (0)

// grows ctx->scratch to at least `bytes`
int dsc_reserve_scratch(dsc_ctx *ctx, size_t bytes);
// checks the launch that was just issued
int dsc_check_launch(dsc_ctx *ctx, const char *what);
// Device-mode calls are asynchronous and share the context's scratch buffers: dsc_enter_stream orders the call
// after the previous one (a no-op wait when both use the same stream), dsc_leave_stream marks its end.
int dsc_enter_stream(dsc_ctx *ctx, cudaStream_t st);
int dsc_leave_stream(dsc_ctx *ctx, cudaStream_t st);
// Offsets of a device-mode call -> ctx->off[*slot_out].d (a ring of pinned + device buffers guarded by events; the
// caller records ctx->off[slot].ev after the kernels that read them and sets .used)
int dsc_upload_offsets(dsc_ctx *ctx, const int64_t *offsets, int n_scans, cudaStream_t st, int *slot_out);
