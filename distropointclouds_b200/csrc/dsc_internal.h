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
    int sc_cfg = 0;                  // scan kernel launch shape (screened default-grid kernels): 0 = 768 thr x 3 loads, 1 = 1024x3, 2 = 1024x2
    size_t persist_max = 0, window_max = 0, persist_set = 0;   // L2 persistence limits of the device
    int l2_persist = 0;              // 1: keep the key slots in the persisting part of L2 during the launch
    int sc_loader = 0;               // scan kernel point loader: 0 = 128-bit register loads, 1-4 = TMA ring shapes (k1_ring)
    int sc_split = -1;               // scan kernel: -1 auto, 0 never split scans into items, 1 always
    int sc_item_points = 131072;     // scan kernel: target points per work item when scans are split
    int st_cfg = 0;                  // streaming kernel shape: 0 = 8 pts/thread, 3 CTAs/SM; 1 = 6 pts/thread, 4 CTAs/SM
    int st_slots = 64;               // key-grid slots of the streaming kernel (L2-resident ring)
    void *sync_buf = nullptr;        // streaming kernel: cursor, per-scan counters, tile table
    size_t sync_cap = 0;
    int pf_dist = 0;                 // bulk L2 prefetch distance in steps (0 = off)
    int hg_cfg = 0;                  // half-grid launch shape: 0 = 1024 thr x 2 loads x 2 sets, 1 = 1024x3x2, 2 = 512x4x2 (A/B runs)
    int k1_impl = 0;                 // 0 auto, 1 chunked L2-key kernels, 2 half-grid, 3 scan-per-CTA, 4 tile kernel
    float last_ms = -1.0f;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;

    // K1 scratch
    uint32_t *keys = nullptr;        // [keys_slots][cells] order-preserving keys (zero == empty)
    size_t keys_bytes = 0;
    int64_t *d_offsets = nullptr;    // device copy of the caller's offsets
    size_t d_offsets_cap = 0;        // in elements
    int64_t *h_offsets = nullptr;    // pinned staging for offsets
    size_t h_offsets_cap = 0;
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
