// dsc_k6_labels.cu -- K6: label-side training tensors built on the device from the rasteriser's output,
// so that a dataset __getitem__ no longer walks JSON cell lists in Python.
//
//   dsc_label_maps     binary mask + adaptive weight map of ScanContextPedestrianDataset.__getitem__
//                      (distro_model/tiny_sc_pred_v2.py:57-85), the binary mask of build_mask_from_json
//                      (tiny_sc_ped.py:62-73) and the clip/wrap mask of ScanContextPedDataset.__getitem__
//                      (distro_model/train_sc_ped.py:175-183)
//   dsc_label_heatmap  centre heat-map: np.maximum over the pedestrians of gaussian2d(center, sigma)
//                      (distro_model/train_sc_ped.py:74-79,165-173)
//
// Numerics.  The per-pedestrian weight min(min(100/num_cells, 5) * (1 + (distance/20)^2), 10) is scalar Python-float
// arithmetic and stays with the caller (labels.py evaluates the reference's very expressions); the device only takes
// the float32 maximum per cell -- the reference's `weight_map[r, s] = max(weight_map[r, s], ped_weight)` on a float32
// array is float32(max) because rounding is monotone.  The Gaussian is a function of the INTEGER squared distance
// d2 only; the caller tabulates exp(-0.5 * d2 / sigma**2) with NumPy for d2 = 0 .. len-1 (up to where it underflows to
// 0) exactly as gaussian2d evaluates it, and the kernel takes table look-ups and float64 maxima: bit-identical to the
// reference on any host, without re-implementing NumPy's exp.
#include "dsc_internal.h"

namespace {

// one thread per covered cell of every box (CSR).  mode 0: cells outside the grid are ignored
// (tiny_sc_pred_v2.py:81, tiny_sc_ped.py:70); mode 1: ring clipped, sector wrapped with np.mod (train_sc_ped.py:181-182)
__global__ void __launch_bounds__(256)
k6_maps(const int32_t *__restrict__ cells, const int64_t *__restrict__ cell_offsets, const float *__restrict__ box_weight,
        const int32_t *__restrict__ box_frame, int n_boxes, int n_frames, int R, int S, int mode,
        float *__restrict__ mask, float *__restrict__ weight)
{
    const int64_t total = cell_offsets[n_boxes] - cell_offsets[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const int64_t c = cell_offsets[0] + i;
        int lo = 0, hi = n_boxes;                        // box of this cell: last b with cell_offsets[b] <= c
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cell_offsets[mid] <= c) lo = mid; else hi = mid;
        }
        const int frame = box_frame ? box_frame[lo] : 0;
        if (frame < 0 || frame >= n_frames) continue;
        int r = cells[2 * c], s = cells[2 * c + 1];
        if (mode == 0) {
            if (r < 0 || r >= R || s < 0 || s >= S) continue;
        } else {
            r = r < 0 ? 0 : (r > R - 1 ? R - 1 : r);
            s %= S;
            if (s < 0) s += S;                           // np.mod: result has the sign of the divisor
        }
        const int64_t o = ((int64_t)frame * R + r) * S + s;
        mask[o] = 1.0f;
        if (weight) {
            const float w = box_weight[lo];
            if (w > 1.0f) atomicMax(reinterpret_cast<unsigned int *>(weight) + o, __float_as_uint(w));   // positive floats order like uints
        }
    }
}

__global__ void __launch_bounds__(256)
k6_fill(float *__restrict__ p, int64_t n, float v)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) p[i] = v;
}

// one CTA per centre: every cell within the table's support takes max(heat, table[d2]).  Values are >= 0, so the
// float64 maximum is an unsigned 64-bit atomicMax on the bit patterns.
__global__ void __launch_bounds__(256)
k6_heat(const int32_t *__restrict__ centers, const int32_t *__restrict__ center_frame, int n_frames, int R, int S,
        const double *__restrict__ table, int table_len, int radius, double *__restrict__ heat)
{
    const int ci = blockIdx.x;
    const int cr = centers[2 * ci], cs = centers[2 * ci + 1];
    const int frame = center_frame ? center_frame[ci] : 0;
    if (frame < 0 || frame >= n_frames) return;
    if (cr < 0 || cr >= R || cs < 0 || cs >= S) return;          // train_sc_ped.py:170
    const int r_lo = cr - radius < 0 ? 0 : cr - radius, r_hi = cr + radius > R - 1 ? R - 1 : cr + radius;
    const int s_lo = cs - radius < 0 ? 0 : cs - radius, s_hi = cs + radius > S - 1 ? S - 1 : cs + radius;
    const int w = s_hi - s_lo + 1, n = (r_hi - r_lo + 1) * w;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int r = r_lo + i / w, s = s_lo + i % w;
        const long long d2 = (long long)(r - cr) * (r - cr) + (long long)(s - cs) * (s - cs);
        if (d2 >= table_len) continue;
        const double g = table[d2];
        if (g > 0.0)
            atomicMax(reinterpret_cast<unsigned long long *>(heat) + ((int64_t)frame * R + r) * S + s,
                      (unsigned long long)__double_as_longlong(g));
    }
}

inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" int dsc_label_maps(dsc_ctx *ctx, const int32_t *cells, const int64_t *cell_offsets, const float *box_weight,
                              const int32_t *box_frame, int n_boxes, int n_frames, int num_ring, int num_sector, int mode,
                              int mem, float *mask_out, float *weight_out, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    if (n_boxes < 0 || n_frames <= 0 || num_ring <= 0 || num_sector <= 0 || (int64_t)num_ring * num_sector > ((int64_t)1 << 26))
        DSC_FAIL(ctx, DSC_ERR_INVALID, "bad sizes");
    if (mode != 0 && mode != 1) DSC_FAIL(ctx, DSC_ERR_INVALID, "mode must be 0 (ignore out-of-grid cells) or 1 (clip ring, wrap sector)");
    if (!mask_out) DSC_FAIL(ctx, DSC_ERR_INVALID, "mask_out is NULL");
    if (n_boxes > 0 && !cell_offsets) DSC_FAIL(ctx, DSC_ERR_INVALID, "cell_offsets is NULL");
    if (weight_out && n_boxes > 0 && !box_weight) DSC_FAIL(ctx, DSC_ERR_INVALID, "box_weight is NULL");
    if (mem != DSC_MEM_HOST && mem != DSC_MEM_DEVICE) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad mem");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t grid_cells = (int64_t)n_frames * num_ring * num_sector;
    const bool host = mem == DSC_MEM_HOST;
    cudaStream_t st = host ? ctx->s_k : (cudaStream_t)stream;
    int rc;
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    // cell_offsets is a HOST array in host mode and a DEVICE array in device mode (it is what dsc_raster_boxes wrote)
    int64_t total = 0;
    std::vector<int64_t> h_off;
    if (host && n_boxes > 0) {
        total = cell_offsets[n_boxes] - cell_offsets[0];
        if (total < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "cell_offsets must be non-decreasing");
        if (total > 0 && !cells) DSC_FAIL(ctx, DSC_ERR_INVALID, "cells is NULL");
    }
    const int32_t *d_cells = cells;
    const int64_t *d_off = cell_offsets;
    const float *d_w = box_weight;
    const int32_t *d_frame = box_frame;
    float *d_mask = mask_out, *d_weight = weight_out;
    if (host) {
        size_t o = 0;
        const size_t o_cells = o;  o = up256(o + (size_t)total * 8);
        const size_t o_off = o;    o = up256(o + ((size_t)n_boxes + 1) * 8);
        const size_t o_w = o;      o = up256(o + (size_t)n_boxes * 4);
        const size_t o_fr = o;     o = up256(o + (size_t)n_boxes * 4);
        const size_t o_mask = o;   o = up256(o + (size_t)grid_cells * 4);
        const size_t o_wt = o;     o = up256(o + (weight_out ? (size_t)grid_cells * 4 : 0));
        if ((rc = dsc_reserve_scratch(ctx, o))) return rc;
        uint8_t *sc = (uint8_t *)ctx->scratch;
        if (n_boxes > 0) {
            h_off.resize((size_t)n_boxes + 1);
            for (int b = 0; b <= n_boxes; ++b) h_off[b] = cell_offsets[b] - cell_offsets[0];
            DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_off, h_off.data(), h_off.size() * 8, cudaMemcpyHostToDevice, st));
            if (total) DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_cells, cells + 2 * cell_offsets[0], (size_t)total * 8, cudaMemcpyHostToDevice, st));
            if (weight_out) DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_w, box_weight, (size_t)n_boxes * 4, cudaMemcpyHostToDevice, st));
            if (box_frame) DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_fr, box_frame, (size_t)n_boxes * 4, cudaMemcpyHostToDevice, st));
        }
        d_cells = (const int32_t *)(sc + o_cells);
        d_off = (const int64_t *)(sc + o_off);
        d_w = (const float *)(sc + o_w);
        d_frame = box_frame ? (const int32_t *)(sc + o_fr) : nullptr;
        d_mask = (float *)(sc + o_mask);
        d_weight = weight_out ? (float *)(sc + o_wt) : nullptr;
    }
    DSC_CUDA(ctx, cudaMemsetAsync(d_mask, 0, (size_t)grid_cells * 4, st));
    const int fill_blocks = (int)((grid_cells + 255) / 256 < (int64_t)ctx->sm_count * 8 ? (grid_cells + 255) / 256 : (int64_t)ctx->sm_count * 8);
    if (d_weight) {
        k6_fill<<<fill_blocks, 256, 0, st>>>(d_weight, grid_cells, 1.0f);       // weight_map = np.ones_like(...)
        if ((rc = dsc_check_launch(ctx, "k6_fill"))) return rc;
    }
    if (n_boxes > 0) {
        k6_maps<<<ctx->sm_count * 4, 256, 0, st>>>(d_cells, d_off, d_w, d_frame, n_boxes, n_frames, num_ring, num_sector, mode,
                                                   d_mask, d_weight);
        if ((rc = dsc_check_launch(ctx, "k6_maps"))) return rc;
    }
    if (!host) return dsc_leave_stream(ctx, st);
    DSC_CUDA(ctx, cudaMemcpyAsync(mask_out, d_mask, (size_t)grid_cells * 4, cudaMemcpyDeviceToHost, st));
    if (weight_out) DSC_CUDA(ctx, cudaMemcpyAsync(weight_out, d_weight, (size_t)grid_cells * 4, cudaMemcpyDeviceToHost, st));
    DSC_CUDA(ctx, cudaStreamSynchronize(st));
    return DSC_OK;
}

extern "C" int dsc_label_heatmap(dsc_ctx *ctx, const int32_t *centers, const int32_t *center_frame, int n_centers,
                                 int n_frames, int num_ring, int num_sector, const double *table, int table_len, int mem,
                                 double *heat_out, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    if (n_centers < 0 || n_frames <= 0 || num_ring <= 0 || num_sector <= 0 || (int64_t)num_ring * num_sector > ((int64_t)1 << 26))
        DSC_FAIL(ctx, DSC_ERR_INVALID, "bad sizes");
    if (!heat_out) DSC_FAIL(ctx, DSC_ERR_INVALID, "heat_out is NULL");
    if (n_centers > 0 && (!centers || !table || table_len <= 0)) DSC_FAIL(ctx, DSC_ERR_INVALID, "centers/table missing");
    if (mem != DSC_MEM_HOST && mem != DSC_MEM_DEVICE) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad mem");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t grid_cells = (int64_t)n_frames * num_ring * num_sector;
    const bool host = mem == DSC_MEM_HOST;
    cudaStream_t st = host ? ctx->s_k : (cudaStream_t)stream;
    int rc;
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    // the Gaussian table is a small HOST array in both modes (like the polar tables)
    size_t o = 0;
    const size_t o_tab = o;   o = up256(o + (size_t)(table_len > 0 ? table_len : 0) * 8);
    const size_t o_cen = o;   o = up256(o + (host ? (size_t)n_centers * 8 : 0));
    const size_t o_fr = o;    o = up256(o + (host ? (size_t)n_centers * 4 : 0));
    const size_t o_heat = o;  o = up256(o + (host ? (size_t)grid_cells * 8 : 0));
    if ((rc = dsc_reserve_scratch(ctx, o))) return rc;
    uint8_t *sc = (uint8_t *)ctx->scratch;
    const int32_t *d_cen = centers, *d_fr = center_frame;
    double *d_heat = heat_out;
    if (n_centers > 0) DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_tab, table, (size_t)table_len * 8, cudaMemcpyHostToDevice, st));
    if (host) {
        if (n_centers > 0) {
            DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_cen, centers, (size_t)n_centers * 8, cudaMemcpyHostToDevice, st));
            if (center_frame) DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_fr, center_frame, (size_t)n_centers * 4, cudaMemcpyHostToDevice, st));
        }
        d_cen = (const int32_t *)(sc + o_cen);
        d_fr = center_frame ? (const int32_t *)(sc + o_fr) : nullptr;
        d_heat = (double *)(sc + o_heat);
    }
    DSC_CUDA(ctx, cudaMemsetAsync(d_heat, 0, (size_t)grid_cells * 8, st));      // heat = np.zeros_like(sc)
    if (n_centers > 0) {
        int radius = 0;
        while ((long long)(radius + 1) * (radius + 1) < table_len) ++radius;    // largest offset with d2 < table_len
        k6_heat<<<n_centers, 256, 0, st>>>(d_cen, d_fr, n_frames, num_ring, num_sector, (const double *)(sc + o_tab), table_len,
                                           radius, d_heat);
        if ((rc = dsc_check_launch(ctx, "k6_heat"))) return rc;
    }
    if (!host) {
        // the table copy above came from pageable memory (staged before the call returned); the scratch it sits in is
        // ordered against later calls by the stream chain
        return dsc_leave_stream(ctx, st);
    }
    DSC_CUDA(ctx, cudaMemcpyAsync(heat_out, d_heat, (size_t)grid_cells * 8, cudaMemcpyDeviceToHost, st));
    DSC_CUDA(ctx, cudaStreamSynchronize(st));
    return DSC_OK;
}
