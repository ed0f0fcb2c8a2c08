// dsc_k1_f64.cu -- K1 for FLOAT64 point clouds: dsc_sc_build_batch_f64.
//
// The reference bins float64 points whenever the cloud went through Open3D's voxel_down_sample
// (np.asarray(d.points) is float64): the live node always (scan_context_node.py:93-100 -> :161-168 ->
// _scan_context, variant B) and the offline generator with use_downsampling=1 (scaner_version2.py:126-133 ->
// ptcloud2sc_vectorized, variant A).  NumPy then runs the polar chain in float64 (dsc_math_f64.cuh).
//   variant A   float64 heights, strict '>' on a float64 grid  -> out: float64 [n][R][S]
//   variant B   heights cast to float32, np.maximum.at on a float32 grid -> out: float32 [n][R][S]
//
// These clouds are small (a down-sampled scan is 20-60 k points, one per call in the live node), so the kernel is
// the simple form: one CTA per (tile of points, scan), atomicMax of order-preserving keys on an L2-resident key grid,
// then a decode pass that also re-zeroes the keys.  Algorithmic bytes per scan: 24*N + (8|4)*R*S.
#include "dsc_internal.h"
#include "dsc_math_f64.cuh"

namespace {

constexpr int F64_THREADS = 256;
constexpr int F64_PPT = 4;
constexpr int F64_TILE = F64_THREADS * F64_PPT;

__device__ __forceinline__ double ld_nc_f64(const void *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

template <int VARIANT>
__global__ void __launch_bounds__(F64_THREADS)
k1f64_build(const uint8_t *__restrict__ pts, const int64_t *__restrict__ offsets, int scan_base, void *__restrict__ keys,
            unsigned long long *__restrict__ stats, int stride_bytes, const dsc::Bin64Cfg cfg)
{
    const int scan_local = blockIdx.y;
    const int scan = scan_base + scan_local;
    const int64_t p0 = offsets[scan];
    const int64_t n = offsets[scan + 1] - p0;
    const int64_t start = (int64_t)blockIdx.x * F64_TILE;
    if (start >= n) return;
    const int64_t cells = (int64_t)cfg.R * cfg.S;
    const int64_t remain = n - start;
    const int in_tile = remain < F64_TILE ? (int)remain : F64_TILE;
    const uint8_t *base = pts + (p0 + start) * (int64_t)stride_bytes;
    unsigned c_nonfinite = 0, c_overflow = 0, c_clamped = 0, c_oor = 0, c_bnd = 0;
    double x[F64_PPT], y[F64_PPT], z[F64_PPT];
#pragma unroll
    for (int k = 0; k < F64_PPT; ++k) {
        const int i = k * F64_THREADS + threadIdx.x;
        if (i < in_tile) {
            const uint8_t *q = base + (int64_t)i * stride_bytes;
            x[k] = ld_nc_f64(q);
            y[k] = ld_nc_f64(q + 8);
            z[k] = ld_nc_f64(q + 16);
        }
    }
#pragma unroll
    for (int k = 0; k < F64_PPT; ++k) {
        const int i = k * F64_THREADS + threadIdx.x;
        if (i >= in_tile) continue;
        const dsc::Bin64Result b = VARIANT == 0 ? dsc::bin64_a(cfg, x[k], y[k], z[k]) : dsc::bin64_b(cfg, x[k], y[k], z[k]);
        c_bnd += (unsigned)b.boundary;
        if (b.status > dsc::PT_CLAMPED) {
            c_nonfinite += (b.status == dsc::PT_NONFINITE);
            c_overflow += (b.status == dsc::PT_SECTOR_OVERFLOW);
            c_oor += (b.status == dsc::PT_OUT_OF_RANGE);
            continue;
        }
        c_clamped += (b.status == dsc::PT_CLAMPED);
        if (VARIANT == 0) {
            const double zv = b.z;
            if (zv != zv) continue;                                  // NaN never wins (scaner_version2.py:107)
            if (!cfg.full_height && !(zv > 0.0)) continue;           // offset mode: cells start at 0, strict '>'
            const unsigned long long key = dsc::double_to_key(zv);
            if (key <= 0x000fffffffffffffull) continue;              // -inf can never beat the -inf init
            atomicMax(reinterpret_cast<unsigned long long *>(keys) + (int64_t)scan_local * cells + b.cell, key);
        } else {
            const float zf = __double2float_rn(b.z);                 // z_used.astype(np.float32), scan_context_node.py:151
            if (!cfg.full_height && !(zf > 0.0f)) continue;          // np.maximum.at against the 0 init
            const uint32_t key = dsc::float_to_key(zf);
            if (key <= 0x007fffffu) continue;                        // -inf (float32 overflow of a huge negative z)
            atomicMax(reinterpret_cast<uint32_t *>(keys) + (int64_t)scan_local * cells + b.cell, key);
        }
    }
    if (stats) {
        unsigned long long *st = stats + (int64_t)scan * DSC_STATS_WORDS;
        const unsigned v0 = __reduce_add_sync(0xffffffffu, c_nonfinite);
        const unsigned v1 = __reduce_add_sync(0xffffffffu, c_overflow);
        const unsigned v2 = __reduce_add_sync(0xffffffffu, c_clamped);
        const unsigned v3 = __reduce_add_sync(0xffffffffu, c_oor);
        const unsigned v4 = __reduce_add_sync(0xffffffffu, c_bnd);
        if ((threadIdx.x & 31) == 0) {
            if (v0) atomicAdd(st + 1, (unsigned long long)v0);
            if (v1) atomicAdd(st + 2, (unsigned long long)v1);
            if (v2) atomicAdd(st + 3, (unsigned long long)v2);
            if (v3) atomicAdd(st + 4, (unsigned long long)v3);
            if (v4) atomicAdd(st + 6, (unsigned long long)v4);
        }
        if (threadIdx.x == 0) atomicAdd(st + 0, (unsigned long long)in_tile);
    }
}

// keys -> grid, re-zeroing the scratch.  Variant B in full-height mode maps every non-finite maximum to 0
// (scan_context_node.py:154-155: a float32 overflow of a huge z would otherwise survive as +inf).
__global__ void __launch_bounds__(256)
k1f64_decode_a(unsigned long long *__restrict__ keys, double *__restrict__ out, int64_t n)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const unsigned long long k = keys[i];
        if (k) keys[i] = 0ull;
        __stcs(out + i, dsc::key_to_double(k));
    }
}
__global__ void __launch_bounds__(256)
k1f64_decode_b(uint32_t *__restrict__ keys, float *__restrict__ out, int64_t n, int full_height)
{
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint32_t k = keys[i];
        if (k) keys[i] = 0u;
        float f = dsc::key_to_float(k);
        if (full_height && !(fabsf(f) < INFINITY)) f = 0.0f;
        __stcs(out + i, f);
    }
}

}  // namespace

static int f64_reserve_keys(dsc_ctx *ctx, size_t need, cudaStream_t st)
{
    if (need <= ctx->keys64_bytes) return DSC_OK;
    if (ctx->keys64) { DSC_CUDA(ctx, cudaDeviceSynchronize()); cudaFree(ctx->keys64); }
    ctx->keys64 = nullptr;
    ctx->keys64_bytes = 0;
    DSC_CUDA(ctx, cudaMalloc(&ctx->keys64, need));
    ctx->keys64_bytes = need;
    DSC_CUDA(ctx, cudaMemsetAsync(ctx->keys64, 0, need, st));     // the decode kernels leave it zeroed
    return DSC_OK;
}

static int f64_build_device(dsc_ctx *ctx, const dsc_params *p, double height_offset, const uint8_t *pts,
                            const int64_t *d_offsets, const int64_t *h_offsets, int n_scans, int stride, void *out,
                            dsc_scan_stats *stats, cudaStream_t st)
{
    const int64_t cells = (int64_t)p->num_ring * p->num_sector;
    const dsc::Bin64Cfg cfg = dsc::make_bin64_cfg(p->num_ring, p->num_sector, p->max_length, p->variant, p->full_height,
                                                  height_offset);
    const size_t ksz = p->variant == DSC_VARIANT_A ? 8 : 4;
    int rc;
    if (stats) DSC_CUDA(ctx, cudaMemsetAsync(stats, 0, sizeof(dsc_scan_stats) * (size_t)n_scans, st));
    int slots = (int)((size_t)48 * 1024 * 1024 / (size_t)(cells * ksz));   // key scratch sized to stay in L2
    if (slots < 1) slots = 1;
    if (slots > n_scans) slots = n_scans;
    if ((rc = f64_reserve_keys(ctx, (size_t)slots * cells * ksz, st))) return rc;
    for (int s0 = 0; s0 < n_scans; s0 += slots) {
        const int g = n_scans - s0 < slots ? n_scans - s0 : slots;
        int64_t max_n = 0;
        for (int s = s0; s < s0 + g; ++s) {
            const int64_t n = h_offsets[s + 1] - h_offsets[s];
            if (n > max_n) max_n = n;
        }
        if (max_n > 0) {
            const int64_t tiles = (max_n + F64_TILE - 1) / F64_TILE;
            if (tiles > 0x7fffffff) DSC_FAIL(ctx, DSC_ERR_INVALID, "scan too large");
            dim3 grid((unsigned)tiles, (unsigned)g);
            if (p->variant == DSC_VARIANT_A)
                k1f64_build<0><<<grid, F64_THREADS, 0, st>>>(pts, d_offsets, s0, ctx->keys64, (unsigned long long *)stats, stride, cfg);
            else
                k1f64_build<1><<<grid, F64_THREADS, 0, st>>>(pts, d_offsets, s0, ctx->keys64, (unsigned long long *)stats, stride, cfg);
            if ((rc = dsc_check_launch(ctx, "k1f64_build"))) return rc;
        }
        const int64_t total = (int64_t)g * cells;
        const int64_t want = (total + 255) / 256, cap = (int64_t)ctx->sm_count * 16;
        const int blocks = (int)(want < cap ? want : cap);
        if (p->variant == DSC_VARIANT_A)
            k1f64_decode_a<<<blocks, 256, 0, st>>>((unsigned long long *)ctx->keys64, (double *)out + (int64_t)s0 * cells, total);
        else
            k1f64_decode_b<<<blocks, 256, 0, st>>>((uint32_t *)ctx->keys64, (float *)out + (int64_t)s0 * cells, total, p->full_height);
        if ((rc = dsc_check_launch(ctx, "k1f64_decode"))) return rc;
    }
    return DSC_OK;
}

void dsc_k1f64_release(dsc_ctx *ctx)
{
    if (ctx->keys64) cudaFree(ctx->keys64);
    ctx->keys64 = nullptr;
    ctx->keys64_bytes = 0;
}

extern "C" int dsc_sc_build_batch_f64(dsc_ctx *ctx, const dsc_params *params, double height_offset, const void *pts,
                                      const int64_t *offsets, int n_scans, int pts_stride_bytes, int mem,
                                      void *out_grids, dsc_scan_stats *stats, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    if (!params) DSC_FAIL(ctx, DSC_ERR_INVALID, "params is NULL");
    if (params->num_ring <= 0 || params->num_sector <= 0 || (int64_t)params->num_ring * params->num_sector > ((int64_t)1 << 26))
        DSC_FAIL(ctx, DSC_ERR_INVALID, "bad num_ring/num_sector");
    if (!(params->max_length > 0.0) || !isfinite(params->max_length)) DSC_FAIL(ctx, DSC_ERR_INVALID, "max_length must be finite and > 0");
    if (params->variant != DSC_VARIANT_A && params->variant != DSC_VARIANT_B) DSC_FAIL(ctx, DSC_ERR_INVALID, "unknown variant");
    if (n_scans < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "n_scans < 0");
    if (n_scans == 0) return DSC_OK;
    if (!offsets || !out_grids) DSC_FAIL(ctx, DSC_ERR_INVALID, "offsets/out_grids is NULL");
    if (pts_stride_bytes != 24 && pts_stride_bytes != 32) DSC_FAIL(ctx, DSC_ERR_INVALID, "pts_stride_bytes must be 24 or 32");
    if (mem != DSC_MEM_HOST && mem != DSC_MEM_DEVICE) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad mem");
    for (int s = 0; s < n_scans; ++s) {
        if (offsets[s + 1] < offsets[s] || offsets[s] < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "offsets must be non-decreasing and >= 0");
        if (offsets[s + 1] - offsets[s] > 0x7fff0000LL) DSC_FAIL(ctx, DSC_ERR_INVALID, "more than 2^31 points in one scan");
    }
    if (offsets[n_scans] > offsets[0] && !pts) DSC_FAIL(ctx, DSC_ERR_INVALID, "pts is NULL");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    const int64_t cells = (int64_t)params->num_ring * params->num_sector;
    const size_t osz = params->variant == DSC_VARIANT_A ? 8 : 4;
    int rc;
    if (mem == DSC_MEM_DEVICE) {
        cudaStream_t st = (cudaStream_t)stream;
        if ((rc = dsc_enter_stream(ctx, st))) return rc;
        int slot = 0;
        if ((rc = dsc_upload_offsets(ctx, offsets, n_scans, st, &slot))) return rc;
        if ((rc = f64_build_device(ctx, params, height_offset, (const uint8_t *)pts, ctx->off[slot].d, offsets, n_scans,
                                   pts_stride_bytes, out_grids, stats, st)))
            return rc;
        DSC_CUDA(ctx, cudaEventRecord(ctx->off[slot].ev, st));
        ctx->off[slot].used = 1;
        return dsc_leave_stream(ctx, st);
    }
    // host buffers: one upload, the kernels, one download (live-node sized inputs; no chunk pipeline needed)
    cudaStream_t st = ctx->s_k;
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    const int64_t first = offsets[0], total = offsets[n_scans] - first;
    size_t o = 0;
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    const size_t o_pts = o;   o = up(o + (size_t)total * pts_stride_bytes);
    const size_t o_out = o;   o = up(o + (size_t)n_scans * cells * osz);
    const size_t o_off = o;   o = up(o + ((size_t)n_scans + 1) * 8);
    const size_t o_st = o;    o = up(o + (stats ? (size_t)n_scans * sizeof(dsc_scan_stats) : 0));
    if ((rc = dsc_reserve_scratch(ctx, o))) return rc;
    uint8_t *sc = (uint8_t *)ctx->scratch;
    std::vector<int64_t> rel((size_t)n_scans + 1);
    for (int s = 0; s <= n_scans; ++s) rel[s] = offsets[s] - first;
    // pageable sources: the copies are staged by the driver before the calls return
    DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_off, rel.data(), rel.size() * 8, cudaMemcpyHostToDevice, st));
    if (total)
        DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_pts, (const uint8_t *)pts + (size_t)first * pts_stride_bytes,
                                      (size_t)total * pts_stride_bytes, cudaMemcpyHostToDevice, st));
    if ((rc = f64_build_device(ctx, params, height_offset, sc + o_pts, (const int64_t *)(sc + o_off), rel.data(), n_scans,
                               pts_stride_bytes, sc + o_out, stats ? (dsc_scan_stats *)(sc + o_st) : nullptr, st)))
        return rc;
    DSC_CUDA(ctx, cudaMemcpyAsync(out_grids, sc + o_out, (size_t)n_scans * cells * osz, cudaMemcpyDeviceToHost, st));
    if (stats) DSC_CUDA(ctx, cudaMemcpyAsync(stats, sc + o_st, (size_t)n_scans * sizeof(dsc_scan_stats), cudaMemcpyDeviceToHost, st));
    DSC_CUDA(ctx, cudaStreamSynchronize(st));
    return DSC_OK;
}
