// dsc_k7_voxel.cu -- K7: deterministic voxel-grid down-sampling, the step the live node runs before
// every scan context: ScanContextComputer._downsample_voxel (scan_context_node.py:93-100) and the offline generator's
// use_downsampling branch (scaner_version2.py:126-133) call Open3D's PointCloud.voxel_down_sample.
//
// Open3D is an un-vendored, unpinned dependency that is not installed here, so PARITY IS UNPINNED for this kernel:
// it restates the algorithm Open3D publishes (PointCloud::VoxelDownSample, unchanged from 0.9 to 0.18):
//     voxel_min_bound = min_bound - voxel_size * 0.5                                (float64)
//     voxel index     = floor((p - voxel_min_bound) / voxel_size) per axis          (true division, float64)
//     output point    = (sum of the voxel's points, accumulated in input order) / count
// The float64 sums are order dependent, so the points of a voxel are added in their input order exactly as Open3D's
// sequential loop adds them: the cloud is STABLY sorted by voxel key first.  What cannot be reproduced is the ORDER of
// the output points -- Open3D iterates an unordered_map -- so the output is sorted by (ix, iy, iz); the scan context
// built from it is a per-cell maximum and does not depend on that order.  Non-finite points are dropped (and counted)
// instead of feeding Open3D's undefined int(floor(nan)).
//
// A down-sampled cloud is one LiDAR frame (~100 k points in, ~30 k out) at 10 Hz, latency- not bandwidth-bound, so the
// whole filter is ONE persistent CTA: bounds -> keys -> LSD radix sort (8-bit digits, only the digits the extent needs;
// per-thread contiguous chunks + digit-major counters make every pass stable) -> segmented sequential sums.
#include "dsc_internal.h"
#include "dsc_math_f64.cuh"

namespace {

constexpr int VX_THREADS = 1024;
constexpr int VX_AXIS_BITS = 21;                 // voxel indices per axis must fit 21 bits (2 M voxels per axis)

struct VoxelScratch {
    unsigned long long *key[2];                  // [n] voxel keys (ping-pong)
    uint32_t *idx[2];                            // [n] original point index (ping-pong)
    uint32_t *cnt;                               // [256][VX_THREADS] digit counters / offsets
    unsigned long long *bounds;                  // [3] order-preserving keys of -min per axis (atomicMax), [3] max
    long long *result;                           // [0] n_out, [1] n_dropped (non-finite), [2] error flag
};

__device__ __forceinline__ void load_point(const uint8_t *pts, int64_t i, int stride, bool f64, double &x, double &y, double &z)
{
    const uint8_t *q = pts + i * (int64_t)stride;
    if (f64) {
        const double *d = reinterpret_cast<const double *>(q);
        x = d[0]; y = d[1]; z = d[2];
    } else {
        const float *f = reinterpret_cast<const float *>(q);
        x = (double)f[0]; y = (double)f[1]; z = (double)f[2];       // Vector3dVector: exact widening
    }
}

// CTA-wide exclusive scan of one value per thread (returns the exclusive prefix; *total = sum)
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *s_warp, uint32_t *total)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = lane < VX_THREADS / 32 ? s_warp[lane] : 0u, winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += t;
        }
        s_warp[lane] = winc - w;
        if (lane == 31) s_warp[32] = winc;
    }
    __syncthreads();
    const uint32_t r = s_warp[warp] + inc - v;
    *total = s_warp[32];
    __syncthreads();
    return r;
}

__global__ void __launch_bounds__(VX_THREADS, 1)
k7_voxel(const uint8_t *__restrict__ pts, const int64_t n, const int stride, const int f64, const double voxel,
         VoxelScratch sc, double *__restrict__ out, const int64_t cap)
{
    __shared__ uint32_t s_warp[33];
    __shared__ unsigned long long s_red[VX_THREADS / 32];
    __shared__ double s_vmin[3];
    __shared__ int s_npass;
    const int tid = threadIdx.x;
    const int64_t chunk = (n + VX_THREADS - 1) / VX_THREADS, c0 = (int64_t)tid * chunk;
    const int64_t c1 = c0 + chunk < n ? c0 + chunk : n;

    // ---- 1. bounds over the finite points (GetMinBound / GetMaxBound) ------------------------------------------
    double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (int64_t i = tid; i < n; i += VX_THREADS) {
        double p[3];
        load_point(pts, i, stride, f64 != 0, p[0], p[1], p[2]);
        if (!(fabs(p[0]) < INFINITY) || !(fabs(p[1]) < INFINITY) || !(fabs(p[2]) < INFINITY)) continue;
#pragma unroll
        for (int a = 0; a < 3; ++a) { mn[a] = fmin(mn[a], p[a]); mx[a] = fmax(mx[a], p[a]); }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        atomicMax(sc.bounds + a, dsc::double_to_key(-mn[a]));
        atomicMax(sc.bounds + 3 + a, dsc::double_to_key(mx[a]));
    }
    __syncthreads();
    if (tid < 3) {
        const unsigned long long kmin = sc.bounds[tid], kmax = sc.bounds[3 + tid];
        const double lo = -dsc::u2d((kmin >> 63) ? (kmin & 0x7fffffffffffffffull) : ~kmin);
        const double hi = dsc::u2d((kmax >> 63) ? (kmax & 0x7fffffffffffffffull) : ~kmax);
        s_vmin[tid] = dsc::dsub(lo, dsc::dmul(voxel, 0.5));                  // voxel_min_bound
        // extent check: every index must fit VX_AXIS_BITS bits
        if (hi >= lo && floor(dsc::ddiv(dsc::dsub(hi, s_vmin[tid]), voxel)) >= (double)(1 << VX_AXIS_BITS)) sc.result[2] = 1;
    }
    __syncthreads();
    if (sc.result[2]) return;

    // ---- 2. voxel keys (non-finite points get the all-ones key: sorted to the end and cut off) -------------------
    unsigned long long kmax = 0;
    uint32_t n_bad = 0;
    for (int64_t i = tid; i < n; i += VX_THREADS) {
        double p[3];
        load_point(pts, i, stride, f64 != 0, p[0], p[1], p[2]);
        unsigned long long key = ~0ull;
        if ((fabs(p[0]) < INFINITY) && (fabs(p[1]) < INFINITY) && (fabs(p[2]) < INFINITY)) {
            const unsigned long long ix = (unsigned long long)floor(dsc::ddiv(dsc::dsub(p[0], s_vmin[0]), voxel));
            const unsigned long long iy = (unsigned long long)floor(dsc::ddiv(dsc::dsub(p[1], s_vmin[1]), voxel));
            const unsigned long long iz = (unsigned long long)floor(dsc::ddiv(dsc::dsub(p[2], s_vmin[2]), voxel));
            key = (ix << (2 * VX_AXIS_BITS)) | (iy << VX_AXIS_BITS) | iz;
            kmax = key > kmax ? key : kmax;
        } else {
            ++n_bad;
        }
        sc.key[0][i] = key;
        sc.idx[0][i] = (uint32_t)i;
    }
    // number of 8-bit digits the largest real key needs (bad keys: one extra pass on the top digit keeps them last)
    for (int o = 16; o; o >>= 1) {
        const unsigned long long t = __shfl_xor_sync(0xffffffffu, kmax, o);
        kmax = t > kmax ? t : kmax;
        n_bad += __shfl_xor_sync(0xffffffffu, n_bad, o);
    }
    if ((tid & 31) == 0) s_red[tid >> 5] = kmax;
    __syncthreads();
    if (tid == 0) {
        unsigned long long m = 0;
        for (int w = 0; w < VX_THREADS / 32; ++w) m = s_red[w] > m ? s_red[w] : m;
        int np = 1;
        while (np < 8 && (m >> (8 * np))) ++np;
        s_npass = np;
    }
    {
        uint32_t tot;
        block_exclusive_scan((tid & 31) == 0 ? n_bad : 0u, s_warp, &tot);
        if (tid == 0) sc.result[1] = (long long)tot;
    }
    __syncthreads();
    const int npass = s_npass;

    // ---- 3. stable LSD radix sort of (key, idx); bad keys carry digit 255 in the extra final pass ------------------
    int cur = 0;
    for (int pass = 0; pass <= npass; ++pass) {
        const bool last = pass == npass;             // final pass: separates the non-finite points (key == ~0) to the end
        const int shift = 8 * pass;
        for (int d = 0; d < 256; ++d) sc.cnt[d * VX_THREADS + tid] = 0u;
        for (int64_t i = c0; i < c1; ++i) {
            const unsigned long long k = sc.key[cur][i];
            const int d = last ? (k == ~0ull ? 1 : 0) : (int)((k >> shift) & 0xffu);
            sc.cnt[d * VX_THREADS + tid] += 1u;      // this thread's private column
        }
        __syncthreads();
        // exclusive scan of cnt in (digit, thread) order: thread t owns entries [256 t, 256 t + 256)
        uint32_t local = 0;
        for (int j = 0; j < 256; ++j) local += sc.cnt[tid * 256 + j];
        uint32_t tot;
        uint32_t base = block_exclusive_scan(local, s_warp, &tot);
        for (int j = 0; j < 256; ++j) {
            const uint32_t v = sc.cnt[tid * 256 + j];
            sc.cnt[tid * 256 + j] = base;
            base += v;
        }
        __syncthreads();
        for (int64_t i = c0; i < c1; ++i) {
            const unsigned long long k = sc.key[cur][i];
            const int d = last ? (k == ~0ull ? 1 : 0) : (int)((k >> shift) & 0xffu);
            const uint32_t o = sc.cnt[d * VX_THREADS + tid]++;
            sc.key[cur ^ 1][o] = k;
            sc.idx[cur ^ 1][o] = sc.idx[cur][i];
        }
        __syncthreads();
        cur ^= 1;
    }
    const unsigned long long *key = sc.key[cur];
    const uint32_t *idx = sc.idx[cur];
    const int64_t n_ok = n - sc.result[1];

    // ---- 4. segments: one output point per distinct key, points added in input order ------------------------------
    uint32_t starts = 0;
    for (int64_t i = c0; i < c1 && i < n_ok; ++i) starts += (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
    uint32_t total;
    uint32_t o = block_exclusive_scan(starts, s_warp, &total);
    if (tid == 0) sc.result[0] = (long long)total;
    for (int64_t i = c0; i < c1 && i < n_ok; ++i) {
        if (!(i == 0 || key[i] != key[i - 1])) continue;
        const unsigned long long k = key[i];
        double ax = 0.0, ay = 0.0, az = 0.0;
        int64_t j = i;
        for (; j < n_ok && key[j] == k; ++j) {                   // AccumulatedPoint::AddPoint, in input order
            double x, y, z;
            load_point(pts, idx[j], stride, f64 != 0, x, y, z);
            ax = dsc::dadd(ax, x); ay = dsc::dadd(ay, y); az = dsc::dadd(az, z);
        }
        if ((int64_t)o < cap) {
            const double c = (double)(j - i);                    // GetAveragePoint: point_ / double(num_of_points_)
            out[3 * (int64_t)o + 0] = dsc::ddiv(ax, c);
            out[3 * (int64_t)o + 1] = dsc::ddiv(ay, c);
            out[3 * (int64_t)o + 2] = dsc::ddiv(az, c);
        }
        ++o;
    }
}

inline size_t up256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

extern "C" int dsc_voxel_downsample(dsc_ctx *ctx, const void *pts, int64_t n_points, int pts_f64, int pts_stride_bytes,
                                    double voxel_size, int mem, double *out_pts, int64_t cap, int64_t *n_out,
                                    int64_t *n_dropped, void *stream)
{
    if (!ctx) return DSC_ERR_INVALID;
    if (n_points < 0 || cap < 0) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad sizes");
    if (!(voxel_size > 0.0) || !isfinite(voxel_size)) DSC_FAIL(ctx, DSC_ERR_INVALID, "voxel_size <= 0");     // Open3D: same error
    if (!n_out) DSC_FAIL(ctx, DSC_ERR_INVALID, "n_out is NULL");
    const int elem = pts_f64 ? 8 : 4;
    if (n_points > 0 && (pts_stride_bytes < 3 * elem || pts_stride_bytes % elem)) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad pts_stride_bytes");
    if (n_points > 0x7fffffffLL) DSC_FAIL(ctx, DSC_ERR_INVALID, "too many points");
    if (mem != DSC_MEM_HOST && mem != DSC_MEM_DEVICE) DSC_FAIL(ctx, DSC_ERR_INVALID, "bad mem");
    if (n_points > 0 && !pts) DSC_FAIL(ctx, DSC_ERR_INVALID, "pts is NULL");
    if (cap > 0 && !out_pts) DSC_FAIL(ctx, DSC_ERR_INVALID, "out_pts is NULL");
    DSC_CUDA(ctx, cudaSetDevice(ctx->device));
    const bool host = mem == DSC_MEM_HOST;
    cudaStream_t st = host ? ctx->s_k : (cudaStream_t)stream;
    int rc;
    if ((rc = dsc_enter_stream(ctx, st))) return rc;
    const size_t n = (size_t)n_points;
    size_t o = 0;
    const size_t o_res = o;   o = up256(o + 64);
    const size_t o_bnd = o;   o = up256(o + 64);
    const size_t o_cnt = o;   o = up256(o + (size_t)256 * VX_THREADS * 4);
    const size_t o_k0 = o;    o = up256(o + n * 8);
    const size_t o_k1 = o;    o = up256(o + n * 8);
    const size_t o_i0 = o;    o = up256(o + n * 4);
    const size_t o_i1 = o;    o = up256(o + n * 4);
    const size_t o_pts = o;   o = up256(o + (host ? n * (size_t)pts_stride_bytes : 0));
    const size_t o_out = o;   o = up256(o + (host ? (size_t)cap * 24 : 0));
    if ((rc = dsc_reserve_scratch(ctx, o))) return rc;
    uint8_t *sc = (uint8_t *)ctx->scratch;
    VoxelScratch vs;
    vs.result = (long long *)(sc + o_res);
    vs.bounds = (unsigned long long *)(sc + o_bnd);
    vs.cnt = (uint32_t *)(sc + o_cnt);
    vs.key[0] = (unsigned long long *)(sc + o_k0);
    vs.key[1] = (unsigned long long *)(sc + o_k1);
    vs.idx[0] = (uint32_t *)(sc + o_i0);
    vs.idx[1] = (uint32_t *)(sc + o_i1);
    DSC_CUDA(ctx, cudaMemsetAsync(sc + o_res, 0, o_cnt - o_res, st));            // result + bounds (key 0 == "nothing yet")
    const uint8_t *d_pts = (const uint8_t *)pts;
    double *d_out = out_pts;
    if (host) {
        if (n) DSC_CUDA(ctx, cudaMemcpyAsync(sc + o_pts, pts, n * (size_t)pts_stride_bytes, cudaMemcpyHostToDevice, st));
        d_pts = sc + o_pts;
        d_out = (double *)(sc + o_out);
    }
    if (n) {
        k7_voxel<<<1, VX_THREADS, 0, st>>>(d_pts, n_points, pts_stride_bytes, pts_f64, voxel_size, vs, d_out, cap);
        if ((rc = dsc_check_launch(ctx, "k7_voxel"))) return rc;
    }
    if (!host) {
        // device mode: n_out / n_dropped are DEVICE int64 pointers, filled stream-ordered; errors surface as n_out = -1
        DSC_CUDA(ctx, cudaMemcpyAsync(n_out, vs.result, 8, cudaMemcpyDeviceToDevice, st));
        if (n_dropped) DSC_CUDA(ctx, cudaMemcpyAsync(n_dropped, vs.result + 1, 8, cudaMemcpyDeviceToDevice, st));
        return dsc_leave_stream(ctx, st);
    }
    long long res[3] = {0, 0, 0};
    DSC_CUDA(ctx, cudaMemcpyAsync(res, vs.result, sizeof res, cudaMemcpyDeviceToHost, st));
    DSC_CUDA(ctx, cudaStreamSynchronize(st));
    if (res[2]) DSC_FAIL(ctx, DSC_ERR_INVALID, "voxel_size is too small for the extent of the cloud (more than 2^21 voxels per axis)");
    *n_out = res[0];
    if (n_dropped) *n_dropped = res[1];
    const int64_t ncopy = res[0] < cap ? res[0] : cap;
    if (ncopy > 0) DSC_CUDA(ctx, cudaMemcpy(out_pts, d_out, (size_t)ncopy * 24, cudaMemcpyDeviceToHost));
    if (res[0] > cap) {
        char buf[128];
        snprintf(buf, sizeof buf, "out_pts capacity %lld < required %lld", (long long)cap, (long long)res[0]);
        DSC_FAIL(ctx, DSC_ERR_CAPACITY, buf);
    }
    return DSC_OK;
}
