// dsc_math.cuh -- per-point polar binning with the reference's arithmetic order and dtype.
//
// Everything that decides a bin is built from individually rounded IEEE operations
// (__fmul_rn / __fadd_rn / __fdiv_rn / __fsqrt_rn: no FMA contraction, no approximations), so the
// result is bit-identical to what NumPy + glibc compute on the host:
//   variant A  scaner_version2.py:85-94   (x/y zero substitution, arctan2, rad2deg, % 360, sqrt, //)
//   variant B  scan_context_node.py:115-145
// atan2f is the fdlibm single-precision algorithm (glibc 2.39 e_atan2f.c / s_atanf.c), the flavour
// that produced the reference's golden test_data/sc_*.npy.
//
// The header also compiles as plain C++ (tests/host_math.cpp builds it with g++
// -ffp-contract=off) so the exact and screened paths can be checked against the oracle on a
// machine without a GPU.  That host build is test infrastructure; libdsc.so only uses the device
// versions.
#pragma once
#include <stdint.h>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define DSC_HD __host__ __device__ __forceinline__
#else
#define DSC_HD inline
#endif

namespace dsc {

// ---- individually rounded primitives ---------------------------------------------------------
#if defined(__CUDA_ARCH__)
DSC_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
DSC_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
DSC_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
DSC_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
DSC_HD float fsqrt(float a) { return __fsqrt_rn(a); }
DSC_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
DSC_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
DSC_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
DSC_HD double ddiv(double a, double b) { return __ddiv_rn(a, b); }
DSC_HD uint32_t f2u(float f) { return __float_as_uint(f); }
DSC_HD float u2f(uint32_t u) { return __uint_as_float(u); }
#else
// host build: relies on -ffp-contract=off (see tests/host_math.cpp)
DSC_HD float fmul(float a, float b) { return a * b; }
DSC_HD float fadd(float a, float b) { return a + b; }
DSC_HD float fsub(float a, float b) { return a - b; }
DSC_HD float fdiv(float a, float b) { return a / b; }
DSC_HD float fsqrt(float a) { return sqrtf(a); }
DSC_HD double dmul(double a, double b) { return a * b; }
DSC_HD double dadd(double a, double b) { return a + b; }
DSC_HD double dsub(double a, double b) { return a - b; }
DSC_HD double ddiv(double a, double b) { return a / b; }
DSC_HD uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
DSC_HD float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }
#endif

DSC_HD int min_i(int a, int b) { return a < b ? a : b; }

// ---- order-preserving float <-> uint key ------------------------------------------------------
// key(a) < key(b)  <=>  a < b for all non-NaN floats; key(-inf) = 0x007fffff, so a zero-filled
// grid is "below everything" and   key <= 0x007fffff  decodes to the reference's empty value 0
// (scaner_version2.py:111-113 maps -inf -> 0; scan_context_node.py:154-155).
DSC_HD uint32_t float_to_key(float z)
{
    const uint32_t u = f2u(z);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
DSC_HD float key_to_float(uint32_t k)
{
    // branch-free (the write-out loops run it 115 200 times per scan): keys with the top bit set came from
    // non-negative floats (flip the top bit back), the others from negative ones (flip every bit back)
    const uint32_t m = (uint32_t)((int32_t)k >> 31);
    const uint32_t f = k ^ (0x80000000u | ~m);
    return k > 0x007fffffu ? u2f(f) : 0.0f;
}

// ---- fdlibm atanf for u >= 0 (finite) ---------------------------------------------------------
DSC_HD float atanf_pos(float u)
{
    const uint32_t ix = f2u(u);
    if (ix >= 0x4c000000u) return fadd(u2f(0x3fc90fdau), u2f(0x33a22168u));  // |u| >= 2^25
    if (ix < 0x31000000u) return u;                                         // |u| < 2^-29
    // argument reduction: v = num / den, result = hi - ((v*(s1+s2) - lo) - v)
    float num, den, hi, lo;
    if (ix < 0x3ee00000u) {            // u < 7/16: no reduction
        num = u; den = 1.0f; hi = 0.0f; lo = 0.0f;
    } else if (ix < 0x3f300000u) {     // 7/16 <= u < 11/16
        num = fsub(fmul(2.0f, u), 1.0f); den = fadd(2.0f, u);
        hi = u2f(0x3eed6338u); lo = u2f(0x31ac3769u);
    } else if (ix < 0x3f980000u) {     // 11/16 <= u < 19/16
        num = fsub(u, 1.0f); den = fadd(u, 1.0f);
        hi = u2f(0x3f490fdau); lo = u2f(0x33222168u);
    } else if (ix < 0x401c0000u) {     // 19/16 <= u < 39/16
        num = fsub(u, 1.5f); den = fadd(1.0f, fmul(1.5f, u));
        hi = u2f(0x3f7b985eu); lo = u2f(0x33140fb4u);
    } else {
        num = -1.0f; den = u;
        hi = u2f(0x3fc90fdau); lo = u2f(0x33a22168u);
    }
    const float v = fdiv(num, den);  // den == 1 reproduces v = u exactly
    const float z = fmul(v, v);
    const float w = fmul(z, z);
    float s1 = fadd(u2f(0x3d4bda59u), fmul(w, u2f(0x3c8569d7u)));  // aT8 + w*aT10
    s1 = fadd(u2f(0x3d886b35u), fmul(w, s1));                      // aT6
    s1 = fadd(u2f(0x3dba2e6eu), fmul(w, s1));                      // aT4
    s1 = fadd(u2f(0x3e124925u), fmul(w, s1));                      // aT2
    s1 = fadd(u2f(0x3eaaaaabu), fmul(w, s1));                      // aT0
    s1 = fmul(z, s1);
    float s2 = fadd(u2f(0xbd6ef16bu), fmul(w, u2f(0xbd15a221u)));  // aT7 + w*aT9
    s2 = fadd(u2f(0xbd9d8795u), fmul(w, s2));                      // aT5
    s2 = fadd(u2f(0xbde38e38u), fmul(w, s2));                      // aT3
    s2 = fadd(u2f(0xbe4ccccdu), fmul(w, s2));                      // aT1
    s2 = fmul(w, s2);
    const float t = fmul(v, fadd(s1, s2));
    // id < 0 in fdlibm returns v - t; with hi = lo = 0 the expression below is -(t - v) = v - t
    return fsub(hi, fsub(fsub(t, lo), v));
}

// fdlibm __ieee754_atan2f for FINITE inputs (zeros allowed).
DSC_HD float atan2f_fdlibm(float y, float x)
{
    const float pi = u2f(0x40490fdbu), pi_lo = u2f(0xb3bbbd2eu), pi_o_2 = u2f(0x3fc90fdbu);
    const uint32_t hx = f2u(x), hy = f2u(y);
    const uint32_t ix = hx & 0x7fffffffu, iy = hy & 0x7fffffffu;
    const bool xneg = (hx >> 31) != 0, yneg = (hy >> 31) != 0;
    if (iy == 0) return xneg ? (yneg ? -pi : pi) : y;   // pi + tiny rounds to pi
    if (ix == 0) return yneg ? -pi_o_2 : pi_o_2;
    const int k = ((int)iy - (int)ix) >> 23;
    float z;
    if (k > 60) z = fadd(pi_o_2, fmul(0.5f, pi_lo));
    else if (xneg && k < -60) z = 0.0f;
    else z = atanf_pos(fabsf(fdiv(y, x)));
    if (!xneg) return yneg ? -z : z;
    const float zl = fsub(z, pi_lo);
    return yneg ? fsub(zl, pi) : fsub(pi, zl);
}

// ---- NumPy's floor_divide (npy_divmod) restated, for non-dyadic gaps --------------------------
DSC_HD double np_floor_divide(double a, double b)
{
    const double mod = fmod(a, b);
    double div = ddiv(dsub(a, mod), b);
    if (mod != 0.0 && ((b < 0) != (mod < 0))) div = dsub(div, 1.0);
    if (div != 0.0) {
        double fd = floor(div);
        if (dsub(div, fd) > 0.5) fd = dadd(fd, 1.0);
        return fd;
    }
    return copysign(0.0, ddiv(a, b));
}
DSC_HD float np_floor_dividef(float a, float b)
{
    const float mod = fmodf(a, b);
    float div = fdiv(fsub(a, mod), b);
    if (mod != 0.0f && ((b < 0) != (mod < 0))) div = fsub(div, 1.0f);
    if (div != 0.0f) {
        float fd = floorf(div);
        if (fsub(div, fd) > 0.5f) fd = fadd(fd, 1.0f);
        return fd;
    }
    return copysignf(0.0f, fdiv(a, b));
}

// ---- binning configuration (filled on the host by make_bin_cfg) -------------------------------
enum { PT_OK = 0, PT_CLAMPED = 1, PT_NONFINITE = 2, PT_SECTOR_OVERFLOW = 3, PT_OUT_OF_RANGE = 4 };

struct BinCfg {
    int32_t R, S;
    int32_t variant;      // 0 = A, 1 = B
    int32_t full_height;
    float   offset;       // custom_height_offset as float32
    int32_t pow2;         // both gaps are powers of two: the `//` reduces to an exact scale + floor
    int32_t gap_f32;      // variant A, non-pow2: float32 floor_divide for the ring gap (bit 0) / the sector gap (bit 1)
    float   inv_gap_ring_f, inv_gap_sector_f;   // exact reciprocals when pow2
    float   gap_ring_f, gap_sector_f;           // float32(gap): variant B divides by these
    double  gap_ring_d, gap_sector_d;
    float   max_length_f;                       // variant B range filter (float32 compare)
    // screened fast path (variant A/B, pow2 gaps only): see bin_screened()
    int32_t fast_ok;
    float   fast_dr, fast_ds;        // half-widths of the "too close to a boundary" bands (bin units)
    float   q90, q180, q360;         // 90/180/360 degrees in sector units (exact: pow2 scaling)
    float   R_f;                     // (float)R
    float   fast_hr, fast_hs;        // 0.5 - fast_dr, 0.5 - fast_ds
    float   atan_c[8];               // atan(u)/u minimax in u^2 on [0,1], pre-scaled to sector units
};

static inline bool is_pow2_double(double g)
{
    if (!(g > 0.0) || !isfinite(g)) return false;
    int e;
    return frexp(g, &e) == 0.5;
}

static inline BinCfg make_bin_cfg(int R, int S, double max_length, int variant, int full_height,
                                  float offset, int gap_f32, int exact_only)
{
    BinCfg c;
    memset(&c, 0, sizeof(c));
    c.R = R; c.S = S; c.variant = variant; c.full_height = full_height; c.offset = offset;
    c.gap_ring_d = max_length / (double)R;
    c.gap_sector_d = 360.0 / (double)S;
    c.gap_ring_f = (float)c.gap_ring_d;
    c.gap_sector_f = (float)c.gap_sector_d;
    c.gap_f32 = gap_f32;
    c.pow2 = is_pow2_double(c.gap_ring_d) && is_pow2_double(c.gap_sector_d) &&
             c.gap_ring_d >= 1.0 / 1048576.0 && c.gap_sector_d >= 1.0 / 1048576.0;
    if (c.pow2) {
        c.inv_gap_ring_f = (float)(1.0 / c.gap_ring_d);
        c.inv_gap_sector_f = (float)(1.0 / c.gap_sector_d);
    }
    c.max_length_f = (float)max_length;
    // Fast path needs pow2 gaps and bin counts small enough that the screening bands (sized from
    // the error budget in bin_screened's comment) stay far below half a bin.
    c.fast_ok = c.pow2 && !exact_only && R <= 4096 && c.gap_sector_d >= 1.0 / 64.0;
    if (c.fast_ok) {
        static const double kAtan[8] = {0.9999993443489075, -0.33329859375953674, 0.19946566224098206,
                                        -0.1390862911939621, 0.09642194956541061, -0.05591229721903801,
                                        0.0218629352748394, -0.004054560791701078};
        const double inv_s = 1.0 / c.gap_sector_d;
        const double k = (180.0 / 3.14159265358979323846) * inv_s;
        for (int i = 0; i < 8; ++i) c.atan_c[i] = (float)(kAtan[i] * k);
        c.q90 = (float)(90.0 * inv_s);
        c.q180 = (float)(180.0 * inv_s);
        c.q360 = (float)(360.0 * inv_s);
        c.R_f = (float)R;
        c.fast_dr = (float)((R + 1) * 1.5e-6);
        c.fast_ds = (float)(5.0e-4 * inv_s);
        c.fast_hr = 0.5f - c.fast_dr;
        c.fast_hs = 0.5f - c.fast_ds;
    }
    return c;
}

struct BinResult {
    int32_t cell;    // ring * S + sector, valid when status <= PT_CLAMPED
    int32_t ring, sector;
    int32_t status;
    float   z;       // height after the mode's offset
};

#define DSC_RAD2DEG_F32 0x42652ee0u  // 180.0f / 3.14159274f: npy_rad2degf's single multiply

// Exact path, variant A  (scaner_version2.py:85-94).
DSC_HD BinResult bin_exact_a(const BinCfg &c, float x, float y, float z)
{
    BinResult o;
    o.cell = 0; o.ring = 0; o.sector = 0;
    o.z = c.full_height ? z : fadd(z, c.offset);   // :73-82, float32 add
    if (x == 0.0f) x = 0.001f;                     // :85-86
    if (y == 0.0f) y = 0.001f;
    const float r = fsqrt(fadd(fmul(x, x), fmul(y, y)));   // :90
    // NaN/Inf in x or y, or x*x+y*y overflowing: the reference indexes with INT_MIN -> IndexError
    if (!(r < INFINITY) || !(fabsf(x) < INFINITY) || !(fabsf(y) < INFINITY)) {
        o.status = PT_NONFINITE;
        return o;
    }
    const float a = atan2f_fdlibm(y, x);                    // :89
    const float d = fmul(a, u2f(DSC_RAD2DEG_F32));
    const float t = (d < 0.0f) ? fadd(d, 360.0f) : d;       // |d| <= 180: fmodf is the identity
    int ring, sector;
    bool clamped;
    if (c.pow2) {
        const float rf = fmul(r, c.inv_gap_ring_f);         // exact scaling (may overflow to inf)
        clamped = rf >= (float)c.R;
        ring = rf >= (float)(c.R - 1) ? c.R - 1 : (int)rf;
        sector = (int)fmul(t, c.inv_gap_sector_f);
    } else {
        // :93-94.  Each `//` runs in the dtype NumPy promotes (float32 array, gap) to: float64 when the gap is an
        // np.float64 (resolution given as np.int64, the class statics), float32 when it is a Python float (weak scalar).
        const double rf = (c.gap_f32 & 1) ? (double)np_floor_dividef(r, c.gap_ring_f) : np_floor_divide((double)r, c.gap_ring_d);
        const double sf = (c.gap_f32 & 2) ? (double)np_floor_dividef(t, c.gap_sector_f) : np_floor_divide((double)t, c.gap_sector_d);
        clamped = rf > (double)(c.R - 1);
        ring = rf < (double)(c.R - 1) ? (int)rf : c.R - 1;
        sector = sf < 2.0e9 ? (int)sf : c.S;
    }
    if (sector >= c.S || sector < 0) {                       // reference: IndexError
        o.status = PT_SECTOR_OVERFLOW;
        return o;
    }
    o.ring = ring;
    o.sector = sector;
    o.cell = ring * c.S + sector;
    o.status = clamped ? PT_CLAMPED : PT_OK;
    return o;
}

// Exact path, variant B  (scan_context_node.py:115-145), float32 input.
DSC_HD BinResult bin_exact_b(const BinCfg &c, float x, float y, float z)
{
    BinResult o;
    o.cell = 0; o.ring = 0; o.sector = 0;
    o.z = z;
    if (!(fabsf(x) < INFINITY) || !(fabsf(y) < INFINITY) || !(fabsf(z) < INFINITY)) {  // :115-117
        o.status = PT_NONFINITE;
        return o;
    }
    o.z = c.full_height ? z : fadd(z, c.offset);            // :120-127
    const float r = fsqrt(fadd(fmul(x, x), fmul(y, y)));    // :130
    const float a = atan2f_fdlibm(y, x);                    // :131
    const float d = fmul(a, u2f(DSC_RAD2DEG_F32));
    const float s = fadd(d, 360.0f);
    const float t = (s >= 360.0f) ? fsub(s, 360.0f) : s;    // fmodf(s, 360) for s in [180, 540]
    if (!(r <= c.max_length_f)) {                           // :134-137
        o.status = PT_OUT_OF_RANGE;
        return o;
    }
    const float rf = floorf(fdiv(r, c.gap_ring_f));         // :142
    const float sf = floorf(fdiv(t, c.gap_sector_f));       // :144
    const int ring = rf > (float)(c.R - 1) ? c.R - 1 : (rf < 0.0f ? 0 : (int)rf);     // :143
    const int sector = sf > (float)(c.S - 1) ? c.S - 1 : (sf < 0.0f ? 0 : (int)sf);   // :145
    o.ring = ring;
    o.sector = sector;
    o.cell = ring * c.S + sector;
    o.status = PT_OK;
    return o;
}

DSC_HD BinResult bin_exact(const BinCfg &c, float x, float y, float z)
{
    return c.variant == 0 ? bin_exact_a(c, x, y, z) : bin_exact_b(c, x, y, z);
}


// ---- screened fast path --------------------------------------------------------------------------
// The exact chain above costs ~100 instructions per point (two IEEE divisions, an IEEE sqrt and 26
// unfused polynomial steps).  Almost every point sits far from any bin boundary, where a much
// cheaper approximation provably lands in the same bin.  screened_core_t() computes
//     qr ~ r / gap_ring        (fma + MUFU.SQRT)                          in ring units
//     qs ~ theta / gap_sector  (MUFU.RCP + 8-term minimax atan, octant fix-up)  in sector units
// and accepts the floor() of each only when the value is at least fast_dr / fast_ds away from every
// integer.  Anything else -- including every NaN/Inf/overflow/underflow, which turn into NaN or land
// on a boundary -- reports "not ok" and is re-done by bin_exact().
//
// floor() and the distance to the bin centre come from one magic-number addition rounded toward minus infinity:
// t = q + 2^23 (add.rm) is 2^23 + floor(q), its low mantissa bits ARE the bin index, and q - (t - (2^23 - 1/2)) is the
// offset from the bin centre, which must stay below 1/2 - band.  Three FADDs and an integer subtract on the FMA/ALU
// pipes; the cvt.rni / cvt.rzi pair they replace (FRND, F2I) runs on the quarter-rate XU pipe next to MUFU.SQRT and
// MUFU.RCP, which was the busiest pipe of k1_scan (42 %, profiles/r2e_k1_scan.*).
//
// Error budget (degrees; true = real-number polar coordinates of the float inputs):
//   reference chain vs true : atan2f < 1 ulp(pi) = 1.4e-5, y/x rounding 1.7e-6, rad2deg multiply
//                             7.6e-6 + constant 6e-6, (+360) rounding 1.5e-5      => < 5e-5
//   fast chain vs true      : polynomial 7.4e-6 (measured in f32), rcp.approx+mul 6.9e-6, three
//                             octant subtractions <= 3e-5                          => < 4.5e-5
//   sum 9.5e-5 deg; band fast_ds = 5e-4 deg (5x).  (The centre offset q - (floor + 1/2) is exact.)
//   ring: sqrt.approx 2^-23 + one multiply + the fma-vs-unfused difference < 3e-7 relative,
//   i.e. < (R+1)*3e-7 ring units; band fast_dr = (R+1)*1.5e-6 (5x).
#if defined(__CUDA_ARCH__)
DSC_HD float approx_sqrt(float s, float) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s)); return r; }
DSC_HD float approx_rcp(float s, float) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(s)); return r; }
DSC_HD float ffma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
DSC_HD float fadd_floor_magic(float a) { return __fadd_rd(a, 8388608.0f); }
#else
// host emulation for tests/host_math.cpp: correctly rounded value perturbed by `noise` * 2^-22
// (|noise| <= 1 covers the documented MUFU error), with flush-to-zero of subnormal operands/results
DSC_HD float approx_sqrt(float s, float noise)
{
    if (s == s && fabsf(s) < 1.17549435e-38f) s = copysignf(0.0f, s);
    float r = (float)(sqrt((double)s) * (1.0 + (double)noise * 2.384185791015625e-07));
    if (fabsf(r) < 1.17549435e-38f) r = copysignf(0.0f, r);
    return r;
}
DSC_HD float approx_rcp(float s, float noise)
{
    if (s == s && fabsf(s) < 1.17549435e-38f) s = copysignf(0.0f, s);
    float r = (float)((1.0 / (double)s) * (1.0 + (double)noise * 2.384185791015625e-07));
    if (fabsf(r) < 1.17549435e-38f) r = copysignf(0.0f, r);
    return r;
}
DSC_HD float ffma(float a, float b, float c) { return (float)((double)a * (double)b + (double)c); }
DSC_HD float fadd_floor_magic(float a) { return (float)(floor((double)a) + 8388608.0); }   // add.rm for 0 <= a < 2^23 (NaN stays NaN)
#endif

#define DSC_FLOOR_MAGIC_U 0x4b000000u     // bits of 2^23
#define DSC_FLOOR_MAGIC_H 8388607.5f      // 2^23 - 1/2

// Core of the screen: true when (ring, sector, beyond) are certain; false sends the point to bin_exact.
// No zero substitution is needed here: a zero coordinate makes u == 0 (or NaN), so the approximate
// angle is an exact multiple of 90 degrees -- a sector boundary -- and the point screens itself out.
// DEF = true bakes the constants of the reference's default grid (160 rings x 720 sectors, i.e. 0.5-degree
// sectors) into the instruction stream as immediates: no constant-bank loads in the streaming loop (an LDC
// there shares the long scoreboard with the in-flight 128-bit point loads and stalls on them).  Only the
// ring scale (8/m for max_length 20, 2/m for 80) stays a run-time value.  Same arithmetic, same values.
namespace defgrid {
constexpr double kScale = (180.0 / 3.14159265358979323846) * 2.0;   // radians -> half-degree sectors
constexpr float c0 = (float)(0.9999993443489075 * kScale), c1 = (float)(-0.33329859375953674 * kScale),
                c2 = (float)(0.19946566224098206 * kScale), c3 = (float)(-0.1390862911939621 * kScale),
                c4 = (float)(0.09642194956541061 * kScale), c5 = (float)(-0.05591229721903801 * kScale),
                c6 = (float)(0.0218629352748394 * kScale), c7 = (float)(-0.004054560791701078 * kScale);
constexpr int R = 160, S = 720;
constexpr float q90 = 180.0f, q180 = 360.0f, q360 = 720.0f, R_f = 160.0f;
constexpr float fast_dr = (float)(161 * 1.5e-6), fast_ds = (float)(5.0e-4 * 2.0);
constexpr float fast_hr = 0.5f - fast_dr, fast_hs = 0.5f - fast_ds;
}  // namespace defgrid

template <bool DEF>
DSC_HD bool screened_core_t(const BinCfg &c, float x, float y, float noise_r, float noise_t, int *ring,
                            int *sector, bool *beyond)
{
    const float ax = fabsf(x), ay = fabsf(y);
    // ring coordinate: sqrt.approx.ftz flushes a subnormal s to 0 (-> boundary 0 -> exact path); an overflowing s is
    // rejected by the explicit test below.  From the middle of ring R on ("beyond": clamped by variant A, dropped by
    // variant B) the value is pinned to R + 1/2, a bin centre, which takes the far zone out of the boundary test.
    const float s = ffma(x, x, y * y);
    const float qr = fminf(approx_sqrt(s, noise_r) * c.inv_gap_ring_f, (DEF ? defgrid::R_f : c.R_f) + 0.5f);
    // sector coordinate
    const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
    const float u = mn * approx_rcp(mx, noise_t);
    const float w = u * u;
    float p = ffma(w, DEF ? defgrid::c7 : c.atan_c[7], DEF ? defgrid::c6 : c.atan_c[6]);
    p = ffma(w, p, DEF ? defgrid::c5 : c.atan_c[5]);
    p = ffma(w, p, DEF ? defgrid::c4 : c.atan_c[4]);
    p = ffma(w, p, DEF ? defgrid::c3 : c.atan_c[3]);
    p = ffma(w, p, DEF ? defgrid::c2 : c.atan_c[2]);
    p = ffma(w, p, DEF ? defgrid::c1 : c.atan_c[1]);
    p = ffma(w, p, DEF ? defgrid::c0 : c.atan_c[0]);
    float qs = p * u;
    if (ay > ax) qs = (DEF ? defgrid::q90 : c.q90) - qs;
    if (x < 0.0f) qs = (DEF ? defgrid::q180 : c.q180) - qs;
    if (y < 0.0f) qs = (DEF ? defgrid::q360 : c.q360) - qs;
    // bin index and offset from the bin centre (NaN compares false everywhere => not ok); qr, qs >= 0 or NaN
    const float tr = fadd_floor_magic(qr), ts = fadd_floor_magic(qs);
    const float er = fabsf(qr - (tr - DSC_FLOOR_MAGIC_H)), es = fabsf(qs - (ts - DSC_FLOOR_MAGIC_H));
    const bool ok = (es < (DEF ? defgrid::fast_hs : c.fast_hs)) && (er < (DEF ? defgrid::fast_hr : c.fast_hr)) &&
                    (s < INFINITY);
    const int Rr = DEF ? defgrid::R : c.R;
    const int ir = (int)(f2u(tr) - DSC_FLOOR_MAGIC_U);   // 0 .. R wherever ok
    *beyond = ir >= Rr;
    *ring = min_i(ir, Rr - 1);
    *sector = (int)(f2u(ts) - DSC_FLOOR_MAGIC_U);        // 0 .. S-1 wherever ok
    return ok;
}

static inline bool is_default_grid(const BinCfg &c)
{
    return c.fast_ok && c.R == defgrid::R && c.S == defgrid::S && c.atan_c[0] == defgrid::c0 &&
           c.atan_c[7] == defgrid::c7 && c.q90 == defgrid::q90 && c.q360 == defgrid::q360 &&
           c.R_f == defgrid::R_f && c.fast_dr == defgrid::fast_dr && c.fast_ds == defgrid::fast_ds &&
           c.fast_hr == defgrid::fast_hr && c.fast_hs == defgrid::fast_hs;
}

DSC_HD bool screened_core(const BinCfg &c, float x, float y, float noise_r, float noise_t, int *ring,
                          int *sector, bool *beyond)
{
    return screened_core_t<false>(c, x, y, noise_r, noise_t, ring, sector, beyond);
}

DSC_HD BinResult bin_screened(const BinCfg &c, float x, float y, float z, float noise_r, float noise_t,
                              bool *slow)
{
    BinResult o;
    o.cell = 0; o.ring = 0; o.sector = 0;
    o.status = PT_OK;
    o.z = c.full_height ? z : fadd(z, c.offset);
    int ring, sector;
    bool beyond;
    bool ok = screened_core(c, x, y, noise_r, noise_t, &ring, &sector, &beyond);
    if (c.variant != 0) ok = ok && (fabsf(z) < INFINITY);              // finite mask (scan_context_node.py:115)
    *slow = !ok;
    if (!ok) return o;
    if (c.variant == 0) {
        o.status = beyond ? PT_CLAMPED : PT_OK;
    } else if (beyond) {
        o.status = PT_OUT_OF_RANGE;
        return o;
    }
    o.ring = ring;
    o.sector = sector;
    o.cell = ring * c.S + sector;
    return o;
}

}  // namespace dsc
