// dsc_math_f64.cuh -- per-point polar binning for FLOAT64 clouds (what Open3D's voxel_down_sample hands the
// reference: scan_context_node.py:93-100 -> :161-168, scaner_version2.py:126-133).  NumPy promotes the whole chain to
// float64; every operation below is an individually rounded IEEE double operation (__dmul_rn / __dadd_rn / ...), so
// the result does not depend on FMA contraction.
//
// float64 arctan2 is the published fdlibm algorithm (e_atan2.c / s_atan.c, < 1 ULP).  Upstream pins no float64
// flavour (NumPy's SVML loop on AVX-512 hosts and glibc's IBM atan2 elsewhere differ in the last bit too); a flavour
// can only move a point to another sector when theta / gap_sector is within a few ULP of an integer, and those
// points are counted (bin64_t::boundary -> dsc_scan_stats.n_boundary).
//
// Compiles as plain C++ too (tests/host_math.cpp, g++ -ffp-contract=off): test infrastructure only.
#pragma once
#include "dsc_math.cuh"

namespace dsc {

#if defined(__CUDA_ARCH__)
DSC_HD double dsqrt(double a) { return __dsqrt_rn(a); }
DSC_HD uint64_t d2u(double d) { return (uint64_t)__double_as_longlong(d); }
DSC_HD double u2d(uint64_t u) { return __longlong_as_double((long long)u); }
#else
DSC_HD double dsqrt(double a) { return sqrt(a); }
DSC_HD uint64_t d2u(double d) { uint64_t u; memcpy(&u, &d, 8); return u; }
DSC_HD double u2d(uint64_t u) { double d; memcpy(&d, &u, 8); return d; }
#endif

// order-preserving double <-> uint64 key (same construction as float_to_key); key(-inf) = 0x000fffffffffffff
DSC_HD uint64_t double_to_key(double z)
{
    const uint64_t u = d2u(z);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
DSC_HD double key_to_double(uint64_t k)
{
    if (k <= 0x000fffffffffffffull) return 0.0;          // empty or -inf: the reference's 0 (scaner_version2.py:111-113)
    return u2d((k >> 63) ? (k & 0x7fffffffffffffffull) : ~k);
}

// fdlibm atan for u >= 0 (finite)
DSC_HD double atan_pos(double u)
{
    const uint32_t ix = (uint32_t)(d2u(u) >> 32);
    if (ix >= 0x44100000u) return dadd(u2d(0x3FF921FB54442D18ull), u2d(0x3C91A62633145C07ull));   // u >= 2^66
    if (ix < 0x3e200000u) return u;                                                                 // u < 2^-29
    double num, den, hi, lo;
    if (ix < 0x3fdc0000u) {            // u < 7/16: no reduction
        num = u; den = 1.0; hi = 0.0; lo = 0.0;
    } else if (ix < 0x3fe60000u) {     // 7/16 <= u < 11/16
        num = dsub(dmul(2.0, u), 1.0); den = dadd(2.0, u);
        hi = u2d(0x3FDDAC670561BB4Full); lo = u2d(0x3C7A2B7F222F65E2ull);
    } else if (ix < 0x3ff30000u) {     // 11/16 <= u < 19/16
        num = dsub(u, 1.0); den = dadd(u, 1.0);
        hi = u2d(0x3FE921FB54442D18ull); lo = u2d(0x3C81A62633145C07ull);
    } else if (ix < 0x40038000u) {     // 19/16 <= u < 39/16
        num = dsub(u, 1.5); den = dadd(1.0, dmul(1.5, u));
        hi = u2d(0x3FEF730BD281F69Bull); lo = u2d(0x3C7007887AF0CBBDull);
    } else {
        num = -1.0; den = u;
        hi = u2d(0x3FF921FB54442D18ull); lo = u2d(0x3C91A62633145C07ull);
    }
    const double v = ddiv(num, den);   // den == 1 reproduces v = u exactly
    const double z = dmul(v, v);
    const double w = dmul(z, z);
    double s1 = dadd(u2d(0x3FA97B4B24760DEBull), dmul(w, u2d(0x3F90AD3AE322DA11ull)));   // aT8 + w*aT10
    s1 = dadd(u2d(0x3FB10D66A0D03D51ull), dmul(w, s1));                                   // aT6
    s1 = dadd(u2d(0x3FB745CDC54C206Eull), dmul(w, s1));                                   // aT4
    s1 = dadd(u2d(0x3FC24924920083FFull), dmul(w, s1));                                   // aT2
    s1 = dadd(u2d(0x3FD555555555550Dull), dmul(w, s1));                                   // aT0
    s1 = dmul(z, s1);
    double s2 = dadd(u2d(0xBFADDE2D52DEFD9Aull), dmul(w, u2d(0xBFA2B4442C6A6C2Full)));   // aT7 + w*aT9
    s2 = dadd(u2d(0xBFB3B0F2AF749A6Dull), dmul(w, s2));                                   // aT5
    s2 = dadd(u2d(0xBFBC71C6FE231671ull), dmul(w, s2));                                   // aT3
    s2 = dadd(u2d(0xBFC999999998EBC4ull), dmul(w, s2));                                   // aT1
    s2 = dmul(w, s2);
    const double t = dmul(v, dadd(s1, s2));
    // id < 0 in fdlibm returns v - t; with hi = lo = 0 the expression below is -(t - v) = v - t
    return dsub(hi, dsub(dsub(t, lo), v));
}

// fdlibm __ieee754_atan2 for FINITE inputs (zeros allowed)
DSC_HD double atan2_fdlibm(double y, double x)
{
    const double pi = u2d(0x400921FB54442D18ull), pi_lo = u2d(0x3CA1A62633145C07ull), pi_o_2 = u2d(0x3FF921FB54442D18ull);
    const uint64_t ux = d2u(x), uy = d2u(y);
    const bool xneg = (ux >> 63) != 0, yneg = (uy >> 63) != 0;
    const uint64_t ax = ux & 0x7fffffffffffffffull, ay = uy & 0x7fffffffffffffffull;
    if (ay == 0) return xneg ? (yneg ? -pi : pi) : y;    // pi + tiny rounds to pi
    if (ax == 0) return yneg ? -pi_o_2 : pi_o_2;
    const int k = ((int)(uint32_t)(ay >> 32) - (int)(uint32_t)(ax >> 32)) >> 20;
    double z;
    if (k > 60) z = dadd(pi_o_2, dmul(0.5, pi_lo));
    else if (xneg && k < -60) z = 0.0;
    else z = atan_pos(fabs(ddiv(y, x)));
    if (!xneg) return yneg ? -z : z;
    const double zl = dsub(z, pi_lo);
    return yneg ? dsub(zl, pi) : dsub(pi, zl);
}

struct Bin64Cfg {
    int32_t R, S, variant, full_height;
    int32_t pow2;                     // both gaps are powers of two: `//` is an exact scale + floor
    double gap_ring, gap_sector, inv_gap_ring, inv_gap_sector;
    double max_length, offset;
    double bnd_eps;                   // S * 2^-50: |q_sector - nearest integer| below this counts as a boundary point
};

static inline Bin64Cfg make_bin64_cfg(int R, int S, double max_length, int variant, int full_height, double offset)
{
    Bin64Cfg c;
    memset(&c, 0, sizeof(c));
    c.R = R; c.S = S; c.variant = variant; c.full_height = full_height;
    c.gap_ring = max_length / (double)R;
    c.gap_sector = 360.0 / (double)S;
    c.pow2 = is_pow2_double(c.gap_ring) && is_pow2_double(c.gap_sector) && c.gap_ring >= 1.0 / 1048576.0 &&
             c.gap_sector >= 1.0 / 1048576.0;
    c.inv_gap_ring = 1.0 / c.gap_ring;
    c.inv_gap_sector = 1.0 / c.gap_sector;
    c.max_length = max_length;
    c.offset = offset;
    c.bnd_eps = (double)S * 0x1p-50;
    return c;
}

struct Bin64Result {
    int32_t cell, status, boundary;
    double z;       // height after the mode's offset (float64; variant B rounds it to float32 afterwards)
};

#define DSC_RAD2DEG_F64 57.295779513082323   // 180.0 / NPY_PI

DSC_HD bool near_int64(double q, double eps) { return fabs(dsub(q, rint(q))) <= eps; }

// variant A on float64 points (scaner_version2.py:73-94 with a float64 ptcloud)
DSC_HD Bin64Result bin64_a(const Bin64Cfg &c, double x, double y, double z)
{
    Bin64Result o;
    o.cell = 0; o.boundary = 0;
    o.z = c.full_height ? z : dadd(z, c.offset);              // :73-82
    if (x == 0.0) x = 0.001;                                  // :85-86
    if (y == 0.0) y = 0.001;
    const double r = dsqrt(dadd(dmul(x, x), dmul(y, y)));     // :90
    if (!(r < INFINITY) || !(fabs(x) < INFINITY) || !(fabs(y) < INFINITY)) {   // reference: INT_MIN index -> IndexError
        o.status = PT_NONFINITE;
        return o;
    }
    const double d = dmul(atan2_fdlibm(y, x), DSC_RAD2DEG_F64);   // :89
    const double t = (d < 0.0) ? dadd(d, 360.0) : (d == 0.0 ? 0.0 : d);   // |d| <= 180: fmod is the identity
    double rf, sf;
    if (c.pow2) {
        rf = floor(dmul(r, c.inv_gap_ring));                  // exact scaling
        sf = floor(dmul(t, c.inv_gap_sector));
    } else {
        rf = np_floor_divide(r, c.gap_ring);                  // :93-94
        sf = np_floor_divide(t, c.gap_sector);
    }
    o.boundary = near_int64(ddiv(t, c.gap_sector), c.bnd_eps);
    const int ring = rf < (double)(c.R - 1) ? (int)rf : c.R - 1;
    const int sector = sf < 2.0e9 ? (int)sf : c.S;
    if (sector >= c.S || sector < 0) {                        // theta rounded to 360.0: reference IndexError
        o.status = PT_SECTOR_OVERFLOW;
        return o;
    }
    o.cell = ring * c.S + sector;
    o.status = rf > (double)(c.R - 1) ? PT_CLAMPED : PT_OK;
    return o;
}

// variant B on float64 points (scan_context_node.py:115-145 with float64 pts_xyz)
DSC_HD Bin64Result bin64_b(const Bin64Cfg &c, double x, double y, double z)
{
    Bin64Result o;
    o.cell = 0; o.boundary = 0; o.z = z;
    if (!(fabs(x) < INFINITY) || !(fabs(y) < INFINITY) || !(fabs(z) < INFINITY)) {   // :115-117
        o.status = PT_NONFINITE;
        return o;
    }
    o.z = c.full_height ? z : dadd(z, c.offset);              // :120-127
    const double r = dsqrt(dadd(dmul(x, x), dmul(y, y)));     // :130
    const double s = dadd(dmul(atan2_fdlibm(y, x), DSC_RAD2DEG_F64), 360.0);   // :131
    const double t = (s >= 360.0) ? dsub(s, 360.0) : s;       // fmod(s, 360) for s in [180, 540]; exact
    if (!(r <= c.max_length)) {                               // :134-137 (r = inf from an overflowing x*x lands here too)
        o.status = PT_OUT_OF_RANGE;
        return o;
    }
    const double qs = ddiv(t, c.gap_sector);
    const double rf = floor(ddiv(r, c.gap_ring)), sf = floor(qs);               // :142,144
    o.boundary = near_int64(qs, c.bnd_eps);
    const int ring = rf > (double)(c.R - 1) ? c.R - 1 : (rf < 0.0 ? 0 : (int)rf);     // :143
    const int sector = sf > (double)(c.S - 1) ? c.S - 1 : (sf < 0.0 ? 0 : (int)sf);   // :145
    o.cell = ring * c.S + sector;
    o.status = PT_OK;
    return o;
}

}  // namespace dsc
