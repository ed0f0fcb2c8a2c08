// host_math.cpp -- compiles distropointclouds_b200/csrc/dsc_math.cuh as plain C++ (g++,
// -ffp-contract=off) so the per-point binning code the CUDA kernels run can be checked against
// the oracle on a machine without a GPU.  TEST INFRASTRUCTURE: not part of libdsc.so.
#include "../distropointclouds_b200/csrc/dsc_math.cuh"

extern "C" {

float hm_atan2f(float y, float x) { return dsc::atan2f_fdlibm(y, x); }

// status/cell/z per point through the exact path
void hm_bin_exact(int R, int S, double max_length, int variant, int full_height, float offset, int gap_f32,
                  const float *xyz, long long n, int stride_floats, int *cell, int *status, float *z)
{
    const dsc::BinCfg c = dsc::make_bin_cfg(R, S, max_length, variant, full_height, offset, gap_f32 == 1 ? 3 : gap_f32, 1);  // 1: both gaps, 4|1: ring only, 2: sector only
    for (long long i = 0; i < n; ++i) {
        const float *p = xyz + i * stride_floats;
        const dsc::BinResult b = dsc::bin_exact(c, p[0], p[1], p[2]);
        cell[i] = b.cell; status[i] = b.status; z[i] = b.z;
    }
}

#ifdef DSC_HAVE_FAST
// screened fast path with an emulated MUFU error model: `noise` holds per-point relative
// perturbations (in units of 2^-22) applied to the approximate reciprocal / rsqrt.
void hm_bin_fast(int R, int S, double max_length, int variant, int full_height, float offset,
                 const float *xyz, long long n, int stride_floats, const float *noise, int *cell, int *status,
                 int *took_slow)
{
    const dsc::BinCfg c = dsc::make_bin_cfg(R, S, max_length, variant, full_height, offset, 0, 0);
    for (long long i = 0; i < n; ++i) {
        const float *p = xyz + i * stride_floats;
        bool slow = false;
        const dsc::BinResult b = dsc::bin_screened(c, p[0], p[1], p[2], noise ? noise[2 * i] : 0.f,
                                                   noise ? noise[2 * i + 1] : 0.f, &slow);
        cell[i] = b.cell; status[i] = b.status; took_slow[i] = slow;
    }
}
#endif

#ifdef DSC_HAVE_FAST
// the compile-time-constant specialisation must agree with the run-time one, bit for bit
long long hm_defgrid_mismatches(double max_length, const float *xyz, long long n, int stride_floats, const float *noise)
{
    const dsc::BinCfg c = dsc::make_bin_cfg(160, 720, max_length, 0, 1, 2.0f, 0, 0);
    if (!dsc::is_default_grid(c)) return -1;
    long long bad = 0;
    for (long long i = 0; i < n; ++i) {
        const float *p = xyz + i * stride_floats;
        int r0, s0, r1, s1; bool b0, b1;
        const float nr = noise ? noise[2 * i] : 0.f, nt = noise ? noise[2 * i + 1] : 0.f;
        const bool k0 = dsc::screened_core_t<false>(c, p[0], p[1], nr, nt, &r0, &s0, &b0);
        const bool k1 = dsc::screened_core_t<true>(c, p[0], p[1], nr, nt, &r1, &s1, &b1);
        if (k0 != k1 || (k0 && (r0 != r1 || s0 != s1 || b0 != b1))) ++bad;
    }
    return bad;
}
#endif

unsigned hm_float_to_key(float z) { return dsc::float_to_key(z); }
float hm_key_to_float(unsigned k) { return dsc::key_to_float(k); }
}

// ---- float64 clouds (dsc_math_f64.cuh) -----------------------------------------------------------------------
#include "../distropointclouds_b200/csrc/dsc_math_f64.cuh"
extern "C" {
double hm_atan2(double y, double x) { return dsc::atan2_fdlibm(y, x); }
void hm_atan2_array(const double *y, const double *x, double *out, long long n)
{
    for (long long i = 0; i < n; ++i) out[i] = dsc::atan2_fdlibm(y[i], x[i]);
}
void hm_bin64(int R, int S, double max_length, int variant, int full_height, double offset, const double *xyz,
              long long n, int stride_doubles, int *cell, int *status, int *boundary, double *z)
{
    const dsc::Bin64Cfg c = dsc::make_bin64_cfg(R, S, max_length, variant, full_height, offset);
    for (long long i = 0; i < n; ++i) {
        const double *p = xyz + i * stride_doubles;
        const dsc::Bin64Result b = variant == 0 ? dsc::bin64_a(c, p[0], p[1], p[2]) : dsc::bin64_b(c, p[0], p[1], p[2]);
        cell[i] = b.cell; status[i] = b.status; boundary[i] = b.boundary; z[i] = b.z;
    }
}
unsigned long long hm_double_to_key(double z) { return dsc::double_to_key(z); }
double hm_key_to_double(unsigned long long k) { return dsc::key_to_double(k); }
}
