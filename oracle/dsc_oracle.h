/* dsc_oracle.h -- CPU oracle for the scan-context hot path (TEST INFRASTRUCTURE ONLY;
 * see dsc_oracle.c for the rules about who may load this). */
#ifndef DSC_ORACLE_H
#define DSC_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_VARIANT_A = 0, ORC_VARIANT_B = 1 };
enum { ORC_PT_OK = 0, ORC_PT_CLAMPED = 1, ORC_PT_NONFINITE = 2, ORC_PT_SECTOR_OVERFLOW = 3,
       ORC_PT_OUT_OF_RANGE = 4 };

typedef struct {
    int32_t num_ring;      /* 160 */
    int32_t num_sector;    /* 720 */
    double  max_length;    /* 20 (offline default) / 80 (live default) */
    int32_t variant;       /* ORC_VARIANT_A: scaner_version2.py:64-115; _B: scan_context_node.py:102-157 */
    int32_t full_height;   /* use_full_height_range */
    float   height_offset; /* custom_height_offset (float32 add) */
    int32_t gap_f32;       /* variant A only: which `//` runs in float32 (Python-int resolution): bit 0 ring gap, bit 1 sector gap */
} orc_params;

typedef struct {
    int64_t n_points;
    int64_t n_nonfinite;       /* variant A: reference raises IndexError; variant B: masked out */
    int64_t n_sector_overflow; /* variant A: theta rounded to 360.0 => IndexError in the reference */
    int64_t n_clamped_ring;    /* variant A: r beyond max_length, clamped into ring R-1 */
    int64_t n_out_of_range;    /* variant B: r > max_length, dropped */
} orc_stats;

float orc_atanf(float x);
float orc_atan2f(float y, float x);
void  orc_atan2f_array(const float *y, const float *x, float *out, int64_t n);

int orc_point_bin(const orc_params *p, float x, float y, float z, float theta_rad_override,
                  int use_override, int *ring, int *sector, float *zval);
int orc_sc_build(const orc_params *p, const float *pts, int64_t n, int stride_floats,
                 const float *theta_override, float *out, orc_stats *st);
int orc_sc_build_batch(const orc_params *p, const float *pts, const int64_t *offsets, int n_scans,
                       int stride_floats, float *out, orc_stats *st, int n_threads);
int orc_max_threads(void);

/* float64 point clouds (dsc_oracle_f64.c): fdlibm double atan2; variant A -> float64 grid, variant B -> float32 grid */
double orc_atan(double x);
double orc_atan2(double y, double x);
void   orc_atan2_array(const double *y, const double *x, double *out, int64_t n);
int orc_point_bin_f64(const orc_params *p, double height_offset, double x, double y, double z,
                      double theta_rad_override, int use_override, int *ring, int *sector, double *zval,
                      int *boundary);
int orc_sc_build_f64(const orc_params *p, double height_offset, const double *pts, int64_t n, int stride_doubles,
                     const double *theta_override, double *out64, float *out32, orc_stats *st, int64_t *n_boundary);

int64_t orc_voxel_downsample(const double *pts, int64_t n, int stride_doubles, double voxel, double *out, int64_t cap,
                             int64_t *n_dropped);

void orc_normalize01(const float *grid, int64_t cells, float *out, float *mn_out, float *mx_out);

int64_t orc_raster_box(int R, int S, const double *r_centers, const double *cos_tab,
                       const double *sin_tab, double x, double y, double l, double w, double cyaw,
                       double syaw, int32_t *cells_out, int64_t cap);
int64_t orc_reconstruct(int R, int S, const double *r_centers, const double *cos_tab,
                        const double *sin_tab, const float *grid, int full_height,
                        double height_offset, double *pts_out, int32_t *bin_ids, int64_t cap);
#ifdef __cplusplus
}
#endif
#endif
