/*
 * dsc.h -- C ABI of the B200-native scan-context engine (libdsc.so).
 *
 * This is the drop-in boundary for the one hot path of armando-genis/DistroPointclouds:
 * packed KITTI-style float32 [x,y,z,i] scans -> Scan Context polar max-height grids, plus the
 * pedestrian-box rasteriser and the inverse reconstructor that work on the same grid.
 *
 * The reference has no FFI layer: its boundary is four Python methods.  Each entry point below
 * names the reference method (file:line, relative to the upstream repo root) it replaces; the
 * ctypes stubs a maintainer adds on the reference side are shown in INTEGRATION.md and shipped
 * in distropointclouds_b200/_ffi.py.
 *
 * Conventions
 *   - plain C types only (no torch / numpy types); all functions return 0 on success and a
 *     negative dsc_status on failure; dsc_last_error(ctx) gives the message.
 *   - `mem` says where the *data* pointers of that call live: DSC_MEM_HOST or DSC_MEM_DEVICE.
 *     Small control arrays (offsets, params, tables) are always host pointers.
 *   - `stream` is a cudaStream_t (as void*).  With DSC_MEM_DEVICE the call is stream-ordered and
 *     asynchronous (consecutive calls on one context are ordered after one another even when they use
 *     different streams: the context's scratch is shared); with DSC_MEM_HOST the call returns after the
 *     results are in the host buffers.
 *   - Bad points are data, not errors: they are dropped and counted in dsc_scan_stats; the
 *     Python shim raises IndexError (as the reference does) when strict mode is on.
 *   - A context belongs to one device and is not thread-safe.  There is no CPU fallback: without
 *     a CUDA device dsc_create fails.
 */
#ifndef DSC_H
#define DSC_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSC_VERSION_STRING "dsc 0.1.0 (sm_100a)"

typedef struct dsc_ctx dsc_ctx;

typedef enum {
    DSC_OK = 0,
    DSC_ERR_INVALID = -1, /* bad argument */
    DSC_ERR_CUDA = -2,    /* CUDA runtime error (message has the cudaError string) */
    DSC_ERR_NOMEM = -3,
    DSC_ERR_CAPACITY = -4 /* caller-provided output buffer too small (required size reported) */
} dsc_status;

enum { DSC_MEM_HOST = 0, DSC_MEM_DEVICE = 1 };

/* DSC_VARIANT_A: ScanContext.ptcloud2sc_vectorized  scaner_version2.py:64-115
 *                (== kitti_dataset.py:47-98 == data_process/scaner.py:55-106)
 * DSC_VARIANT_B: ScanContextComputer._scan_context  scan_context_node.py:102-157 */
enum { DSC_VARIANT_A = 0, DSC_VARIANT_B = 1 };

/* dsc_params.flags */
enum {
    DSC_FLAG_GAP_F32 = 1,    /* variant A: num_ring/num_sector were Python ints, so NumPy does the
                                `//` in float32 instead of float64 (matters for non-dyadic gaps) */
    DSC_FLAG_EXACT_ONLY = 2, /* disable the screened fast path; every point takes the fdlibm path */
    DSC_FLAG_GAP_RING_F32 = 4,   /* as DSC_FLAG_GAP_F32, for the ring gap only (max_length / num_ring a Python float) */
    DSC_FLAG_GAP_SECTOR_F32 = 8  /* ... for the sector gap only (360 / num_sector a Python float) */
};

/* The reference keeps these as class statics (scaner_version2.py:28-47, kitti_dataset.py:28-41,
 * reconstruct_pointcloud.py:32-45) / constructor args (scan_context_node.py:31-45). */
typedef struct {
    int32_t num_ring;      /* ring_res[0],   default 160 */
    int32_t num_sector;    /* sector_res[0], default 720 */
    double  max_length;    /* 20 offline, 80 live */
    int32_t variant;       /* DSC_VARIANT_A / DSC_VARIANT_B */
    int32_t full_height;   /* use_full_height_range */
    float   height_offset; /* custom_height_offset */
    uint32_t flags;
} dsc_params;

/* Per-scan counters (device- or host-resident like the grids). */
typedef struct {
    int64_t n_points;
    int64_t n_nonfinite;       /* A: reference raises IndexError; B: masked out (:115-117) */
    int64_t n_sector_overflow; /* A: theta % 360 rounded to 360.0 -> sector == num_sector -> IndexError */
    int64_t n_clamped_ring;    /* A: r beyond max_length, clamped into ring num_ring-1 (:93) */
    int64_t n_out_of_range;    /* B: r > max_length, dropped (:134-137) */
    int64_t n_slow_path;       /* points that failed the fast-path screen and took the fdlibm path */
    int64_t n_boundary;        /* float64 input only: points whose sector coordinate theta/gap_sector lies within
                                  num_sector * 2^-50 of an integer -- the only points on which two faithful float64
                                  arctan2 implementations (NumPy-SVML, glibc, fdlibm) can pick different sectors */
} dsc_scan_stats;

/* ---- lifetime ------------------------------------------------------------------------------ */
const char *dsc_version(void);
int dsc_create(int device, dsc_ctx **out);
int dsc_destroy(dsc_ctx *ctx);
const char *dsc_last_error(const dsc_ctx *ctx);
/* Device time in ms of the kernels of the last DSC_MEM_DEVICE call issued with timing enabled
 * (dsc_set_timing(ctx,1)); the call then synchronises the stream.  -1 if not available. */
int dsc_set_timing(dsc_ctx *ctx, int enabled);
float dsc_last_kernel_ms(const dsc_ctx *ctx);
/* Engine options (debugging / A-B measurements; defaults are the measured best, see profiles/r1_notes.md and profiles/r2_notes.md):
 *   "k1_impl"        0 = auto (k1_scan; generic kernels for 12-byte stride / odd shapes / few small scans),
 *                    1 = generic L2-key kernels only, 3 = k1_scan whenever the shape allows
 *   "sc_cfg"         k1_scan launch shape: -1 = auto (768 threads x 6 loads per register set for whole scans, 768 x 3 when
 *                    scans are cut into work items), 0 = 768 x 6, 1 = 768 x 3, 2 = 1024 x 2
 *   "sc_smem_rows"   resident ring rows of k1_scan (-1 = as many as the carve-out allows)
 *   "sc_window"      first ring of k1_scan's resident window (-1 = chosen per call by k1_probe from a ring histogram
 *                    of sampled points; the grids never depend on it)
 *   "sc_split"       -1 = auto, 0 = never cut scans into work items, 1 = always
 *   "sc_item_points" target points per work item when scans are cut (default 131072)
 *   "no_defgrid"     1 = do not use the default-grid specialisation of k1_scan */
int dsc_set_option(dsc_ctx *ctx, const char *name, int64_t value);
/* Number of engine kernels launched since the context was created. */
int64_t dsc_kernel_launches(const dsc_ctx *ctx);

/* Pinned host memory for the DSC_MEM_HOST paths (pageable memory works too, more slowly). */
int dsc_host_alloc(dsc_ctx *ctx, size_t bytes, void **out);
int dsc_host_free(dsc_ctx *ctx, void *p);

/* ---- K1: scan-context build ------------------------------------------------------------------
 * Replaces load_velo_scan + ptcloud2sc_vectorized (scaner_version2.py:57-115) /
 * _scan_context (scan_context_node.py:102-157) for a batch of scans.
 *   pts              packed points of all scans back to back; row i starts at byte
 *                    i*pts_stride_bytes and holds float32 x,y,z (and intensity when the stride
 *                    is 16 -- the .bin layout, which is what enables 128-bit loads)
 *   offsets          HOST int64[n_scans+1], first point of each scan (in points); offsets[0]
 *                    need not be 0
 *   pts_stride_bytes 16 or 12
 *   out_grids        float32 [n_scans][num_ring][num_sector], C order; empty cells are 0
 *   out_norm         optional (NULL to skip): to_tensor01 of each grid (tiny_sc_ped.py:54-60)
 *   stats            optional dsc_scan_stats[n_scans] in the same memory space as out_grids
 */
int dsc_sc_build_batch(dsc_ctx *ctx, const dsc_params *params, const void *pts,
                       const int64_t *offsets, int n_scans, int pts_stride_bytes, int mem,
                       float *out_grids, float *out_norm, dsc_scan_stats *stats, void *stream);

/* ---- K1, float64 points ------------------------------------------------------------------------
 * The same two methods when the cloud is float64 -- what Open3D's voxel_down_sample returns, i.e. the live node's
 * only call path (scan_context_node.py:93-100 -> :161-168 -> _scan_context, variant B) and the offline generator
 * with use_downsampling=1 (scaner_version2.py:126-133 -> ptcloud2sc_vectorized, variant A).  NumPy then runs the
 * polar chain in float64; float64 arctan2 is the fdlibm algorithm (see dsc_scan_stats.n_boundary).
 *   pts              float64 rows x,y,z; row i starts at byte i*pts_stride_bytes (24, or 32 for padded rows)
 *   height_offset    custom_height_offset as float64 (the add is a float64 add; params->height_offset is ignored)
 *   out_grids        variant A: float64 [n_scans][num_ring][num_sector] (float64 heights, scaner_version2.py:97-108)
 *                    variant B: float32 [n_scans][num_ring][num_sector] (heights cast to float32 before the
 *                    reduction, scan_context_node.py:150-151)
 *   params->flags    ignored (there is no screened path and no float32 `//` for float64 input) */
int dsc_sc_build_batch_f64(dsc_ctx *ctx, const dsc_params *params, double height_offset, const void *pts,
                           const int64_t *offsets, int n_scans, int pts_stride_bytes, int mem,
                           void *out_grids, dsc_scan_stats *stats, void *stream);

/* ---- K4: per-grid min/max -> [0,1]  (to_tensor01, tiny_sc_ped.py:54-60;
 *      scaner_version2.py:402-407; scan_context_node.py:171-173) ------------------------------ */
int dsc_sc_normalize01(dsc_ctx *ctx, const float *grids, int n_grids, int64_t cells_per_grid,
                       int mem, float *out, void *stream);

/* ---- cell-centre tables ----------------------------------------------------------------------
 * r_centers[num_ring], cos_tab[num_sector], sin_tab[num_sector] (HOST float64), produced by the
 * caller with NumPy exactly as the reference does (kitti_dataset.py:219-229,
 * reconstruct_pointcloud.py:52-59), so the engine only performs individually rounded f64
 * mul/add/compare and stays bit-identical to the reference on any host.  Cached in the context. */
int dsc_set_polar_tables(dsc_ctx *ctx, int num_ring, int num_sector, const double *r_centers,
                         const double *cos_tab, const double *sin_tab);

/* ---- K2: pedestrian box -> covered cells  (cells_covered_by_box, kitti_dataset.py:204-253) ----
 *   boxes        HOST-or-device float64 [n_boxes][6] = x, y, l, w, cos(yaw), sin(yaw)
 *   cell_offsets int64 [n_boxes+1] (out): CSR offsets into cells_out
 *   cells_out    int32 [cap][2] (out): (ring, sector) rows, row-major order within each box
 *   dense_mask   optional uint8 [num_ring*num_sector] (out): 1 where any box covers the cell
 *                (build_mask_from_json, tiny_sc_ped.py:62-73)
 * Returns DSC_ERR_CAPACITY (with cell_offsets filled) when cap is too small. */
int dsc_raster_boxes(dsc_ctx *ctx, const double *boxes, int n_boxes, int mem,
                     int64_t *cell_offsets, int32_t *cells_out, int64_t cap, uint8_t *dense_mask,
                     void *stream);

/* ---- K3: grid -> points  (scan_context_to_pointcloud, reconstruct_pointcloud.py:47-77;
 *      loop twin scaner_version2.py:216-262) ---------------------------------------------------
 *   grids        float32 [n_grids][num_ring][num_sector]
 *   pt_offsets   int64 [n_grids+1] (out): CSR offsets into pts_out / bin_ids_out
 *   pts_out      float64 [cap][3] (out), row-major cell order per grid
 *   bin_ids_out  int32 [cap] (out): ring*num_sector + sector */
int dsc_reconstruct_batch(dsc_ctx *ctx, const float *grids, int n_grids, int full_height,
                          double height_offset, int mem, int64_t *pt_offsets, double *pts_out,
                          int32_t *bin_ids_out, int64_t cap, void *stream);

/* ---- K5: covered-cell helpers  (reconstruct_pointcloud.py:104-130, :153-164) -----------------
 * dsc_covered_mask: mask_out[i] = bin_ids[i] in {clip(r)*S + clip(s)}   (np.isin recolour)
 * dsc_gather_cells: points of the listed cells (clip, cell centre, grid height) */
int dsc_covered_mask(dsc_ctx *ctx, const int32_t *bin_ids, int64_t n_ids, const int32_t *cells,
                     int64_t n_cells, int mem, uint8_t *mask_out, void *stream);
int dsc_gather_cells(dsc_ctx *ctx, const float *grid, const int32_t *cells, int64_t n_cells,
                     int mem, double *pts_out, void *stream);

/* ---- K6: label-side training tensors ----------------------------------------
 * Built on the device from the rasteriser's CSR output, so a dataset __getitem__ no longer walks JSON cell lists.
 *
 * dsc_label_maps: binary mask (+ optional adaptive weight map) of a batch of frames.
 *   mode 0  cells outside the grid are ignored: ScanContextPedestrianDataset.__getitem__
 *           (distro_model/tiny_sc_pred_v2.py:57-85) and build_mask_from_json (tiny_sc_ped.py:62-73)
 *   mode 1  ring clipped to [0, R-1], sector wrapped with np.mod: ScanContextPedDataset.__getitem__
 *           (distro_model/train_sc_ped.py:175-183)
 *   cells / cell_offsets   the CSR lists dsc_raster_boxes writes ((ring, sector) int32 rows; int64 [n_boxes+1]);
 *                          both live where `mem` says
 *   box_weight   float32 [n_boxes]: min(min(100/num_cells, 5) * (1 + (distance/20)^2), 10) evaluated by the caller in
 *                Python floats (tiny_sc_pred_v2.py:64-76) -- scalar host arithmetic, like a4/a6; may be NULL when
 *                weight_out is NULL
 *   box_frame    int32 [n_boxes] frame of each box, or NULL (all frame 0)
 *   mask_out     float32 [n_frames][num_ring][num_sector]: 1 where any box covers the cell
 *   weight_out   optional float32, same shape: max(1, largest box_weight covering the cell)
 *
 * dsc_label_heatmap: centre heat-map, np.maximum over the pedestrians of gaussian2d(center, sigma)
 * (distro_model/train_sc_ped.py:74-79, :165-173; centres outside the grid are skipped, no sector wrap).
 *   centers      int32 [n_centers][2] (center_ring_idx, center_sector_idx); center_frame int32 [n_centers] or NULL
 *   table        HOST float64 [table_len]: exp(-0.5 * d2 / sigma**2) for the integer squared distances d2 = 0 ..
 *                table_len-1, tabulated by the caller with NumPy exactly as gaussian2d evaluates it (cut where it
 *                underflows to 0), so the heat-map is bit-identical to the reference's on any host
 *   heat_out     float64 [n_frames][num_ring][num_sector] */
int dsc_label_maps(dsc_ctx *ctx, const int32_t *cells, const int64_t *cell_offsets, const float *box_weight,
                   const int32_t *box_frame, int n_boxes, int n_frames, int num_ring, int num_sector, int mode,
                   int mem, float *mask_out, float *weight_out, void *stream);
int dsc_label_heatmap(dsc_ctx *ctx, const int32_t *centers, const int32_t *center_frame, int n_centers,
                      int n_frames, int num_ring, int num_sector, const double *table, int table_len, int mem,
                      double *heat_out, void *stream);

/* ---- K7: voxel-grid down-sampling ------------------------------------------------
 * What the live node runs before every scan context: ScanContextComputer._downsample_voxel
 * (scan_context_node.py:93-100), and the offline generator's use_downsampling branch (scaner_version2.py:126-133).
 * Both call Open3D's PointCloud.voxel_down_sample, an un-vendored dependency that is not installed here: PARITY
 * UNPINNED.  The kernel restates Open3D's published algorithm (float64: voxel_min_bound = min_bound - voxel/2; voxel
 * index = floor((p - voxel_min_bound) / voxel); output = sum of the voxel's points in input order / count) and is
 * deterministic.  Only the ORDER of the output points is ours (sorted by voxel index; Open3D iterates a hash map) -- the
 * scan context built from them does not depend on it.  Non-finite points are dropped and counted.
 *   pts       float32 (pts_f64 = 0; widened exactly, as Vector3dVector does) or float64 rows x,y,z
 *   out_pts   float64 [cap][3]; *n_out = number of voxels (DSC_ERR_CAPACITY when cap is too small)
 *   n_out / n_dropped   HOST int64 with DSC_MEM_HOST, DEVICE int64 with DSC_MEM_DEVICE (stream-ordered) */
int dsc_voxel_downsample(dsc_ctx *ctx, const void *pts, int64_t n_points, int pts_f64, int pts_stride_bytes,
                         double voxel_size, int mem, double *out_pts, int64_t cap, int64_t *n_out,
                         int64_t *n_dropped, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* DSC_H */
