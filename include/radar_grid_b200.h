/*
 * radar_grid_b200.h — C ABI of the B200-native radar gridding engine (libradargrid_b200.so).
 *
 * This is the drop-in boundary for the reference's gridding hot path.  The reference
 * (jgmarti84/radar-processor) is pure Python and has no FFI layer; the interface it exposes for this
 * path is the Python function API of `radar_grid` (reference src/radar_grid/__init__.py:39-82).  Every
 * entry point below names the reference function whose arithmetic it replaces; the Python mirror of
 * that API (radar-processor_b200/radar_grid_b200) binds these symbols with ctypes, and INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * Conventions
 *   - plain C: opaque handles, pointers + sizes, no C++/torch types.
 *   - every function returns an rg_status; rg_last_error() gives the message (thread-local).
 *     RG_ERR_INVALID maps to the ValueError the reference raises for the same misuse.
 *   - "memspace" says where the caller's buffers live: RG_HOST (pageable or pinned host memory; the
 *     library does the H2D/D2H copies itself) or RG_DEVICE (CUDA device pointers on the context's
 *     device; zero-copy, asynchronous on the context's stream).
 *   - voxel index = (iz*ny + iy)*nx + ix, gate index = ray*ngates + bin — the reference's flattening
 *     (reference compute.py:188-190, utils.py:35-38).
 *   - there is NO CPU fallback: every compute entry point fails with RG_ERR_CUDA without a device.
 */
#ifndef RADAR_GRID_B200_H
#define RADAR_GRID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RG_ABI_VERSION 2
#define RG_MAX_FIELDS 8      /* fields gridded per neighbour-table pass (one index load serves all) */
#define RG_MAX_RULES 8       /* fused QC range rules per apply call */
#define RG_MAX_SLICES 4      /* CAPPI / PPI planes per apply call */

typedef struct rg_context rg_context;     /* one per (device, stream); not shared between threads */
typedef struct rg_geometry rg_geometry;   /* device-resident neighbour table of one z-slab */

typedef enum rg_status {
    RG_OK = 0,
    RG_ERR_INVALID = 1,       /* bad argument  -> ValueError in the Python mirror */
    RG_ERR_CUDA = 2,          /* CUDA runtime failure / no device -> RuntimeError */
    RG_ERR_NOMEM = 3,
    RG_ERR_UNSUPPORTED = 4
} rg_status;

typedef enum rg_memspace { RG_HOST = 0, RG_DEVICE = 1 } rg_memspace;

/* reference compute.py:82-87 */
/* BARNES2 / CRESSMAN / NEAREST are the reference's weighting functions (compute.py:82-87; 'nearest' = weight 1 for EVERY
 * gate in the ROI).  DIST2 stores float32(d^2) in the weight slot: the table of the nearest-GATE gridding that
 * process_radar_to_cog asks pyart for (weighting_function='nearest', processor.py:152-163), applied with
 * rg_apply_args.reference_order = 2. */
typedef enum rg_weighting { RG_W_BARNES2 = 0, RG_W_CRESSMAN = 1, RG_W_NEAREST = 2, RG_W_DIST2 = 3 } rg_weighting;

/* Target grid.  Axes are float32 linspace(lo, hi, n) inclusive, computed with NumPy's formula
 * (reference compute.py:184-186).  [z_begin, z_end) selects the z-slab a geometry / grid covers; the
 * whole grid is z_begin = 0, z_end = nz. */
typedef struct rg_grid_spec {
    int32_t nz, ny, nx;
    int32_t z_begin, z_end;
    int32_t reserved_;
    double z_min, z_max, y_min, y_max, x_min, x_max;
} rg_grid_spec;

typedef struct rg_geometry_info {
    int64_t n_rows;            /* voxels in the slab = (z_end - z_begin)*ny*nx */
    int64_t n_pairs;           /* CSR non-zeros */
    int64_t n_gates;           /* length every field array must have */
    int64_t n_empty_rows;
    int64_t max_row_len;
    int64_t n_gates_binned;    /* gates that passed the TOA / domain cull (0 when imported from CSR) */
    int64_t n_candidates;      /* distance tests done by the build (0 when imported) */
    int64_t device_bytes;      /* HBM held by this geometry */
    double build_ms;           /* device time of the build (0 when imported) */
    double cell_size;
    rg_grid_spec grid;
} rg_geometry_info;

/* ---- library / context --------------------------------------------------------------------------- */
int rg_abi_version(void);
const char* rg_last_error(void);
int rg_device_count(int32_t* count);

/* `stream` is a cudaStream_t to run on (e.g. torch's current stream) or NULL for a private stream. */
int rg_context_create(int32_t device, void* stream, rg_context** out);
int rg_context_destroy(rg_context* ctx);
int rg_context_set_stream(rg_context* ctx, void* stream);
int rg_context_synchronize(rg_context* ctx);
/* the cudaStream_t the context launches on (for event / stream interop with the caller's framework) */
int rg_context_get_stream(const rg_context* ctx, void** stream);
/* number of this library's kernels launched through the context so far (bench.py: gpu_launches) */
int rg_context_kernel_launches(const rg_context* ctx, int64_t* count);
/* options: "group_width" (0 = auto, 4/8/16/32 lanes per voxel column), "apply_variant" (0 = auto: lane-group
 * kernel, pairs from the warp-slice copy of the table when two or more fields are gridded, else from the CSR copy;
 * 1 = lane-group kernel over the CSR copy; 2 = thread-per-column kernel over the interleaved table copy; 3 = lane-group
 * kernel with the generic product path; 4 = lane-group kernel over the warp-slice copy),
 * "timing" (1 = record CUDA events around every pack / apply launch, read with rg_context_kernel_time),
 * "sort_rows" (default 1: rg_geometry_build orders every row by gate id) */
int rg_context_set_option(rg_context* ctx, const char* key, int64_t value);
/* Sum of the device time of the timed launches since the last reset: which = 0 (K4 pack), 1 (K5 apply).
 * Synchronises the stream. */
int rg_context_kernel_time(rg_context* ctx, int32_t which, double* total_ms, int64_t* count, int32_t reset);

/* pinned host memory for the end-to-end path */
int rg_host_alloc(void** ptr, int64_t bytes);
int rg_host_free(void* ptr);

/* NumPy-compatible float32 linspace (host arithmetic; testable without a GPU) */
int rg_linspace_f32(double start, double stop, int32_t num, float* out);

/* ---- gate coordinates ----------------------------------------------------------------------------- */
/* Cartesian gate coordinates from the scan's polar description, on the device, so that a table build ships
 * (n_gates_per_ray + 2 n_rays) * 4 bytes instead of 12 bytes per gate.  Replaces the arrays the reference reads through
 * get_gate_coordinates (reference utils.py:12-38: radar.gate_x / gate_y / gate_z, which pyart derives with its 4/3-earth
 * antenna_to_cartesian): in float64
 *     z = sqrt(r^2 + R^2 + 2 r R sin(el)) - R,  s = R asin(r cos(el) / (R + z)),  x = s sin(az),  y = s cos(az),
 *     R = 4/3 * 6 371 000 m,
 * rounded once to float32; gate id = ray * n_bins + bin.  range_m[n_bins] in metres, azimuth_deg / elevation_deg[n_rays]
 * in degrees (float32, host or device per `memspace_in`); x, y, z [n_rays * n_bins] float32 in `memspace_out`. */
int rg_gate_coordinates(rg_context* ctx, const float* range_m, const float* azimuth_deg, const float* elevation_deg,
                        int64_t n_rays, int64_t n_bins, int32_t memspace_in, float* x, float* y, float* z,
                        int32_t memspace_out);

/* ---- neighbour table (K1-K3) ---------------------------------------------------------------------- */
/* Replaces compute_grid_geometry / _process_single_level (reference compute.py:106-284, 18-103):
 * TOA cull `gate_z - radar_altitude <= toa` (float32), counting-sort binning into a uniform cell grid,
 * warp-per-voxel float64 test ((dx*dx + dy*dy) + dz*dz) < r*r with r = max(min_radius, |voxel|*beam_factor),
 * CSR of uint32 gate ids + float32 weights.  Row order is deterministic (cell-major, gate id within a
 * cell), not the reference's KD-tree order; row *sets* and weights are what match. */
int rg_geometry_build(rg_context* ctx,
                      const float* gate_x, const float* gate_y, const float* gate_z, int64_t n_gates,
                      int32_t memspace, const rg_grid_spec* grid, double radar_altitude,
                      double min_radius, double beam_factor, int32_t weighting, double toa,
                      rg_geometry** out);

/* Level census for balanced z-slabs (no reference counterpart; the reference's only parallelism is a Pool over z-levels,
 * compute.py:203-222): the binning pass and the counting pass of rg_geometry_build over every `column_stride`-th
 * column in x and y; pairs_per_level (host, z_end - z_begin entries) receives the pair count of every level, scaled
 * to the full plane (exact for column_stride = 1).  Nothing is kept on the device. */
int rg_geometry_level_pairs(rg_context* ctx,
                            const float* gate_x, const float* gate_y, const float* gate_z, int64_t n_gates,
                            int32_t memspace, const rg_grid_spec* grid, double radar_altitude,
                            double min_radius, double beam_factor, double toa, int32_t column_stride,
                            int64_t* pairs_per_level);

/* Import an existing table (e.g. one the reference built and saved with save_geometry, reference
 * geometry.py:94-118).  indptr has n_rows+1 entries of `indptr_bits` (32 or 64) bits; row order is kept. */
int rg_geometry_from_csr(rg_context* ctx, const rg_grid_spec* grid, const void* indptr, int32_t indptr_bits,
                         const int32_t* gate_indices, const float* weights, int64_t n_gates,
                         int32_t memspace, rg_geometry** out);

int rg_geometry_get_info(const rg_geometry* geom, rg_geometry_info* info);

/* State of the table's column-pair ("duo") copy, the layout rg_apply reads on tables of short rows, where neighbouring
 * columns share most of their gates (the reference re-streams the whole table per field, interpolate.py:107-142):
 * > 0 = built, that many 384-byte slots; 0 = not asked for yet; < 0 = not available for this table (rows not sorted by
 * gate id, as in an imported reference table, or no room).  Context option "duo": 0 never, 1 auto (default: mean row
 * below 96 pairs), 2 whenever the table allows it. */
int rg_geometry_duo_slots(const rg_geometry* geom, int64_t* n_slots);
/* Export as the reference's GridGeometry arrays (reference geometry.py:14-52).  Any pointer may be NULL. */
int rg_geometry_export_csr(rg_context* ctx, const rg_geometry* geom, void* indptr, int32_t indptr_bits,
                           int32_t* gate_indices, float* weights, int32_t memspace);
int rg_geometry_destroy(rg_geometry* geom);

/* ---- products (K6) -------------------------------------------------------------------------------- */
typedef enum rg_product_kind {
    RG_PROD_COLMAX = 1,        /* column_max   reference products.py:420-490 (np.nanmax over z_lo..z_hi) */
    RG_PROD_COLMIN = 2,        /* column_min   reference products.py:493-535 */
    RG_PROD_COLMEAN = 3,       /* column_mean  reference products.py:538-580 */
    RG_PROD_LEVEL = 4,         /* CAPPI        reference products.py:317-415: level pick or 2-level blend */
    RG_PROD_BEAM = 5           /* PPI          reference products.py:168-314: beam-following slice */
} rg_product_kind;

typedef enum rg_blend_mode {
    RG_BLEND_PICK = 0,         /* out = grid[z_lo]                                   (float32 out) */
    RG_BLEND_F32 = 1,          /* out = f32(w_lo)*v_lo + f32(w_hi)*v_hi in float32   (float32 out) */
    RG_BLEND_F64 = 2,          /* same in float64, rounded to float32 at the end     (float32 out) */
    RG_BLEND_F64_OUT64 = 3     /* same in float64, float64 output (PPI 'linear')     (float64 out) */
} rg_blend_mode;

/* Image form of a 2-D product, written by the same epilogue that finishes the product so that only ny*nx*4 bytes per
 * field have to leave the GPU: GridFilter thresholds (reference filters.py:631-746, applied in order, each replacing
 * the selected pixels by its fill value), then the colormap of reference geotiff.py:70-145 -- no-data test (== fill_value,
 * or NaN when has_fill_value is 0), Normalize(vmin, vmax, clip=True) in the plane's arithmetic type, LUT index
 * int(x * lut_entries) with 1.0 mapped to the last entry and NaN to the "bad" entry lut_entries + 2, alpha 0 for
 * no-data.  `lut` holds lut_entries + 3 RGBA byte quadruples, already scaled ((lut * 255).astype(uint8): matplotlib's
 * Colormap._lut including its under / over / bad rows), in the memory space of the call. */
#define RG_MAX_IMAGE_FILTERS 4
typedef struct rg_image {
    int32_t n_filters;                               /* 0..RG_MAX_IMAGE_FILTERS */
    int32_t filter_kind[RG_MAX_IMAGE_FILTERS];       /* rg_plane_filter_kind */
    double filter_a[RG_MAX_IMAGE_FILTERS];
    double filter_b[RG_MAX_IMAGE_FILTERS];
    double filter_fill[RG_MAX_IMAGE_FILTERS];
    double vmin, vmax;                               /* vmin <= vmax, finite */
    double fill_value;
    int32_t has_fill_value;
    int32_t lut_entries;                             /* N (256 for matplotlib colormaps), 1..4096 */
    const uint8_t* lut;                              /* [(N + 3) * 4] */
    uint8_t* out;                                    /* [n_fields][ny*nx][4] */
} rg_image;

/* One 2-D product.  `out` receives n_fields planes of ny*nx elements, field-major (may be NULL when `image` is set:
 * then only the RGBA image is written).  z indices are GLOBAL level indices (0..nz-1), inclusive. */
typedef struct rg_product {
    int32_t kind;              /* rg_product_kind */
    int32_t mode;              /* rg_blend_mode (LEVEL); for BEAM: 0 = 'linear', 1 = 'nearest', 2 = radar_processor's own collapse
                                * (processor.py:513-528): level closest to r sin(el) + r^2 / (2 * 8.49e6) on float64 axes, no
                                * out-of-grid NaN; uses sin_elev only */
    int32_t z_lo, z_hi;
    int32_t earth_curvature;   /* BEAM */
    int32_t partial;           /* 0: the finished product.  1: this z-slab's TERM of it, to be combined across slabs by one
                                * all-reduce: COLMAX / COLMIN write -inf / +inf where the slab has no data (all-reduce MAX /
                                * MIN); LEVEL and BEAM write the sum over the levels the slab OWNS of weight * value, -0.0
                                * for levels of other slabs (all-reduce SUM; a float64 blend stays float64: RG_BLEND_F64
                                * then has a float64 plane), NaN where the reference's result is NaN for every slab */
    double w_lo, w_hi;         /* LEVEL blend weights */
    double sin_elev;           /* BEAM: np.sin(np.radians(elev)) */
    double cos_elev_clamped;   /* BEAM: np.maximum(np.cos(np.radians(elev)), 0.01) */
    double tan_elev;           /* BEAM, flat earth: np.tan(np.radians(elev)) */
    double ke_re;              /* BEAM: ke * EARTH_RADIUS */
    double ke_re_sq;           /* BEAM: ke_re ** 2 */
    void* out;
    const rg_image* image;     /* NULL, or the RGBA form of this product (not with partial) */
} rg_product;

/* Products of existing 3-D grids: n_fields grids of (z_end-z_begin)*ny*nx float32 each. */
int rg_products(rg_context* ctx, const rg_grid_spec* grid, int32_t n_fields, const float* const* grids,
                int32_t n_products, const rg_product* products, int32_t memspace);

/* GridFilter on a 2-D product (reference filters.py:631-746): out = in with the selected elements replaced
 * by fill_value.  elem_bits is 32 (float) or 64 (double: PPI 'linear' planes). */
typedef enum rg_plane_filter_kind {
    RG_PF_BELOW = 1,           /* apply_below:   in <  a            */
    RG_PF_ABOVE = 2,           /* apply_above:   in >  a            */
    RG_PF_OUTSIDE = 3,         /* apply_outside_range: in < a || in > b */
    RG_PF_INVALID = 4,         /* apply_invalid: NaN or Inf          */
    RG_PF_BELOW_EQUAL = 5      /* in <= a: np.ma.masked_less_equal(arr, vmin), radar_processor's re-mask (processor.py:541-546) */
} rg_plane_filter_kind;
int rg_plane_filter(rg_context* ctx, const void* in, void* out, int64_t n, int32_t elem_bits, int32_t kind,
                    double a, double b, double fill_value, int32_t memspace);

/* ---- interpolation (K4 + K5, products fused as the epilogue) --------------------------------------- */
/* Fused gate-mask rule: exclude gate g from the fields in `field_bits` when
 * (use_lo && values[g] < lo) || (use_hi && values[g] > hi)  — GateFilter.exclude_below / exclude_above /
 * exclude_outside (reference filters.py:114-211); NaN compares false exactly as there. */
typedef struct rg_qc_rule {
    const float* values;       /* n_gates values of the QC field (may alias one of the gridded fields) */
    float lo, hi;
    int32_t use_lo, use_hi;
    uint32_t field_bits;
    uint32_t reserved_;
} rg_qc_rule;

typedef struct rg_apply_args {
    int32_t n_fields;                      /* 1..RG_MAX_FIELDS */
    int32_t n_rules;                       /* 0..RG_MAX_RULES */
    int32_t n_products;                    /* COLMAX/COLMIN/COLMEAN at most one each, <= RG_MAX_SLICES slices */
    int32_t reference_order;               /* 0: fast fused path.  1: np.add.reduceat summation order -> bit-exact vs reference.
                                            * 2: nearest-gate gridding on a RG_W_DIST2 table: every voxel takes, per field, the
                                            *    value of the closest gate whose value is not masked (ties: lowest gate id);
                                            *    Py-ART map_gates_to_grid, weighting_function='nearest' */
    uint32_t mask_invalid_bits;            /* bit f: NaN/Inf of field f are masked (np.ma.masked_invalid) */
    float fill_value;
    const float* const* fields;            /* [n_fields] -> float32[n_gates] */
    const uint8_t* const* masks;           /* NULL, or [n_fields] -> uint8[n_gates] (nonzero = excluded) / NULL */
    const rg_qc_rule* rules;
    float* const* grid_out;                /* NULL (no 3-D output), or [n_fields] -> float32[n_rows] / NULL */
    const rg_product* products;
} rg_apply_args;

/* Replaces apply_geometry / apply_geometry_multi (reference interpolate.py:15-142):
 * out[v] = sum_j w_j v_j m_j / sum_j w_j m_j in float32, fill_value when the row is empty or the
 * effective weight sum is not > 0.  All n_fields share one pass over the table.  2-D products are
 * computed in the epilogue, so a products-only call never writes a 3-D grid to HBM. */
int rg_apply(rg_context* ctx, const rg_geometry* geom, const rg_apply_args* args, int32_t memspace);

#ifdef __cplusplus
}
#endif
#endif /* RADAR_GRID_B200_H */
