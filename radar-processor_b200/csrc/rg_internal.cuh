// Internal declarations shared by the translation units of libradargrid_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <mutex>
#include <string>
#include <vector>

#include "radar_grid_b200.h"

namespace rg {

// ---- error plumbing ---------------------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int status, const std::string& msg);

#define RG_CUDA(expr)                                                                              \
    do {                                                                                           \
        cudaError_t rg_e_ = (expr);                                                                \
        if (rg_e_ != cudaSuccess)                                                                  \
            return ::rg::fail(rg_e_ == cudaErrorMemoryAllocation ? RG_ERR_NOMEM : RG_ERR_CUDA,     \
                              std::string(#expr) + ": " + cudaGetErrorString(rg_e_));              \
    } while (0)

#define RG_TRY(expr)                          \
    do {                                      \
        int rg_s_ = (expr);                   \
        if (rg_s_ != RG_OK) return rg_s_;     \
    } while (0)

// ---- constants --------------------------------------------------------------------------------------
// A masked gate value is stored as this bit pattern inside the packed gate records (a NaN payload no
// arithmetic produces); genuine NaNs are canonicalised to kCanonNaN by the pack kernel so that an
// *unmasked* NaN still propagates exactly as in the reference (interpolate.py:78-82).
constexpr uint32_t kMaskedBits = 0xFFFFFFFFu;
constexpr uint32_t kCanonNaN = 0x7FC00000u;
constexpr uint32_t kInvalidCell = 0xFFFFFFFFu;

#ifndef RG_APPLY_THREADS
#define RG_APPLY_THREADS 128
#endif
constexpr int kApplyThreads = RG_APPLY_THREADS;   // threads per CTA of the column-tile apply kernel
constexpr uint32_t kHeavyRow = 512;      // rows longer than this are reduced by whole CTAs in their own launch (heavy_rows_kernel)
constexpr uint32_t kHeavyChunk = 4096;   // pairs of a heavy row one CTA of kHeavyThreads reduces
constexpr int kHeavyThreads = 256;
constexpr int kSellThreads = 64;         // threads per CTA of the thread-per-column apply kernel
constexpr uint32_t kSellCap = 192;       // pairs of a row kept in the interleaved copy; the rest is read from the CSR

// Device-resident neighbour table of one z-slab.
struct Geometry {
    int device = 0;
    rg_grid_spec grid{};
    int64_t n_rows = 0, n_pairs = 0, n_gates = 0, ncol = 0;
    int32_t n_levels = 0;
    uint32_t* indptr = nullptr;          // [n_rows + 1]
    uint2* pairs = nullptr;              // [n_pairs] {gate id, float32 weight bits}
    // Sliced, lane-interleaved, compacted copy of the first kSellCap pairs of every row (see rg_apply.cu):
    uint2* sell = nullptr;               // [n_sell]
    uint32_t* slice_base = nullptr;      // [n_levels * slices_per_level + 1]
    int64_t n_sell = 0, slices_per_level = 0;
    // Warp-slice copy for the column-group kernel: the rows of the 32/W adjacent columns one warp sums at one level
    // form a slice; slot j of a slice holds pairs [jW, jW+W) of each of its rows, lane order, idle entries filled
    // with the all-masked record.  One slot is one 256-byte coalesced load of the warp.
    // One copy per group width in use (W = 4, 8, 16, 32 -> index 0..3), built on first use under quad_mu and kept
    // until the table is destroyed, so that contexts on other streams can share it.
    struct QuadCopy {
        uint2* quads = nullptr;          // [n_slots][32]
        uint32_t* ptr = nullptr;         // [n_levels * ny * quads_x + 1]: (first slot << 1) | slice has a heavy row
        int64_t n_slots = 0;
        int32_t quads_x = 0;
    } quad[4];
    // Column-pair ("duo") copy for the multi-field kernel of rg_duo.cu: a group of 4 lanes owns the two columns (x, 2yp)
    // and (x, 2yp + 1); the rows of the two columns at one level are merged by gate id into entries {gate, w0, w1}
    // (w = 0 where the gate is not in that column's row), so every gate record is gathered once for both columns.
    // The 8 groups of a warp (8 adjacent x) at one level form a slice; slot j of a slice holds entries [4j, 4j+4) of
    // each of its groups in lane order: 32 gate ids (128 B) followed by 32 weight pairs (256 B).  Idle lanes point at
    // the all-masked record with absent weights (-0.0).  Built on first use under quad_mu; n_slots < 0: not available for
    // this table (rows not sorted by gate id, or no room).
    struct DuoCopy {
        uint32_t* slots = nullptr;       // [n_slots][96]
        uint32_t* ptr = nullptr;         // [n_levels * nyp * qx + 1]: (first slot << 1) | slice has a heavy row
        int64_t n_slots = 0;
        int32_t qx = 0, nyp = 0;
    } duo;
    std::mutex quad_mu;
    // Rows longer than kHeavyRow (voxels next to the radar see the first gates of every ray): sorted row ids, the
    // chunks (<= kHeavyChunk pairs) they are cut into, built once per table under quad_mu (ensure_heavy).
    uint32_t* heavy_rows = nullptr;      // [n_heavy] ascending local row ids
    uint32_t* heavy_first = nullptr;     // [n_heavy + 1] first chunk of every heavy row
    uint2* heavy_chunks = nullptr;       // [n_heavy_chunks] {first pair, number of pairs}
    int64_t n_heavy = -1, n_heavy_chunks = 0;   // -1: not looked for yet
    float* x_ax = nullptr;               // [nx]   float32 linspace axes (reference compute.py:184-186)
    float* y_ax = nullptr;               // [ny]
    float* z_ax = nullptr;               // [nz]   (full grid)
    rg_geometry_info info{};
};

// Per-(device, stream) state: the stream and grow-only scratch buffers.
struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
};

// Optional per-kernel device timing (bench.py roofline): CUDA events recorded around the pack / apply launches.
struct KernelTimer {
    std::vector<cudaEvent_t> start, stop;
};
enum { kTimerPack = 0, kTimerApply = 1, kTimerCount = 2 };

struct Context {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int64_t launches = 0;
    int64_t apply_variant = 0;           // 0 = auto
    int64_t group_width = 0;             // 0 = auto
    int sm_count = 148;
    int64_t timing = 0;                  // record events around pack/apply launches
    int64_t sort_rows = 1;               // geometry build: order every row by gate id
    KernelTimer timers[kTimerCount];
    Scratch records;                     // packed gate records
    unsigned long long tex_a = 0, tex_b = 0;   // texture objects over the record arrays (RG_TEX builds)
    const void* tex_a_ptr = nullptr;
    const void* tex_b_ptr = nullptr;
    int tex_fields = 0;
    int64_t tex_gates = -1;
    Scratch stage_in;                    // H2D staging of fields / masks / rule values
    Scratch stage_out;                   // device-side outputs of a host-memspace call
    Scratch misc;
    Scratch luts;                        // colormap LUTs of the image products of a call
    Scratch heavy;                       // per-chunk partial sums of the heavy rows: float[n_heavy_chunks][2 * F]
    Scratch flags;                       // uint32: epoch of the last volume that held an UNMASKED non-finite value (see PackParams)
    uint32_t epoch = 0;                  // counts rg_apply calls of this context
    int64_t duo = 1;                     // option "duo": 0 = never use the column-pair kernel, 1 = auto (tables of short rows), 2 = whenever the table allows it
};

int ensure(Context* ctx, Scratch& s, size_t bytes);
void timer_begin(Context* ctx, int which);
void timer_end(Context* ctx, int which);

// ---- kernel-facing parameter blocks -----------------------------------------------------------------
constexpr int RG_MAX_IMAGES = 3;          // products of one call that also leave as RGBA images
// RGBA form of a product plane (rg_image): thresholds, then the colormap index arithmetic of geotiff.py:70-145
struct ImageParams {
    int32_t on;
    int32_t n_filters;
    int32_t kind[RG_MAX_IMAGE_FILTERS];
    double a[RG_MAX_IMAGE_FILTERS], b[RG_MAX_IMAGE_FILTERS], fill[RG_MAX_IMAGE_FILTERS];
    double vmin, vmax, fill_value;
    int32_t has_fill, lut_n;
    const uchar4* lut;                    // device, lut_n + 3 entries
    uchar4* out;                          // [n_fields][ncol]
};

struct SliceParams {                      // one RG_PROD_LEVEL / RG_PROD_BEAM product
    int32_t kind;                         // 0 = unused, RG_PROD_LEVEL, RG_PROD_BEAM
    int32_t mode;                         // LEVEL: rg_blend_mode.  BEAM: 0 linear, 1 nearest
    int32_t z_lo, z_hi;
    int32_t curvature;
    int32_t partial;                      // z-slab term: only the levels this slab owns contribute, the others add -0.0
    double w_lo, w_hi;
    double sin_e, cos_c, tan_e, ke_re, ke_re_sq;
    void* out;
    int32_t image;                        // index into ProductParams::images, -1 = none
    int32_t pad_;
};

struct ProductParams {
    int32_t any;                          // any product requested
    int32_t cmax_on, cmax_z0, cmax_z1;
    int32_t cmin_on, cmin_z0, cmin_z1;
    int32_t cmean_on, cmean_z0, cmean_z1;
    uint32_t cmax_w, cmin_w, cmean_w;     // z1 - z0 + 1 when on, 0 when off
    int32_t n_slices;
    int32_t nz_full;
    int32_t own_z0, own_z1;               // global levels [own_z0, own_z1) of the slab being gridded (partial products)
    int32_t cmax_partial, cmin_partial;   // no-data pixels are written as -inf / +inf, ready for all-reduce(MAX / MIN)
    float* cmax_out;
    float* cmin_out;
    float* cmean_out;
    int32_t cmax_image, cmin_image, cmean_image, n_images;   // index into images, -1 = none
    ImageParams images[RG_MAX_IMAGES];
    double z_min, z_max, z_step;
    double x_min, x_max, y_min, y_max;    // float64 axes of the closest-level beam product (BEAM mode 2)
    int32_t nx, ny;
    const float* x_ax;
    const float* y_ax;
    SliceParams slices[RG_MAX_SLICES];
    // The same products as a short op list for kernels that keep the per-(field, column) state in shared memory:
    // only requested products cost instructions.  kind 1 max, 2 min, 3 mean, 4 capture (uniform levels),
    // 5 capture (per-column levels, BEAM).  slot = first state word.
    int32_t n_ops;
    struct Op { int32_t kind, z0, z1, slot; uint32_t w; int32_t k; } ops[3 + RG_MAX_SLICES];
    // shared-memory slots of the per-(field, column) product state (-1 = unused)
    int32_t n_state_words;
    int32_t slot_cmax, slot_cmin, slot_cmean, slot_slice[RG_MAX_SLICES];
};

struct ApplyParams {
    const uint32_t* indptr;
    const uint2* pairs;
    const float* records;                 // [n_gates + 1][FA]  fields 0..3 (last record: all masked)
    const float* records_b;               // [n_gates + 1][FB]  fields 4..7
    const uint2* sell;                    // interleaved copy of the table (thread-per-column kernel)
    const uint32_t* slice_base;
    int64_t slices_per_level;
    const uint2* quads;                   // warp-slice copy of the table (column-group kernel)
    const uint32_t* quad_ptr;
    int32_t quads_x;
    uint32_t null_gate;                   // index of the all-masked record (= n_gates)
    const uint32_t* heavy_rows;           // rows summed by heavy_rows_kernel (sorted), their chunk ranges and partial sums
    const uint32_t* heavy_first;
    const uint2* heavy_chunks;
    float* heavy_part;                    // [n_heavy_chunks][2 * n_fields]: sum(w*v) per field, then sum(w) per field
    int32_t n_heavy, n_heavy_chunks;
    const uint32_t* duo;                  // column-pair copy of the table (rg_duo.cu), nullptr = not in use
    const uint32_t* duo_ptr;
    int32_t duo_qx, duo_nyp;
    const uint32_t* nonfinite;            // == epoch: this volume holds an unmasked non-finite value (zero weights must not touch it)
    uint32_t epoch;
    unsigned long long tex_a, tex_b;      // texture objects over records / records_b (RG_TEX builds)
    int64_t ncol;                         // ny*nx
    int32_t nx, ny;
    int32_t z_begin;                      // global index of local level 0
    int32_t lz_first, lz_last;            // local levels [first, last) this launch walks
    int32_t n_fields;
    float fill;
    float* grid_out[RG_MAX_FIELDS];
    ProductParams prod;
};

struct RecSrc {                           // where the apply kernels gather gate records from
    const float* a;
    const float* b;
    unsigned long long tex_a, tex_b;
    uint32_t null_gate;
};

struct PackParams {
    int64_t n_gates;
    int32_t n_fields;
    int32_t n_rules;
    uint32_t invalid_bits;
    const float* fields[RG_MAX_FIELDS];
    const uint8_t* masks[RG_MAX_FIELDS];
    const float* rule_values[RG_MAX_RULES];
    float rule_lo[RG_MAX_RULES], rule_hi[RG_MAX_RULES];
    int32_t rule_use_lo[RG_MAX_RULES], rule_use_hi[RG_MAX_RULES];
    uint32_t rule_bits[RG_MAX_RULES];
    float* records;
    float* records_b;
    // An unmasked NaN / inf value propagates into every voxel whose row holds the gate (interpolate.py:78-82).  The
    // column-pair kernel multiplies such a value by a ZERO weight for the column that does not hold the gate, so it has to
    // know: the pack kernel stamps *nonfinite with the call's epoch when it sees one, and the kernel then takes its
    // predicated path.
    uint32_t* nonfinite;
    uint32_t epoch;
};

// ---- launchers (defined in the .cu files) -----------------------------------------------------------
int bind_record_textures(Context* ctx, const float* rec_a, const float* rec_b, int n_fields, int64_t n_gates);
int records_width(int n_fields);          // floats per packed gate record: 1, 2, 4 or 8
size_t records_b_offset(int n_fields, int64_t n_gates);   // byte offset of array B inside the record buffer 
int launch_pack(Context* ctx, const PackParams& p);
int launch_apply(Context* ctx, const Geometry* g, const ApplyParams& p, bool reference_order);
int launch_apply_nearest(Context* ctx, const Geometry* g, const ApplyParams& p);
int launch_products(Context* ctx, const rg_grid_spec& grid, int n_fields, const float* const* grids_dev,
                    const ProductParams& prod);
int build_geometry_device(Context* ctx, const float* gx, const float* gy, const float* gz, int64_t n_gates,
                          double radar_altitude, double min_radius, double beam_factor, int weighting,
                          double toa, Geometry* out, int col_stride = 1, int64_t* level_pairs_host = nullptr);
int finalize_geometry_stats(Context* ctx, Geometry* g);
int build_sell(Context* ctx, Geometry* g);
int ensure_quads(Context* ctx, Geometry* g, int W, const Geometry::QuadCopy** out);
int ensure_heavy(Context* ctx, Geometry* g);
int ensure_duo(Context* ctx, Geometry* g, const Geometry::DuoCopy** out);
int launch_duo(Context* ctx, const ApplyParams& p);      // rg_duo.cu: the column-pair kernel alone (no heavy launch, no timers)
int exclusive_scan_u32(Context* ctx, const uint32_t* in, uint32_t* out, int64_t n, unsigned long long* tmp, uint64_t* total_host);
void linspace_f32(double start, double stop, int num, float* out);

}  // namespace rg
