// Device code shared by the apply kernels (rg_apply.cu, rg_duo.cu): build-time switches, the gate-record layout, the
// per-column product state (K6) and the record-gather helpers.  Everything here is header-only device code.
#pragma once

#include <math.h>

#include "rg_internal.cuh"

namespace rg {

// Build-time switches: every design decision of K5 can be re-measured (tools/gpu_ab.sh, build.py --variant X -DRG_...).
#ifndef RG_PREFETCH
#define RG_PREFETCH 1          // levels of look-ahead for the L2 prefetch of the pair stream (0 = off)
#endif
#ifndef RG_TREDUCE
#define RG_TREDUCE 1           // 1: reduce-scatter the row sums inside a group (14 shuffles), 0: plain butterfly (6F)
#endif
#ifndef RG_UNROLL
#define RG_UNROLL 4            // pairs (and their gathers) in flight per lane
#endif
#ifndef RG_MINBLOCKS
#define RG_MINBLOCKS (1024 / RG_APPLY_THREADS)     // 64 registers per thread: 32 resident warps per SM
#endif
#ifndef RG_TEX
#define RG_TEX 0               // 1: gather the gate records through the texture path (tex1Dfetch) instead of LDG
#endif

#ifndef RG_PDL
#define RG_PDL 1               // heavy_rows_kernel -> column kernel chained by programmatic dependent launch (cfg3 0.659 -> 0.650 ms,
                               //    cfg1 0.0525 -> 0.0465 ms)
#endif
#ifndef RG_ILPRE
#define RG_ILPRE 0             // 1: slice-copy passes load the first batch of the NEXT level's pair slots into registers right after
                               //    the current level's gathers, so that they travel while the row sums are reduced and stored.
                               //    Measured: cfg3 0.760 vs 0.633 ms, cfg1 0.0462 vs 0.0432 ms: off
#endif
#ifndef RG_PRELOAD
#define RG_PRELOAD 0           // N > 0: one- and two-field passes over the CSR copy keep the first N pairs of the NEXT level's row in
                               //    registers.  Measured with N = 3: cfg1 0.0466 vs 0.0465 ms, one field of cfg3 0.565 vs 0.540 ms: off.
#endif
#ifndef RG_MASKBITS
#define RG_MASKBITS 0          // 1: odd field counts keep a mask-bit word in the free record slot (see Layout): packed FFMA2 value sums
                               //    and R2P predicates, 14 % fewer instructions -- and 16 % SLOWER (0.768 vs 0.663 ms): the kernel is
                               //    bound by the bytes its gathers pull through L1 (8 instead of 4 for array B) and by load latency,
                               //    not by issue slots (profiles/r02_maskbits_vs_marker.md).  0 (default): marker values only.
#endif

// Record layout per field count F (shared by K4, K5 and the exact kernel): array A = float[G+1][FA] holds fields 0..3,
// array B = float[G+1][FB] fields 4..7 (an interleaved 32-byte record lost twice: as two 128-bit loads in round 1 and
// as one 256-bit load, LDG.E.256, in round 2 -- 0.855 vs 0.643 ms).
//   MB layouts (odd F >= 3: one float of the last vector is free): masked values are stored as +0.0 and the free slot,
//   v[F], holds one mask bit per field (bit f + SH set = field f masked).  sum(w*v) then needs no predicate and pairs
//   up into packed FFMA2s (fma.rn.f32x2, sm_100+), and the predicates of the sum(w) adds come out of the mask word
//   with ONE R2P instead of an ISETP per field: 9 instead of 15 instructions per pair at five fields.
//   Other F: a masked value is the bit pattern kMaskedBits (ISETP + predicated FADD + FFMA per field).
template <int F>
struct Layout {
    static constexpr int FP = F == 1 ? 1 : F == 2 ? 2 : F <= 4 ? 4 : 8;
    static constexpr int FA = F == 1 ? 1 : F == 2 ? 2 : 4;
    static constexpr int FB = F <= 4 ? 0 : F == 5 ? (RG_MASKBITS ? 2 : 1) : F == 6 ? 2 : 4;
    static constexpr int NV = FA + FB;           // floats gathered per gate
    static constexpr bool MB = RG_MASKBITS && (F == 3 || F == 5 || F == 7);   // mask bits in slot F = NV - 1
    static constexpr int SH = F <= 5 ? 1 : 0;    // R2P fills P1.. in one instruction; bit 0 would cost two more
};

// ------------------------------------------------------------------------------------------------------
// K6  per-column product state, shared by the fused epilogue and the stand-alone kernel so that both
//     give bit-identical planes
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double beam_target_z(const SliceParams& s, float x, float y)
{
    // products.py:235  horizontal_dist = sqrt(xx**2 + yy**2) in float32
    const float h = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)));
    if (s.curvature) {
        // products.py:80-87 in float64 (float32 array / np.float64 scalar promotes)
        const double sr = __ddiv_rn((double)h, s.cos_c);
        const double a = __dadd_rn(__dmul_rn(sr, sr), s.ke_re_sq);
        const double b = __dmul_rn(__dmul_rn(__dmul_rn(2.0, sr), s.ke_re), s.sin_e);
        return __dadd_rn(__dsub_rn(__dsqrt_rn(__dadd_rn(a, b)), s.ke_re), 0.0);
    }
    return __dadd_rn(__dmul_rn((double)h, s.tan_e), 0.0);   // products.py:165
}

__device__ __forceinline__ int clampi(double v, int lo, int hi)
{
    if (!(v >= (double)lo)) return lo;     // also catches NaN
    if (v > (double)hi) return hi;
    return (int)v;
}

// RGBA form of one finished product value (rg_image): GridFilter thresholds in the plane's own type (filters.py:660,
// 689, 721, 746), then reference geotiff.py:122-143 with matplotlib's arithmetic: Normalize(vmin, vmax, clip=True) --
// np.clip against the float64 limits promotes to float64, so clip, subtract and divide are float64 operations -- and
// Colormap.__call__: xa = x * N, xa == N -> N - 1, NaN -> the "bad" row N + 2, astype(int) truncation.  The LUT rows are
// already (lut * 255).astype(uint8); alpha is 0 for no-data pixels.
template <typename T>
__device__ __forceinline__ void emit_image(const ImageParams& im, size_t o, T v)
{
#pragma unroll
    for (int i = 0; i < RG_MAX_IMAGE_FILTERS; ++i) {
        if (i < im.n_filters) {
            const T a = (T)im.a[i], b = (T)im.b[i];
            bool hit;
            switch (im.kind[i]) {
                case RG_PF_BELOW: hit = v < a; break;
                case RG_PF_ABOVE: hit = v > a; break;
                case RG_PF_OUTSIDE: hit = (v < a) || (v > b); break;
                case RG_PF_BELOW_EQUAL: hit = v <= a; break;
                default: hit = isnan(v) || isinf(v); break;
            }
            if (hit) v = (T)im.fill[i];
        }
    }
    const bool nodata = im.has_fill ? (v == (T)im.fill_value) : isnan(v);
    int idx;
    if (isnan(v)) {
        idx = im.lut_n + 2;
    } else {
        double x = 0.0;
        if (im.vmin != im.vmax) {
            const double d = (double)v;
            const double c = d < im.vmin ? im.vmin : (d > im.vmax ? im.vmax : d);
            x = __ddiv_rn(__dsub_rn(c, im.vmin), __dsub_rn(im.vmax, im.vmin));
        }
        double xa = __dmul_rn(x, (double)im.lut_n);
        if (xa == (double)im.lut_n) xa = (double)(im.lut_n - 1);
        idx = xa < 0.0 ? im.lut_n : (xa >= (double)im.lut_n ? im.lut_n + 1 : (int)xa);
    }
    uchar4 c = __ldg(im.lut + idx);
    if (nodata) c.w = 0;
    im.out[o] = c;
}

// np.linspace(start, stop, num)[i] in float64 (arange * step + start, the last element is stop itself)
__device__ __forceinline__ double linspace_f64(double start, double stop, int num, int i)
{
    if (num <= 1) return start;
    if (i == num - 1) return stop;
    return __dadd_rn(__dmul_rn((double)i, __ddiv_rn(__dsub_rn(stop, start), (double)(num - 1))), start);
}

// radar_processor's own PPI collapse (reference processor.py:513-528, utils.py:369-378): the level closest to
// z_target = r sin(el) + r^2 / (2 * 8.49e6), r = sqrt(X^2 + Y^2) on the grid's float64 axes; argmin keeps the first
// minimum; no out-of-grid masking.
__device__ __forceinline__ int closest_level(const ProductParams& pp, const SliceParams& s, int64_t col)
{
    const int iy = (int)(col / pp.nx), ix = (int)(col - (int64_t)iy * pp.nx);
    const double X = linspace_f64(pp.x_min, pp.x_max, pp.nx, ix), Y = linspace_f64(pp.y_min, pp.y_max, pp.ny, iy);
    const double r = __dsqrt_rn(__dadd_rn(__dmul_rn(X, X), __dmul_rn(Y, Y)));
    const double zt = __dadd_rn(__dmul_rn(r, s.sin_e), __ddiv_rn(__dmul_rn(r, r), 2.0 * 8.49e6));
    int best = 0;
    double best_d = fabs(__dsub_rn(zt, linspace_f64(pp.z_min, pp.z_max, pp.nz_full, 0)));
    for (int k = 1; k < pp.nz_full; ++k) {
        const double d = fabs(__dsub_rn(zt, linspace_f64(pp.z_min, pp.z_max, pp.nz_full, k)));
        if (d < best_d) { best_d = d; best = k; }
    }
    return best;
}

struct ColumnState {
    float cmax, cmin, msum;
    int mcnt;
    float s_lo[RG_MAX_SLICES], s_hi[RG_MAX_SLICES];
    int zz[RG_MAX_SLICES];               // captured levels: lo | hi << 16

    __device__ __forceinline__ void init(const ProductParams& pp, float x, float y, int64_t col = 0)
    {
        cmax = cmin = __uint_as_float(kCanonNaN);
        msum = 0.f;
        mcnt = 0;
#pragma unroll
        for (int k = 0; k < RG_MAX_SLICES; ++k) {
            s_lo[k] = s_hi[k] = __uint_as_float(kCanonNaN);
            zz[k] = 0x7FFF7FFF;          // unused slice: a level no grid has (nz <= 32767)
            if (k < pp.n_slices) {
                const SliceParams& s = pp.slices[k];
                int lo, hi;
                if (s.kind == RG_PROD_BEAM && s.mode == 2) {
                    lo = hi = closest_level(pp, s, col);
                } else if (s.kind == RG_PROD_BEAM) {
                    const double tz = beam_target_z(s, x, y);
                    const double zf = __ddiv_rn(__dsub_rn(tz, pp.z_min), pp.z_step);
                    if (s.mode == 1) {           // 'nearest'  products.py:259-264
                        lo = hi = clampi(rint(zf), 0, pp.nz_full - 1);
                    } else {                     // 'linear'   products.py:279-292
                        const double fl = floor(zf);
                        lo = clampi(fl, 0, pp.nz_full - 1);
                        hi = clampi(fl + 1.0, 0, pp.nz_full - 1);
                    }
                } else {
                    lo = s.z_lo;
                    hi = s.z_hi;
                }
                zz[k] = lo | (hi << 16);
            }
        }
    }

    // v is the finished voxel value of GLOBAL level z; levels must arrive in ascending order
    __device__ __forceinline__ void update(const ProductParams& pp, int z, float v)
    {
        const bool ok = !isnan(v);
        // z0 <= z <= z1 as one unsigned compare against the range width (0 when the product is off)
        if ((unsigned)(z - pp.cmax_z0) < pp.cmax_w && ok) cmax = isnan(cmax) ? v : fmaxf(cmax, v);
        if ((unsigned)(z - pp.cmin_z0) < pp.cmin_w && ok) cmin = isnan(cmin) ? v : fminf(cmin, v);
        if ((unsigned)(z - pp.cmean_z0) < pp.cmean_w) {
            // np.nanmean: NaN -> 0, sequential float32 adds along z, count of non-NaN
            msum = __fadd_rn(msum, ok ? v : 0.f);
            mcnt += ok ? 1 : 0;
        }
#pragma unroll
        for (int k = 0; k < RG_MAX_SLICES; ++k) {
            if (z == (zz[k] & 0xFFFF)) s_lo[k] = v;
            if (z == (zz[k] >> 16)) s_hi[k] = v;
        }
    }

    // Per-(field, column) words kept in shared memory by the thread-per-column kernel: word w of field f of
    // thread t lives at sm[(w * n_fields + f) * blockDim.x + t].  zz[] is per column and stays in registers.
    __device__ __forceinline__ void store_words(const ProductParams& pp, float* sm, int f, int n_fields) const
    {
        const int stride = blockDim.x, o = f * stride + threadIdx.x;
        if (pp.slot_cmax >= 0) sm[pp.slot_cmax * n_fields * stride + o] = cmax;
        if (pp.slot_cmin >= 0) sm[pp.slot_cmin * n_fields * stride + o] = cmin;
        if (pp.slot_cmean >= 0) {
            sm[pp.slot_cmean * n_fields * stride + o] = msum;
            sm[(pp.slot_cmean + 1) * n_fields * stride + o] = __int_as_float(mcnt);
        }
#pragma unroll
        for (int k = 0; k < RG_MAX_SLICES; ++k) {
            if (pp.slot_slice[k] >= 0) {
                sm[pp.slot_slice[k] * n_fields * stride + o] = s_lo[k];
                sm[(pp.slot_slice[k] + 1) * n_fields * stride + o] = s_hi[k];
            }
        }
    }

    // Update the shared-memory state of (field f, this thread) with the finished value v of GLOBAL level z.
    // Same arithmetic as update(), driven by the op list so that only requested products cost anything.
    static __device__ __forceinline__ void update_words(const ProductParams& pp, float* sm, int f, int n_fields, int z, float v)
    {
        const int stride = blockDim.x, o = f * stride + threadIdx.x;
        const bool ok = !isnan(v);
        for (int i = 0; i < pp.n_ops; ++i) {
            const ProductParams::Op& op = pp.ops[i];
            float* w0 = sm + op.slot * n_fields * stride + o;
            switch (op.kind) {
                case 1:
                    if ((unsigned)(z - op.z0) < op.w && ok) { const float c = *w0; *w0 = isnan(c) ? v : fmaxf(c, v); }
                    break;
                case 2:
                    if ((unsigned)(z - op.z0) < op.w && ok) { const float c = *w0; *w0 = isnan(c) ? v : fminf(c, v); }
                    break;
                case 3:
                    if ((unsigned)(z - op.z0) < op.w) {
                        float* w1 = w0 + n_fields * stride;
                        *w0 = __fadd_rn(*w0, ok ? v : 0.f);
                        *w1 = __int_as_float(__float_as_int(*w1) + (ok ? 1 : 0));
                    }
                    break;
                case 4:                                   // warp-uniform levels: nothing to do on most levels
                    if (z == op.z0) *w0 = v;
                    if (z == op.z1) w0[n_fields * stride] = v;
                    break;
                default: {                                // per-column levels, parked next to the state words
                    const int zz = __float_as_int(sm[(pp.n_state_words * n_fields + op.k) * stride + threadIdx.x]);
                    if (z == (zz & 0xFFFF)) *w0 = v;
                    if (z == (zz >> 16)) w0[n_fields * stride] = v;
                    break;
                }
            }
        }
    }

    // The same for the NK (field, column) states one lane finishes at a level, ops in the OUTER loop: the op's parameters are
    // read once and the NK shared-memory updates of an op are independent of each other.  (One call of update_words per
    // state put 3 x n_ops dependent constant-bank reads, address computations and shared-memory round trips in sequence:
    // a request with a PPI cost 1.44 instead of 0.65 ms at cfg3.)
    template <int NK>
    static __device__ __forceinline__ void update_words_n(const ProductParams& pp, float* sm, int n_fields, int z, const float (&v)[NK],
                                                          const bool (&on)[NK])
    {
        const int stride = blockDim.x;
        for (int i = 0; i < pp.n_ops; ++i) {
            const ProductParams::Op& op = pp.ops[i];
            const int kind = op.kind, z0 = op.z0, z1 = op.z1;
            float* const w0 = sm + op.slot * n_fields * stride + threadIdx.x;          // state word of field 0 of this lane
            if (kind <= 3) {
                if ((unsigned)(z - z0) < op.w) {                                      // warp-uniform
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        if (!on[k]) continue;
                        float* const w = w0 + k * stride;
                        const bool ok = !isnan(v[k]);
                        if (kind == 1) {
                            if (ok) { const float c = *w; *w = isnan(c) ? v[k] : fmaxf(c, v[k]); }
                        } else if (kind == 2) {
                            if (ok) { const float c = *w; *w = isnan(c) ? v[k] : fminf(c, v[k]); }
                        } else {
                            float* const w1 = w + n_fields * stride;
                            *w = __fadd_rn(*w, ok ? v[k] : 0.f);
                            *w1 = __int_as_float(__float_as_int(*w1) + (ok ? 1 : 0));
                        }
                    }
                }
            } else if (kind == 4) {                                                   // warp-uniform levels: nothing to do on most levels
                if (z == z0 || z == z1) {
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        if (!on[k]) continue;
                        if (z == z0) w0[k * stride] = v[k];
                        if (z == z1) w0[(k + n_fields) * stride] = v[k];
                    }
                }
            } else {                                                                  // per-column levels, parked next to the state words
                const int zz = __float_as_int(sm[(pp.n_state_words * n_fields + op.k) * stride + threadIdx.x]);
                const bool lo = z == (zz & 0xFFFF), hi = z == (zz >> 16);
                if (lo || hi) {
#pragma unroll
                    for (int k = 0; k < NK; ++k) {
                        if (!on[k]) continue;
                        if (lo) w0[k * stride] = v[k];
                        if (hi) w0[(k + n_fields) * stride] = v[k];
                    }
                }
            }
        }
    }

    // the captured level pair of every slice is per column; kernels that keep no ColumnState in registers park it too
    __device__ __forceinline__ void store_levels(const ProductParams& pp, float* sm, int n_fields) const
    {
        const int stride = blockDim.x;
#pragma unroll
        for (int k = 0; k < RG_MAX_SLICES; ++k)
            if (pp.slot_slice[k] >= 0) sm[(pp.n_state_words * n_fields + k) * stride + threadIdx.x] = __int_as_float(zz[k]);
    }

    __device__ __forceinline__ void load_levels(const ProductParams& pp, const float* sm, int n_fields)
    {
        const int stride = blockDim.x;
#pragma unroll
        for (int k = 0; k < RG_MAX_SLICES; ++k)
            zz[k] = pp.slot_slice[k] >= 0 ? __float_as_int(sm[(pp.n_state_words * n_fields + k) * stride + threadIdx.x]) : 0x7FFF7FFF;
    }

    __device__ __forceinline__ void load_words(const ProductParams& pp, const float* sm, int f, int n_fields)
    {
        const int stride = blockDim.x, o = f * stride + threadIdx.x;
        if (pp.slot_cmax >= 0) cmax = sm[pp.slot_cmax * n_fields * stride + o];
        if (pp.slot_cmin >= 0) cmin = sm[pp.slot_cmin * n_fields * stride + o];
        if (pp.slot_cmean >= 0) {
            msum = sm[pp.slot_cmean * n_fields * stride + o];
            mcnt = __float_as_int(sm[(pp.slot_cmean + 1) * n_fields * stride + o]);
        }
#pragma unroll
        for (int k = 0; k < RG_MAX_SLICES; ++k) {
            if (pp.slot_slice[k] >= 0) {
                s_lo[k] = sm[pp.slot_slice[k] * n_fields * stride + o];
                s_hi[k] = sm[(pp.slot_slice[k] + 1) * n_fields * stride + o];
            }
        }
    }

    // planes are [field][ncol]
    __device__ __forceinline__ void write(const ProductParams& pp, int field, int64_t col, int64_t ncol, float x,
                                          float y) const
    {
        const size_t o = (size_t)field * (size_t)ncol + (size_t)col;
        const float qnan = __uint_as_float(kCanonNaN);
        // partial (z-slab) planes carry "no data in this slab" as the neutral element of the collective that follows:
        // -inf / +inf for all-reduce(MAX / MIN), -0.0 for all-reduce(SUM) (x + -0.0 == x for every x, signed zeros included)
        auto put = [&](void* out, int image, auto val) {
            using T = decltype(val);
            if (out != nullptr) reinterpret_cast<T*>(out)[o] = val;
            if (image >= 0) emit_image<T>(pp.images[image], o, val);
        };
        if (pp.cmax_on) put(pp.cmax_out, pp.cmax_image, pp.cmax_partial && isnan(cmax) ? __uint_as_float(0xFF800000u) : cmax);
        if (pp.cmin_on) put(pp.cmin_out, pp.cmin_image, pp.cmin_partial && isnan(cmin) ? __uint_as_float(0x7F800000u) : cmin);
        if (pp.cmean_on) {
            // _divide_by_count: true_divide(float32 sum, intp count) evaluates in float64, stored as float32
            put(pp.cmean_out, pp.cmean_image, (float)__ddiv_rn((double)msum, (double)mcnt));
        }
#pragma unroll
        for (int k = 0; k < RG_MAX_SLICES; ++k) {
            if (k < pp.n_slices) {
                const SliceParams& s = pp.slices[k];
                const int lo = zz[k] & 0xFFFF, hi = zz[k] >> 16;
                const bool own_lo = !s.partial || (lo >= pp.own_z0 && lo < pp.own_z1);
                const bool own_hi = !s.partial || (hi >= pp.own_z0 && hi < pp.own_z1);
                if (s.kind == RG_PROD_BEAM && s.mode == 2) {
                    put(s.out, s.image, own_lo ? s_lo[k] : -0.0f);
                } else if (s.kind == RG_PROD_BEAM) {
                    const double tz = beam_target_z(s, x, y);
                    const double zf = __ddiv_rn(__dsub_rn(tz, pp.z_min), pp.z_step);
                    if (s.mode == 1) {
                        const double zi = rint(zf);
                        const bool ok = zi >= 0.0 && zi < (double)pp.nz_full;
                        put(s.out, s.image, !ok ? qnan : own_lo ? s_lo[k] : -0.0f);                  // products.py:263-272
                    } else {
                        const double w_hi = __dsub_rn(zf, floor(zf));                      // products.py:287-288
                        const double w_lo = __dsub_rn(1.0, w_hi);
                        const double t_lo = own_lo ? __dmul_rn(w_lo, (double)s_lo[k]) : -0.0;
                        const double t_hi = own_hi ? __dmul_rn(w_hi, (double)s_hi[k]) : -0.0;
                        double r = __dadd_rn(t_lo, t_hi);
                        if (tz < pp.z_min || tz > pp.z_max) r = (double)qnan;              // products.py:306-309
                        put(s.out, s.image, r);
                    }
                } else if (s.mode == RG_BLEND_PICK) {
                    put(s.out, s.image, own_lo ? s_lo[k] : -0.0f);
                } else if (s.mode == RG_BLEND_F32) {
                    const float t_lo = own_lo ? __fmul_rn((float)s.w_lo, s_lo[k]) : -0.0f;
                    const float t_hi = own_hi ? __fmul_rn((float)s.w_hi, s_hi[k]) : -0.0f;
                    put(s.out, s.image, __fadd_rn(t_lo, t_hi));
                } else {
                    const double t_lo = own_lo ? __dmul_rn(s.w_lo, (double)s_lo[k]) : -0.0;
                    const double t_hi = own_hi ? __dmul_rn(s.w_hi, (double)s_hi[k]) : -0.0;
                    const double r = __dadd_rn(t_lo, t_hi);
                    // a partial float64 blend stays float64 until the ranks' terms have been added
                    if (s.mode == RG_BLEND_F64_OUT64 || s.partial) put(s.out, s.image, r);
                    else put(s.out, s.image, (float)r);
                }
            }
        }
    }
};

#ifndef RG_FASTDIV
#define RG_FASTDIV 1           // fast path: a * rcp(b) refined once (<= 1 ulp off IEEE) instead of the IEEE division sequence
#endif
__device__ __forceinline__ float fast_div(float a, float b)
{
#if RG_FASTDIV && !defined(RG_EMU)
    // b is a positive, finite sum of weights here: hardware reciprocal (1 ulp), then one residual correction of
    // the quotient -- within 1 ulp of the IEEE quotient, far inside the 1e-5 relative bar of the fast path
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b));
#if RG_FASTDIV == 2
    r = fmaf(fmaf(-b, r, 1.0f), r, r);
#endif
    const float q = a * r;
    // an infinite sum (an UNMASKED +-inf value, interpolate.py:78-82) stays infinite: the residual would be inf - inf
    return fabsf(q) <= 3.402823466e+38f ? fmaf(fmaf(-b, q, a), r, q) : q;
#else
    return __fdiv_rn(a, b);
#endif
}

__device__ __forceinline__ void prefetch_l2(const void* ptr)
{
#ifndef RG_EMU
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
#else
    (void)ptr;
#endif
}

template <int N>
__device__ __forceinline__ void load_vec(const float* __restrict__ base, uint32_t gate, float* v)
{
    if constexpr (N == 1) {
        v[0] = __ldg(base + gate);
    } else if constexpr (N == 2) {
        const float2 t = __ldg(reinterpret_cast<const float2*>(base) + gate);
        v[0] = t.x; v[1] = t.y;
    } else if constexpr (N == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(base) + gate);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
}

#if RG_TEX
template <int N>
__device__ __forceinline__ void tex_vec(cudaTextureObject_t t, uint32_t gate, float* v)
{
    if constexpr (N == 1) {
        v[0] = tex1Dfetch<float>(t, (int)gate);
    } else if constexpr (N == 2) {
        const float2 q = tex1Dfetch<float2>(t, (int)gate);
        v[0] = q.x; v[1] = q.y;
    } else {
        const float4 q = tex1Dfetch<float4>(t, (int)gate);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    }
}
#endif

template <int F>
__device__ __forceinline__ void load_record(const RecSrc& r, uint32_t gate, float (&v)[Layout<F>::NV])
{
#if RG_TEX == 1
    tex_vec<Layout<F>::FA>(r.tex_a, gate, v);
    if constexpr (Layout<F>::FB > 0) tex_vec<Layout<F>::FB>(r.tex_b, gate, v + Layout<F>::FA);
#elif RG_TEX == 2                                            // array A through LSU, the narrow array B through TEX
    load_vec<Layout<F>::FA>(r.a, gate, v);
    if constexpr (Layout<F>::FB > 0) tex_vec<Layout<F>::FB>(r.tex_b, gate, v + Layout<F>::FA);
#elif RG_TEX == 3                                            // the other way round
    tex_vec<Layout<F>::FA>(r.tex_a, gate, v);
    if constexpr (Layout<F>::FB > 0) load_vec<Layout<F>::FB>(r.b, gate, v + Layout<F>::FA);
#else
    load_vec<Layout<F>::FA>(r.a, gate, v);
    if constexpr (Layout<F>::FB > 0) load_vec<Layout<F>::FB>(r.b, gate, v + Layout<F>::FA);
#endif
}

// (a0, a1) += w * (v0, v1): one packed FFMA2 (fma.rn.f32x2, sm_100+) when the operands sit in register pairs, which
// the 256-bit record load and the accumulator arrays give for free
__device__ __forceinline__ void fma2(float& a0, float& a1, float w, float v0, float v1)
{
#ifndef RG_EMU
    unsigned long long acc, vv, ww;
    asm("mov.b64 %0, {%1,%2};" : "=l"(acc) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(vv) : "f"(v0), "f"(v1));
    asm("mov.b64 %0, {%1,%1};" : "=l"(ww) : "f"(w));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ww), "l"(vv));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc));
#else
    a0 = fmaf(w, v0, a0);
    a1 = fmaf(w, v1, a1);
#endif
}

}  // namespace rg
