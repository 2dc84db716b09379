// K4 (gate-mask fusion + record packing), K5 (CSR gather-weighted mean) and K6 (COLMAX / CAPPI / PPI,
// fused as the epilogue of K5 or run stand-alone on an existing 3-D grid).
//
// Reference arithmetic being replaced (paths relative to the reference repo):
//   src/radar_grid/interpolate.py:59-104   apply_geometry
//   src/radar_grid/filters.py:114-211      GateFilter.exclude_below / exclude_above / exclude_outside
//   src/radar_grid/products.py:168-314     constant_elevation_ppi (+ compute_beam_height :70-89, _flat :164-165)
//   src/radar_grid/products.py:317-415     constant_altitude_ppi
//   src/radar_grid/products.py:420-580     column_max / column_min / column_mean
//
// Data layout in HBM
//   pairs   uint2[P]        {gate id, float32 weight bits} - one 8-byte stream, read exactly once per apply; every
//                           row is sorted by gate id, i.e. it is a few runs of consecutive bins of the same ray
//   indptr  uint32[V+1]     row v = (lz*ny + iy)*nx + ix
//   records A float[G+1][FA], B float[G+1][FB]   the F field values of a gate: fields 0..3 in A, 4..7 in B; a masked
//                           value is the bit pattern kMaskedBits; record G is all-masked (target of idle lanes).
//                           The records are small (20 B * G for five fields) and re-read ~P/G times: they live in
//                           L1/L2 while the pair stream goes past them with evict-first loads.
// Mapping (apply_columns_kernel)
//   A CTA owns a small 2-D patch of voxel columns; a group of W lanes owns one column and walks its levels bottom-up.
//   Lane j of the group takes pairs j, j+W, ... of the row, so neighbouring lanes gather neighbouring gate records.
//   Because a group sees every level of its column, the column products (running max, the two levels of a CAPPI /
//   PPI blend) are kept on chip and written once: a products-only request never writes the 3-D grid.
//   DESIGN.md section 6 has the measured history of every decision in this file.

#include <math.h>

#include <type_traits>

#include "rg_internal.cuh"
#include "rg_device.cuh"

namespace rg {

// The kernels of this file are instantiated per field count (1..8); build.py compiles the file twice, in parallel:
// RG_PART 1 = the common code and field counts 1..4, RG_PART 2 = field counts 5..8 only.  0 (the emulator build) = all.
#ifndef RG_PART
#define RG_PART 0
#endif
#define RG_LO (RG_PART != 2)
#define RG_HI (RG_PART != 1)
void launch_pack_hi(Context* ctx, const PackParams& p, bool vec);
void launch_heavy_hi(Context* ctx, const ApplyParams& p);
int launch_columns_hi(Context* ctx, const ApplyParams& p, int W);
int launch_sell_hi(Context* ctx, const ApplyParams& p);

// ------------------------------------------------------------------------------------------------------
// K4  pack: gate masks (field mask | masked_invalid | fused QC range rules) + AoS records
// ------------------------------------------------------------------------------------------------------
// One gate: field mask | masked_invalid | fused range rules -> the gate's record values (see Layout).
template <int F>
__device__ __forceinline__ void pack_gate(const PackParams& p, bool null_gate, uint32_t excluded, const float (&val)[F],
                                          const uint32_t (&msk)[F], float (&out)[Layout<F>::NV])
{
    using L = Layout<F>;
    uint32_t mask_bits = 0;
#pragma unroll
    for (int f = 0; f < L::NV; ++f) out[f] = __uint_as_float(kMaskedBits);
#pragma unroll
    for (int f = 0; f < F; ++f) {
        uint32_t bits = kMaskedBits;
        if (!null_gate) {
            const float v = val[f];
            bool masked = ((excluded >> f) & 1u) || msk[f] != 0;
            if ((p.invalid_bits >> f) & 1u) masked |= !isfinite(v);          // np.ma.masked_invalid
            bits = masked ? kMaskedBits : (isnan(v) ? kCanonNaN : __float_as_uint(v));
            if (!masked && !isfinite(v) && p.nonfinite != nullptr) *p.nonfinite = p.epoch;   // rare; every writer stores the same word
        }
        if constexpr (L::MB) {
            const bool m = bits == kMaskedBits;
            mask_bits |= m ? (1u << (f + L::SH)) : 0u;
            out[f] = m ? 0.f : __uint_as_float(bits);
        } else {
            out[f] = __uint_as_float(bits);
        }
    }
    if constexpr (L::MB) out[F] = __uint_as_float(mask_bits);
}

__device__ __forceinline__ uint32_t rule_hits(const PackParams& p, int r, float q)
{
    // one exclusion bit per field from the fused range rules (filters.py:133-134, 156-157, 208-209)
    const bool hit = (p.rule_use_lo[r] && q < p.rule_lo[r]) || (p.rule_use_hi[r] && q > p.rule_hi[r]);
    return hit ? p.rule_bits[r] : 0u;
}

// Scalar form: one thread per gate, any pointer alignment.  Also writes the all-masked record n_gates.
template <int F>
__global__ void __launch_bounds__(256) pack_records_kernel(const __grid_constant__ PackParams p, int64_t first)
{
    using L = Layout<F>;
    constexpr int FA = L::FA, FB = L::FB, NV = L::NV;
    const int64_t g = first + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > p.n_gates) return;
    const bool null_gate = g == p.n_gates;       // record n_gates is all-masked: a harmless target for idle lanes
    uint32_t excluded = 0;
    for (int r = 0; r < p.n_rules && !null_gate; ++r) excluded |= rule_hits(p, r, __ldg(p.rule_values[r] + g));
    float val[F];
    uint32_t msk[F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        val[f] = null_gate ? 0.f : __ldg(p.fields[f] + g);
        msk[f] = (!null_gate && p.masks[f] != nullptr) ? __ldg(p.masks[f] + g) : 0u;
    }
    float out[NV];
    pack_gate<F>(p, null_gate, excluded, val, msk, out);
    auto store = [](float* dst, const float* v, auto n) {
        constexpr int N = decltype(n)::value;
        if constexpr (N == 1) dst[0] = v[0];
        else if constexpr (N == 2) *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1]);
        else *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
    };
    store(p.records + (size_t)g * FA, out, std::integral_constant<int, FA>{});
    if constexpr (FB > 0) store(p.records_b + (size_t)g * FB, out + FA, std::integral_constant<int, FB>{});
}

// Vector form: four consecutive gates per thread -- 128-bit loads of every field (32-bit of every mask), and the four
// records leave as whole 128-bit stores (array B of a five-field pass: ONE store for the four gates).  HBM-bound
// streaming: 4 F G bytes in, 4 (FA + FB) G out.  Needs 16-byte aligned field / rule pointers and 4-byte aligned masks
// (launch_pack checks and otherwise takes the scalar kernel); the last G mod 4 gates and the all-masked record are
// written by a scalar launch.
template <int F>
__global__ void __launch_bounds__(256) pack_records4_kernel(const __grid_constant__ PackParams p, int64_t n_quads)
{
    using L = Layout<F>;
    constexpr int FA = L::FA, FB = L::FB, NV = L::NV;
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_quads) return;
    uint32_t excluded[4] = {0u, 0u, 0u, 0u};
    for (int r = 0; r < p.n_rules; ++r) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.rule_values[r]) + q);
        excluded[0] |= rule_hits(p, r, t.x); excluded[1] |= rule_hits(p, r, t.y);
        excluded[2] |= rule_hits(p, r, t.z); excluded[3] |= rule_hits(p, r, t.w);
    }
    float val[4][F];
    uint32_t msk[4][F];
#pragma unroll
    for (int f = 0; f < F; ++f) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p.fields[f]) + q);
        val[0][f] = t.x; val[1][f] = t.y; val[2][f] = t.z; val[3][f] = t.w;
        const uint32_t m = p.masks[f] != nullptr ? __ldg(reinterpret_cast<const uint32_t*>(p.masks[f]) + q) : 0u;
        msk[0][f] = m & 0xFFu; msk[1][f] = (m >> 8) & 0xFFu; msk[2][f] = (m >> 16) & 0xFFu; msk[3][f] = m >> 24;
    }
    float a[4 * FA], b[FB > 0 ? 4 * FB : 1];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        float out[NV];
        pack_gate<F>(p, false, excluded[k], val[k], msk[k], out);
#pragma unroll
        for (int i = 0; i < FA; ++i) a[k * FA + i] = out[i];
#pragma unroll
        for (int i = 0; i < FB; ++i) b[k * FB + i] = out[FA + i];
    }
    float4* da = reinterpret_cast<float4*>(p.records + (size_t)q * 4 * FA);
#pragma unroll
    for (int i = 0; i < FA; ++i) da[i] = make_float4(a[4 * i], a[4 * i + 1], a[4 * i + 2], a[4 * i + 3]);
    if constexpr (FB > 0) {
        float4* db = reinterpret_cast<float4*>(p.records_b + (size_t)q * 4 * FB);
#pragma unroll
        for (int i = 0; i < FB; ++i) db[i] = make_float4(b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]);
    }
}

#if RG_LO
// (Re)create the texture objects over the record arrays when the buffer or the field count changed.
int bind_record_textures(Context* ctx, const float* rec_a, const float* rec_b, int n_fields, int64_t n_gates)
{
#if RG_TEX && !defined(RG_EMU)
    if (ctx->tex_a_ptr == rec_a && ctx->tex_b_ptr == rec_b && ctx->tex_fields == n_fields && ctx->tex_gates == n_gates) return RG_OK;
    if (ctx->tex_a) { cudaDestroyTextureObject(ctx->tex_a); ctx->tex_a = 0; }
    if (ctx->tex_b) { cudaDestroyTextureObject(ctx->tex_b); ctx->tex_b = 0; }
    const int fa = n_fields == 1 ? 1 : n_fields == 2 ? 2 : 4;
    const int fb = n_fields <= 4 ? 0 : n_fields == 5 ? (RG_MASKBITS ? 2 : 1) : n_fields == 6 ? 2 : 4;
    auto make = [&](const float* ptr, int nf, unsigned long long* out) -> int {
        cudaResourceDesc rd{};
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = const_cast<float*>(ptr);
        rd.res.linear.desc = cudaCreateChannelDesc(32, nf >= 2 ? 32 : 0, nf >= 4 ? 32 : 0, nf >= 4 ? 32 : 0, cudaChannelFormatKindFloat);
        rd.res.linear.sizeInBytes = (size_t)(n_gates + 1) * nf * sizeof(float);
        cudaTextureDesc td{};
        td.readMode = cudaReadModeElementType;
        cudaTextureObject_t t = 0;
        RG_CUDA(cudaCreateTextureObject(&t, &rd, &td, nullptr));
        *out = (unsigned long long)t;
        return RG_OK;
    };
    RG_TRY(make(rec_a, fa, &ctx->tex_a));
    if (fb > 0) RG_TRY(make(rec_b, fb, &ctx->tex_b));
    ctx->tex_a_ptr = rec_a; ctx->tex_b_ptr = rec_b; ctx->tex_fields = n_fields; ctx->tex_gates = n_gates;
#else
    (void)ctx; (void)rec_a; (void)rec_b; (void)n_fields; (void)n_gates;
#endif
    return RG_OK;
}

int records_width(int n_fields) { return n_fields <= 1 ? 1 : n_fields == 2 ? 2 : n_fields <= 4 ? 4 : 8; }

size_t records_b_offset(int n_fields, int64_t n_gates)
{
    const int fa = n_fields == 1 ? 1 : n_fields == 2 ? 2 : 4;
    return (((size_t)(n_gates + 1) * fa * sizeof(float)) + 511) & ~(size_t)511;   // texture-bindable
}

#endif  // RG_LO

#ifndef RG_PACK4
#define RG_PACK4 0             // 1: vector pack kernel (four gates per thread) when the pointers allow it; measured 46 vs 42 us for the
                               //    one-gate-per-thread kernel at cfg3 (its loads are already coalesced and all issued before the first use)
#endif

template <int F>
static void launch_pack_f(Context* ctx, const PackParams& p, bool vec)
{
    int64_t first = 0;
    if (vec && p.n_gates >= 4) {
        const int64_t n_quads = p.n_gates / 4;
        pack_records4_kernel<F><<<(unsigned)((n_quads + 255) / 256), 256, 0, ctx->stream>>>(p, n_quads);
        ctx->launches++;
        first = n_quads * 4;
    }
    const int64_t rest = p.n_gates + 1 - first;              // the tail and the all-masked record n_gates
    pack_records_kernel<F><<<(unsigned)((rest + 255) / 256), 256, 0, ctx->stream>>>(p, first);
    ctx->launches++;
}

#if RG_HI
void launch_pack_hi(Context* ctx, const PackParams& p, bool vec)
{
    switch (p.n_fields) {
        case 5: launch_pack_f<5>(ctx, p, vec); break;
        case 6: launch_pack_f<6>(ctx, p, vec); break;
        case 7: launch_pack_f<7>(ctx, p, vec); break;
        default: launch_pack_f<8>(ctx, p, vec); break;
    }
}
#endif

#if RG_LO
int launch_pack(Context* ctx, const PackParams& p)
{
    bool vec = RG_PACK4 != 0;
    for (int f = 0; f < p.n_fields; ++f)
        vec = vec && ((uintptr_t)p.fields[f] % 16 == 0) && (p.masks[f] == nullptr || (uintptr_t)p.masks[f] % 4 == 0);
    for (int r = 0; r < p.n_rules; ++r) vec = vec && ((uintptr_t)p.rule_values[r] % 16 == 0);
    timer_begin(ctx, kTimerPack);
    switch (p.n_fields) {
        case 1: launch_pack_f<1>(ctx, p, vec); break;
        case 2: launch_pack_f<2>(ctx, p, vec); break;
        case 3: launch_pack_f<3>(ctx, p, vec); break;
        case 4: launch_pack_f<4>(ctx, p, vec); break;
        default: launch_pack_hi(ctx, p, vec); break;
    }
    timer_end(ctx, kTimerPack);
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}
#endif  // RG_LO

#if RG_LO
// Stand-alone products over existing grids: one thread per (column, field), coalesced along x.
struct ProductsKernelParams {
    const float* grids[RG_MAX_FIELDS];
    int64_t ncol;
    int32_t nx, n_levels, z_begin, n_fields;
    ProductParams prod;
};

__global__ void __launch_bounds__(256) products_kernel(const __grid_constant__ ProductsKernelParams p)
{
    const int64_t col = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int field = blockIdx.y;
    if (col >= p.ncol) return;
    const float x = __ldg(p.prod.x_ax + (int)(col % p.nx));
    const float y = __ldg(p.prod.y_ax + (int)(col / p.nx));
    ColumnState st;
    st.init(p.prod, x, y, col);
    const float* __restrict__ grid = p.grids[field];
    for (int lz = 0; lz < p.n_levels; ++lz) {
        const float v = __ldcs(grid + (size_t)lz * (size_t)p.ncol + (size_t)col);
        st.update(p.prod, p.z_begin + lz, v);
    }
    st.write(p.prod, field, col, p.ncol, x, y);
}

int launch_products(Context* ctx, const rg_grid_spec& grid, int n_fields, const float* const* grids_dev,
                    const ProductParams& prod)
{
    ProductsKernelParams kp{};
    for (int f = 0; f < n_fields; ++f) kp.grids[f] = grids_dev[f];
    kp.ncol = (int64_t)grid.ny * grid.nx;
    kp.nx = grid.nx;
    kp.n_levels = grid.z_end - grid.z_begin;
    kp.z_begin = grid.z_begin;
    kp.n_fields = n_fields;
    kp.prod = prod;
    if (kp.ncol == 0 || n_fields == 0) return RG_OK;
    dim3 blocks((unsigned)((kp.ncol + 255) / 256), (unsigned)n_fields);
    products_kernel<<<blocks, 256, 0, ctx->stream>>>(kp);
    ctx->launches++;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

#endif  // RG_LO

// ------------------------------------------------------------------------------------------------------
// K5  fast path: column-tile CSR gather, W lanes per column, products in the epilogue
//
// Gate records are split in two arrays, A = float[G][FA] (fields 0..3) and B = float[G][FB] (fields 4..7):
// 20 bytes per gate for five fields, every float is a real field, and eight consecutive gates fill exactly one
// 128-byte line of A.  (An interleaved 32-byte record and a two-lanes-per-record scheme were measured and lost,
// see DESIGN.md section 6.)  The build-time switches below exist so that each design decision can be re-measured.
// ------------------------------------------------------------------------------------------------------
// Reduce-scatter of F (sum_wv, sum_w) units over the 8 lanes of a group: three halving exchanges
// (4 + 2 + 1 units) instead of a full butterfly per value; lane g ends up with the totals of field g.
template <int F>
__device__ __forceinline__ void group8_reduce_scatter(const float (&swv)[F], const float (&sw)[F], int gl, float& a, float& b)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    float u_wv[8], u_w[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) { u_wv[f] = f < F ? swv[f] : 0.f; u_w[f] = f < F ? sw[f] : 0.f; }
    float t_wv[4], t_w[4];
    {
        const bool up = gl & 4;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (j < F) {                       // fields j and j+4 take part only if they exist
                const float send_wv = up ? u_wv[j] : u_wv[j + 4];
                const float send_w = up ? u_w[j] : u_w[j + 4];
                t_wv[j] = (up ? u_wv[j + 4] : u_wv[j]) + __shfl_xor_sync(kFull, send_wv, 4);
                t_w[j] = (up ? u_w[j + 4] : u_w[j]) + __shfl_xor_sync(kFull, send_w, 4);
            } else {
                t_wv[j] = 0.f; t_w[j] = 0.f;
            }
        }
    }
    float s_wv[2], s_w[2];
    {
        const bool up = gl & 2;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            if (j < F) {
                const float send_wv = up ? t_wv[j] : t_wv[j + 2];
                const float send_w = up ? t_w[j] : t_w[j + 2];
                s_wv[j] = (up ? t_wv[j + 2] : t_wv[j]) + __shfl_xor_sync(kFull, send_wv, 2);
                s_w[j] = (up ? t_w[j + 2] : t_w[j]) + __shfl_xor_sync(kFull, send_w, 2);
            } else {
                s_wv[j] = 0.f; s_w[j] = 0.f;
            }
        }
    }
    const bool up = gl & 1;
    a = (up ? s_wv[1] : s_wv[0]) + __shfl_xor_sync(kFull, up ? s_wv[0] : s_wv[1], 1);
    b = (up ? s_w[1] : s_w[0]) + __shfl_xor_sync(kFull, up ? s_w[0] : s_w[1], 1);
}

// Same idea for groups of 4 lanes: two halving exchanges; lane g ends up with the totals of field g and, when there
// are more than four fields, of field g + 4.
template <int F, int NO>
__device__ __forceinline__ void group4_reduce_scatter(const float (&swv)[F], const float (&sw)[F], int gl, float (&a)[NO], float (&b)[NO])
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    float u_wv[8], u_w[8];
#pragma unroll
    for (int f = 0; f < 8; ++f) { u_wv[f] = f < F ? swv[f] : 0.f; u_w[f] = f < F ? sw[f] : 0.f; }
    // step 1 (lane distance 2): units {0,1,4,5} stay in the lower pair of lanes, {2,3,6,7} in the upper pair
    float t_wv[4], t_w[4];
    {
        const bool up = gl & 2;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int lo = j < 2 ? j : j + 2, hi = lo + 2;       // (0,2) (1,3) (4,6) (5,7)
            if (lo < F && !(F == 5 && lo == 4)) {
                t_wv[j] = (up ? u_wv[hi] : u_wv[lo]) + __shfl_xor_sync(kFull, up ? u_wv[lo] : u_wv[hi], 2);
                t_w[j] = (up ? u_w[hi] : u_w[lo]) + __shfl_xor_sync(kFull, up ? u_w[lo] : u_w[hi], 2);
            } else {
                t_wv[j] = 0.f; t_w[j] = 0.f;
            }
        }
    }
    // step 2 (lane distance 1): even unit of each pair to the even lane
    const bool up = gl & 1;
    a[0] = (up ? t_wv[1] : t_wv[0]) + __shfl_xor_sync(kFull, up ? t_wv[0] : t_wv[1], 1);
    b[0] = (up ? t_w[1] : t_w[0]) + __shfl_xor_sync(kFull, up ? t_w[0] : t_w[1], 1);
    if constexpr (F == 5) {
        // a single unit in the upper half: a plain butterfly (no selects) is cheaper than scattering it
        float x = u_wv[4], y = u_w[4];
        x += __shfl_xor_sync(kFull, x, 2);
        y += __shfl_xor_sync(kFull, y, 2);
        a[1] = x + __shfl_xor_sync(kFull, x, 1);
        b[1] = y + __shfl_xor_sync(kFull, y, 1);
    } else if constexpr (NO > 1) {
        a[1] = (up ? t_wv[3] : t_wv[2]) + __shfl_xor_sync(kFull, up ? t_wv[2] : t_wv[3], 1);
        b[1] = (up ? t_w[3] : t_w[2]) + __shfl_xor_sync(kFull, up ? t_w[2] : t_w[3], 1);
    }
}


template <int F, int NV>
__device__ __forceinline__ void accumulate(float w, const float (&v)[NV], float (&swv)[F], float (&sw)[F])
{
    if constexpr (Layout<F>::MB) {
        // masked values are +0.0, so sum(w*v) needs no predicate (interpolate.py:78-82: masked gates contribute 0 to both
        // sums).  Bit-identical to the predicated form: x + w*0 = x.
#pragma unroll
        for (int k = 0; k + 1 < F; k += 2) fma2(swv[k], swv[k + 1], w, v[k], v[k + 1]);
        swv[F - 1] = fmaf(w, v[F - 1], swv[F - 1]);
        const uint32_t mb = __float_as_uint(v[F]);
#pragma unroll
        for (int f = 0; f < F; ++f)                            // one R2P + a predicated FADD per field
            if (!((mb >> (f + Layout<F>::SH)) & 1u)) sw[f] = __fadd_rn(sw[f], w);
    } else {
#pragma unroll
        for (int f = 0; f < F; ++f) {
            const bool m = __float_as_uint(v[f]) == kMaskedBits;   // interpolate.py:78-79
            if (!m) {                                              // predicated: ISETP + FADD + FFMA per field
                sw[f] = __fadd_rn(sw[f], w);
                swv[f] = fmaf(w, v[f], swv[f]);
            }
        }
    }
}

#ifndef RG_HEADBATCH
#define RG_HEADBATCH 2         // 1: the pair loads of the first 2U-1 slots of a row are issued together; 2: same, and exactly
                               //    as many slots as the longest row of the warp needs (one code path per count)
#endif
#ifndef RG_QSMEM
#define RG_QSMEM 1             // 1: when a lane owns two fields, the COLMAX / level-pick state of the PSIG 2 path lives in shared
                               //    memory and the grid pointers are re-read from the parameter bank at the store (fewer
                               //    registers live across the gathers; measured: 0.82 -> 0.71 ms at five fields, W = 4)
#endif
#ifndef RG_TILE2D
#define RG_TILE2D 1            // CTA = 8 x 4 patch of columns (1) or 32 consecutive columns (0)
#endif
#ifndef RG_TAIL
#define RG_TAIL 1              // 1: the tail of a row is ONE predicated batch (idle slots read the all-masked record)
#endif

// Sum pairs [p, e) with stride `step`, RG_UNROLL pairs (and their gathers) in flight per lane.
template <int F>
__device__ __forceinline__ void gather_run(const uint2* __restrict__ pairs, const RecSrc& rec, uint32_t p, uint32_t e,
                                           uint32_t step, float (&swv)[F], float (&sw)[F])
{
    constexpr int NV = Layout<F>::NV;
    constexpr int U = RG_UNROLL;
    while (p < e && e - p > (U - 1) * step) {
        uint2 a[U];
#pragma unroll
        for (int j = 0; j < U; ++j) a[j] = __ldcs(pairs + p + j * step);
        float v[U][NV];
#pragma unroll
        for (int j = 0; j < U; ++j) load_record<F>(rec, a[j].x, v[j]);
#pragma unroll
        for (int j = 0; j < U; ++j) accumulate<F, NV>(__uint_as_float(a[j].y), v[j], swv, sw);
        p += U * step;
    }
#if RG_TAIL
    if (p < e) {
        // fewer than U pairs left for this lane: still one batch with all its loads in flight together
        // (idle slots read the all-masked record)
        uint2 a[U - 1];
#pragma unroll
        for (int j = 0; j < U - 1; ++j) a[j] = p + j * step < e ? __ldcs(pairs + p + j * step) : make_uint2(rec.null_gate, 0u);
        float v[U - 1][NV];
#pragma unroll
        for (int j = 0; j < U - 1; ++j) load_record<F>(rec, a[j].x, v[j]);
#pragma unroll
        for (int j = 0; j < U - 1; ++j) accumulate<F, NV>(__uint_as_float(a[j].y), v[j], swv, sw);
    }
#else
    while (p < e) {
        const uint2 a0 = __ldcs(pairs + p);
        float v0[NV];
        load_record<F>(rec, a0.x, v0);
        accumulate<F, NV>(__uint_as_float(a0.y), v0, swv, sw);
        p += step;
    }
#endif
}

// Rows far longer than a group is wide (the voxels next to the radar see the first gates of every ray: 3 600 pairs at
// cfg1, 10 800 at cfg3 / cfg5) are cut into chunks of kHeavyChunk pairs and every chunk is reduced by a whole CTA in
// its own small launch, ahead of the column kernel, which then only adds the chunks' partial sums (fixed chunk and
// thread partition: deterministic).  Round 1 summed these rows with one warp inside the column kernel while the rest of
// the patch waited: 0.33 of the roofline at cfg1, 0.09 on the lowest z-slab of cfg5.
template <int F>
__global__ void __launch_bounds__(kHeavyThreads) heavy_rows_kernel(const __grid_constant__ ApplyParams p)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    __shared__ float part[kHeavyThreads / 32][2 * F];
#if RG_PDL && !defined(RG_EMU)
    asm volatile("griddepcontrol.launch_dependents;");           // the column kernel may start now; it waits only where it needs us
#endif
    const uint2 ch = __ldg(p.heavy_chunks + blockIdx.x);
    const RecSrc rec{p.records, p.records_b, p.tex_a, p.tex_b, p.null_gate};
    float swv[F], sw[F];
#pragma unroll
    for (int f = 0; f < F; ++f) { swv[f] = 0.f; sw[f] = 0.f; }
    gather_run<F>(p.pairs, rec, ch.x + threadIdx.x, ch.x + ch.y, kHeavyThreads, swv, sw);
#pragma unroll
    for (int f = 0; f < F; ++f) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            swv[f] += __shfl_xor_sync(kFull, swv[f], off);
            sw[f] += __shfl_xor_sync(kFull, sw[f], off);
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
#pragma unroll
        for (int f = 0; f < F; ++f) { part[warp][f] = swv[f]; part[warp][F + f] = sw[f]; }
    }
    __syncthreads();
    if (threadIdx.x < 2 * F) {
        float t = part[0][threadIdx.x];
#pragma unroll
        for (int w = 1; w < kHeavyThreads / 32; ++w) t += part[w][threadIdx.x];
        p.heavy_part[(size_t)blockIdx.x * (2 * F) + threadIdx.x] = t;
    }
}

// PSIG: 0 = no products, 1 = any product list (op list over shared-memory state), 2 = the operational request
// "COLMAX and/or one level pick/blend (CAPPI)" with its three state words in registers.
// IL: the pairs come from the warp-slice copy of the table (one coalesced 256-byte load per slot, no per-lane row
// bounds, the slot count of a level is warp-uniform) instead of the CSR copy.
#ifndef RG_MINBLOCKS_F1
#define RG_MINBLOCKS_F1 RG_MINBLOCKS   // CTAs per SM asked of the ONE-field kernel: 10 (48 registers, no spills, 40 warps) is the candidate
#endif
template <int F, int W, int PSIG, bool IL>
__global__ void __launch_bounds__(kApplyThreads, F == 1 ? RG_MINBLOCKS_F1 : RG_MINBLOCKS) apply_columns_kernel(const __grid_constant__ ApplyParams p)
{
    constexpr bool PROD = PSIG == 1;
    static_assert(!IL || RG_TILE2D == 1, "the warp-slice copy assumes the groups of a warp are adjacent in x");
    // lane gl of a group finishes field gl and, in groups narrower than the field count, field gl + W as well
    constexpr int NO = W < F ? 2 : 1;
    static_assert(W * NO >= F, "every field needs an owner lane");
    constexpr unsigned kFull = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int gl = threadIdx.x & (W - 1);                      // lane within the column group
#if RG_TILE2D
    // A CTA covers a TX x TY patch of columns instead of a 1-D run: the neighbour sets of its columns overlap in
    // both directions, so more of its gate-record gathers hit L1.  Groups of one warp stay adjacent in x.
    constexpr int kGroups = kApplyThreads / W, TX = kGroups >= 8 ? 8 : kGroups, TY = kGroups / TX;
    const int tiles_x = (p.nx + TX - 1) / TX;
    const int g_in_cta = threadIdx.x / W;
    const int cx = (int)(blockIdx.x % tiles_x) * TX + (g_in_cta % TX);
    const int cy = (int)(blockIdx.x / tiles_x) * TY + (g_in_cta / TX);
    const bool col_ok = cx < p.nx && cy < p.ny;
    const int64_t col = col_ok ? (int64_t)cy * p.nx + cx : 0;
#else
    const int64_t col = (int64_t)blockIdx.x * (kApplyThreads / W) + threadIdx.x / W;
    const bool col_ok = col < p.ncol;
    const int cx = 0, cy = 0;                                  // only used by the warp-slice path (2-D tiles)
#endif
    const bool owner = col_ok && gl < F;                       // lane gl finishes field gl

    const uint32_t* __restrict__ indptr = p.indptr;
    const uint2* __restrict__ pairs = p.pairs;
    const RecSrc rec{p.records, p.records_b, p.tex_a, p.tex_b, p.null_gate};

    // PSIG 2: running max and the two captured levels in registers
    constexpr bool QS = RG_QSMEM && NO > 1;    // two fields per owner lane: the registers are needed for the gathers
    // ... in shared memory ([word][k][thread]): touched once per level by the owner lanes, and out of the way of the
    // registers the gathers need
    extern __shared__ float sm_state[];
    float* const q_state = sm_state;           // dynamic shared memory: 3 * NO words per thread (launch_columns)
    if constexpr (PSIG == 2 && QS) {
#pragma unroll
        for (int i = 0; i < 3 * NO; ++i) q_state[i * kApplyThreads + threadIdx.x] = __uint_as_float(kCanonNaN);
    }
    float q_max[NO], q_lo[NO], q_hi[NO];
#pragma unroll
    for (int k = 0; k < NO; ++k) q_max[k] = q_lo[k] = q_hi[k] = __uint_as_float(kCanonNaN);
    // PSIG 1: per-lane product state lives in shared memory: [word][field of the lane][thread]
    if constexpr (PROD) {
        float x = 0.f, y = 0.f;
        if (owner) {
            x = __ldg(p.prod.x_ax + (int)(col % p.nx));
            y = __ldg(p.prod.y_ax + (int)(col / p.nx));
        }
        ColumnState st;
        st.init(p.prod, x, y, col);
#pragma unroll
        for (int k = 0; k < NO; ++k) st.store_words(p.prod, sm_state, k, NO);
        st.store_levels(p.prod, sm_state, NO);
    }

    float* out[NO];
#pragma unroll
    for (int k = 0; k < NO; ++k) out[k] = !QS && owner && gl + k * W < F ? p.grid_out[gl + k * W] : nullptr;
    // Row bounds run two levels ahead of the sums, so that the L2 prefetch of level z+1 (issued while level z is
    // summed) never waits for the bounds load it depends on.
    // IL: (bs, be) are the slice's two quad_ptr words, identical in all lanes of the warp
    const uint2* __restrict__ quads = p.quads;
    const uint32_t* __restrict__ quad_ptr = p.quad_ptr;
    const int qx = cx / (32 / W);
    const bool slice_ok = cy < p.ny && qx < p.quads_x;
    // The bounds words of consecutive levels are a fixed stride apart: walk a pointer instead of re-deriving the index.
    const uint32_t* bp;
    size_t bstride;
    bool b_ok;
    if constexpr (IL) {
        bp = quad_ptr + ((size_t)p.lz_first * (size_t)p.ny + (size_t)cy) * (size_t)p.quads_x + (size_t)qx;
        bstride = (size_t)p.ny * (size_t)p.quads_x;
        b_ok = slice_ok;
    } else {
        bp = indptr + (size_t)p.lz_first * (size_t)p.ncol + (size_t)col;
        bstride = (size_t)p.ncol;
        b_ok = col_ok;
    }
    auto bounds = [&](const uint32_t* ptr, int lz, uint32_t& bs, uint32_t& be) {
        bs = be = 0;
        if (b_ok && lz < p.lz_last) {
            bs = __ldg(ptr);
            be = __ldg(ptr + 1);
        }
    };
    uint32_t s_next, e_next, s_next2, e_next2;
    bounds(bp, p.lz_first, s_next, e_next);
    bounds(bp + bstride, p.lz_first + 1, s_next2, e_next2);
    bp += 2 * bstride;                                         // bounds of the level two ahead of the loop variable
    // Narrow passes over the CSR copy (one or two fields: few registers, short rows, latency-bound) keep the first NP
    // pairs of the NEXT level's row in registers, loaded while the current level is summed: the dependent chain of a
    // level shrinks from pairs -> records to records only.
    constexpr int NP = (!IL && F <= 2) ? RG_PRELOAD : 0;
    uint2 pre[NP > 0 ? NP : 1];
    if constexpr (NP > 0) {
#pragma unroll
        for (int j = 0; j < NP; ++j)
            pre[j] = s_next + gl + j * W < e_next ? __ldcs(pairs + s_next + gl + j * W) : make_uint2(rec.null_gate, 0u);
    }

    constexpr int PH = (IL && RG_ILPRE) ? 2 * RG_UNROLL - 1 : 0;
    uint2 ipre[PH > 0 ? PH : 1];
    auto il_preload = [&](uint32_t sn, uint32_t en) {          // (sn, en): the slice's bounds words of the level to preload
        const uint32_t p0 = sn >> 1, mm = (en >> 1) - p0;
        const uint2* nb = quads + (size_t)p0 * 32 + lane;
#pragma unroll
        for (int j = 0; j < PH; ++j) ipre[j] = (uint32_t)j < mm ? __ldcs(nb + j * 32) : make_uint2(rec.null_gate, 0u);
    };
    if constexpr (PH > 0) il_preload(s_next, e_next);

    size_t row = (size_t)p.lz_first * (size_t)p.ncol + (size_t)col - (size_t)p.ncol;
    for (int lz = p.lz_first; lz < p.lz_last; ++lz) {
        const uint32_t s = s_next, e = e_next;
        row += (size_t)p.ncol;
        s_next = s_next2;
        e_next = e_next2;
        bounds(bp, lz + 2, s_next2, e_next2);
        bp += IL ? (size_t)p.ny * (size_t)p.quads_x : (size_t)p.ncol;
        uint2 cur[NP > 0 ? NP : 1];
        if constexpr (NP > 0) {
#pragma unroll
            for (int j = 0; j < NP; ++j) cur[j] = pre[j];
#pragma unroll
            for (int j = 0; j < NP; ++j)
                pre[j] = s_next + gl + j * W < e_next ? __ldcs(pairs + s_next + gl + j * W) : make_uint2(rec.null_gate, 0u);
        }
        uint2 icur[PH > 0 ? PH : 1];
        if constexpr (PH > 0) {
#pragma unroll
            for (int j = 0; j < PH; ++j) icur[j] = ipre[j];
        }
#if RG_PREFETCH > 0
        if constexpr (IL) {   // level z+1's slots are contiguous: one 128-byte line per lane
            const uint32_t l0 = (s_next >> 1) * 2u + (uint32_t)lane;
            if (l0 < (e_next >> 1) * 2u) prefetch_l2(quads + (size_t)l0 * 16);
        } else {   // pull the pair lines of level z+1 from HBM into L2: one 128-byte line (16 pairs) per lane of the group
            const uint32_t q = s_next + 16u * gl;
            if (q < e_next) prefetch_l2(pairs + q);
        }
#endif

        // IL: m slots for the whole warp, hv = some row of the slice is heavy (kept out of the slice copy)
        const uint32_t il_p0 = s >> 1, il_m = (e >> 1) - il_p0;
        const bool il_hv = s & 1u;
        const uint32_t len = IL ? il_m + (il_hv ? 1u : 0u) : e - s;
        float a[NO], b[NO];
#pragma unroll
        for (int k = 0; k < NO; ++k) a[k] = b[k] = 0.f;
        // warps whose four rows are all empty (outside the radar range, above the highest sweep) skip the sums
        if (IL ? len != 0 : __any_sync(kFull, len != 0)) {     // IL: len is the same in all lanes
            float swv[F], sw[F];
#pragma unroll
            for (int f = 0; f < F; ++f) { swv[f] = 0.f; sw[f] = 0.f; }

            uint32_t cs = s, ce = e;                           // CSR bounds of this lane's row
            if constexpr (IL) {
                cs = ce = 0;
                if (il_hv && col_ok) {
                    cs = __ldg(indptr + row);
                    ce = __ldg(indptr + row + 1);
                }
            }
            bool heavy_mine = ce - cs > kHeavyRow;
            if constexpr (W < 32) {
                // heavy rows were summed by heavy_rows_kernel: the group's first lane adds up the chunks' partial sums
                if (heavy_mine && gl == 0) {
#if RG_PDL && !defined(RG_EMU)
                    asm volatile("griddepcontrol.wait;" ::: "memory");   // heavy_rows_kernel complete, its sums visible
#endif
                    const uint32_t r32 = (uint32_t)row;
                    int lo = 0, hi = p.n_heavy - 1;
                    while (lo < hi) {                              // sorted list of a few hundred rows at most
                        const int mid = (lo + hi) >> 1;
                        if (__ldg(p.heavy_rows + mid) < r32) lo = mid + 1; else hi = mid;
                    }
                    const uint32_t c0 = __ldg(p.heavy_first + lo), c1 = __ldg(p.heavy_first + lo + 1);
                    for (uint32_t c = c0; c < c1; ++c) {
                        const float* hp = p.heavy_part + (size_t)c * (2 * F);
#pragma unroll
                        for (int f = 0; f < F; ++f) { swv[f] += __ldcg(hp + f); sw[f] += __ldcg(hp + F + f); }   // L2: written by the heavy kernel
                    }
                }
            } else {
                heavy_mine = false;
            }

#if RG_HEADBATCH == 2
            {
                // Head of the row, sized exactly: m = the largest number of pairs any lane of the warp has to take
                // (one warp reduction), then ONE code path per m <= 7 that issues all m pair loads together and
                // gathers in chunks of at most U.  No slot is executed that no lane needs.
                constexpr int U = RG_UNROLL, H = 2 * U - 1, NV = Layout<F>::NV;
                const uint32_t lim = heavy_mine ? s : e;               // heavy rows were summed by the whole warp
                uint32_t need = 0, m;
                const uint2* il_base = nullptr;
                if constexpr (IL) {
                    m = il_m;                                          // warp-uniform by construction
                    il_base = quads + (size_t)il_p0 * 32 + lane;
                } else {
                    need = lim > s + gl ? (lim - s - gl + W - 1) / W : 0u;
                    m = __reduce_max_sync(kFull, need);
                }
                auto head = [&](auto mm, auto first_tag) {
                    constexpr int M = decltype(mm)::value;
                    constexpr bool FIRST = decltype(first_tag)::value && PH > 0;     // the batch preloaded during the previous level
                    uint2 hp[M];
#pragma unroll
                    for (int j = 0; j < M; ++j) {
                        if constexpr (IL) hp[j] = FIRST ? icur[j < PH ? j : 0] : __ldcs(il_base + j * 32);
                        else hp[j] = (uint32_t)j < need ? (j < NP ? cur[j < NP ? j : 0] : __ldcs(pairs + s + gl + j * W)) : make_uint2(rec.null_gate, 0u);
                    }
                    constexpr int C0 = M < U ? M : U;
                    {
                        float v[C0][NV];
#pragma unroll
                        for (int j = 0; j < C0; ++j) load_record<F>(rec, hp[j].x, v[j]);
#pragma unroll
                        for (int j = 0; j < C0; ++j) accumulate<F, NV>(__uint_as_float(hp[j].y), v[j], swv, sw);
                    }
                    if constexpr (M > U) {
                        float v[M - U][NV];
#pragma unroll
                        for (int j = 0; j < M - U; ++j) load_record<F>(rec, hp[U + j].x, v[j]);
#pragma unroll
                        for (int j = 0; j < M - U; ++j) accumulate<F, NV>(__uint_as_float(hp[U + j].y), v[j], swv, sw);
                    }
                };
                static_assert(H == 7 || H == 5 || H == 3, "head sizes are written out below");
                bool first = PH > 0;
                if constexpr (IL) {
                    // the slot count is warp-uniform: full batches of H first, then ONE exactly sized batch, so that
                    // long rows never run slots that only hold padding
                    if (PH > 0 && m > (uint32_t)H) {
                        head(std::integral_constant<int, H>{}, std::true_type{});
                        il_base += H * 32;
                        m -= (uint32_t)H;
                        first = false;
                    }
                    while (m > (uint32_t)H) {
                        head(std::integral_constant<int, H>{}, std::false_type{});
                        il_base += H * 32;
                        m -= (uint32_t)H;
                    }
                }
                auto head_k = [&](auto mm) {
                    if constexpr (PH > 0) { if (first) { head(mm, std::true_type{}); return; } }
                    head(mm, std::false_type{});
                };
                switch (m < (uint32_t)H ? m : (uint32_t)H) {
                    case 0: break;
                    case 1: head_k(std::integral_constant<int, 1>{}); break;
                    case 2: head_k(std::integral_constant<int, 2>{}); break;
                    case 3: head_k(std::integral_constant<int, 3>{}); break;
                    case 4: if constexpr (H >= 4) head_k(std::integral_constant<int, 4>{}); break;
                    case 5: if constexpr (H >= 5) head_k(std::integral_constant<int, 5>{}); break;
                    case 6: if constexpr (H >= 6) head_k(std::integral_constant<int, 6>{}); break;
                    default: if constexpr (H >= 7) head_k(std::integral_constant<int, 7>{}); break;
                }
                if constexpr (PH > 0) il_preload(s_next, e_next);      // the next level's first batch travels during the reduce
                if constexpr (!IL) {
                    if (m > (uint32_t)H) gather_run<F>(pairs, rec, min(s + gl + H * W, lim), lim, W, swv, sw);
                }
            }
#elif RG_HEADBATCH
            static_assert(!IL, "the warp-slice path is written for RG_HEADBATCH 2");
            {
                // Head of the row: the pair loads of the first 2U-1 slots of every lane are issued together, so
                // that the second batch of gathers does not wait for another trip to L2 (most rows fit entirely).
                constexpr int U = RG_UNROLL, H = 2 * U - 1, NV = Layout<F>::NV;
                const uint32_t lim = heavy_mine ? s : e;               // heavy rows were summed by the whole warp
                uint2 hp[H];
#pragma unroll
                for (int j = 0; j < H; ++j) hp[j] = s + gl + j * W < lim ? __ldcs(pairs + s + gl + j * W) : make_uint2(rec.null_gate, 0u);
                {
                    float v[U][NV];
#pragma unroll
                    for (int j = 0; j < U; ++j) load_record<F>(rec, hp[j].x, v[j]);
#pragma unroll
                    for (int j = 0; j < U; ++j) accumulate<F, NV>(__uint_as_float(hp[j].y), v[j], swv, sw);
                }
                if (__any_sync(kFull, s + gl + U * W < lim)) {
                    float v[U - 1][NV];
#pragma unroll
                    for (int j = 0; j < U - 1; ++j) load_record<F>(rec, hp[U + j].x, v[j]);
#pragma unroll
                    for (int j = 0; j < U - 1; ++j) accumulate<F, NV>(__uint_as_float(hp[U + j].y), v[j], swv, sw);
                }
                if (__any_sync(kFull, s + gl + H * W < lim)) gather_run<F>(pairs, rec, min(s + gl + H * W, lim), lim, W, swv, sw);
            }
#else
            static_assert(!IL, "the warp-slice path is written for RG_HEADBATCH 2");
            if (!heavy_mine) gather_run<F>(pairs, rec, s + gl, e, W, swv, sw);
#endif

            if constexpr (RG_TREDUCE && W >= 8) {
                // plain butterfly down to 8 lanes, then reduce-scatter: lane f of the group gets field f
#pragma unroll
                for (int f = 0; f < F; ++f) {
#pragma unroll
                    for (int off = W / 2; off >= 8; off >>= 1) {
                        swv[f] += __shfl_xor_sync(kFull, swv[f], off);
                        sw[f] += __shfl_xor_sync(kFull, sw[f], off);
                    }
                }
                group8_reduce_scatter<F>(swv, sw, gl & 7, a[0], b[0]);
            } else if constexpr (RG_TREDUCE && W == 4) {
                group4_reduce_scatter<F, NO>(swv, sw, gl, a, b);
            } else {
                // butterfly inside the group: afterwards every lane of the group holds the row sums
#pragma unroll
                for (int f = 0; f < F; ++f) {
#pragma unroll
                    for (int off = W / 2; off >= 1; off >>= 1) {
                        swv[f] += __shfl_xor_sync(kFull, swv[f], off);
                        sw[f] += __shfl_xor_sync(kFull, sw[f], off);
                    }
                }
#pragma unroll
                for (int f = 0; f < F; ++f)
                    if (gl + (f / W) * W == f) { a[f / W] = swv[f]; b[f / W] = sw[f]; }
            }
        }

        // EMPTY: the whole slice has no pairs at this level (above the highest sweep, beyond the last gate): every
        // output is the fill value; with a NaN fill the running maximum is not touched either.
        auto finish = [&](auto empty_tag) {
            constexpr bool EMPTY = decltype(empty_tag)::value;
            if constexpr (PROD) {                              // the generic product list: both states of the lane in one pass
                float vv[NO];
                bool on[NO];
#pragma unroll
                for (int k = 0; k < NO; ++k) {
                    on[k] = owner && gl + k * W < F;
                    vv[k] = !EMPTY && b[k] > 0.f ? fast_div(a[k], b[k]) : p.fill;              // interpolate.py:99-102
                    if (on[k]) {
                        float* const dst = QS ? p.grid_out[gl + k * W] : out[k];
                        if (dst != nullptr) __stcs(dst + row, vv[k]);
                    }
                }
                ColumnState::update_words_n<NO>(p.prod, sm_state, NO, p.z_begin + lz, vv, on);
                return;
            }
#pragma unroll
            for (int k = 0; k < NO; ++k) {
                if (owner && gl + k * W < F) {
                    const float v = !EMPTY && b[k] > 0.f ? fast_div(a[k], b[k]) : p.fill;      // interpolate.py:99-102
                    // QS: pointer from the parameter bank, no register held across levels
                    float* const dst = QS ? p.grid_out[gl + k * W] : out[k];
                    if (dst != nullptr) __stcs(dst + row, v);
                    if constexpr (PSIG == 2) {
                        const int z = p.z_begin + lz;
                        if constexpr (QS) {
                            float* const qs = q_state + k * kApplyThreads + threadIdx.x;
                            if (!EMPTY || !isnan(p.fill)) {
                                if ((unsigned)(z - p.prod.cmax_z0) < p.prod.cmax_w && !isnan(v)) {
                                    const float c = qs[0];
                                    qs[0] = isnan(c) ? v : fmaxf(c, v);
                                }
                            }
                            if (z == p.prod.slices[0].z_lo || z == p.prod.slices[0].z_hi) {     // uniform: two levels of the column
                                if (z == p.prod.slices[0].z_lo) qs[NO * kApplyThreads] = v;
                                if (z == p.prod.slices[0].z_hi) qs[2 * NO * kApplyThreads] = v;
                            }
                        } else {
                            if ((unsigned)(z - p.prod.cmax_z0) < p.prod.cmax_w && !isnan(v)) q_max[k] = isnan(q_max[k]) ? v : fmaxf(q_max[k], v);
                            if (z == p.prod.slices[0].z_lo) q_lo[k] = v;
                            if (z == p.prod.slices[0].z_hi) q_hi[k] = v;
                        }
                    }
                }
            }
        };
        if constexpr (PH > 0) { if (len == 0) il_preload(s_next, e_next); }   // empty level: nothing was summed, preload here
        if (IL && len == 0) finish(std::true_type{});
        else finish(std::false_type{});
    }
    if constexpr (PSIG == 2) {
#pragma unroll
        for (int k = 0; k < NO; ++k) {
            if (owner && gl + k * W < F) {
                ColumnState st;
                if constexpr (QS) {
                    const float* const qs = q_state + k * kApplyThreads + threadIdx.x;
                    st.cmax = qs[0];
                    st.s_lo[0] = qs[NO * kApplyThreads];
                    st.s_hi[0] = qs[2 * NO * kApplyThreads];
                } else {
                    st.cmax = q_max[k];
                    st.s_lo[0] = q_lo[k];
                    st.s_hi[0] = q_hi[k];
                }
                st.zz[0] = p.prod.slices[0].z_lo | (p.prod.slices[0].z_hi << 16);   // ownership test of a partial (z-slab) blend
                st.write(p.prod, gl + k * W, col, p.ncol, 0.f, 0.f);   // only cmax and the LEVEL slice are on: x, y unused
            }
        }
    }

    if constexpr (PROD) {
        if (owner) {
            const float x = __ldg(p.prod.x_ax + (int)(col % p.nx));
            const float y = __ldg(p.prod.y_ax + (int)(col / p.nx));
#pragma unroll
            for (int k = 0; k < NO; ++k) {
                if (gl + k * W < F) {
                    ColumnState st;
                    st.load_words(p.prod, sm_state, k, NO);
                    st.load_levels(p.prod, sm_state, NO);
                    st.write(p.prod, gl + k * W, col, p.ncol, x, y);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// K5  thread-per-column path over the interleaved ("sliced, compacted") copy of the table.
//
// A warp owns 32 consecutive columns, one per lane, and walks the levels bottom-up.  For one level the warp's 32
// rows form a slice whose pairs are stored step-major (see rg_geometry.cu), so step k is ONE coalesced load of
// the active lanes' k-th pairs; every lane then gathers its gate record and accumulates its own row in
// registers.  No cross-lane reduction, no per-row shuffles, 32 rows share the per-level bookkeeping, and the
// grid stores are 128-byte coalesced per field.  The pair addresses depend only on the row lengths, so several
// steps are in flight before the first gather returns.  The few rows longer than kSellCap continue in the CSR
// copy with the whole warp striding over the remainder.
// ------------------------------------------------------------------------------------------------------
template <int F, bool PROD>
__global__ void __launch_bounds__(kSellThreads) apply_sell_kernel(const __grid_constant__ ApplyParams p)
{
    extern __shared__ float sm_state[];                        // [n_state_words][F][kSellThreads]
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int NV = Layout<F>::NV;
    constexpr int kDepth = 4;
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t col = (int64_t)blockIdx.x * kSellThreads + threadIdx.x;
    const bool col_ok = col < p.ncol;
    const int64_t slice_in_level = ((int64_t)blockIdx.x * kSellThreads + (threadIdx.x & ~31)) >> 5;

    const uint32_t* __restrict__ indptr = p.indptr;
    const uint2* __restrict__ sell = p.sell;
    const RecSrc rec{p.records, p.records_b, p.tex_a, p.tex_b, p.null_gate};

    float x = 0.f, y = 0.f;
    ColumnState st;
    if constexpr (PROD) {
        if (col_ok) {
            x = __ldg(p.prod.x_ax + (int)(col % p.nx));
            y = __ldg(p.prod.y_ax + (int)(col / p.nx));
        }
        st.init(p.prod, x, y, col);
#pragma unroll
        for (int f = 0; f < F; ++f) st.store_words(p.prod, sm_state, f, F);
        st.store_levels(p.prod, sm_state, F);
    }

    uint32_t s_next = 0, e_next = 0;
    if (col_ok && p.lz_first < p.lz_last) {
        const size_t row = (size_t)p.lz_first * (size_t)p.ncol + (size_t)col;
        s_next = __ldg(indptr + row);
        e_next = __ldg(indptr + row + 1);
    }

    for (int lz = p.lz_first; lz < p.lz_last; ++lz) {
        const uint32_t s = s_next, e = e_next;
        const size_t row = (size_t)lz * (size_t)p.ncol + (size_t)col;
        if (col_ok && lz + 1 < p.lz_last) {
            s_next = __ldg(indptr + row + (size_t)p.ncol);
            e_next = __ldg(indptr + row + (size_t)p.ncol + 1);
        }
        const uint32_t len = e - s;
        const uint32_t n = min(len, kSellCap);
        const uint32_t* sbp = p.slice_base + (size_t)lz * (size_t)p.slices_per_level + (size_t)slice_in_level;
        uint32_t base = __ldg(sbp);
        const uint32_t kmax = __reduce_max_sync(kFull, n);
#if RG_PREFETCH > 0
        if (lz + RG_PREFETCH < p.lz_last) {                   // next level's slice: HBM -> L2 while this one is summed
            const uint32_t b0 = __ldg(sbp + (size_t)RG_PREFETCH * (size_t)p.slices_per_level);
            const uint32_t b1 = __ldg(sbp + (size_t)RG_PREFETCH * (size_t)p.slices_per_level + 1);
            for (uint32_t q = b0 + 16u * lane; q < b1; q += 16u * 32u) prefetch_l2(sell + q);
        }
#endif

        float swv[F], sw[F];
#pragma unroll
        for (int f = 0; f < F; ++f) { swv[f] = 0.f; sw[f] = 0.f; }

        for (uint32_t k = 0; k < kmax; k += kDepth) {
            uint32_t addr[kDepth];
            bool act[kDepth];
#pragma unroll
            for (int j = 0; j < kDepth; ++j) {
                act[j] = k + j < n;
                const unsigned m = __ballot_sync(kFull, act[j]);
                addr[j] = base + __popc(m & lt_mask);
                base += __popc(m);
            }
            // idle lanes point at the all-masked record n_gates: no divergent control flow in the hot loop
            uint2 pr[kDepth];
#pragma unroll
            for (int j = 0; j < kDepth; ++j) pr[j] = act[j] ? __ldcs(sell + addr[j]) : make_uint2(p.null_gate, 0u);
            float v[kDepth][NV];
#pragma unroll
            for (int j = 0; j < kDepth; ++j) load_record<F>(rec, pr[j].x, v[j]);
#pragma unroll
            for (int j = 0; j < kDepth; ++j) accumulate<F, NV>(__uint_as_float(pr[j].y), v[j], swv, sw);
        }

        // rows longer than the interleaved copy holds: the whole warp strides over the rest in the CSR copy
        unsigned heavy = __ballot_sync(kFull, len > kSellCap);
        while (heavy) {
            const int src = __ffs(heavy) - 1;
            heavy &= heavy - 1;
            const uint32_t hs = __shfl_sync(kFull, s, src) + kSellCap;
            const uint32_t he = __shfl_sync(kFull, e, src);
            float hwv[F], hw[F];
#pragma unroll
            for (int f = 0; f < F; ++f) { hwv[f] = 0.f; hw[f] = 0.f; }
            gather_run<F>(p.pairs, rec, hs + lane, he, 32, hwv, hw);
#pragma unroll
            for (int f = 0; f < F; ++f) {
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    hwv[f] += __shfl_xor_sync(kFull, hwv[f], off);
                    hw[f] += __shfl_xor_sync(kFull, hw[f], off);
                }
                if (lane == src) { swv[f] += hwv[f]; sw[f] += hw[f]; }
            }
        }

        if (col_ok) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                const float val = sw[f] > 0.f ? fast_div(swv[f], sw[f]) : p.fill;    // interpolate.py:99-102
                float* out = p.grid_out[f];
                if (out != nullptr) __stcs(out + row, val);
                if constexpr (PROD) {
                    ColumnState::update_words(p.prod, sm_state, f, F, p.z_begin + lz, val);
                }
            }
        }
    }

    if constexpr (PROD) {
        if (col_ok) {
#pragma unroll
            for (int f = 0; f < F; ++f) {
                st.load_words(p.prod, sm_state, f, F);
                st.write(p.prod, f, col, p.ncol, x, y);
            }
        }
    }
}


// ------------------------------------------------------------------------------------------------------
// K5  reference-order path: reproduces np.add.reduceat's summation order, so that on the reference's own
//     table (KD-tree row order) the grid is bit-identical to interpolate.py's.  One thread per row.
//     reduceat(seg) = a[0] + pairwise_sum(a[1:])   with NumPy's pairwise_sum: < 8 elements sequential,
//     <= 128 eight interleaved accumulators, otherwise split at (n/2 rounded down to a multiple of 8).
// ------------------------------------------------------------------------------------------------------
#if RG_LO
template <int FP>
struct RowTerms {
    const uint2* pairs;
    const float* rec;
    int field;
    bool weights_only;
    int stride;        // floats per gate in the array `rec` points at
    int offset;        // position of the field inside its record
    const float* mask_rec;   // array holding the mask-bit word (MB layouts), nullptr: masked values carry kMaskedBits
    int mask_stride;
    int mask_slot;
    int mask_shift;    // Layout<F>::SH
    __device__ __forceinline__ float at(uint32_t i) const
    {
        const uint2 pr = __ldg(pairs + i);
        const float v = __ldg(rec + (size_t)pr.x * stride + offset);
        const bool m = mask_rec != nullptr
                           ? (__float_as_uint(__ldg(mask_rec + (size_t)pr.x * mask_stride + mask_slot)) >> (field + mask_shift)) & 1u
                           : __float_as_uint(v) == kMaskedBits;
        const float we = m ? 0.f : __uint_as_float(pr.y);
        if (weights_only) return we;
        return __fmul_rn(we, m ? 0.f : v);                                  // interpolate.py:82
    }
};

template <int FP>
__device__ float pairwise_leaf(const RowTerms<FP>& t, uint32_t off, uint32_t n)
{
    if (n < 8) {
        float res = 0.f;
        for (uint32_t i = 0; i < n; ++i) res = __fadd_rn(res, t.at(off + i));
        return res;
    }
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = t.at(off + j);
    uint32_t i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], t.at(off + i + j));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, t.at(off + i));
    return res;
}

template <int FP>
__device__ float pairwise_sum(const RowTerms<FP>& t, uint32_t off, uint32_t n)
{
    if (n <= 128) return pairwise_leaf<FP>(t, off, n);
    // explicit post-order walk of NumPy's recursion
    struct Frame { uint32_t off, n; int stage; float left; };
    Frame st[40];
    int sp = 0;
    st[0] = {off, n, 0, 0.f};
    float ret = 0.f;
    while (sp >= 0) {
        Frame& f = st[sp];
        if (f.n <= 128) {
            ret = pairwise_leaf<FP>(t, f.off, f.n);
            --sp;
            continue;
        }
        uint32_t n2 = f.n / 2;
        n2 -= n2 % 8;
        if (f.stage == 0) {
            f.stage = 1;
            st[sp + 1] = {f.off, n2, 0, 0.f};
            ++sp;
        } else if (f.stage == 1) {
            f.left = ret;
            f.stage = 2;
            st[sp + 1] = {f.off + n2, f.n - n2, 0, 0.f};
            ++sp;
        } else {
            ret = __fadd_rn(f.left, ret);
            --sp;
        }
    }
    return ret;
}

template <int FP>
__global__ void __launch_bounds__(128) apply_reference_order_kernel(const __grid_constant__ ApplyParams p, int64_t n_rows)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint32_t s = __ldg(p.indptr + row), e = __ldg(p.indptr + row + 1);
    for (int f = 0; f < p.n_fields; ++f) {
        float* out = p.grid_out[f];
        if (out == nullptr) continue;
        float v = p.fill;
        if (e > s) {
            const int nf = p.n_fields;                             // see Layout<F>
            const int fa = nf == 1 ? 1 : nf == 2 ? 2 : 4;
            const int fb = nf <= 4 ? 0 : nf == 5 ? (RG_MASKBITS ? 2 : 1) : nf == 6 ? 2 : 4;
            const bool mb = RG_MASKBITS && (nf == 3 || nf == 5 || nf == 7);   // mask word in slot nf: A[3], B[1] or B[3]
            RowTerms<FP> t{p.pairs, f < fa ? p.records : p.records_b, f, false, f < fa ? fa : fb, f < fa ? f : f - fa,
                           mb ? (nf == 3 ? p.records : p.records_b) : nullptr, nf == 3 ? fa : fb, mb ? (nf == 3 ? 3 : nf - fa) : -1,
                           nf <= 5 ? 1 : 0};
            float swv = t.at(s);
            if (e - s > 1) swv = __fadd_rn(swv, pairwise_sum<FP>(t, s + 1, e - s - 1));
            t.weights_only = true;
            float sw = t.at(s);
            if (e - s > 1) sw = __fadd_rn(sw, pairwise_sum<FP>(t, s + 1, e - s - 1));
            if (sw > 0.f) v = __fdiv_rn(swv, sw);
        }
        out[row] = v;
    }
}

// ------------------------------------------------------------------------------------------------------
// Nearest-gate gridding (the interpolation stage process_radar_to_cog asks Py-ART for: map_gates_to_grid with
// weighting_function='nearest', reference processor.py:152-163).  The table was built with RG_W_DIST2, i.e. its weight
// slot holds float32(d^2); per field a voxel takes the value of the closest gate whose value is not masked; among equal
// distances the lowest gate id wins (Py-ART keeps the first gate of its scan that reaches the minimum, `dist2 <
// min_dist2`).  One warp per row; not a fused path: the grid is written, products come from products_kernel.
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) apply_nearest_kernel(const __grid_constant__ ApplyParams p, int64_t n_rows)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (row >= n_rows) return;
    const uint32_t s = __ldg(p.indptr + row), e = __ldg(p.indptr + row + 1);
    const int nf = p.n_fields;
    const int fa = nf == 1 ? 1 : nf == 2 ? 2 : 4;
    const int fb = nf <= 4 ? 0 : nf == 5 ? (RG_MASKBITS ? 2 : 1) : nf == 6 ? 2 : 4;
    const bool mb = RG_MASKBITS && (nf == 3 || nf == 5 || nf == 7);
    for (int f = 0; f < nf; ++f) {
        float* out = p.grid_out[f];
        if (out == nullptr) continue;
        const float* rec = f < fa ? p.records : p.records_b;
        const int stride = f < fa ? fa : fb, off = f < fa ? f : f - fa;
        float best_d = __uint_as_float(0x7F800000u);          // +inf
        uint32_t best_g = 0xFFFFFFFFu;
        float best_v = 0.f;
        for (uint32_t i = s + lane; i < e; i += 32) {
            const uint2 pr = __ldg(p.pairs + i);
            const float v = __ldg(rec + (size_t)pr.x * stride + off);
            bool masked;
            if (mb) {
                const float* mrec = nf == 3 ? p.records : p.records_b;
                const uint32_t bits = __float_as_uint(__ldg(mrec + (size_t)pr.x * (nf == 3 ? fa : fb) + (nf == 3 ? 3 : nf - fa)));
                masked = (bits >> (f + (nf <= 5 ? 1 : 0))) & 1u;
            } else {
                masked = __float_as_uint(v) == kMaskedBits;
            }
            const float d = __uint_as_float(pr.y);
            if (!masked && (d < best_d || (d == best_d && pr.x < best_g))) { best_d = d; best_g = pr.x; best_v = v; }
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            const float od = __shfl_xor_sync(kFull, best_d, o);
            const uint32_t og = __shfl_xor_sync(kFull, best_g, o);
            const float ov = __shfl_xor_sync(kFull, best_v, o);
            if (og != 0xFFFFFFFFu && (best_g == 0xFFFFFFFFu || od < best_d || (od == best_d && og < best_g))) {
                best_d = od; best_g = og; best_v = ov;
            }
        }
        if (lane == 0) out[row] = best_g == 0xFFFFFFFFu ? p.fill : best_v;
    }
}

int launch_apply_nearest(Context* ctx, const Geometry* g, const ApplyParams& p)
{
    if (g->n_rows == 0) return RG_OK;
    apply_nearest_kernel<<<(unsigned)((g->n_rows + 3) / 4), 128, 0, ctx->stream>>>(p, g->n_rows);
    ctx->launches++;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

#endif  // RG_LO

// ------------------------------------------------------------------------------------------------------
// dispatch
// ------------------------------------------------------------------------------------------------------
// Launch the column kernel.  When heavy_rows_kernel was launched just before it, the two are chained by programmatic
// dependent launch: the heavy kernel lets its dependents start right away (griddepcontrol.launch_dependents) and only the
// few groups that pick up a heavy row's partial sums wait for it (griddepcontrol.wait), so its ~10 us of latency-bound
// work run under the column kernel instead of in front of it.
template <typename K>
static void launch_column_kernel(K kernel, unsigned blocks, size_t smem, Context* ctx, const ApplyParams& p)
{
#if RG_PDL && !defined(RG_EMU)
    if (p.n_heavy_chunks > 0) {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(blocks);
        cfg.blockDim = dim3(kApplyThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = ctx->stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaLaunchKernelEx(&cfg, kernel, p);
        return;
    }
#endif
    kernel<<<blocks, kApplyThreads, smem, ctx->stream>>>(p);
}

template <int F, int W>
static void launch_columns(Context* ctx, const ApplyParams& p)
{
    const int cols_per_cta = kApplyThreads / W;
#if RG_TILE2D
    const int TX = cols_per_cta >= 8 ? 8 : cols_per_cta, TY = cols_per_cta / TX;
    const unsigned blocks = (unsigned)(((p.nx + TX - 1) / TX) * ((p.ny + TY - 1) / TY));
#else
    const unsigned blocks = (unsigned)((p.ncol + cols_per_cta - 1) / cols_per_cta);
#endif
    constexpr int NO = W < F ? 2 : 1;                             // fields per owner lane, as in the kernel
    const size_t smem = (size_t)(p.prod.n_state_words * NO + RG_MAX_SLICES) * kApplyThreads * sizeof(float);
    const ProductParams& pp = p.prod;
    const bool simple = pp.any && !pp.cmin_on && !pp.cmean_on && pp.n_slices <= 1 &&
                        (pp.n_slices == 0 || pp.slices[0].kind == RG_PROD_LEVEL) && ctx->apply_variant != 3;
    const size_t smem2 = RG_QSMEM && NO > 1 ? (size_t)3 * NO * kApplyThreads * sizeof(float) : 0;   // PSIG 2 state (QS)
    if (p.quads != nullptr) {
#if RG_TILE2D
        if (!pp.any) launch_column_kernel(apply_columns_kernel<F, W, 0, true>, blocks, 0, ctx, p);
        else if (simple) launch_column_kernel(apply_columns_kernel<F, W, 2, true>, blocks, smem2, ctx, p);
        else launch_column_kernel(apply_columns_kernel<F, W, 1, true>, blocks, smem, ctx, p);
        return;
#endif
    }
    if (!pp.any) launch_column_kernel(apply_columns_kernel<F, W, 0, false>, blocks, 0, ctx, p);
    else if (simple) launch_column_kernel(apply_columns_kernel<F, W, 2, false>, blocks, smem2, ctx, p);
    else launch_column_kernel(apply_columns_kernel<F, W, 1, false>, blocks, smem, ctx, p);
}

template <int F>
static int launch_columns_w(Context* ctx, const ApplyParams& p, int W)
{
    if (W == 4) { launch_columns<F, 4>(ctx, p); return RG_OK; }
    if (W <= 8) launch_columns<F, 8>(ctx, p);
    else if (W == 16) launch_columns<F, 16>(ctx, p);
    else launch_columns<F, 32>(ctx, p);
    return RG_OK;
}

template <int F>
static int launch_sell(Context* ctx, const ApplyParams& p)
{
    const unsigned blocks = (unsigned)((p.ncol + kSellThreads - 1) / kSellThreads);
    const size_t smem = p.prod.any ? (size_t)(p.prod.n_state_words * F + RG_MAX_SLICES) * kSellThreads * sizeof(float) : 0;
    if (p.prod.any) {
        if (smem > 48 * 1024)
            RG_CUDA(cudaFuncSetAttribute(apply_sell_kernel<F, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        apply_sell_kernel<F, true><<<blocks, kSellThreads, smem, ctx->stream>>>(p);
    } else {
        apply_sell_kernel<F, false><<<blocks, kSellThreads, 0, ctx->stream>>>(p);
    }
    return RG_OK;
}

template <int F>
static void launch_heavy(Context* ctx, const ApplyParams& p)
{
    heavy_rows_kernel<F><<<(unsigned)p.n_heavy_chunks, kHeavyThreads, 0, ctx->stream>>>(p);
}

#if RG_HI
void launch_heavy_hi(Context* ctx, const ApplyParams& p)
{
    switch (p.n_fields) {
        case 5: launch_heavy<5>(ctx, p); break;
        case 6: launch_heavy<6>(ctx, p); break;
        case 7: launch_heavy<7>(ctx, p); break;
        default: launch_heavy<8>(ctx, p); break;
    }
}

int launch_columns_hi(Context* ctx, const ApplyParams& p, int W)
{
    switch (p.n_fields) {
        case 5: return launch_columns_w<5>(ctx, p, W);
        case 6: return launch_columns_w<6>(ctx, p, W);
        case 7: return launch_columns_w<7>(ctx, p, W);
        default: return launch_columns_w<8>(ctx, p, W);
    }
}

int launch_sell_hi(Context* ctx, const ApplyParams& p)
{
    switch (p.n_fields) {
        case 5: return launch_sell<5>(ctx, p);
        case 6: return launch_sell<6>(ctx, p);
        case 7: return launch_sell<7>(ctx, p);
        default: return launch_sell<8>(ctx, p);
    }
}
#endif  // RG_HI

#if RG_LO
// Pairs from the warp-slice copy?  It pays when the kernel is bound by instruction issue and the L1 data pipe: several
// fields per pair, or one field over short rows (cfg1: rows of 16 pairs, 0.0439 vs 0.0466 ms: the per-lane bounds and
// votes of the CSR path weigh as much as the sums there).  A single-field pass over long rows (cfg5) is HBM-bound and
// better off without the padding and the second copy of a table that fills the GPU.
#ifndef RG_DUO_MIN_FIELDS
#define RG_DUO_MIN_FIELDS 1    // fewest fields of a pass that takes the column-pair kernel on its own (option duo = 2: any).  Measured on the
                               //    cfg3 table, duo against column-group kernel: F=1 0.451 / 0.550 ms, 2: 0.441 / 0.478, 3: 0.497 / 0.566, 4: 0.514 / 0.536
#endif
static bool use_slices(const Context* ctx, const Geometry* g, int n_fields)
{
#if RG_TILE2D && RG_HEADBATCH == 2
    if (ctx->apply_variant == 4) return true;
    if (ctx->apply_variant == 1) return false;
    if (n_fields >= 2) return true;
    const int64_t nonempty = g->info.n_rows - g->info.n_empty_rows;
    return nonempty > 0 && (double)g->info.n_pairs / (double)nonempty < 24.0;
#else
    return false;
#endif
}

static int pick_group_width(const Context* ctx, const Geometry* g, int n_fields, bool slices)
{
    int W = (int)ctx->group_width;
    if (W == 0) {
        const int64_t nonempty = g->info.n_rows - g->info.n_empty_rows;
        const double avg = nonempty > 0 ? (double)g->info.n_pairs / (double)nonempty : 0.0;
        // measured on B200: CSR copy: cfg1 (avg 16) W=4, cfg3 (avg 40) W=8; slice copy: cfg3 W=4 (0.71 vs 0.81 ms)
        if (slices) W = avg < 96.0 ? 4 : avg < 320.0 ? 8 : avg < 1200.0 ? 16 : 32;
        else W = avg < 24.0 ? 4 : avg < 160.0 ? 8 : avg < 600.0 ? 16 : 32;
        if (!slices && W < 8 && n_fields > 4) W = 8;
    }
    if (W != 4 && W != 8 && W != 16 && W != 32) W = 8;
    return W;
}

int launch_apply(Context* ctx, const Geometry* g, const ApplyParams& p, bool reference_order)
{
    // An empty z-slab (z_begin == z_end: more ranks than levels) has no rows, but a fused products request must still
    // get its planes: the column kernel walks zero levels and writes the initial state (NaN; 0/0 for COLMEAN).
    const bool empty_slab = g->n_rows == 0;
    if (empty_slab && (reference_order || !p.prod.any || g->ncol == 0)) return RG_OK;
    if (reference_order) {
        const unsigned blocks = (unsigned)((g->n_rows + 127) / 128);
        switch (records_width(p.n_fields)) {
            case 1: apply_reference_order_kernel<1><<<blocks, 128, 0, ctx->stream>>>(p, g->n_rows); break;
            case 2: apply_reference_order_kernel<2><<<blocks, 128, 0, ctx->stream>>>(p, g->n_rows); break;
            case 4: apply_reference_order_kernel<4><<<blocks, 128, 0, ctx->stream>>>(p, g->n_rows); break;
            default: apply_reference_order_kernel<8><<<blocks, 128, 0, ctx->stream>>>(p, g->n_rows); break;
        }
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
        return RG_OK;
    }
    if (ctx->apply_variant == 2) {                      // thread-per-column over the interleaved copy (kept for A/B)
        timer_begin(ctx, kTimerApply);
        int st = RG_OK;
        switch (p.n_fields) {
            case 1: st = launch_sell<1>(ctx, p); break;
            case 2: st = launch_sell<2>(ctx, p); break;
            case 3: st = launch_sell<3>(ctx, p); break;
            case 4: st = launch_sell<4>(ctx, p); break;
            case 5: case 6: case 7: case 8: st = launch_sell_hi(ctx, p); break;
            default: return fail(RG_ERR_INVALID, "n_fields must be 1..8");
        }
        timer_end(ctx, kTimerApply);
        ctx->launches++;
        RG_TRY(st);
        RG_CUDA(cudaGetLastError());
        return RG_OK;
    }
    const bool slices = !empty_slab && use_slices(ctx, g, p.n_fields);
    const int W = pick_group_width(ctx, g, p.n_fields, slices);
    ApplyParams q = p;
    q.quads = nullptr;
    q.quad_ptr = nullptr;
    q.quads_x = 0;
    q.heavy_rows = nullptr; q.heavy_first = nullptr; q.heavy_chunks = nullptr; q.heavy_part = nullptr;
    q.n_heavy = 0; q.n_heavy_chunks = 0;
    // Rows that overlap with their neighbour's: the column-pair kernel (rg_duo.cu) over the merged rows of two adjacent columns.
    // Auto: tables whose mean row is below 160 pairs (group widths 4 and 8).  Measured against the column-group kernel: cfg3
    // (mean 40 pairs, five fields) 0.604 / 0.640 ms; cfg1 (16 pairs, one field) 0.0378 / 0.0444 ms; cfg5 (110 pairs, one field,
    // 226 GB of table slab by slab) 35.2 / 41.9 ms.  The decision depends on the row lengths only, so the z-slabs of one grid
    // take the same kernel and stay bit-identical to the unsharded pass.
    bool duo = false;
    q.duo = nullptr; q.duo_ptr = nullptr; q.duo_qx = 0; q.duo_nyp = 0;
    if (!empty_slab && ctx->duo != 0 && ctx->apply_variant == 0 && ctx->group_width == 0 &&
        (ctx->duo >= 2 || (W <= 8 && p.n_fields >= RG_DUO_MIN_FIELDS))) {
        const Geometry::DuoCopy* dc = nullptr;
        RG_TRY(ensure_duo(ctx, const_cast<Geometry*>(g), &dc));
        if (dc != nullptr) {                            // nullptr: the table does not lend itself to it
            q.duo = dc->slots; q.duo_ptr = dc->ptr; q.duo_qx = dc->qx; q.duo_nyp = dc->nyp;
            duo = true;
        }
    }
    if ((W < 32 || duo) && !empty_slab) {               // a 32-lane group IS the whole warp: nothing is "heavy" for it
        Geometry* gm = const_cast<Geometry*>(g);
        RG_TRY(ensure_heavy(ctx, gm));
        if (g->n_heavy > 0) {
            RG_TRY(ensure(ctx, ctx->heavy, (size_t)g->n_heavy_chunks * 2 * p.n_fields * sizeof(float)));
            q.heavy_rows = g->heavy_rows; q.heavy_first = g->heavy_first; q.heavy_chunks = g->heavy_chunks;
            q.heavy_part = (float*)ctx->heavy.ptr;
            q.n_heavy = (int32_t)g->n_heavy; q.n_heavy_chunks = (int32_t)g->n_heavy_chunks;
        }
    }
#if RG_TILE2D && RG_HEADBATCH == 2
    if (slices && !duo) {                               // pairs from the warp-slice copy (built on first use)
        const Geometry::QuadCopy* qc = nullptr;
        RG_TRY(ensure_quads(ctx, const_cast<Geometry*>(g), W, &qc));
        if (qc != nullptr) {                            // nullptr: no room for the copy, read the CSR copy
            q.quads = qc->quads;
            q.quad_ptr = qc->ptr;
            q.quads_x = qc->quads_x;
        }
    }
#endif
    timer_begin(ctx, kTimerApply);
    if (q.n_heavy_chunks > 0) {
        switch (q.n_fields) {
            case 1: launch_heavy<1>(ctx, q); break;
            case 2: launch_heavy<2>(ctx, q); break;
            case 3: launch_heavy<3>(ctx, q); break;
            case 4: launch_heavy<4>(ctx, q); break;
            case 5: case 6: case 7: case 8: launch_heavy_hi(ctx, q); break;
            default: return fail(RG_ERR_INVALID, "n_fields must be 1..8");
        }
        ctx->launches++;
    }
    if (duo) {
        RG_TRY(launch_duo(ctx, q));
    } else switch (q.n_fields) {
        case 1: launch_columns_w<1>(ctx, q, W); break;
        case 2: launch_columns_w<2>(ctx, q, W); break;
        case 3: launch_columns_w<3>(ctx, q, W); break;
        case 4: launch_columns_w<4>(ctx, q, W); break;
        case 5: case 6: case 7: case 8: launch_columns_hi(ctx, q, W); break;
        default: return fail(RG_ERR_INVALID, "n_fields must be 1..8");
    }
    timer_end(ctx, kTimerApply);
    ctx->launches++;
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

#endif  // RG_LO

}  // namespace rg
