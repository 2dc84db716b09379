// C ABI of libradargrid_b200.so (see include/radar_grid_b200.h): handles, argument validation,
// host<->device staging for RG_HOST callers, and the launch sequence of one apply:
//     [H2D fields/masks] -> K4 pack -> K5 apply (+K6 epilogue) -> [D2H grids/planes]

#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <string>
#include <vector>

#include "rg_internal.cuh"

namespace rg {

static thread_local std::string g_last_error;

void set_error(const std::string& msg) { g_last_error = msg; }
int fail(int status, const std::string& msg)
{
    g_last_error = msg;
    return status;
}

int ensure(Context* ctx, Scratch& s, size_t bytes)
{
    if (bytes <= s.bytes) return RG_OK;
    if (s.ptr) {
        RG_CUDA(cudaStreamSynchronize(ctx->stream));
        RG_CUDA(cudaFree(s.ptr));
        s.ptr = nullptr;
        s.bytes = 0;
    }
    const size_t want = bytes + bytes / 8 + 256;
    RG_CUDA(cudaMalloc(&s.ptr, want));
    s.bytes = want;
    return RG_OK;
}

void timer_begin(Context* ctx, int which)
{
    if (!ctx->timing) return;
    cudaEvent_t a = nullptr, b = nullptr;
    if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
    ctx->timers[which].start.push_back(a);
    ctx->timers[which].stop.push_back(b);
    cudaEventRecord(a, ctx->stream);
}

void timer_end(Context* ctx, int which)
{
    if (!ctx->timing || ctx->timers[which].stop.empty()) return;
    cudaEventRecord(ctx->timers[which].stop.back(), ctx->stream);
}

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

#define RG_ENTER(ctx_)                                                                   \
    if ((ctx_) == nullptr) return ::rg::fail(RG_ERR_INVALID, "context is NULL");         \
    ::rg::DeviceGuard rg_guard_((ctx_)->device);                                         \
    if (!rg_guard_.ok) return ::rg::fail(RG_ERR_CUDA, "cannot select the context's CUDA device")

int check_grid(const rg_grid_spec* g)
{
    if (g == nullptr) return fail(RG_ERR_INVALID, "grid spec is NULL");
    if (g->nz <= 0 || g->ny <= 0 || g->nx <= 0) return fail(RG_ERR_INVALID, "grid_shape entries must be positive");
    if (g->z_begin < 0 || g->z_end > g->nz || g->z_begin > g->z_end)
        return fail(RG_ERR_INVALID, "z-slab [z_begin, z_end) must lie inside [0, nz]");
    if (g->nz > 32767) return fail(RG_ERR_UNSUPPORTED, "nz > 32767 is not supported");
    return RG_OK;
}

__global__ void __launch_bounds__(256) interleave_kernel(const int32_t* __restrict__ idx, const float* __restrict__ w,
                                                         uint2* __restrict__ pairs, int64_t n, uint32_t n_gates,
                                                         unsigned int* __restrict__ bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t g = idx[i];
    if (g < 0 || (uint32_t)g >= n_gates) {
        atomicAdd(bad, 1u);
        pairs[i] = make_uint2(0u, 0u);
        return;
    }
    pairs[i] = make_uint2((uint32_t)g, __float_as_uint(w[i]));
}

__global__ void __launch_bounds__(256) deinterleave_kernel(const uint2* __restrict__ pairs, int32_t* __restrict__ idx,
                                                           float* __restrict__ w, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 p = pairs[i];
    if (idx) idx[i] = (int32_t)p.x;
    if (w) w[i] = __uint_as_float(p.y);
}

__global__ void __launch_bounds__(256) widen_indptr_kernel(const uint32_t* __restrict__ in, int64_t* __restrict__ out, int64_t n)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int64_t)in[i];
}

__global__ void __launch_bounds__(256) narrow_indptr_kernel(const int64_t* __restrict__ in, uint32_t* __restrict__ out, int64_t n,
                                                            unsigned int* __restrict__ bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t v = in[i];
    if (v < 0 || v >= 0xFFFFFFFFll || (i > 0 && in[i - 1] > v)) atomicAdd(bad, 1u);
    out[i] = (uint32_t)v;
}

__global__ void __launch_bounds__(256) check_indptr32_kernel(const uint32_t* __restrict__ in, int64_t n, unsigned int* __restrict__ bad)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || i == 0) return;
    if (in[i - 1] > in[i]) atomicAdd(bad, 1u);
}

void free_geometry(Geometry* g)
{
    if (!g) return;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(g->device);
    cudaFree(g->indptr);
    cudaFree(g->pairs);
    cudaFree(g->sell);
    for (auto& qc : g->quad) { cudaFree(qc.quads); cudaFree(qc.ptr); }
    cudaFree(g->duo.slots);
    cudaFree(g->duo.ptr);
    cudaFree(g->slice_base);
    cudaFree(g->heavy_rows);
    cudaFree(g->heavy_first);
    cudaFree(g->heavy_chunks);
    cudaFree(g->x_ax);
    cudaFree(g->y_ax);
    cudaFree(g->z_ax);
    if (prev >= 0) cudaSetDevice(prev);
    delete g;
}

int upload_axes(Context* ctx, Geometry* g)
{
    const rg_grid_spec& gs = g->grid;
    std::vector<float> xa(gs.nx), ya(gs.ny), za(gs.nz);
    linspace_f32(gs.x_min, gs.x_max, gs.nx, xa.data());
    linspace_f32(gs.y_min, gs.y_max, gs.ny, ya.data());
    linspace_f32(gs.z_min, gs.z_max, gs.nz, za.data());
    RG_CUDA(cudaMalloc(&g->x_ax, gs.nx * sizeof(float)));
    RG_CUDA(cudaMalloc(&g->y_ax, gs.ny * sizeof(float)));
    RG_CUDA(cudaMalloc(&g->z_ax, gs.nz * sizeof(float)));
    RG_CUDA(cudaMemcpyAsync(g->x_ax, xa.data(), gs.nx * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    RG_CUDA(cudaMemcpyAsync(g->y_ax, ya.data(), gs.ny * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    RG_CUDA(cudaMemcpyAsync(g->z_ax, za.data(), gs.nz * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    return RG_OK;
}

// Translate the public product list into the kernel parameter block.  `x_ax`/`y_ax` are device axes.
// `plane_ptrs[i]` (device) replaces products[i].out when non-null (host-memspace staging).
// Device-side placement of the image products of one call: LUT and RGBA output per product (nullptr = no image).
struct ImageStage {
    std::vector<const uchar4*> lut;
    std::vector<uchar4*> out;
};

int check_image(const rg_product& pr)
{
    const rg_image& im = *pr.image;
    if (pr.partial) return fail(RG_ERR_INVALID, "a partial (z-slab term) product has no image form");
    if (im.n_filters < 0 || im.n_filters > RG_MAX_IMAGE_FILTERS) return fail(RG_ERR_INVALID, "image: n_filters must be 0..4");
    for (int i = 0; i < im.n_filters; ++i)
        if (im.filter_kind[i] < RG_PF_BELOW || im.filter_kind[i] > RG_PF_BELOW_EQUAL) return fail(RG_ERR_INVALID, "image: unknown filter kind");
    if (!isfinite(im.vmin) || !isfinite(im.vmax)) return fail(RG_ERR_INVALID, "image: vmin and vmax must be finite");
    if (im.vmin > im.vmax) return fail(RG_ERR_INVALID, "minvalue must be less than or equal to maxvalue");   // Normalize's ValueError
    if (im.lut_entries < 1 || im.lut_entries > 4096) return fail(RG_ERR_INVALID, "image: lut_entries must be 1..4096");
    if (!im.lut || !im.out) return fail(RG_ERR_INVALID, "image: lut / out pointer is NULL");
    return RG_OK;
}

// Stage the LUTs (and, for host callers, device buffers for the RGBA planes) of the image products.
int stage_images(Context* ctx, int n_products, const rg_product* products, bool host, int n_fields, int64_t ncol, ImageStage* st)
{
    st->lut.assign(std::max(n_products, 0), nullptr);
    st->out.assign(std::max(n_products, 0), nullptr);
    size_t lut_bytes = 0, out_bytes = 0;
    int n_images = 0;
    for (int i = 0; i < n_products; ++i) {
        if (!products[i].image) continue;
        RG_TRY(check_image(products[i]));
        if (++n_images > RG_MAX_IMAGES) return fail(RG_ERR_UNSUPPORTED, "at most 3 image products per call");
        lut_bytes += (((size_t)products[i].image->lut_entries + 3) * 4 + 255) & ~(size_t)255;
        out_bytes += (((size_t)n_fields * (size_t)ncol * 4) + 255) & ~(size_t)255;
    }
    if (n_images == 0 || !host) {
        for (int i = 0; i < n_products; ++i)
            if (products[i].image) { st->lut[i] = (const uchar4*)products[i].image->lut; st->out[i] = (uchar4*)products[i].image->out; }
        return RG_OK;
    }
    RG_TRY(ensure(ctx, ctx->luts, lut_bytes + out_bytes));
    char* cur = (char*)ctx->luts.ptr;
    for (int i = 0; i < n_products; ++i) {
        if (!products[i].image) continue;
        const size_t nb = ((size_t)products[i].image->lut_entries + 3) * 4;
        RG_CUDA(cudaMemcpyAsync(cur, products[i].image->lut, nb, cudaMemcpyHostToDevice, ctx->stream));
        st->lut[i] = (const uchar4*)cur;
        cur += (nb + 255) & ~(size_t)255;
    }
    for (int i = 0; i < n_products; ++i) {
        if (!products[i].image) continue;
        st->out[i] = (uchar4*)cur;
        cur += (((size_t)n_fields * (size_t)ncol * 4) + 255) & ~(size_t)255;
    }
    return RG_OK;
}

int fetch_images(Context* ctx, int n_products, const rg_product* products, int n_fields, int64_t ncol, const ImageStage& st)
{
    for (int i = 0; i < n_products; ++i)
        if (products[i].image)
            RG_CUDA(cudaMemcpyAsync(products[i].image->out, st.out[i], (size_t)n_fields * (size_t)ncol * 4, cudaMemcpyDeviceToHost, ctx->stream));
    return RG_OK;
}

int make_product_params(const rg_grid_spec& gs, int n_products, const rg_product* products, const float* x_ax,
                        const float* y_ax, void* const* plane_ptrs, const ImageStage* images, ProductParams* pp,
                        int* z_need_lo, int* z_need_hi)
{
    memset(pp, 0, sizeof(*pp));
    pp->cmax_image = pp->cmin_image = pp->cmean_image = -1;
    auto add_image = [&](int i) -> int {
        const rg_product& pr = products[i];
        if (!pr.image || !images) return -1;
        const rg_image& im = *pr.image;
        ImageParams& ip = pp->images[pp->n_images];
        ip.on = 1;
        ip.n_filters = im.n_filters;
        for (int k = 0; k < RG_MAX_IMAGE_FILTERS; ++k) {
            ip.kind[k] = im.filter_kind[k]; ip.a[k] = im.filter_a[k]; ip.b[k] = im.filter_b[k]; ip.fill[k] = im.filter_fill[k];
        }
        ip.vmin = im.vmin; ip.vmax = im.vmax; ip.fill_value = im.fill_value;
        ip.has_fill = im.has_fill_value; ip.lut_n = im.lut_entries;
        ip.lut = images->lut[i];
        ip.out = images->out[i];
        return pp->n_images++;
    };
    pp->nz_full = gs.nz;
    pp->own_z0 = gs.z_begin;
    pp->own_z1 = gs.z_end;
    pp->z_min = gs.z_min;
    pp->z_max = gs.z_max;
    // products.py:256,276: (z_max - z_min) / (nz - 1) if nz > 1 else 1.0
    pp->z_step = gs.nz > 1 ? (gs.z_max - gs.z_min) / (double)(gs.nz - 1) : 1.0;
    pp->x_ax = x_ax;
    pp->y_ax = y_ax;
    pp->nx = gs.nx; pp->ny = gs.ny;
    pp->x_min = gs.x_min; pp->x_max = gs.x_max; pp->y_min = gs.y_min; pp->y_max = gs.y_max;
    int lo = gs.nz, hi = -1;
    if (n_products < 0) return fail(RG_ERR_INVALID, "n_products < 0");
    if (n_products > 0 && products == nullptr) return fail(RG_ERR_INVALID, "products is NULL");
    for (int i = 0; i < n_products; ++i) {
        const rg_product& pr = products[i];
        void* out = pr.out == nullptr ? nullptr : (plane_ptrs ? plane_ptrs[i] : pr.out);
        if (out == nullptr && !pr.image) return fail(RG_ERR_INVALID, "product output pointer is NULL");
        const int z0 = std::max(0, pr.z_lo), z1 = std::min(gs.nz - 1, pr.z_hi);
        switch (pr.kind) {
            case RG_PROD_COLMAX:
                if (pp->cmax_on) return fail(RG_ERR_UNSUPPORTED, "at most one COLMAX per call");
                pp->cmax_on = 1; pp->cmax_z0 = z0; pp->cmax_z1 = z1; pp->cmax_out = (float*)out;
                pp->cmax_partial = pr.partial != 0;
                pp->cmax_image = add_image(i);
                lo = std::min(lo, z0); hi = std::max(hi, z1);
                break;
            case RG_PROD_COLMIN:
                if (pp->cmin_on) return fail(RG_ERR_UNSUPPORTED, "at most one COLMIN per call");
                pp->cmin_on = 1; pp->cmin_z0 = z0; pp->cmin_z1 = z1; pp->cmin_out = (float*)out;
                pp->cmin_partial = pr.partial != 0;
                pp->cmin_image = add_image(i);
                lo = std::min(lo, z0); hi = std::max(hi, z1);
                break;
            case RG_PROD_COLMEAN:
                if (pp->cmean_on) return fail(RG_ERR_UNSUPPORTED, "at most one COLMEAN per call");
                pp->cmean_on = 1; pp->cmean_z0 = z0; pp->cmean_z1 = z1; pp->cmean_out = (float*)out;
                pp->cmean_image = add_image(i);
                lo = std::min(lo, z0); hi = std::max(hi, z1);
                break;
            case RG_PROD_LEVEL:
            case RG_PROD_BEAM: {
                if (pp->n_slices >= RG_MAX_SLICES) return fail(RG_ERR_UNSUPPORTED, "too many LEVEL/BEAM products in one call");
                SliceParams& s = pp->slices[pp->n_slices++];
                s.kind = pr.kind;
                s.mode = pr.mode;
                s.partial = pr.partial != 0;
                s.out = out;
                s.image = add_image(i);
                if (pr.kind == RG_PROD_LEVEL) {
                    if (pr.mode < RG_BLEND_PICK || pr.mode > RG_BLEND_F64_OUT64) return fail(RG_ERR_INVALID, "bad blend mode");
                    if (pr.z_lo < 0 || pr.z_lo >= gs.nz || (pr.mode != RG_BLEND_PICK && (pr.z_hi < 0 || pr.z_hi >= gs.nz)))
                        return fail(RG_ERR_INVALID, "LEVEL product: level index outside the grid");
                    s.z_lo = pr.z_lo;
                    s.z_hi = pr.mode == RG_BLEND_PICK ? pr.z_lo : pr.z_hi;
                    s.w_lo = pr.w_lo; s.w_hi = pr.w_hi;
                    lo = std::min(lo, std::min(s.z_lo, s.z_hi)); hi = std::max(hi, std::max(s.z_lo, s.z_hi));
                } else {
                    if (pr.mode < 0 || pr.mode > 2) return fail(RG_ERR_INVALID, "BEAM product: mode must be 0 (linear), 1 (nearest) or 2 (closest level)");
                    s.curvature = pr.earth_curvature;
                    s.sin_e = pr.sin_elev; s.cos_c = pr.cos_elev_clamped; s.tan_e = pr.tan_elev;
                    s.ke_re = pr.ke_re; s.ke_re_sq = pr.ke_re_sq;
                    lo = 0; hi = gs.nz - 1;
                }
                break;
            }
            default:
                return fail(RG_ERR_INVALID, "unknown product kind");
        }
    }
    pp->any = n_products > 0 ? 1 : 0;
    pp->cmax_w = pp->cmax_on && pp->cmax_z1 >= pp->cmax_z0 ? (uint32_t)(pp->cmax_z1 - pp->cmax_z0 + 1) : 0u;
    pp->cmin_w = pp->cmin_on && pp->cmin_z1 >= pp->cmin_z0 ? (uint32_t)(pp->cmin_z1 - pp->cmin_z0 + 1) : 0u;
    pp->cmean_w = pp->cmean_on && pp->cmean_z1 >= pp->cmean_z0 ? (uint32_t)(pp->cmean_z1 - pp->cmean_z0 + 1) : 0u;
    int words = 0;
    pp->slot_cmax = pp->cmax_on ? words++ : -1;
    pp->slot_cmin = pp->cmin_on ? words++ : -1;
    pp->slot_cmean = pp->cmean_on ? words : -1;
    if (pp->cmean_on) words += 2;
    for (int k = 0; k < RG_MAX_SLICES; ++k) {
        pp->slot_slice[k] = k < pp->n_slices ? words : -1;
        if (k < pp->n_slices) words += 2;
    }
    pp->n_state_words = words;
    pp->n_ops = 0;
    if (pp->cmax_w) pp->ops[pp->n_ops++] = {1, pp->cmax_z0, pp->cmax_z1, pp->slot_cmax, pp->cmax_w, 0};
    if (pp->cmin_w) pp->ops[pp->n_ops++] = {2, pp->cmin_z0, pp->cmin_z1, pp->slot_cmin, pp->cmin_w, 0};
    if (pp->cmean_w) pp->ops[pp->n_ops++] = {3, pp->cmean_z0, pp->cmean_z1, pp->slot_cmean, pp->cmean_w, 0};
    for (int k = 0; k < pp->n_slices; ++k) {
        const SliceParams& sl = pp->slices[k];
        if (sl.kind == RG_PROD_BEAM) pp->ops[pp->n_ops++] = {5, 0, 0, pp->slot_slice[k], 0u, k};
        else pp->ops[pp->n_ops++] = {4, sl.z_lo, sl.z_hi, pp->slot_slice[k], 0u, k};
    }
    *z_need_lo = lo;
    *z_need_hi = hi;
    return RG_OK;
}

size_t product_plane_bytes(const rg_product& pr, int n_fields, int64_t ncol)
{
    const bool f64 = (pr.kind == RG_PROD_BEAM && pr.mode == 0) || (pr.kind == RG_PROD_LEVEL && pr.mode == RG_BLEND_F64_OUT64) ||
                     (pr.kind == RG_PROD_LEVEL && pr.mode == RG_BLEND_F64 && pr.partial);
    return (size_t)n_fields * (size_t)ncol * (f64 ? 8 : 4);
}

template <typename T>
__global__ void __launch_bounds__(256) plane_filter_kernel(const T* __restrict__ in, T* __restrict__ out, int64_t n, int kind,
                                                           T a, T b, T fill)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const T v = in[i];
    bool hit;
    switch (kind) {
        case RG_PF_BELOW: hit = v < a; break;                    // filters.py:660
        case RG_PF_ABOVE: hit = v > a; break;                    // filters.py:689
        case RG_PF_OUTSIDE: hit = (v < a) || (v > b); break;     // filters.py:721
        case RG_PF_BELOW_EQUAL: hit = v <= a; break;             // processor.py:543 (masked_less_equal)
        default: hit = isnan(v) || isinf(v); break;              // filters.py:746
    }
    out[i] = hit ? fill : v;
}

// pyart's antenna_to_cartesian (4/3 effective earth radius), float64 with one rounding per operation, then float32
__global__ void __launch_bounds__(256) gate_xyz_kernel(const float* __restrict__ range_m, const float* __restrict__ az_deg,
                                                       const float* __restrict__ el_deg, int64_t n_rays, int64_t n_bins,
                                                       float* __restrict__ x, float* __restrict__ y, float* __restrict__ z)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_rays * n_bins) return;
    const int64_t ray = g / n_bins, bin = g - ray * n_bins;
    const double R = 4.0 / 3.0 * 6371000.0;
    const double deg = 3.14159265358979323846 / 180.0;              // np.radians: x * (pi / 180)
    const double r = (double)__ldg(range_m + bin);
    const double el = __dmul_rn((double)__ldg(el_deg + ray), deg), az = __dmul_rn((double)__ldg(az_deg + ray), deg);
    const double zz = __dsub_rn(__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(r, r), __dmul_rn(R, R)),
                                                     __dmul_rn(__dmul_rn(__dmul_rn(2.0, r), R), sin(el)))), R);
    const double s = __dmul_rn(R, asin(__ddiv_rn(__dmul_rn(r, cos(el)), __dadd_rn(R, zz))));
    x[g] = (float)__dmul_rn(s, sin(az));
    y[g] = (float)__dmul_rn(s, cos(az));
    z[g] = (float)zz;
}

inline size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace
}  // namespace rg

using namespace rg;

extern "C" {

int rg_abi_version(void) { return RG_ABI_VERSION; }

const char* rg_last_error(void) { return g_last_error.c_str(); }

int rg_device_count(int32_t* count)
{
    if (!count) return fail(RG_ERR_INVALID, "count is NULL");
    int n = 0;
    const cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(RG_ERR_CUDA, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = n;
    return RG_OK;
}

int rg_context_create(int32_t device, void* stream, rg_context** out)
{
    if (!out) return fail(RG_ERR_INVALID, "out is NULL");
    *out = nullptr;
    int n = 0;
    RG_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return fail(RG_ERR_CUDA, "no such CUDA device (this library has no CPU fallback)");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(RG_ERR_CUDA, "cannot select CUDA device");
    Context* ctx = new (std::nothrow) Context();
    if (!ctx) return fail(RG_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        const cudaError_t e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) { delete ctx; return fail(RG_ERR_CUDA, cudaGetErrorString(e)); }
        ctx->own_stream = true;
    }
    cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (const char* e = getenv("RADAR_GRID_B200_DUO")) {        // operations / A-B knob: the default of option "duo"
        const int v = atoi(e);
        if (v >= 0 && v <= 2) ctx->duo = v;
    }
    *out = reinterpret_cast<rg_context*>(ctx);
    return RG_OK;
}

int rg_context_destroy(rg_context* c)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    if (!ctx) return RG_OK;
    DeviceGuard guard(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (Scratch* s : {&ctx->records, &ctx->stage_in, &ctx->stage_out, &ctx->misc, &ctx->heavy, &ctx->luts, &ctx->flags})
        if (s->ptr) cudaFree(s->ptr);
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return RG_OK;
}

int rg_context_set_stream(rg_context* c, void* stream)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream) { cudaStreamDestroy(ctx->stream); ctx->own_stream = false; }
    if (stream) {
        ctx->stream = (cudaStream_t)stream;
    } else {
        RG_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        ctx->own_stream = true;
    }
    return RG_OK;
}

int rg_context_synchronize(rg_context* c)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    return RG_OK;
}

int rg_context_get_stream(const rg_context* c, void** stream)
{
    const Context* ctx = reinterpret_cast<const Context*>(c);
    if (!ctx || !stream) return fail(RG_ERR_INVALID, "NULL argument");
    *stream = (void*)ctx->stream;
    return RG_OK;
}

int rg_context_kernel_launches(const rg_context* c, int64_t* count)
{
    const Context* ctx = reinterpret_cast<const Context*>(c);
    if (!ctx || !count) return fail(RG_ERR_INVALID, "NULL argument");
    *count = ctx->launches;
    return RG_OK;
}

int rg_context_set_option(rg_context* c, const char* key, int64_t value)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    if (!ctx || !key) return fail(RG_ERR_INVALID, "NULL argument");
    if (strcmp(key, "apply_variant") == 0) ctx->apply_variant = value;
    else if (strcmp(key, "timing") == 0) ctx->timing = value;
    else if (strcmp(key, "sort_rows") == 0) ctx->sort_rows = value;
    else if (strcmp(key, "duo") == 0) {
        if (value < 0 || value > 2) return fail(RG_ERR_INVALID, "duo must be 0 (off), 1 (auto) or 2 (whenever the table allows it)");
        ctx->duo = value;
    }
    else if (strcmp(key, "group_width") == 0) {
        if (value != 0 && value != 4 && value != 8 && value != 16 && value != 32)
            return fail(RG_ERR_INVALID, "group_width must be 0, 4, 8, 16 or 32");
        ctx->group_width = value;
    } else return fail(RG_ERR_INVALID, std::string("unknown option: ") + key);
    return RG_OK;
}

int rg_context_kernel_time(rg_context* c, int32_t which, double* total_ms, int64_t* count, int32_t reset)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    if (which < 0 || which >= kTimerCount || !total_ms || !count) return fail(RG_ERR_INVALID, "bad argument");
    KernelTimer& t = ctx->timers[which];
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    double sum = 0.0;
    for (size_t i = 0; i < t.start.size(); ++i) {
        float ms = 0.f;
        RG_CUDA(cudaEventElapsedTime(&ms, t.start[i], t.stop[i]));
        sum += ms;
    }
    *total_ms = sum;
    *count = (int64_t)t.start.size();
    if (reset) {
        for (size_t i = 0; i < t.start.size(); ++i) { cudaEventDestroy(t.start[i]); cudaEventDestroy(t.stop[i]); }
        t.start.clear();
        t.stop.clear();
    }
    return RG_OK;
}

int rg_host_alloc(void** ptr, int64_t bytes)
{
    if (!ptr || bytes < 0) return fail(RG_ERR_INVALID, "bad argument");
    *ptr = nullptr;
    RG_CUDA(cudaHostAlloc(ptr, (size_t)std::max<int64_t>(bytes, 1), cudaHostAllocDefault));
    return RG_OK;
}

int rg_host_free(void* ptr)
{
    if (ptr) RG_CUDA(cudaFreeHost(ptr));
    return RG_OK;
}

int rg_linspace_f32(double start, double stop, int32_t num, float* out)
{
    if (num < 0 || (num > 0 && !out)) return fail(RG_ERR_INVALID, "bad argument");
    linspace_f32(start, stop, num, out);
    return RG_OK;
}

// ---- gate coordinates ---------------------------------------------------------------------------------
int rg_gate_coordinates(rg_context* c, const float* range_m, const float* azimuth_deg, const float* elevation_deg,
                        int64_t n_rays, int64_t n_bins, int32_t memspace_in, float* x, float* y, float* z, int32_t memspace_out)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    if (n_rays < 0 || n_bins < 0 || n_rays * n_bins >= 0xFFFFFFFFll) return fail(RG_ERR_INVALID, "n_rays * n_bins out of range");
    if ((memspace_in != RG_DEVICE && memspace_in != RG_HOST) || (memspace_out != RG_DEVICE && memspace_out != RG_HOST))
        return fail(RG_ERR_INVALID, "bad memspace");
    const int64_t n = n_rays * n_bins;
    if (n == 0) return RG_OK;
    if (!range_m || !azimuth_deg || !elevation_deg || !x || !y || !z) return fail(RG_ERR_INVALID, "NULL pointer");
    const float *dr = range_m, *da = azimuth_deg, *de = elevation_deg;
    if (memspace_in == RG_HOST) {
        RG_TRY(ensure(ctx, ctx->stage_in, align256((size_t)n_bins * 4) + 2 * align256((size_t)n_rays * 4)));
        char* cur = (char*)ctx->stage_in.ptr;
        RG_CUDA(cudaMemcpyAsync(cur, range_m, (size_t)n_bins * 4, cudaMemcpyHostToDevice, ctx->stream));
        dr = (const float*)cur; cur += align256((size_t)n_bins * 4);
        RG_CUDA(cudaMemcpyAsync(cur, azimuth_deg, (size_t)n_rays * 4, cudaMemcpyHostToDevice, ctx->stream));
        da = (const float*)cur; cur += align256((size_t)n_rays * 4);
        RG_CUDA(cudaMemcpyAsync(cur, elevation_deg, (size_t)n_rays * 4, cudaMemcpyHostToDevice, ctx->stream));
        de = (const float*)cur;
    }
    float *dx = x, *dy = y, *dz = z;
    if (memspace_out == RG_HOST) {
        RG_TRY(ensure(ctx, ctx->stage_out, 3 * align256((size_t)n * 4)));
        dx = (float*)ctx->stage_out.ptr;
        dy = (float*)((char*)dx + align256((size_t)n * 4));
        dz = (float*)((char*)dy + align256((size_t)n * 4));
    }
    gate_xyz_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(dr, da, de, n_rays, n_bins, dx, dy, dz);
    ctx->launches++;
    RG_CUDA(cudaGetLastError());
    if (memspace_out == RG_HOST) {
        RG_CUDA(cudaMemcpyAsync(x, dx, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        RG_CUDA(cudaMemcpyAsync(y, dy, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
        RG_CUDA(cudaMemcpyAsync(z, dz, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    if (memspace_in == RG_HOST || memspace_out == RG_HOST) RG_CUDA(cudaStreamSynchronize(ctx->stream));
    return RG_OK;
}

// ---- geometry ---------------------------------------------------------------------------------------
int rg_geometry_build(rg_context* c, const float* gate_x, const float* gate_y, const float* gate_z, int64_t n_gates,
                      int32_t memspace, const rg_grid_spec* grid, double radar_altitude, double min_radius,
                      double beam_factor, int32_t weighting, double toa, rg_geometry** out)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    if (!out) return fail(RG_ERR_INVALID, "out is NULL");
    *out = nullptr;
    RG_TRY(check_grid(grid));
    if (n_gates < 0 || n_gates >= 0xFFFFFFFFll) return fail(RG_ERR_INVALID, "n_gates out of range");
    if (n_gates > 0 && (!gate_x || !gate_y || !gate_z)) return fail(RG_ERR_INVALID, "gate coordinate pointer is NULL");
    if (weighting != RG_W_BARNES2 && weighting != RG_W_CRESSMAN && weighting != RG_W_NEAREST && weighting != RG_W_DIST2)
        return fail(RG_ERR_INVALID, "Unknown weighting function");
    if (!(min_radius >= 0.0) || !(beam_factor >= 0.0) || !isfinite(min_radius) || !isfinite(beam_factor))
        return fail(RG_ERR_INVALID, "min_radius and beam_factor must be finite and >= 0");

    const float *dx = gate_x, *dy = gate_y, *dz = gate_z;
    float* staged = nullptr;
    if (memspace == RG_HOST && n_gates > 0) {
        RG_CUDA(cudaMalloc(&staged, (size_t)n_gates * 3 * sizeof(float)));
        const size_t nb = (size_t)n_gates * sizeof(float);
        cudaError_t e = cudaMemcpyAsync(staged, gate_x, nb, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(staged + n_gates, gate_y, nb, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(staged + 2 * n_gates, gate_z, nb, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { cudaFree(staged); return fail(RG_ERR_CUDA, cudaGetErrorString(e)); }
        dx = staged; dy = staged + n_gates; dz = staged + 2 * n_gates;
    } else if (memspace != RG_DEVICE && memspace != RG_HOST) {
        return fail(RG_ERR_INVALID, "bad memspace");
    }

    Geometry* g = new (std::nothrow) Geometry();
    if (!g) { cudaFree(staged); return fail(RG_ERR_NOMEM, "out of host memory"); }
    g->device = ctx->device;
    g->grid = *grid;
    const int st = build_geometry_device(ctx, dx, dy, dz, n_gates, radar_altitude, min_radius, beam_factor, weighting, toa, g);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(staged);
    if (st != RG_OK) { free_geometry(g); return st; }
    *out = reinterpret_cast<rg_geometry*>(g);
    return RG_OK;
}

int rg_geometry_level_pairs(rg_context* c, const float* gate_x, const float* gate_y, const float* gate_z, int64_t n_gates,
                            int32_t memspace, const rg_grid_spec* grid, double radar_altitude, double min_radius,
                            double beam_factor, double toa, int32_t column_stride, int64_t* pairs_per_level)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    RG_TRY(check_grid(grid));
    if (!pairs_per_level) return fail(RG_ERR_INVALID, "pairs_per_level is NULL");
    if (column_stride < 1) return fail(RG_ERR_INVALID, "column_stride must be >= 1");
    if (n_gates < 0 || n_gates >= 0xFFFFFFFFll) return fail(RG_ERR_INVALID, "n_gates out of range");
    if (n_gates > 0 && (!gate_x || !gate_y || !gate_z)) return fail(RG_ERR_INVALID, "gate coordinate pointer is NULL");
    if (memspace != RG_DEVICE && memspace != RG_HOST) return fail(RG_ERR_INVALID, "bad memspace");
    if (!(min_radius >= 0.0) || !(beam_factor >= 0.0) || !isfinite(min_radius) || !isfinite(beam_factor))
        return fail(RG_ERR_INVALID, "min_radius and beam_factor must be finite and >= 0");
    const float *dx = gate_x, *dy = gate_y, *dz = gate_z;
    float* staged = nullptr;
    if (memspace == RG_HOST && n_gates > 0) {
        RG_CUDA(cudaMalloc(&staged, (size_t)n_gates * 3 * sizeof(float)));
        const size_t nb = (size_t)n_gates * sizeof(float);
        cudaError_t e = cudaMemcpyAsync(staged, gate_x, nb, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(staged + n_gates, gate_y, nb, cudaMemcpyHostToDevice, ctx->stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(staged + 2 * n_gates, gate_z, nb, cudaMemcpyHostToDevice, ctx->stream);
        if (e != cudaSuccess) { cudaFree(staged); return fail(RG_ERR_CUDA, cudaGetErrorString(e)); }
        dx = staged; dy = staged + n_gates; dz = staged + 2 * n_gates;
    }
    Geometry* g = new (std::nothrow) Geometry();
    if (!g) { cudaFree(staged); return fail(RG_ERR_NOMEM, "out of host memory"); }
    g->device = ctx->device;
    g->grid = *grid;
    const int st = build_geometry_device(ctx, dx, dy, dz, n_gates, radar_altitude, min_radius, beam_factor, RG_W_NEAREST, toa, g,
                                         column_stride, pairs_per_level);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(staged);
    free_geometry(g);
    return st;
}

int rg_geometry_from_csr(rg_context* c, const rg_grid_spec* grid, const void* indptr, int32_t indptr_bits,
                         const int32_t* gate_indices, const float* weights, int64_t n_gates, int32_t memspace,
                         rg_geometry** out)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    if (!out) return fail(RG_ERR_INVALID, "out is NULL");
    *out = nullptr;
    RG_TRY(check_grid(grid));
    if (!indptr) return fail(RG_ERR_INVALID, "indptr is NULL");
    if (indptr_bits != 32 && indptr_bits != 64) return fail(RG_ERR_INVALID, "indptr_bits must be 32 or 64");
    if (n_gates < 0 || n_gates >= 0xFFFFFFFFll) return fail(RG_ERR_INVALID, "n_gates out of range");
    if (memspace != RG_DEVICE && memspace != RG_HOST) return fail(RG_ERR_INVALID, "bad memspace");

    const int64_t ncol = (int64_t)grid->ny * grid->nx;
    const int64_t n_rows = ncol * (grid->z_end - grid->z_begin);
    const size_t ip_bytes = (size_t)(n_rows + 1) * (indptr_bits / 8);

    Geometry* g = new (std::nothrow) Geometry();
    if (!g) return fail(RG_ERR_NOMEM, "out of host memory");
    g->device = ctx->device;
    g->grid = *grid;
    g->n_rows = n_rows; g->ncol = ncol; g->n_levels = grid->z_end - grid->z_begin; g->n_gates = n_gates;

    auto bail = [&](int st) { free_geometry(g); return st; };
#define RG_CUDA_G(expr)                                                                                   \
    do {                                                                                                  \
        cudaError_t e_ = (expr);                                                                          \
        if (e_ != cudaSuccess) return bail(fail(e_ == cudaErrorMemoryAllocation ? RG_ERR_NOMEM : RG_ERR_CUDA, \
                                                std::string(#expr) + ": " + cudaGetErrorString(e_)));    \
    } while (0)

    // indptr -> device uint32
    unsigned int* bad = nullptr;
    RG_CUDA_G(cudaMalloc(&bad, sizeof(unsigned int)));
    struct Free { void* p; ~Free() { cudaFree(p); } } free_bad{bad};
    RG_CUDA_G(cudaMemsetAsync(bad, 0, sizeof(unsigned int), ctx->stream));
    RG_CUDA_G(cudaMalloc(&g->indptr, (size_t)(n_rows + 1) * sizeof(uint32_t)));
    void* ip_tmp = nullptr;
    Free free_ip{nullptr};
    const void* ip_dev = indptr;
    if (memspace == RG_HOST || indptr_bits == 64) {
        if (memspace == RG_HOST) {
            RG_CUDA_G(cudaMalloc(&ip_tmp, ip_bytes));
            free_ip.p = ip_tmp;
            RG_CUDA_G(cudaMemcpyAsync(ip_tmp, indptr, ip_bytes, cudaMemcpyHostToDevice, ctx->stream));
            ip_dev = ip_tmp;
        }
    }
    const unsigned ipb = (unsigned)((n_rows + 1 + 255) / 256);
    if (indptr_bits == 64) {
        narrow_indptr_kernel<<<ipb, 256, 0, ctx->stream>>>((const int64_t*)ip_dev, g->indptr, n_rows + 1, bad);
    } else {
        RG_CUDA_G(cudaMemcpyAsync(g->indptr, ip_dev, ip_bytes, cudaMemcpyDeviceToDevice, ctx->stream));
        check_indptr32_kernel<<<ipb, 256, 0, ctx->stream>>>(g->indptr, n_rows + 1, bad);
    }
    ctx->launches++;
    uint32_t first = 0, last = 0;
    RG_CUDA_G(cudaMemcpyAsync(&first, g->indptr, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    RG_CUDA_G(cudaMemcpyAsync(&last, g->indptr + n_rows, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    unsigned int nbad = 0;
    RG_CUDA_G(cudaMemcpyAsync(&nbad, bad, sizeof(nbad), cudaMemcpyDeviceToHost, ctx->stream));
    RG_CUDA_G(cudaStreamSynchronize(ctx->stream));
    if (nbad != 0 || first != 0) return bail(fail(RG_ERR_INVALID, "indptr must start at 0, be non-decreasing and below 2^32-1"));
    const int64_t n_pairs = (int64_t)last;
    if (n_pairs > 0 && (!gate_indices || !weights)) return bail(fail(RG_ERR_INVALID, "gate_indices / weights is NULL"));
    g->n_pairs = n_pairs;

    RG_CUDA_G(cudaMalloc(&g->pairs, std::max<size_t>((size_t)n_pairs, 1) * sizeof(uint2)));
    if (n_pairs > 0) {
        const int32_t* idx_dev = gate_indices;
        const float* w_dev = weights;
        void* tmp = nullptr;
        Free free_tmp{nullptr};
        if (memspace == RG_HOST) {
            RG_CUDA_G(cudaMalloc(&tmp, (size_t)n_pairs * 8));
            free_tmp.p = tmp;
            RG_CUDA_G(cudaMemcpyAsync(tmp, gate_indices, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
            RG_CUDA_G(cudaMemcpyAsync((char*)tmp + (size_t)n_pairs * 4, weights, (size_t)n_pairs * 4, cudaMemcpyHostToDevice, ctx->stream));
            idx_dev = (const int32_t*)tmp;
            w_dev = (const float*)((char*)tmp + (size_t)n_pairs * 4);
        }
        interleave_kernel<<<(unsigned)((n_pairs + 255) / 256), 256, 0, ctx->stream>>>(idx_dev, w_dev, g->pairs, n_pairs,
                                                                                     (uint32_t)n_gates, bad);
        ctx->launches++;
        RG_CUDA_G(cudaMemcpyAsync(&nbad, bad, sizeof(nbad), cudaMemcpyDeviceToHost, ctx->stream));
        RG_CUDA_G(cudaStreamSynchronize(ctx->stream));
        if (nbad != 0) return bail(fail(RG_ERR_INVALID, "gate_indices contains values outside [0, n_gates)"));
    }
    {
        const int st = upload_axes(ctx, g);
        if (st != RG_OK) return bail(st);
    }
    g->info.n_rows = n_rows; g->info.n_pairs = n_pairs; g->info.n_gates = n_gates; g->info.grid = *grid;
    g->info.device_bytes = (int64_t)(((size_t)n_rows + 1) * 4 + (size_t)n_pairs * 8);
    {
        const int st = finalize_geometry_stats(ctx, g);
        if (st != RG_OK) return bail(st);
    }
#undef RG_CUDA_G
    *out = reinterpret_cast<rg_geometry*>(g);
    return RG_OK;
}

int rg_geometry_get_info(const rg_geometry* geom, rg_geometry_info* info)
{
    const Geometry* g = reinterpret_cast<const Geometry*>(geom);
    if (!g || !info) return fail(RG_ERR_INVALID, "NULL argument");
    *info = g->info;
    return RG_OK;
}

int rg_geometry_duo_slots(const rg_geometry* geom, int64_t* n_slots)
{
    const Geometry* g = reinterpret_cast<const Geometry*>(geom);
    if (!g || !n_slots) return fail(RG_ERR_INVALID, "NULL argument");
    *n_slots = g->duo.n_slots;
    return RG_OK;
}

int rg_geometry_export_csr(rg_context* c, const rg_geometry* geom, void* indptr, int32_t indptr_bits,
                           int32_t* gate_indices, float* weights, int32_t memspace)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    const Geometry* g = reinterpret_cast<const Geometry*>(geom);
    if (!g) return fail(RG_ERR_INVALID, "geometry is NULL");
    if (g->device != ctx->device) return fail(RG_ERR_INVALID, "geometry lives on another device");
    if (indptr_bits != 32 && indptr_bits != 64) return fail(RG_ERR_INVALID, "indptr_bits must be 32 or 64");
    if (memspace != RG_DEVICE && memspace != RG_HOST) return fail(RG_ERR_INVALID, "bad memspace");
    const int64_t n1 = g->n_rows + 1, np = g->n_pairs;
    const cudaMemcpyKind kind = memspace == RG_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;

    if (indptr) {
        if (indptr_bits == 32) {
            if (g->n_pairs > 0x7FFFFFFFll)
                return fail(RG_ERR_UNSUPPORTED, "table has more than 2^31-1 pairs: export indptr as 64-bit");
            RG_CUDA(cudaMemcpyAsync(indptr, g->indptr, (size_t)n1 * 4, kind, ctx->stream));
        } else {
            int64_t* dst = (int64_t*)indptr;
            if (memspace == RG_HOST) {
                RG_TRY(ensure(ctx, ctx->misc, (size_t)n1 * 8));
                dst = (int64_t*)ctx->misc.ptr;
            }
            widen_indptr_kernel<<<(unsigned)((n1 + 255) / 256), 256, 0, ctx->stream>>>(g->indptr, dst, n1);
            ctx->launches++;
            RG_CUDA(cudaGetLastError());
            if (memspace == RG_HOST) RG_CUDA(cudaMemcpyAsync(indptr, dst, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
        }
        RG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    if ((gate_indices || weights) && np > 0) {
        int32_t* di = gate_indices;
        float* dw = weights;
        if (memspace == RG_HOST) {
            RG_TRY(ensure(ctx, ctx->misc, (size_t)np * 8));
            di = (int32_t*)ctx->misc.ptr;
            dw = (float*)((char*)ctx->misc.ptr + (size_t)np * 4);
        }
        deinterleave_kernel<<<(unsigned)((np + 255) / 256), 256, 0, ctx->stream>>>(g->pairs, gate_indices ? di : nullptr,
                                                                                  weights ? dw : nullptr, np);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
        if (memspace == RG_HOST) {
            if (gate_indices) RG_CUDA(cudaMemcpyAsync(gate_indices, di, (size_t)np * 4, cudaMemcpyDeviceToHost, ctx->stream));
            if (weights) RG_CUDA(cudaMemcpyAsync(weights, dw, (size_t)np * 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
        RG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return RG_OK;
}

int rg_geometry_destroy(rg_geometry* geom)
{
    free_geometry(reinterpret_cast<Geometry*>(geom));
    return RG_OK;
}

// ---- products on existing grids -----------------------------------------------------------------------
int rg_products(rg_context* c, const rg_grid_spec* grid, int32_t n_fields, const float* const* grids,
                int32_t n_products, const rg_product* products, int32_t memspace)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    RG_TRY(check_grid(grid));
    if (n_fields < 1 || n_fields > RG_MAX_FIELDS) return fail(RG_ERR_INVALID, "n_fields must be 1..8");
    if (!grids) return fail(RG_ERR_INVALID, "grids is NULL");
    if (memspace != RG_DEVICE && memspace != RG_HOST) return fail(RG_ERR_INVALID, "bad memspace");
    if (n_products <= 0) return RG_OK;
    const int64_t ncol = (int64_t)grid->ny * grid->nx;
    const int64_t n_rows = ncol * (grid->z_end - grid->z_begin);

    // axes
    std::vector<float> xa(grid->nx), ya(grid->ny);
    linspace_f32(grid->x_min, grid->x_max, grid->nx, xa.data());
    linspace_f32(grid->y_min, grid->y_max, grid->ny, ya.data());
    const size_t ax_bytes = align256((size_t)(grid->nx + grid->ny) * sizeof(float));

    size_t in_bytes = 0, out_bytes = 0;
    std::vector<size_t> plane_off(n_products);
    if (memspace == RG_HOST) {
        in_bytes = align256((size_t)n_rows * 4) * n_fields;
        for (int i = 0; i < n_products; ++i) {
            plane_off[i] = out_bytes;
            out_bytes += align256(product_plane_bytes(products[i], n_fields, ncol));
        }
    }
    RG_TRY(ensure(ctx, ctx->stage_in, ax_bytes + in_bytes));
    RG_TRY(ensure(ctx, ctx->stage_out, out_bytes));
    float* ax_dev = (float*)ctx->stage_in.ptr;
    RG_CUDA(cudaMemcpyAsync(ax_dev, xa.data(), grid->nx * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    RG_CUDA(cudaMemcpyAsync(ax_dev + grid->nx, ya.data(), grid->ny * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));

    const float* gdev[RG_MAX_FIELDS] = {};
    std::vector<void*> planes;
    for (int f = 0; f < n_fields; ++f) {
        if (!grids[f]) return fail(RG_ERR_INVALID, "grid pointer is NULL");
        if (memspace == RG_HOST) {
            float* d = (float*)((char*)ctx->stage_in.ptr + ax_bytes + (size_t)f * align256((size_t)n_rows * 4));
            RG_CUDA(cudaMemcpyAsync(d, grids[f], (size_t)n_rows * 4, cudaMemcpyHostToDevice, ctx->stream));
            gdev[f] = d;
        } else {
            gdev[f] = grids[f];
        }
    }
    if (memspace == RG_HOST) {
        planes.resize(n_products);
        for (int i = 0; i < n_products; ++i) planes[i] = (char*)ctx->stage_out.ptr + plane_off[i];
    }
    ProductParams pp;
    int zlo, zhi;
    ImageStage images;
    RG_TRY(stage_images(ctx, n_products, products, memspace == RG_HOST, n_fields, ncol, &images));
    RG_TRY(make_product_params(*grid, n_products, products, ax_dev, ax_dev + grid->nx,
                               memspace == RG_HOST ? planes.data() : nullptr, &images, &pp, &zlo, &zhi));
    RG_TRY(launch_products(ctx, *grid, n_fields, gdev, pp));
    if (memspace == RG_HOST) {
        for (int i = 0; i < n_products; ++i) {
            if (!products[i].out) continue;                 // image-only product
            RG_CUDA(cudaMemcpyAsync(products[i].out, planes[i], product_plane_bytes(products[i], n_fields, ncol),
                                    cudaMemcpyDeviceToHost, ctx->stream));
        }
        RG_TRY(fetch_images(ctx, n_products, products, n_fields, ncol, images));
    }
    // xa/ya are stack-owned and host outputs must be complete on return
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    return RG_OK;
}

int rg_plane_filter(rg_context* c, const void* in, void* out, int64_t n, int32_t elem_bits, int32_t kind, double a,
                    double b, double fill_value, int32_t memspace)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    if (n < 0 || (n > 0 && (!in || !out))) return fail(RG_ERR_INVALID, "bad argument");
    if (elem_bits != 32 && elem_bits != 64) return fail(RG_ERR_INVALID, "elem_bits must be 32 or 64");
    if (kind < RG_PF_BELOW || kind > RG_PF_BELOW_EQUAL) return fail(RG_ERR_INVALID, "unknown plane filter kind");
    if (memspace != RG_DEVICE && memspace != RG_HOST) return fail(RG_ERR_INVALID, "bad memspace");
    if (n == 0) return RG_OK;
    const size_t bytes = (size_t)n * (elem_bits / 8);
    const void* din = in;
    void* dout = out;
    if (memspace == RG_HOST) {
        RG_TRY(ensure(ctx, ctx->stage_in, bytes));
        RG_TRY(ensure(ctx, ctx->stage_out, bytes));
        RG_CUDA(cudaMemcpyAsync(ctx->stage_in.ptr, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
        din = ctx->stage_in.ptr;
        dout = ctx->stage_out.ptr;
    }
    const unsigned blocks = (unsigned)((n + 255) / 256);
    if (elem_bits == 32)
        plane_filter_kernel<float><<<blocks, 256, 0, ctx->stream>>>((const float*)din, (float*)dout, n, kind, (float)a, (float)b,
                                                                    (float)fill_value);
    else
        plane_filter_kernel<double><<<blocks, 256, 0, ctx->stream>>>((const double*)din, (double*)dout, n, kind, a, b, fill_value);
    ctx->launches++;
    RG_CUDA(cudaGetLastError());
    if (memspace == RG_HOST) {
        RG_CUDA(cudaMemcpyAsync(out, dout, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        RG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return RG_OK;
}

// ---- apply ------------------------------------------------------------------------------------------------
int rg_apply(rg_context* c, const rg_geometry* geom, const rg_apply_args* a, int32_t memspace)
{
    Context* ctx = reinterpret_cast<Context*>(c);
    RG_ENTER(ctx);
    const Geometry* g = reinterpret_cast<const Geometry*>(geom);
    if (!g || !a) return fail(RG_ERR_INVALID, "NULL argument");
    if (g->device != ctx->device) return fail(RG_ERR_INVALID, "geometry lives on another device");
    if (memspace != RG_DEVICE && memspace != RG_HOST) return fail(RG_ERR_INVALID, "bad memspace");
    const int F = a->n_fields;
    if (F < 1 || F > RG_MAX_FIELDS) return fail(RG_ERR_INVALID, "n_fields must be 1..8");
    if (a->n_rules < 0 || a->n_rules > RG_MAX_RULES) return fail(RG_ERR_INVALID, "n_rules must be 0..8");
    if (!a->fields) return fail(RG_ERR_INVALID, "fields is NULL");
    if (a->n_rules > 0 && !a->rules) return fail(RG_ERR_INVALID, "rules is NULL");
    const int64_t G = g->n_gates, V = g->n_rows, ncol = g->ncol;
    const bool host = memspace == RG_HOST;
    const int FP = records_width(F);

    // ---- inputs: stage to the device when the caller's buffers are host memory
    PackParams pk{};
    pk.n_gates = G;
    pk.n_fields = F;
    pk.n_rules = a->n_rules;
    pk.invalid_bits = a->mask_invalid_bits;
    size_t in_bytes = 0;
    const size_t fbytes = align256((size_t)G * 4), mbytes = align256((size_t)G);
    if (host) {
        in_bytes = (size_t)F * fbytes;
        for (int f = 0; f < F; ++f)
            if (a->masks && a->masks[f]) in_bytes += mbytes;
        in_bytes += (size_t)a->n_rules * fbytes;
        RG_TRY(ensure(ctx, ctx->stage_in, in_bytes));
    }
    char* in_cur = (char*)ctx->stage_in.ptr;
    for (int f = 0; f < F; ++f) {
        if (!a->fields[f]) return fail(RG_ERR_INVALID, "field pointer is NULL");
        if (host) {
            RG_CUDA(cudaMemcpyAsync(in_cur, a->fields[f], (size_t)G * 4, cudaMemcpyHostToDevice, ctx->stream));
            pk.fields[f] = (const float*)in_cur;
            in_cur += fbytes;
        } else {
            pk.fields[f] = a->fields[f];
        }
    }
    for (int f = 0; f < F; ++f) {
        const uint8_t* m = a->masks ? a->masks[f] : nullptr;
        if (m && host) {
            RG_CUDA(cudaMemcpyAsync(in_cur, m, (size_t)G, cudaMemcpyHostToDevice, ctx->stream));
            pk.masks[f] = (const uint8_t*)in_cur;
            in_cur += mbytes;
        } else {
            pk.masks[f] = m;
        }
    }
    for (int r = 0; r < a->n_rules; ++r) {
        const rg_qc_rule& q = a->rules[r];
        if (!q.values) return fail(RG_ERR_INVALID, "rule values pointer is NULL");
        pk.rule_values[r] = q.values;
        if (host) {
            int alias = -1;
            for (int f = 0; f < F; ++f) if (a->fields[f] == q.values) alias = f;
            for (int r2 = 0; r2 < r && alias < 0; ++r2)
                if (a->rules[r2].values == q.values) { pk.rule_values[r] = pk.rule_values[r2]; alias = -2; }
            if (alias >= 0) pk.rule_values[r] = pk.fields[alias];
            else if (alias == -1) {
                RG_CUDA(cudaMemcpyAsync(in_cur, q.values, (size_t)G * 4, cudaMemcpyHostToDevice, ctx->stream));
                pk.rule_values[r] = (const float*)in_cur;
                in_cur += fbytes;
            }
        }
        pk.rule_lo[r] = q.lo; pk.rule_hi[r] = q.hi;
        pk.rule_use_lo[r] = q.use_lo; pk.rule_use_hi[r] = q.use_hi;
        pk.rule_bits[r] = q.field_bits;
    }
    RG_TRY(ensure(ctx, ctx->records, (size_t)(G + 1) * FP * 4 + 1024));
    pk.records = (float*)ctx->records.ptr;
    pk.records_b = (float*)((char*)ctx->records.ptr + records_b_offset(F, G));
    if (ctx->flags.ptr == nullptr || ctx->epoch == 0xFFFFFFFFu) {   // first call (or the epoch counter wrapped): clear the stamp
        RG_TRY(ensure(ctx, ctx->flags, 256));
        RG_CUDA(cudaMemsetAsync(ctx->flags.ptr, 0, 256, ctx->stream));
        ctx->epoch = 0;
    }
    pk.nonfinite = (uint32_t*)ctx->flags.ptr;
    pk.epoch = ++ctx->epoch;
    RG_TRY(launch_pack(ctx, pk));

    // ---- outputs
    if (a->reference_order < 0 || a->reference_order > 2) return fail(RG_ERR_INVALID, "reference_order must be 0, 1 or 2");
    const bool nearest_gate = a->reference_order == 2;
    const bool ref_order = a->reference_order != 0;          // both exact modes are un-fused: grid, then stand-alone products
    const int n_products = a->n_products;
    bool want_grid[RG_MAX_FIELDS] = {};
    bool any_grid = false;
    for (int f = 0; f < F; ++f) {
        want_grid[f] = a->grid_out && a->grid_out[f];
        any_grid |= want_grid[f];
    }
    const size_t gbytes = align256((size_t)V * 4);
    size_t out_bytes = 0;
    std::vector<size_t> plane_off(std::max(n_products, 0));
    // device 3-D grids are needed when the caller is on the host, or (reference order + products) for every field
    bool dev_grid_tmp[RG_MAX_FIELDS] = {};
    for (int f = 0; f < F; ++f) {
        dev_grid_tmp[f] = (host && want_grid[f]) || (ref_order && n_products > 0 && !(want_grid[f] && !host));
        if (dev_grid_tmp[f]) out_bytes += gbytes;
    }
    for (int i = 0; i < n_products; ++i)
        if (!a->products || (!a->products[i].out && !a->products[i].image)) return fail(RG_ERR_INVALID, "product output pointer is NULL");
    ImageStage images;
    RG_TRY(stage_images(ctx, n_products, a->products, host, F, ncol, &images));
    if (host) {
        for (int i = 0; i < n_products; ++i) {
            plane_off[i] = out_bytes;
            out_bytes += align256(product_plane_bytes(a->products[i], F, ncol));
        }
    }
    RG_TRY(ensure(ctx, ctx->stage_out, out_bytes));

    ApplyParams ap{};
    ap.indptr = g->indptr;
    ap.pairs = g->pairs;
    ap.records = pk.records;
    ap.records_b = pk.records_b;
    if (ctx->apply_variant == 2 && g->sell == nullptr)      // the interleaved copy is only built for the A/B kernel
        RG_TRY(build_sell(ctx, const_cast<Geometry*>(g)));
    ap.sell = g->sell;
    ap.slice_base = g->slice_base;
    ap.slices_per_level = g->slices_per_level;
    ap.null_gate = (uint32_t)G;
    ap.nonfinite = pk.nonfinite;
    ap.epoch = pk.epoch;
    RG_TRY(bind_record_textures(ctx, pk.records, pk.records_b, F, G));
    ap.tex_a = ctx->tex_a;
    ap.tex_b = ctx->tex_b;
    ap.ncol = ncol;
    ap.nx = g->grid.nx;
    ap.ny = g->grid.ny;
    ap.z_begin = g->grid.z_begin;
    ap.n_fields = F;
    ap.fill = a->fill_value;
    char* out_cur = (char*)ctx->stage_out.ptr;
    float* grid_dev[RG_MAX_FIELDS] = {};
    for (int f = 0; f < F; ++f) {
        if (dev_grid_tmp[f]) { grid_dev[f] = (float*)out_cur; out_cur += gbytes; }
        else if (want_grid[f]) grid_dev[f] = a->grid_out[f];
        ap.grid_out[f] = grid_dev[f];
    }
    std::vector<void*> planes;
    if (host && n_products > 0) {
        planes.resize(n_products);
        for (int i = 0; i < n_products; ++i) planes[i] = (char*)ctx->stage_out.ptr + plane_off[i];
    }
    int zlo = 0, zhi = -1;
    RG_TRY(make_product_params(g->grid, n_products, a->products, g->x_ax, g->y_ax, host ? planes.data() : nullptr,
                               &images, &ap.prod, &zlo, &zhi));

    if (ref_order) {
        // exact mode: un-fused.  Grid with NumPy's summation order, then the stand-alone product kernel.
        ap.lz_first = 0;
        ap.lz_last = g->n_levels;
        ProductParams pp = ap.prod;
        ap.prod.any = 0;
        bool any = false;
        for (int f = 0; f < F; ++f) any |= ap.grid_out[f] != nullptr;
        if (any) RG_TRY(nearest_gate ? launch_apply_nearest(ctx, g, ap) : launch_apply(ctx, g, ap, true));
        if (n_products > 0) RG_TRY(launch_products(ctx, g->grid, F, grid_dev, pp));
    } else {
        // walk only the levels somebody needs
        int first = 0, last = g->n_levels;
        if (!any_grid) {
            first = std::max(0, zlo - g->grid.z_begin);
            last = std::min(g->n_levels, zhi - g->grid.z_begin + 1);
            if (last < first) last = first;
        }
        ap.lz_first = first;
        ap.lz_last = last;
        if (any_grid || n_products > 0) RG_TRY(launch_apply(ctx, g, ap, false));
    }

    if (host) {
        for (int f = 0; f < F; ++f)
            if (want_grid[f]) RG_CUDA(cudaMemcpyAsync(a->grid_out[f], grid_dev[f], (size_t)V * 4, cudaMemcpyDeviceToHost, ctx->stream));
        for (int i = 0; i < n_products; ++i)
            if (a->products[i].out)
                RG_CUDA(cudaMemcpyAsync(a->products[i].out, planes[i], product_plane_bytes(a->products[i], F, ncol),
                                        cudaMemcpyDeviceToHost, ctx->stream));
        RG_TRY(fetch_images(ctx, n_products, a->products, F, ncol, images));
        RG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return RG_OK;
}

}  // extern "C"
