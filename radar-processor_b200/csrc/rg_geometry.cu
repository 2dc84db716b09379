// K1-K3: the per-voxel radius-of-influence neighbour table, built on the GPU.
//
// Reference being replaced: src/radar_grid/compute.py:18-103 (_process_single_level) and :106-284
// (compute_grid_geometry).  There, every z-level worker rebuilds a cKDTree over the TOA-valid gates and a
// Python loop ball-queries it voxel by voxel; the tree is only a candidate generator, the neighbour set is
//     { gate : ((dx*dx + dy*dy) + dz*dz) < r*r },   r = max(min_radius, sqrt((x*x + y*y) + z*z) * beam_factor)
// evaluated in float64 on float32-valued coordinates (compute.py:46-47, 69-74).  Here:
//   K1  counting-sort binning of the gates into a uniform cell grid (deterministic: gate id order inside a cell)
//   K2  one warp per voxel scans the cell rows its sphere touches (x-extent clipped per row) and counts the
//       gates that pass the same float64 test, with every operation individually rounded (no FMA contraction)
//   --  exclusive scan of the counts -> indptr
//   K3  the same scan again, writing {gate id, float32 weight} pairs (compute.py:82-87)
// The cell scan is a superset generator exactly like the KD-tree is, so the sets are identical.

#include <math.h>

#include <algorithm>
#include <vector>

#include "rg_internal.cuh"

namespace rg {

namespace {

constexpr int kScanTile = 4096;           // elements per CTA in the scan kernels (1024 threads x 4)
constexpr uint32_t kLongRun = 24;         // candidate runs at least this long are scanned by the whole warp
constexpr int64_t kMaxCells = 48ll << 20;

struct CellGrid {
    double ox, oy, oz, cell, inv_cell;
    int32_t ncx, ncy, ncz;
};

struct BinParams {
    const float* gx;
    const float* gy;
    const float* gz;
    int64_t n_gates;
    float radar_alt_f32, toa_f32;
    CellGrid cg;
    uint32_t* cell_of;                    // [n_gates]  cell id or kInvalidCell
    uint32_t* cell_count;                 // [ncell]
    unsigned long long* n_nonfinite;
};

__device__ __forceinline__ int cell_coord(double v, double origin, double inv_cell)
{
    // monotone in v; clamped so that far-away gates cannot overflow the int conversion
    const double f = floor(__dmul_rn(__dsub_rn(v, origin), inv_cell));
    return (int)fmin(fmax(f, -2.0), 2147483000.0);
}

// K1a: TOA cull (compute.py:182,193: float32 subtraction, float32 compare) + cell id + histogram
__global__ void __launch_bounds__(256) bin_count_kernel(const __grid_constant__ BinParams p)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= p.n_gates) return;
    const float x = __ldg(p.gx + g), y = __ldg(p.gy + g);
    const float zr = __fsub_rn(__ldg(p.gz + g), p.radar_alt_f32);
    uint32_t cell = kInvalidCell;
    if (zr <= p.toa_f32) {
        if (!(isfinite(x) && isfinite(y) && isfinite(zr))) {
            atomicAdd(p.n_nonfinite, 1ull);           // the reference's cKDTree refuses non-finite data
        } else {
            const int cx = cell_coord((double)x, p.cg.ox, p.cg.inv_cell);
            const int cy = cell_coord((double)y, p.cg.oy, p.cg.inv_cell);
            const int cz = cell_coord((double)zr, p.cg.oz, p.cg.inv_cell);
            if (cx >= 0 && cx < p.cg.ncx && cy >= 0 && cy < p.cg.ncy && cz >= 0 && cz < p.cg.ncz) {
                cell = (uint32_t)(((int64_t)cz * p.cg.ncy + cy) * p.cg.ncx + cx);
                atomicAdd(p.cell_count + cell, 1u);
            }
        }
    }
    p.cell_of[g] = cell;
}

// K1c: unordered scatter of gate ids into their cell's range
__global__ void __launch_bounds__(256) bin_scatter_kernel(const uint32_t* __restrict__ cell_of, int64_t n_gates,
                                                          const uint32_t* __restrict__ cell_start,
                                                          uint32_t* __restrict__ cell_fill, uint32_t* __restrict__ slot_ids)
{
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_gates) return;
    const uint32_t c = cell_of[g];
    if (c == kInvalidCell) return;
    const uint32_t pos = cell_start[c] + atomicAdd(cell_fill + c, 1u);
    slot_ids[pos] = (uint32_t)g;
}

// K1d: make the order inside every cell deterministic (ascending gate id) and gather the coordinates
__global__ void __launch_bounds__(256) bin_rank_kernel(const uint32_t* __restrict__ slot_ids, int64_t n_binned,
                                                       const uint32_t* __restrict__ cell_of,
                                                       const uint32_t* __restrict__ cell_start,
                                                       const float* __restrict__ gx, const float* __restrict__ gy,
                                                       const float* __restrict__ gz, float radar_alt_f32,
                                                       float4* __restrict__ sorted)
{
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= n_binned) return;
    const uint32_t id = slot_ids[slot];
    const uint32_t c = cell_of[id];
    const uint32_t s = cell_start[c], e = cell_start[c + 1];
    uint32_t rank = 0;
    for (uint32_t j = s; j < e; ++j) rank += __ldg(slot_ids + j) < id ? 1u : 0u;
    sorted[s + rank] = make_float4(__ldg(gx + id), __ldg(gy + id), __fsub_rn(__ldg(gz + id), radar_alt_f32),
                                   __uint_as_float(id));
}

// ---- exclusive scan of uint32 counts (64-bit total) ---------------------------------------------------
__device__ __forceinline__ unsigned long long block_reduce_u64(unsigned long long v, unsigned long long* smem)
{
    for (int off = 16; off >= 1; off >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, off);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) smem[w] = v;
    __syncthreads();
    unsigned long long t = (threadIdx.x < (blockDim.x >> 5)) ? smem[threadIdx.x] : 0ull;
    if (w == 0) {
        for (int off = 16; off >= 1; off >>= 1) t += __shfl_xor_sync(0xFFFFFFFFu, t, off);
        if (l == 0) smem[0] = t;
    }
    __syncthreads();
    t = smem[0];
    __syncthreads();
    return t;
}

__global__ void __launch_bounds__(1024) scan_tile_sums_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                              unsigned long long* __restrict__ tile_sums)
{
    __shared__ unsigned long long smem[32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * 4;
    unsigned long long v = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (base + k < n) v += in[base + k];
    const unsigned long long t = block_reduce_u64(v, smem);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = t;
}

// single CTA: exclusive scan of the tile sums in place, grand total to *total
__global__ void __launch_bounds__(1024) scan_tile_offsets_kernel(unsigned long long* __restrict__ tile_sums, int64_t n_tiles,
                                                                 unsigned long long* __restrict__ total)
{
    __shared__ unsigned long long warp_tot[32];
    __shared__ unsigned long long carry, chunk_total;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int64_t base = 0; base < n_tiles; base += 1024) {
        const int64_t i = base + threadIdx.x;
        const unsigned long long v = i < n_tiles ? tile_sums[i] : 0ull;
        unsigned long long incl = v;
        for (int off = 1; off < 32; off <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
            if (l >= off) incl += t;
        }
        if (l == 31) warp_tot[w] = incl;
        __syncthreads();
        if (w == 0) {
            const unsigned long long t = warp_tot[l];
            unsigned long long ti = t;
            for (int off = 1; off < 32; off <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, ti, off);
                if (l >= off) ti += u;
            }
            warp_tot[l] = ti - t;          // exclusive prefix of the warp totals
            if (l == 31) chunk_total = ti;
        }
        __syncthreads();
        if (i < n_tiles) tile_sums[i] = carry + warp_tot[w] + (incl - v);
        __syncthreads();
        if (threadIdx.x == 0) carry += chunk_total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(1024) scan_apply_kernel(const uint32_t* __restrict__ in, int64_t n,
                                                          const unsigned long long* __restrict__ tile_offsets,
                                                          uint32_t* __restrict__ out)
{
    __shared__ uint32_t warp_tot[32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * 4;
    uint32_t v[4];
    uint32_t mine = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = base + k < n ? in[base + k] : 0u;
        mine += v[k];
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    uint32_t incl = mine;
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (l >= off) incl += t;
    }
    if (l == 31) warp_tot[w] = incl;
    __syncthreads();
    if (w == 0) {
        const uint32_t t = warp_tot[l];
        uint32_t ti = t;
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, ti, off);
            if (l >= off) ti += u;
        }
        warp_tot[l] = ti - t;
    }
    __syncthreads();
    uint32_t run = (uint32_t)tile_offsets[blockIdx.x] + warp_tot[w] + (incl - mine);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

__global__ void scan_write_total_kernel(const unsigned long long* total, uint32_t* out_last)
{
    *out_last = (uint32_t)(*total);
}

}  // namespace

// out[0..n] = exclusive scan of in[0..n-1]; *total_host = sum.  `tmp` must hold ceil(n/4096)+1 u64.
int exclusive_scan_u32(Context* ctx, const uint32_t* in, uint32_t* out, int64_t n, unsigned long long* tmp,
                       uint64_t* total_host)
{
    const int64_t n_tiles = (n + kScanTile - 1) / kScanTile;
    unsigned long long* total_dev = tmp + n_tiles;
    if (n_tiles > 0) {
        scan_tile_sums_kernel<<<(unsigned)n_tiles, 1024, 0, ctx->stream>>>(in, n, tmp);
        ctx->launches++;
    }
    scan_tile_offsets_kernel<<<1, 1024, 0, ctx->stream>>>(tmp, n_tiles, total_dev);
    ctx->launches++;
    if (n_tiles > 0) {
        scan_apply_kernel<<<(unsigned)n_tiles, 1024, 0, ctx->stream>>>(in, n, tmp, out);
        ctx->launches++;
    }
    scan_write_total_kernel<<<1, 1, 0, ctx->stream>>>(total_dev, out + n);
    ctx->launches++;
    RG_CUDA(cudaGetLastError());
    unsigned long long t = 0;
    RG_CUDA(cudaMemcpyAsync(&t, total_dev, sizeof(t), cudaMemcpyDeviceToHost, ctx->stream));
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    *total_host = t;
    return RG_OK;
}

namespace {

// ---- K2 / K3: warp-per-voxel neighbour search -----------------------------------------------------------
struct NeighbourParams {
    const float4* sorted;                 // binned gates {x, y, z - radar_alt, id bits}
    const uint32_t* cell_start;
    CellGrid cg;
    const float* x_ax;
    const float* y_ax;
    const float* z_ax;
    int32_t nx, ny, z_begin;
    int64_t ncol, n_rows;
    double min_radius, beam_factor;
    int32_t weighting;
    uint32_t* counts;                     // K2 output
    // K2 in level-census mode (rg_geometry_level_pairs): every `col_stride`-th column in x and y is counted and the
    // counts are added up per level instead of being stored per row
    unsigned long long* level_pairs;      // [n_levels] or nullptr
    int32_t col_stride, nxs, nys;
    const uint32_t* indptr;               // K3 input
    uint2* pairs;                         // K3 output
    unsigned long long* n_candidates;
};

__device__ __forceinline__ float neighbour_weight(int weighting, double d2, double r2)
{
    if (weighting == RG_W_BARNES2)        // compute.py:83
        return __double2float_rn(__dadd_rn(exp(__ddiv_rn(-d2, __ddiv_rn(r2, 4.0))), 1e-5));
    if (weighting == RG_W_CRESSMAN)       // compute.py:85
        return __double2float_rn(__ddiv_rn(__dsub_rn(r2, d2), __dadd_rn(r2, d2)));
    if (weighting == RG_W_DIST2) return __double2float_rn(d2);     // nearest-gate tables carry the squared distance
    return 1.0f;                          // compute.py:87 ('nearest' = every gate in the ROI, weight 1)
}

template <bool FILL>
__global__ void __launch_bounds__(256) neighbour_kernel(const __grid_constant__ NeighbourParams p)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    __shared__ unsigned long long cta_candidates;
    if (threadIdx.x == 0) cta_candidates = 0;
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    unsigned long long tested = 0;

    if (row < p.n_rows) {
        int lz, iy, ix;
        if (!FILL && p.level_pairs != nullptr) {
            const int64_t per_level = (int64_t)p.nxs * p.nys;
            lz = (int)(row / per_level);
            const int64_t c = row - (int64_t)lz * per_level;
            iy = (int)(c / p.nxs) * p.col_stride;
            ix = (int)(c % p.nxs) * p.col_stride;
        } else {
            lz = (int)(row / p.ncol);
            const int64_t c = row - (int64_t)lz * p.ncol;
            iy = (int)(c / p.nx);
            ix = (int)(c - (int64_t)iy * p.nx);
        }
        // compute.py:43,189-190: float32 axis values widened to float64
        const double vx = (double)__ldg(p.x_ax + ix);
        const double vy = (double)__ldg(p.y_ax + iy);
        const double vz = (double)__ldg(p.z_ax + p.z_begin + lz);
        // compute.py:46-47
        const double dist = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(vx, vx), __dmul_rn(vy, vy)), __dmul_rn(vz, vz)));
        const double r = fmax(p.min_radius, __dmul_rn(dist, p.beam_factor));
        const double r2 = __dmul_rn(r, r);
        // conservative (padded) search radius for the cell scan; the exact test below decides membership
        const double rp = r * (1.0 + 1e-9) + 1e-3;
        const CellGrid& cg = p.cg;
        const int czlo = max(0, min(cg.ncz - 1, cell_coord(vz - rp, cg.oz, cg.inv_cell)));
        const int czhi = max(0, min(cg.ncz - 1, cell_coord(vz + rp, cg.oz, cg.inv_cell)));
        const int cylo = max(0, min(cg.ncy - 1, cell_coord(vy - rp, cg.oy, cg.inv_cell)));
        const int cyhi = max(0, min(cg.ncy - 1, cell_coord(vy + rp, cg.oy, cg.inv_cell)));
        const int yspan = cyhi - cylo + 1;
        const int n_scan = (czhi - czlo + 1) * yspan;

        uint32_t found = 0;                               // warp-uniform running count
        uint32_t out_base = 0;
        if (FILL) out_base = __ldg(p.indptr + row);

        auto test_candidate = [&](bool in, uint32_t k) {
            bool keep = false;
            float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
            double d2 = 0.0;
            if (in) {
                g = __ldg(p.sorted + k);
                // compute.py:69-74, float64, each operation rounded on its own
                const double dx = __dsub_rn((double)g.x, vx);
                const double dy = __dsub_rn((double)g.y, vy);
                const double dz = __dsub_rn((double)g.z, vz);
                d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                keep = d2 < r2;
            }
            const unsigned km = __ballot_sync(kFull, keep);
            if (FILL && keep) {
                const uint32_t pos = out_base + found + __popc(km & lt_mask);
                p.pairs[pos] = make_uint2(__float_as_uint(g.w), __float_as_uint(neighbour_weight(p.weighting, d2, r2)));
            }
            found += __popc(km);
        };

        for (int base = 0; base < n_scan; base += 32) {
            const int i = base + lane;
            uint32_t s = 0, e = 0;
            if (i < n_scan) {
                const int cz = czlo + i / yspan;
                const int cy = cylo + i % yspan;
                const double zlo = cg.oz + cz * cg.cell, ylo = cg.oy + cy * cg.cell;
                const double dzm = fmax(0.0, fmax(zlo - vz, vz - (zlo + cg.cell)));
                const double dym = fmax(0.0, fmax(ylo - vy, vy - (ylo + cg.cell)));
                const double h2 = rp * rp - dzm * dzm - dym * dym;
                if (h2 > 0.0) {
                    const double hx = sqrt(h2) + 1e-3;
                    const int cxlo = max(0, min(cg.ncx - 1, cell_coord(vx - hx, cg.ox, cg.inv_cell)));
                    const int cxhi = max(0, min(cg.ncx - 1, cell_coord(vx + hx, cg.ox, cg.inv_cell)));
                    const size_t rb = ((size_t)cz * cg.ncy + cy) * (size_t)cg.ncx;
                    s = __ldg(p.cell_start + rb + cxlo);
                    e = __ldg(p.cell_start + rb + cxhi + 1);
                }
            }
            const uint32_t len = e - s;
            tested += len;

            // long candidate runs: all 32 lanes stride over the run (coalesced 16-byte loads)
            unsigned longm = __ballot_sync(kFull, len >= kLongRun);
            while (longm) {
                const int src = __ffs(longm) - 1;
                longm &= longm - 1;
                const uint32_t hs = __shfl_sync(kFull, s, src);
                const uint32_t he = __shfl_sync(kFull, e, src);
                for (uint32_t k0 = hs; k0 < he; k0 += 32) test_candidate(k0 + lane < he, k0 + lane);
            }
            // short runs: every lane walks its own run, in lockstep
            const uint32_t slen = len < kLongRun ? len : 0u;
            const uint32_t smax = __reduce_max_sync(kFull, slen);
            for (uint32_t j = 0; j < smax; ++j) test_candidate(j < slen, s + j);
        }
        if (!FILL && lane == 0) {
            if (p.level_pairs != nullptr) { if (found) atomicAdd(p.level_pairs + lz, (unsigned long long)found); }
            else p.counts[row] = found;
        }
    }

    if (!FILL) {
        for (int off = 16; off >= 1; off >>= 1) tested += __shfl_xor_sync(kFull, tested, off);
        if (lane == 0 && tested) atomicAdd(&cta_candidates, tested);
        __syncthreads();
        if (threadIdx.x == 0 && cta_candidates) atomicAdd(p.n_candidates, cta_candidates);
    }
}

// ---- row ordering ----------------------------------------------------------------------------------------
// The cell scan emits a row's gates cell by cell.  Sorting every row by gate id turns it into runs of
// consecutive bins of the same ray, so that neighbouring lanes of the apply kernel gather neighbouring gate
// records (fewer 128-byte lines per load instruction, better L1 reuse).  One warp per row, bitonic network in
// shared memory; rows longer than kSortCap (a handful next to the radar) keep the scan order.
constexpr int kSortCap = 1024;

__global__ void __launch_bounds__(128) sort_rows_kernel(const uint32_t* __restrict__ indptr, uint2* __restrict__ pairs,
                                                        int64_t n_rows)
{
    __shared__ uint2 buf[4][kSortCap];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int64_t row = (int64_t)blockIdx.x * 4 + w;
    if (row >= n_rows) return;
    const uint32_t s = indptr[row], e = indptr[row + 1];
    const uint32_t len = e - s;
    if (len < 2 || len > (uint32_t)kSortCap) return;
    uint32_t n2 = 2;
    while (n2 < len) n2 <<= 1;
    uint2* b = buf[w];
    for (uint32_t i = lane; i < n2; i += 32) b[i] = i < len ? pairs[s + i] : make_uint2(0xFFFFFFFFu, 0u);
    __syncwarp();
    for (uint32_t k = 2; k <= n2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = lane; i < n2; i += 32) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const uint2 a = b[i], c = b[ixj];
                    const bool ascending = (i & k) == 0;
                    if ((a.x > c.x) == ascending) { b[i] = c; b[ixj] = a; }
                }
            }
            __syncwarp();
        }
    }
    for (uint32_t i = lane; i < len; i += 32) pairs[s + i] = b[i];
}

// ---- row statistics ----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) row_stats_kernel(const uint32_t* __restrict__ indptr, int64_t n_rows,
                                                        unsigned long long* __restrict__ n_empty,
                                                        unsigned int* __restrict__ max_len)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t len = 0;
    bool empty = false;
    if (row < n_rows) {
        len = indptr[row + 1] - indptr[row];
        empty = len == 0;
    }
    const unsigned em = __ballot_sync(0xFFFFFFFFu, empty);
    const uint32_t wmax = __reduce_max_sync(0xFFFFFFFFu, len);
    if ((threadIdx.x & 31) == 0) {
        if (em) atomicAdd(n_empty, (unsigned long long)__popc(em));
        if (wmax) atomicMax(max_len, wmax);
    }
}

// ---- interleaved copy of the table for the thread-per-column apply kernel ---------------------------------
// A slice = 32 consecutive columns of one level = 32 consecutive CSR rows.  Inside a slice the pairs are stored
// step-major: all rows' element 0, then all rows' element 1, ... with rows that have run out simply skipped, so
// lane r of a warp reads element k of row r at  base + sum_{k'<k} active(k') + rank_k(r)  — the active lanes of
// every step read consecutive addresses (perfectly coalesced), nothing is padded, and the copy holds exactly
// sum(min(len, kSellCap)) pairs.
__global__ void __launch_bounds__(128) sell_count_kernel(const uint32_t* __restrict__ indptr, int64_t ncol, int n_levels,
                                                         int64_t slices_per_level, uint32_t* __restrict__ slice_count)
{
    const int lane = threadIdx.x & 31;
    const int64_t slice = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (slice >= slices_per_level * n_levels) return;
    const int64_t lz = slice / slices_per_level, sb = slice - lz * slices_per_level;
    const int64_t col = sb * 32 + lane;
    uint32_t n = 0;
    if (col < ncol) {
        const size_t row = (size_t)lz * (size_t)ncol + (size_t)col;
        n = min(indptr[row + 1] - indptr[row], kSellCap);
    }
    for (int off = 16; off >= 1; off >>= 1) n += __shfl_xor_sync(0xFFFFFFFFu, n, off);
    if (lane == 0) slice_count[slice] = n;
}

__global__ void __launch_bounds__(128) sell_fill_kernel(const uint32_t* __restrict__ indptr, const uint2* __restrict__ pairs,
                                                        int64_t ncol, int n_levels, int64_t slices_per_level,
                                                        const uint32_t* __restrict__ slice_base, uint2* __restrict__ sell)
{
    const int lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    const int64_t slice = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (slice >= slices_per_level * n_levels) return;
    const int64_t lz = slice / slices_per_level, sb = slice - lz * slices_per_level;
    const int64_t col = sb * 32 + lane;
    uint32_t s = 0, n = 0;
    if (col < ncol) {
        const size_t row = (size_t)lz * (size_t)ncol + (size_t)col;
        s = indptr[row];
        n = min(indptr[row + 1] - s, kSellCap);
    }
    uint32_t base = slice_base[slice];
    const uint32_t kmax = __reduce_max_sync(0xFFFFFFFFu, n);
    for (uint32_t k = 0; k < kmax; ++k) {
        const bool act = k < n;
        const unsigned m = __ballot_sync(0xFFFFFFFFu, act);
        if (act) sell[base + __popc(m & lt_mask)] = pairs[s + k];
        base += __popc(m);
    }
}

// ------------------------------------------------------------------------------------------------------
// Warp-slice copy of the table for the column-group apply kernel (see rg_internal.cuh / rg_apply.cu).
// Rows longer than kHeavyRow stay out of it (the kernel sums them from the CSR copy with the whole warp).
// ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) quad_count_kernel(const uint32_t* __restrict__ indptr, int nx, int64_t n_slices, int R, int W,
                                                         int quads_x, uint32_t heavy_len, uint32_t* __restrict__ counts,
                                                         uint8_t* __restrict__ heavy)
{
    const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n_slices) return;
    const int64_t line = q / quads_x;                          // lz * ny + cy
    const int qx = (int)(q - line * quads_x);
    uint32_t m = 0;
    uint8_t hv = 0;
    for (int k = 0; k < R; ++k) {
        const int cx = qx * R + k;
        if (cx >= nx) break;
        const size_t row = (size_t)line * (size_t)nx + (size_t)cx;
        const uint32_t len = indptr[row + 1] - indptr[row];
        if (len > heavy_len) hv = 1;
        else m = max(m, (len + (uint32_t)W - 1u) / (uint32_t)W);
    }
    counts[q] = m;
    heavy[q] = hv;
}

__global__ void __launch_bounds__(128) quad_fill_kernel(const uint32_t* __restrict__ indptr, const uint2* __restrict__ pairs, int nx,
                                                        int64_t n_slices, int R, int W, int quads_x, uint32_t heavy_len,
                                                        const uint32_t* __restrict__ offs, const uint32_t* __restrict__ counts,
                                                        const uint8_t* __restrict__ heavy, uint32_t null_gate,
                                                        uint2* __restrict__ quads, uint32_t* __restrict__ quad_ptr)
{
    const int lane = threadIdx.x & 31;
    const int64_t q = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    if (q >= n_slices) return;
    const int64_t line = q / quads_x;
    const int qx = (int)(q - line * quads_x);
    const uint32_t m = counts[q], o = offs[q];
    if (lane == 0) {
        quad_ptr[q] = (o << 1) | heavy[q];
        if (q == n_slices - 1) quad_ptr[n_slices] = (o + m) << 1;
    }
    const int cx = qx * R + lane / W, k0 = lane % W;
    uint32_t s = 0, len = 0;
    if (cx < nx) {
        const size_t row = (size_t)line * (size_t)nx + (size_t)cx;
        s = indptr[row];
        len = indptr[row + 1] - s;
        if (len > heavy_len) len = 0;
    }
    uint2* dst = quads + (size_t)o * 32 + lane;
    for (uint32_t j = 0; j < m; ++j) {
        const uint32_t k = j * (uint32_t)W + (uint32_t)k0;
        dst[(size_t)j * 32] = k < len ? pairs[s + k] : make_uint2(null_gate, 0u);
    }
}

// ------------------------------------------------------------------------------------------------------
// Column-pair ("duo") copy of the table for apply_duo_kernel (rg_duo.cu; layout in rg_internal.cuh).
// One thread per group = the columns (x, 2yp) and (x, 2yp + 1) at one level; the 8 groups of a slice sit in 8 consecutive
// threads.  The two rows are merged by gate id (both are sorted: sort_rows_kernel; a table imported in another order
// raises `unsorted` and does not get a duo copy).  Rows longer than kHeavyRow stay out (heavy_rows_kernel sums them).
// ------------------------------------------------------------------------------------------------------
struct DuoGroup {
    uint32_t s0, e0, s1, e1;
    int k;                                 // group within the slice
    int64_t q;                             // slice
    bool heavy;
};

__device__ __forceinline__ DuoGroup duo_group(const uint32_t* __restrict__ indptr, int nx, int ny, int nyp, int qxn, int64_t n_groups,
                                              uint32_t heavy_len)
{
    DuoGroup g{0, 0, 0, 0, 0, 0, false};
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    g.q = t >> 3;
    g.k = (int)(t & 7);
    if (t >= n_groups) return g;
    const int64_t line = g.q / qxn;                            // lz * nyp + yp
    const int qx = (int)(g.q - line * qxn);
    const int64_t lz = line / nyp;
    const int yp = (int)(line - lz * nyp);
    const int x = qx * 8 + g.k, y0 = 2 * yp;
    if (x >= nx) return g;
    const size_t row0 = ((size_t)lz * (size_t)ny + (size_t)y0) * (size_t)nx + (size_t)x;
    g.s0 = indptr[row0];
    g.e0 = indptr[row0 + 1];
    if (g.e0 - g.s0 > heavy_len) { g.heavy = true; g.e0 = g.s0; }
    if (y0 + 1 < ny) {
        g.s1 = indptr[row0 + (size_t)nx];
        g.e1 = indptr[row0 + (size_t)nx + 1];
        if (g.e1 - g.s1 > heavy_len) { g.heavy = true; g.e1 = g.s1; }
    }
    return g;
}

__global__ void __launch_bounds__(256) duo_count_kernel(const uint32_t* __restrict__ indptr, const uint2* __restrict__ pairs, int nx, int ny,
                                                        int nyp, int qxn, int64_t n_groups, uint32_t heavy_len,
                                                        uint32_t* __restrict__ counts, uint8_t* __restrict__ heavy,
                                                        unsigned int* __restrict__ unsorted)
{
    const DuoGroup g = duo_group(indptr, nx, ny, nyp, qxn, n_groups, heavy_len);
    uint32_t i = g.s0, j = g.s1, n = 0;
    uint32_t prev0 = 0, prev1 = 0;
    bool bad = false;
    while (i < g.e0 || j < g.e1) {
        const uint32_t a = i < g.e0 ? pairs[i].x : 0xFFFFFFFFu, b = j < g.e1 ? pairs[j].x : 0xFFFFFFFFu;
        if (a <= b) { bad |= i > g.s0 && a <= prev0; prev0 = a; ++i; }
        if (b <= a) { bad |= j > g.s1 && b <= prev1; prev1 = b; ++j; }
        ++n;
    }
    if (bad) atomicAdd(unsorted, 1u);
    uint32_t m = (n + 3u) / 4u;
    uint32_t hv = g.heavy ? 1u : 0u;
#pragma unroll
    for (int off = 1; off <= 4; off <<= 1) {
        m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, off));
        hv |= __shfl_xor_sync(0xFFFFFFFFu, hv, off);
    }
    if (g.k == 0 && g.q * 8 < n_groups) {
        counts[g.q] = m;
        heavy[g.q] = (uint8_t)hv;
    }
}

__global__ void __launch_bounds__(256) duo_fill_kernel(const uint32_t* __restrict__ indptr, const uint2* __restrict__ pairs, int nx, int ny,
                                                       int nyp, int qxn, int64_t n_groups, uint32_t heavy_len,
                                                       const uint32_t* __restrict__ offs, const uint32_t* __restrict__ counts,
                                                       const uint8_t* __restrict__ heavy, uint32_t null_gate,
                                                       uint32_t* __restrict__ slots, uint32_t* __restrict__ duo_ptr)
{
    const DuoGroup g = duo_group(indptr, nx, ny, nyp, qxn, n_groups, heavy_len);
    if (g.q * 8 >= n_groups) return;
    const uint32_t m = counts[g.q], o = offs[g.q];
    if (g.k == 0) {
        duo_ptr[g.q] = (o << 1) | heavy[g.q];
        if ((g.q + 1) * 8 >= n_groups) duo_ptr[g.q + 1] = (o + m) << 1;
    }
    const uint32_t absent = 0x80000000u;                       // -0.0f: see rg_duo.cu
    uint32_t i = g.s0, j = g.s1, n = 0;
    auto put = [&](uint32_t gate, uint32_t w0, uint32_t w1) {
        uint32_t* slot = slots + (size_t)(o + (n >> 2)) * 96;
        const uint32_t l = (uint32_t)g.k * 4u + (n & 3u);
        slot[l] = gate;
        slot[32 + 2 * l] = w0;
        slot[32 + 2 * l + 1] = w1;
        ++n;
    };
    while (i < g.e0 || j < g.e1) {
        const uint2 a = i < g.e0 ? pairs[i] : make_uint2(0xFFFFFFFFu, 0u), b = j < g.e1 ? pairs[j] : make_uint2(0xFFFFFFFFu, 0u);
        if (a.x == b.x) { put(a.x, a.y, b.y); ++i; ++j; }
        else if (a.x < b.x) { put(a.x, a.y, absent); ++i; }
        else { put(b.x, absent, b.y); ++j; }
    }
    while (n < m * 4u) put(null_gate, absent, absent);        // idle lanes: the all-masked record
}

// ---- heavy rows --------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) heavy_list_kernel(const uint32_t* __restrict__ indptr, int64_t n_rows, uint32_t heavy_len,
                                                         uint32_t capacity, uint32_t* __restrict__ rows, unsigned int* __restrict__ count)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_rows) return;
    const uint32_t s = indptr[row], e = indptr[row + 1];
    if (e - s > heavy_len) {
        const unsigned int k = atomicAdd(count, 1u);
        if (k < capacity) { rows[3 * k] = (uint32_t)row; rows[3 * k + 1] = s; rows[3 * k + 2] = e; }
    }
}

template <typename T>
struct DevBuf {
    T* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    cudaError_t alloc(size_t n) { return cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)); }
};

}  // namespace

void linspace_f32(double start, double stop, int num, float* out)
{
    // numpy.linspace(start, stop, num, dtype=float32): float64 arange * step + start, last = stop, then cast
    if (num <= 0) return;
    const int div = num - 1;
    const double delta = stop - start;
    if (div > 0) {
        const double step = delta / div;
        for (int i = 0; i < num; ++i) {
            double y = (double)i;
            if (step == 0.0) y = (y / div) * delta;
            else y = y * step;
            out[i] = (float)(y + start);
        }
        out[num - 1] = (float)stop;
    } else {
        out[0] = (float)(0.0 * delta + start);
    }
}

int build_sell(Context* ctx, Geometry* g)
{
    g->slices_per_level = (g->ncol + 31) / 32;
    const int64_t n_slices = g->slices_per_level * g->n_levels;
    if (g->sell) { cudaFree(g->sell); g->sell = nullptr; }
    if (g->slice_base) { cudaFree(g->slice_base); g->slice_base = nullptr; }
    RG_CUDA(cudaMalloc(&g->slice_base, ((size_t)n_slices + 1) * sizeof(uint32_t)));
    DevBuf<uint32_t> counts;
    DevBuf<unsigned long long> tmp;
    RG_CUDA(counts.alloc((size_t)n_slices));
    RG_CUDA(tmp.alloc((size_t)(n_slices / kScanTile + 4)));
    const unsigned blocks = (unsigned)((n_slices + 3) / 4);
    if (n_slices > 0) {
        sell_count_kernel<<<blocks, 128, 0, ctx->stream>>>(g->indptr, g->ncol, g->n_levels, g->slices_per_level, counts.p);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
    }
    uint64_t total = 0;
    RG_TRY(exclusive_scan_u32(ctx, counts.p, g->slice_base, n_slices, tmp.p, &total));
    g->n_sell = (int64_t)total;
    RG_CUDA(cudaMalloc(&g->sell, std::max<size_t>((size_t)total, 1) * sizeof(uint2)));
    if (n_slices > 0 && total > 0) {
        sell_fill_kernel<<<blocks, 128, 0, ctx->stream>>>(g->indptr, g->pairs, g->ncol, g->n_levels, g->slices_per_level,
                                                         g->slice_base, g->sell);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
    }
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    g->info.device_bytes = (int64_t)(((size_t)g->n_rows + 1) * 4 + (size_t)g->n_pairs * 8 + (size_t)total * 8 + ((size_t)n_slices + 1) * 4);
    return RG_OK;
}

int ensure_quads(Context* ctx, Geometry* g, int W, const Geometry::QuadCopy** out)
{
    const int slot = W == 4 ? 0 : W == 8 ? 1 : W == 16 ? 2 : 3;
    std::lock_guard<std::mutex> lock(g->quad_mu);
    Geometry::QuadCopy& qc = g->quad[slot];
    *out = &qc;
    if (qc.ptr != nullptr) return RG_OK;
    if (qc.n_slots < 0) { *out = nullptr; return RG_OK; }      // did not fit earlier
    const int R = 32 / W, nx = g->grid.nx, ny = g->grid.ny;
    const int quads_x = (nx + R - 1) / R;
    const int64_t n_slices = (int64_t)g->n_levels * ny * quads_x;
    // the kernel hands rows longer than kHeavyRow to the whole warp; a 32-lane group IS the whole warp
    const uint32_t heavy_len = W == 32 ? 0xFFFFFFFFu : kHeavyRow;
    DevBuf<uint32_t> counts, offs, qptr;
    DevBuf<uint8_t> heavy;
    DevBuf<unsigned long long> tmp;
    RG_CUDA(qptr.alloc((size_t)n_slices + 1));
    RG_CUDA(cudaMemsetAsync(qptr.p, 0, ((size_t)n_slices + 1) * sizeof(uint32_t), ctx->stream));
    RG_CUDA(counts.alloc((size_t)n_slices));
    RG_CUDA(offs.alloc((size_t)n_slices + 1));
    RG_CUDA(heavy.alloc((size_t)n_slices));
    RG_CUDA(tmp.alloc((size_t)(n_slices / kScanTile + 4)));
    if (n_slices > 0) {
        quad_count_kernel<<<(unsigned)((n_slices + 255) / 256), 256, 0, ctx->stream>>>(g->indptr, nx, n_slices, R, W, quads_x, heavy_len,
                                                                                      counts.p, heavy.p);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
    }
    uint64_t total = 0;
    RG_TRY(exclusive_scan_u32(ctx, counts.p, offs.p, n_slices, tmp.p, &total));
    if (total >= (1ull << 31)) return fail(RG_ERR_UNSUPPORTED, "warp-slice copy of the table exceeds 2^31 slots; use more z-slabs");
    // The copy is an accelerator, not a requirement: when it does not fit next to the table (a z-slab sized to fill
    // the GPU), the kernel keeps reading the CSR copy.
    size_t free_b = 0, total_b = 0;
    RG_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t need_b = (size_t)total * 32 * sizeof(uint2);
    if (need_b + (size_t)(1ull << 30) > free_b) {
        qc.n_slots = -1;                                       // remembered: do not try again for this table
        *out = nullptr;
        return RG_OK;
    }
    DevBuf<uint2> quads;
    RG_CUDA(quads.alloc((size_t)total * 32));
    if (n_slices > 0) {
        quad_fill_kernel<<<(unsigned)((n_slices + 3) / 4), 128, 0, ctx->stream>>>(g->indptr, g->pairs, nx, n_slices, R, W, quads_x, heavy_len,
                                                                                 offs.p, counts.p, heavy.p, (uint32_t)g->n_gates, quads.p, qptr.p);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
    }
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    qc.quads = quads.p; quads.p = nullptr;
    qc.ptr = qptr.p; qptr.p = nullptr;
    qc.n_slots = (int64_t)total;
    qc.quads_x = quads_x;
    g->info.device_bytes += (int64_t)((size_t)total * 32 * sizeof(uint2) + ((size_t)n_slices + 1) * sizeof(uint32_t));
    return RG_OK;
}

int ensure_duo(Context* ctx, Geometry* g, const Geometry::DuoCopy** out)
{
    std::lock_guard<std::mutex> lock(g->quad_mu);
    Geometry::DuoCopy& dc = g->duo;
    *out = &dc;
    if (dc.ptr != nullptr) return RG_OK;
    *out = nullptr;
    if (dc.n_slots < 0) return RG_OK;                          // found unsuitable earlier
    const int nx = g->grid.nx, ny = g->grid.ny;
    const int nyp = (ny + 1) / 2, qxn = (nx + 7) / 8;
    const int64_t n_slices = (int64_t)g->n_levels * nyp * qxn, n_groups = n_slices * 8;
    if (n_slices == 0 || g->n_pairs == 0) { dc.n_slots = -1; return RG_OK; }
    DevBuf<uint32_t> counts, offs, dptr;
    DevBuf<uint8_t> heavy;
    DevBuf<unsigned long long> tmp;
    DevBuf<unsigned int> unsorted;
    RG_CUDA(dptr.alloc((size_t)n_slices + 1));
    RG_CUDA(counts.alloc((size_t)n_slices));
    RG_CUDA(offs.alloc((size_t)n_slices + 1));
    RG_CUDA(heavy.alloc((size_t)n_slices));
    RG_CUDA(tmp.alloc((size_t)(n_slices / kScanTile + 4)));
    RG_CUDA(unsorted.alloc(1));
    RG_CUDA(cudaMemsetAsync(unsorted.p, 0, sizeof(unsigned int), ctx->stream));
    const unsigned blocks = (unsigned)((n_groups + 255) / 256);
    duo_count_kernel<<<blocks, 256, 0, ctx->stream>>>(g->indptr, g->pairs, nx, ny, nyp, qxn, n_groups, kHeavyRow, counts.p, heavy.p,
                                                     unsorted.p);
    ctx->launches++;
    RG_CUDA(cudaGetLastError());
    uint64_t total = 0;
    RG_TRY(exclusive_scan_u32(ctx, counts.p, offs.p, n_slices, tmp.p, &total));
    unsigned int n_unsorted = 0;
    RG_CUDA(cudaMemcpy(&n_unsorted, unsorted.p, sizeof(n_unsorted), cudaMemcpyDeviceToHost));
    // Not for this table: rows in another order than by gate id (an imported table), more slots than a 31-bit slot index
    // holds, or no room next to the table.  Whether the merge pays is decided by the caller from the row lengths alone
    // (launch_apply: tables of short rows, where the pass is bound by latency and L1 traffic and the kernel won every
    // measurement -- cfg1: 0.89 entries per pair, 0.0383 against 0.0445 ms; cfg3: 0.74, 0.606 against 0.642 ms), so that
    // all z-slabs of one grid take the same kernel and stay bit-identical to the unsharded pass.
    size_t free_b = 0, total_b = 0;
    RG_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const size_t need_b = (size_t)total * 96 * sizeof(uint32_t);
    if (n_unsorted != 0 || total >= (1ull << 31) || need_b + (size_t)(1ull << 30) > free_b) {
        dc.n_slots = -1;
        return RG_OK;
    }
    DevBuf<uint32_t> slots;
    RG_CUDA(slots.alloc((size_t)total * 96));
    duo_fill_kernel<<<blocks, 256, 0, ctx->stream>>>(g->indptr, g->pairs, nx, ny, nyp, qxn, n_groups, kHeavyRow, offs.p, counts.p, heavy.p,
                                                    (uint32_t)g->n_gates, slots.p, dptr.p);
    ctx->launches++;
    RG_CUDA(cudaGetLastError());
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    dc.slots = slots.p; slots.p = nullptr;
    dc.ptr = dptr.p; dptr.p = nullptr;
    dc.n_slots = (int64_t)total;
    dc.qx = qxn;
    dc.nyp = nyp;
    g->info.device_bytes += (int64_t)(need_b + ((size_t)n_slices + 1) * sizeof(uint32_t));
    *out = &dc;
    return RG_OK;
}

int ensure_heavy(Context* ctx, Geometry* g)
{
    std::lock_guard<std::mutex> lock(g->quad_mu);
    if (g->n_heavy >= 0) return RG_OK;
    if (g->n_rows >= 0xFFFFFFFFll) return fail(RG_ERR_UNSUPPORTED, "more than 2^32-1 rows in one slab; use more z-slabs");
    if (g->n_rows == 0 || (uint64_t)g->info.max_row_len <= kHeavyRow) { g->n_heavy = 0; return RG_OK; }
    DevBuf<unsigned int> count;
    DevBuf<uint32_t> rows;
    RG_CUDA(count.alloc(1));
    uint32_t capacity = 1u << 16;
    struct Heavy { uint32_t row, s, e; };
    std::vector<Heavy> found;
    for (;;) {                                                   // second round only if the first list was too small
        if (rows.p) { cudaFree(rows.p); rows.p = nullptr; }
        RG_CUDA(rows.alloc((size_t)3 * capacity));
        RG_CUDA(cudaMemsetAsync(count.p, 0, sizeof(unsigned int), ctx->stream));
        heavy_list_kernel<<<(unsigned)((g->n_rows + 255) / 256), 256, 0, ctx->stream>>>(g->indptr, g->n_rows, kHeavyRow, capacity, rows.p, count.p);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
        unsigned int n_found = 0;
        RG_CUDA(cudaMemcpyAsync(&n_found, count.p, sizeof(n_found), cudaMemcpyDeviceToHost, ctx->stream));
        RG_CUDA(cudaStreamSynchronize(ctx->stream));
        if (n_found <= capacity) {
            found.resize(n_found);
            if (n_found) RG_CUDA(cudaMemcpy(found.data(), rows.p, (size_t)n_found * sizeof(Heavy), cudaMemcpyDeviceToHost));
            break;
        }
        capacity = n_found;
    }
    // the atomic append order is arbitrary; nothing downstream may depend on it
    std::sort(found.begin(), found.end(), [](const Heavy& a, const Heavy& b) { return a.row < b.row; });
    const size_t n = found.size();
    std::vector<uint32_t> host_rows(n), ptr(2 * n);
    for (size_t i = 0; i < n; ++i) { host_rows[i] = found[i].row; ptr[2 * i] = found[i].s; ptr[2 * i + 1] = found[i].e; }
    std::vector<uint32_t> first(n + 1, 0);
    std::vector<uint2> chunks;
    for (size_t i = 0; i < n; ++i) {
        first[i] = (uint32_t)chunks.size();
        for (uint32_t s = ptr[2 * i]; s < ptr[2 * i + 1]; s += kHeavyChunk)
            chunks.push_back(make_uint2(s, std::min(kHeavyChunk, ptr[2 * i + 1] - s)));
    }
    first[n] = (uint32_t)chunks.size();
    if (n) {
        RG_CUDA(cudaMalloc(&g->heavy_rows, n * sizeof(uint32_t)));
        RG_CUDA(cudaMalloc(&g->heavy_first, (n + 1) * sizeof(uint32_t)));
        RG_CUDA(cudaMalloc(&g->heavy_chunks, chunks.size() * sizeof(uint2)));
        RG_CUDA(cudaMemcpy(g->heavy_rows, host_rows.data(), n * sizeof(uint32_t), cudaMemcpyHostToDevice));
        RG_CUDA(cudaMemcpy(g->heavy_first, first.data(), (n + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice));
        RG_CUDA(cudaMemcpy(g->heavy_chunks, chunks.data(), chunks.size() * sizeof(uint2), cudaMemcpyHostToDevice));
    }
    g->n_heavy_chunks = (int64_t)chunks.size();
    g->n_heavy = (int64_t)n;
    return RG_OK;
}

int finalize_geometry_stats(Context* ctx, Geometry* g)
{
    DevBuf<unsigned long long> stat;
    RG_CUDA(stat.alloc(2));
    RG_CUDA(cudaMemsetAsync(stat.p, 0, 2 * sizeof(unsigned long long), ctx->stream));
    if (g->n_rows > 0) {
        row_stats_kernel<<<(unsigned)((g->n_rows + 255) / 256), 256, 0, ctx->stream>>>(
            g->indptr, g->n_rows, stat.p, reinterpret_cast<unsigned int*>(stat.p + 1));
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
    }
    unsigned long long h[2] = {0, 0};
    RG_CUDA(cudaMemcpyAsync(h, stat.p, sizeof(h), cudaMemcpyDeviceToHost, ctx->stream));
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    g->info.n_empty_rows = (int64_t)h[0];
    g->info.max_row_len = (int64_t)(h[1] & 0xFFFFFFFFull);
    return RG_OK;
}

int build_geometry_device(Context* ctx, const float* gx, const float* gy, const float* gz, int64_t n_gates,
                          double radar_altitude, double min_radius, double beam_factor, int weighting, double toa,
                          Geometry* out, int col_stride, int64_t* level_pairs_host)
{
    const rg_grid_spec& gs = out->grid;
    const int n_levels = gs.z_end - gs.z_begin;
    const int64_t ncol = (int64_t)gs.ny * gs.nx;
    const int64_t n_rows = ncol * n_levels;

    // host copies of the axes (tiny) to size the cell grid
    std::vector<float> xa(gs.nx), ya(gs.ny), za(gs.nz);
    linspace_f32(gs.x_min, gs.x_max, gs.nx, xa.data());
    linspace_f32(gs.y_min, gs.y_max, gs.ny, ya.data());
    linspace_f32(gs.z_min, gs.z_max, gs.nz, za.data());
    auto lohi = [](const std::vector<float>& a, int b, int e, double& lo, double& hi) {
        lo = 1e300; hi = -1e300;
        for (int i = b; i < e; ++i) { lo = std::min(lo, (double)a[i]); hi = std::max(hi, (double)a[i]); }
        if (b >= e) lo = hi = 0.0;
    };
    double xlo, xhi, ylo, yhi, zlo, zhi;
    lohi(xa, 0, gs.nx, xlo, xhi);
    lohi(ya, 0, gs.ny, ylo, yhi);
    // the cell grid is laid out for the WHOLE grid even when only a z-slab is built, so that every slab scans the
    // gates in the same order and a slab's rows are bit-identical to the same rows of a full build
    lohi(za, 0, gs.nz, zlo, zhi);
    const double ax = std::max(fabs(xlo), fabs(xhi)), ay = std::max(fabs(ylo), fabs(yhi)), az = std::max(fabs(zlo), fabs(zhi));
    const double rmax = std::max(min_radius, sqrt(ax * ax + ay * ay + az * az) * beam_factor);
    const double pad = rmax * (1.0 + 1e-6) + 1.0;

    CellGrid cg{};
    cg.ox = xlo - pad; cg.oy = ylo - pad; cg.oz = zlo - pad;
    const double ex = (xhi + pad) - cg.ox, ey = (yhi + pad) - cg.oy, ez = (zhi + pad) - cg.oz;
    double cell = std::max(2.0 * min_radius, 1e-3);
    cell = std::max(cell, std::max(ex, std::max(ey, ez)) / 2048.0);
    for (;;) {
        const double cells = (floor(ex / cell) + 1) * (floor(ey / cell) + 1) * (floor(ez / cell) + 1);
        if (cells <= (double)kMaxCells) break;
        cell *= 1.2;
    }
    cg.cell = cell;
    cg.inv_cell = 1.0 / cell;
    cg.ncx = (int)floor(ex / cell) + 1;
    cg.ncy = (int)floor(ey / cell) + 1;
    cg.ncz = (int)floor(ez / cell) + 1;
    const int64_t ncell = (int64_t)cg.ncx * cg.ncy * cg.ncz;

    // the build's device time; destroyed on every exit path, the early error returns of RG_CUDA / RG_TRY included
    struct EventPair {
        cudaEvent_t a = nullptr, b = nullptr;
        ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
    } evs;
    RG_CUDA(cudaEventCreate(&evs.a));
    RG_CUDA(cudaEventCreate(&evs.b));
    cudaEvent_t ev0 = evs.a, ev1 = evs.b;
    RG_CUDA(cudaEventRecord(ev0, ctx->stream));

    // ---- K1 binning
    DevBuf<uint32_t> cell_of, cell_count, cell_start, slot_ids;
    DevBuf<float4> sorted;
    DevBuf<unsigned long long> scan_tmp, counters;
    const int64_t scan_len = std::max(ncell, n_rows);
    RG_CUDA(cell_of.alloc((size_t)n_gates));
    RG_CUDA(cell_count.alloc((size_t)ncell));
    RG_CUDA(cell_start.alloc((size_t)ncell + 1));
    RG_CUDA(scan_tmp.alloc((size_t)(scan_len / kScanTile + 4)));
    RG_CUDA(counters.alloc(2));
    RG_CUDA(cudaMemsetAsync(cell_count.p, 0, (size_t)ncell * sizeof(uint32_t), ctx->stream));
    RG_CUDA(cudaMemsetAsync(counters.p, 0, 2 * sizeof(unsigned long long), ctx->stream));

    BinParams bp{};
    bp.gx = gx; bp.gy = gy; bp.gz = gz; bp.n_gates = n_gates;
    bp.radar_alt_f32 = (float)radar_altitude;     // compute.py:182: float32 array - Python float stays float32
    bp.toa_f32 = (float)toa;                      // compute.py:193: compared in float32 (NEP 50 weak scalar)
    bp.cg = cg; bp.cell_of = cell_of.p; bp.cell_count = cell_count.p; bp.n_nonfinite = counters.p;
    if (n_gates > 0) {
        bin_count_kernel<<<(unsigned)((n_gates + 255) / 256), 256, 0, ctx->stream>>>(bp);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
    }
    uint64_t n_binned = 0;
    RG_TRY(exclusive_scan_u32(ctx, cell_count.p, cell_start.p, ncell, scan_tmp.p, &n_binned));
    unsigned long long nonfinite = 0;
    RG_CUDA(cudaMemcpy(&nonfinite, counters.p, sizeof(nonfinite), cudaMemcpyDeviceToHost));
    if (nonfinite != 0) {
        return fail(RG_ERR_INVALID, "data must be finite, check for nan or inf values (gate coordinates below toa)");
    }
    RG_CUDA(slot_ids.alloc((size_t)n_binned));
    RG_CUDA(sorted.alloc((size_t)n_binned));
    if (n_binned > 0) {
        RG_CUDA(cudaMemsetAsync(cell_count.p, 0, (size_t)ncell * sizeof(uint32_t), ctx->stream));   // reuse as fill cursor
        bin_scatter_kernel<<<(unsigned)((n_gates + 255) / 256), 256, 0, ctx->stream>>>(cell_of.p, n_gates, cell_start.p,
                                                                                       cell_count.p, slot_ids.p);
        bin_rank_kernel<<<(unsigned)((n_binned + 255) / 256), 256, 0, ctx->stream>>>(
            slot_ids.p, (int64_t)n_binned, cell_of.p, cell_start.p, gx, gy, gz, bp.radar_alt_f32, sorted.p);
        ctx->launches += 2;
        RG_CUDA(cudaGetLastError());
    }

    // ---- axes on the device (kept by the geometry: the product epilogues need x/y)
    RG_CUDA(cudaMalloc(&out->x_ax, std::max(1, gs.nx) * sizeof(float)));
    RG_CUDA(cudaMalloc(&out->y_ax, std::max(1, gs.ny) * sizeof(float)));
    RG_CUDA(cudaMalloc(&out->z_ax, std::max(1, gs.nz) * sizeof(float)));
    RG_CUDA(cudaMemcpyAsync(out->x_ax, xa.data(), gs.nx * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    RG_CUDA(cudaMemcpyAsync(out->y_ax, ya.data(), gs.ny * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    RG_CUDA(cudaMemcpyAsync(out->z_ax, za.data(), gs.nz * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    RG_CUDA(cudaStreamSynchronize(ctx->stream));   // xa/ya/za are stack-owned

    if (level_pairs_host != nullptr) {
        // level census: K2 over every col_stride-th column, pairs added up per level and scaled to the full plane
        const int st = std::max(1, col_stride);
        const int nxs = (gs.nx + st - 1) / st, nys = (gs.ny + st - 1) / st;
        DevBuf<unsigned long long> per_level;
        RG_CUDA(per_level.alloc((size_t)std::max(1, n_levels)));
        RG_CUDA(cudaMemsetAsync(per_level.p, 0, (size_t)std::max(1, n_levels) * sizeof(unsigned long long), ctx->stream));
        NeighbourParams cp{};
        cp.sorted = sorted.p; cp.cell_start = cell_start.p; cp.cg = cg;
        cp.x_ax = out->x_ax; cp.y_ax = out->y_ax; cp.z_ax = out->z_ax;
        cp.nx = gs.nx; cp.ny = gs.ny; cp.z_begin = gs.z_begin; cp.ncol = ncol;
        cp.n_rows = (int64_t)n_levels * nxs * nys;
        cp.min_radius = min_radius; cp.beam_factor = beam_factor; cp.weighting = weighting;
        cp.level_pairs = per_level.p; cp.col_stride = st; cp.nxs = nxs; cp.nys = nys;
        cp.n_candidates = counters.p + 1;
        if (cp.n_rows > 0) {
            neighbour_kernel<false><<<(unsigned)((cp.n_rows + 7) / 8), 256, 0, ctx->stream>>>(cp);
            ctx->launches++;
            RG_CUDA(cudaGetLastError());
        }
        std::vector<unsigned long long> h(std::max(1, n_levels));
        RG_CUDA(cudaMemcpyAsync(h.data(), per_level.p, (size_t)std::max(1, n_levels) * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
        RG_CUDA(cudaStreamSynchronize(ctx->stream));
        const double scale = (double)ncol / (double)std::max<int64_t>(1, (int64_t)nxs * nys);
        for (int l = 0; l < n_levels; ++l) level_pairs_host[l] = (int64_t)llround((double)h[l] * scale);
        return RG_OK;
    }

    // ---- K2 count
    DevBuf<uint32_t> counts;
    RG_CUDA(counts.alloc((size_t)n_rows));
    RG_CUDA(cudaMalloc(&out->indptr, ((size_t)n_rows + 1) * sizeof(uint32_t)));
    NeighbourParams np{};
    np.sorted = sorted.p; np.cell_start = cell_start.p; np.cg = cg;
    np.x_ax = out->x_ax; np.y_ax = out->y_ax; np.z_ax = out->z_ax;
    np.nx = gs.nx; np.ny = gs.ny; np.z_begin = gs.z_begin; np.ncol = ncol; np.n_rows = n_rows;
    np.min_radius = min_radius; np.beam_factor = beam_factor; np.weighting = weighting;
    np.counts = counts.p; np.n_candidates = counters.p + 1;
    const unsigned nb = (unsigned)((n_rows + 7) / 8);
    if (n_rows > 0) {
        neighbour_kernel<false><<<nb, 256, 0, ctx->stream>>>(np);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
    }
    uint64_t n_pairs = 0;
    RG_TRY(exclusive_scan_u32(ctx, counts.p, out->indptr, n_rows, scan_tmp.p, &n_pairs));
    if (n_pairs >= 0xFFFFFFFFull) {
        return fail(RG_ERR_UNSUPPORTED,
                    "neighbour table of this slab exceeds 2^32-1 pairs; split the grid into thinner z-slabs (z_begin/z_end)");
    }

    // ---- K3 fill
    RG_CUDA(cudaMalloc(&out->pairs, std::max<size_t>((size_t)n_pairs, 1) * sizeof(uint2)));
    np.indptr = out->indptr; np.pairs = out->pairs;
    if (n_rows > 0 && n_pairs > 0) {
        neighbour_kernel<true><<<nb, 256, 0, ctx->stream>>>(np);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
    }
    if (ctx->sort_rows && n_rows > 0 && n_pairs > 0) {
        sort_rows_kernel<<<(unsigned)((n_rows + 3) / 4), 128, 0, ctx->stream>>>(out->indptr, out->pairs, n_rows);
        ctx->launches++;
        RG_CUDA(cudaGetLastError());
    }
    out->n_rows = n_rows; out->n_pairs = (int64_t)n_pairs; out->ncol = ncol; out->n_levels = n_levels;
    RG_CUDA(cudaEventRecord(ev1, ctx->stream));
    RG_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0.f;
    RG_CUDA(cudaEventElapsedTime(&ms, ev0, ev1));
    unsigned long long cand = 0;
    RG_CUDA(cudaMemcpy(&cand, counters.p + 1, sizeof(cand), cudaMemcpyDeviceToHost));

    out->n_rows = n_rows; out->n_pairs = (int64_t)n_pairs; out->n_gates = n_gates; out->ncol = ncol; out->n_levels = n_levels;
    out->info.n_rows = n_rows; out->info.n_pairs = (int64_t)n_pairs; out->info.n_gates = n_gates;
    out->info.n_gates_binned = (int64_t)n_binned; out->info.n_candidates = (int64_t)cand;
    out->info.build_ms = ms; out->info.cell_size = cell; out->info.grid = gs;
    out->info.device_bytes = (int64_t)(((size_t)n_rows + 1) * 4 + (size_t)n_pairs * 8);
    return finalize_geometry_stats(ctx, out);
}

}  // namespace rg
