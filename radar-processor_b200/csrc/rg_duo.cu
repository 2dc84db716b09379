// K5, column-pair ("duo") form: the multi-field gather-weighted mean over the merged rows of two adjacent columns.
//
// Reference arithmetic being replaced (paths relative to the reference repo):
//   src/radar_grid/interpolate.py:59-104   apply_geometry (one field at a time, two P-long gathers per field)
//
// Why: the column-group kernel of rg_apply.cu gathers a gate record once per PAIR, and its time is the bytes those
// gathers pull through the L1 data pipe.  But neighbouring voxels see almost the same gates: at
// cfg3 the rows of the columns (x, y) and (x, y + 1) share 61 % of their gates.  The duo copy of the table (ensure_duo,
// rg_geometry.cu) merges the two rows by gate id into entries {gate, w0, w1}; a lane gathers the record ONCE and adds
// w0 * v / w1 * v to the sums of both columns (per field one predicate, 2 FFMA, 2 FADD; the packed fma.rn.f32x2 form is
// used in the mask-bit record layout only: ptxas does not predicate a packed instruction in place).  Per pair that is
// 0.57 x the record gathers and, with 12-byte entries over 1.6 pairs each, 0.85 x the table bytes of the warp-slice copy.
// The kernel is latency-bound (issue 58 %, L1 61 %, DRAM 54 %): what it needs is loads in flight -- batches of 7 entry
// loads, gathers in chunks of 4, and the first batch of the next level loaded before the reduce and the epilogue of the
// current one (DESIGN.md section 6 has the measured history of every choice).
//
// Mapping: a group of 4 lanes owns the columns (x, 2yp) and (x, 2yp + 1); the 8 groups of a warp are adjacent in x; the
// 4 warps of a CTA take 4 consecutive yp: an 8 x 8 patch of columns per CTA.  Lane j of a group takes entries j, j + 4,
// ... of the merged row.  After the sums a reduce-scatter over the group leaves lane gl with column gl >> 1 and the
// fields [(gl & 1) * HF, ...): that lane divides, stores the grid values and updates the column products.
//
// An absent weight is stored as -0.0: a finite value times -0.0 adds a zero to both sums, exactly as if the pair did not
// exist.  A NON-finite unmasked value would turn that zero into NaN, so the pack kernel reports such volumes
// (PackParams::nonfinite) and the kernel then takes a predicated path that skips absent weights (bit pattern test).

#include <type_traits>

#include "rg_internal.cuh"
#include "rg_device.cuh"

namespace rg {

#ifndef RG_DUO_H
#define RG_DUO_H 7             // entry loads (gate id + weight pair) issued together
#endif
#ifndef RG_DUO_U
#define RG_DUO_U 4             // record gathers in flight per lane
#endif
#ifndef RG_DUO_PIPE
#define RG_DUO_PIPE 2          // 1: the entries of the NEXT chunk of U slots (of this level, or the first chunk of the next level)
                               //    are loaded while the current chunk's records are gathered and summed: the entry loads (23 % of
                               //    all stall samples of the batch form, profiles/r02_duo_batch_h7u4_apply.md) leave the dependent chain.
                               // 2: batches of H as in 0; the first batch of the NEXT LEVEL is loaded right after the last sums of the
                               //    current one and travels during the reduce and the epilogue (0.617 -> 0.606 ms)
                               // 0: batches of H entry loads, gathers in chunks of U (the first version)
                               // (a form 3 -- the previous level's quotients, stores and product state executed under the first gathers
                               //  of the next level -- was built and measured at 0.71-0.78 ms: spills, and a predicated first chunk
                               //  that runs whether the level needs it or not; removed)
#endif
#ifndef RG_DUO_ENTRY_LD
#define RG_DUO_ENTRY_LD 2      // how the entries (read once, by one warp) are loaded: 0 = ld.global.cs (evict first), 1 = ld.global.cg
                               //    (L2 only), 2 = ld.global.L1::no_allocate -- they should not push gate records out of L1
#endif
#ifndef RG_DUO_MINBLOCKS
#define RG_DUO_MINBLOCKS 5     // CTAs of 128 threads per SM: 96 registers per thread, 20 warps
#endif

constexpr uint32_t kAbsentWeight = 0x80000000u;    // -0.0f

__device__ __forceinline__ uint32_t duo_ld_gate(const uint32_t* p)
{
#if RG_DUO_ENTRY_LD == 1 && !defined(RG_EMU)
    return __ldcg(p);
#elif RG_DUO_ENTRY_LD == 2 && !defined(RG_EMU)
    uint32_t r;
    asm volatile("ld.global.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
#else
    return __ldcs(p);
#endif
}

__device__ __forceinline__ float2 duo_ld_weights(const uint32_t* p)
{
#if RG_DUO_ENTRY_LD == 1 && !defined(RG_EMU)
    return __ldcg(reinterpret_cast<const float2*>(p));
#elif RG_DUO_ENTRY_LD == 2 && !defined(RG_EMU)
    float2 r;
    asm volatile("ld.global.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(r.x), "=f"(r.y) : "l"(p));
    return r;
#else
    return __ldcs(reinterpret_cast<const float2*>(p));
#endif
}

// (a0, a1) += (w0, w1) * v: one packed FFMA2 (fma.rn.f32x2, sm_100+) with the value broadcast; the weight pair is the
// register pair the 64-bit entry load wrote, the accumulators sit in neighbouring registers.  Only used UNPREDICATED:
// ptxas does not predicate a packed instruction in place (it writes a temporary pair and copies or selects it back, two
// to four extra instructions per packed one -- measured: 0.946 ms for the first version of this kernel).
__device__ __forceinline__ void fma2w(float& a0, float& a1, float w0, float w1, float v)
{
#ifndef RG_EMU
    unsigned long long acc, vv, ww;
    asm("mov.b64 %0, {%1,%2};" : "=l"(acc) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%2};" : "=l"(ww) : "f"(w0), "f"(w1));
    asm("mov.b64 %0, {%1,%1};" : "=l"(vv) : "f"(v));
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(ww), "l"(vv));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(a0), "=f"(a1) : "l"(acc));
#else
    a0 = fmaf(w0, v, a0);
    a1 = fmaf(w1, v, a1);
#endif
}

// One entry into the sums of both columns (a0/s0: column 0, a1/s1: column 1).
//   marker layout: per field ISETP + four predicated scalar instructions (2 FFMA, 2 FADD);
//   mask-bit layout (Layout<F>::MB, build switch RG_MASKBITS): masked values are +0.0, so sum(w*v) is one unpredicated
//   FFMA2 per field, and the predicates of the sum(w) adds come out of the mask word with one R2P;
//   CAREFUL: the volume holds an unmasked NaN / inf somewhere: a column that does not hold the gate must not see the
//   value (absent weights are skipped by their bit pattern).
template <int F, int NV, bool CAREFUL>
__device__ __forceinline__ void duo_accumulate(float w0, float w1, const float (&v)[NV], float (&a0)[F], float (&a1)[F],
                                               float (&s0)[F], float (&s1)[F])
{
    using L = Layout<F>;
    auto masked = [&](int f) -> bool {
        if constexpr (L::MB) return (__float_as_uint(v[F]) >> (f + L::SH)) & 1u;
        else return __float_as_uint(v[f]) == kMaskedBits;        // interpolate.py:78-79
    };
    if constexpr (CAREFUL) {
        const bool in0 = __float_as_uint(w0) != kAbsentWeight, in1 = __float_as_uint(w1) != kAbsentWeight;
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (!masked(f)) {
                if (in0) { a0[f] = fmaf(w0, v[f], a0[f]); s0[f] = __fadd_rn(s0[f], w0); }
                if (in1) { a1[f] = fmaf(w1, v[f], a1[f]); s1[f] = __fadd_rn(s1[f], w1); }
            }
        }
    } else if constexpr (L::MB) {
#pragma unroll
        for (int f = 0; f < F; ++f) fma2w(a0[f], a1[f], w0, w1, v[f]);
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (!masked(f)) {
                s0[f] = __fadd_rn(s0[f], w0);
                s1[f] = __fadd_rn(s1[f], w1);
            }
        }
    } else {
#pragma unroll
        for (int f = 0; f < F; ++f) {
            if (!masked(f)) {
                a0[f] = fmaf(w0, v[f], a0[f]);
                a1[f] = fmaf(w1, v[f], a1[f]);
                s0[f] = __fadd_rn(s0[f], w0);
                s1[f] = __fadd_rn(s1[f], w1);
            }
        }
    }
}

// PSIG: 0 = no products, 1 = any product list (op list over shared-memory state), 2 = "COLMAX and/or one level
// pick / blend" with three state words per (lane, field) in shared memory -- the same three forms as apply_columns_kernel.
template <int F, int PSIG>
__global__ void __launch_bounds__(kApplyThreads, RG_DUO_MINBLOCKS) apply_duo_kernel(const __grid_constant__ ApplyParams p)
{
    constexpr unsigned kFull = 0xFFFFFFFFu;
    constexpr int NV = Layout<F>::NV;
    constexpr int HF = (F + 1) / 2;                            // fields one lane finishes
    constexpr bool PROD = PSIG == 1;
    constexpr int H = RG_DUO_H, U = RG_DUO_U;
    const int lane = threadIdx.x & 31;
    const int gl = lane & 3;                                   // lane within the group
    const int qx = (int)(blockIdx.x % (unsigned)p.duo_qx);
    const int yp = (int)(blockIdx.x / (unsigned)p.duo_qx) * (kApplyThreads / 32) + (int)(threadIdx.x >> 5);
    const int cx = qx * 8 + (lane >> 2);
    const int cy = 2 * yp + (gl >> 1);                         // the column this lane finishes
    const bool col_ok = cx < p.nx && cy < p.ny;
    const int64_t col = col_ok ? (int64_t)cy * p.nx + cx : 0;
    const int f0 = (gl & 1) * HF;                              // ... and its first field
    const bool slice_ok = yp < p.duo_nyp;
    const bool careful = __ldg(p.nonfinite) == p.epoch;

    const RecSrc rec{p.records, p.records_b, p.tex_a, p.tex_b, p.null_gate};
    extern __shared__ float sm_state[];
    if constexpr (PSIG == 2) {
#pragma unroll
        for (int i = 0; i < 3 * HF; ++i) sm_state[i * kApplyThreads + threadIdx.x] = __uint_as_float(kCanonNaN);
    }
    if constexpr (PROD) {
        float x = 0.f, y = 0.f;
        if (col_ok) {
            x = __ldg(p.prod.x_ax + cx);
            y = __ldg(p.prod.y_ax + cy);
        }
        ColumnState st;
        st.init(p.prod, x, y, col);
#pragma unroll
        for (int k = 0; k < HF; ++k) st.store_words(p.prod, sm_state, k, HF);
        st.store_levels(p.prod, sm_state, HF);
    }

    // the slice's bounds words, two levels ahead of the sums (identical in all lanes of the warp)
    const uint32_t* __restrict__ slots = p.duo;
    const size_t bstride = (size_t)p.duo_nyp * (size_t)p.duo_qx;
    const uint32_t* bp = p.duo_ptr + ((size_t)p.lz_first * (size_t)p.duo_nyp + (size_t)(slice_ok ? yp : 0)) * (size_t)p.duo_qx + (size_t)qx;
    auto bounds = [&](const uint32_t* ptr, int lz, uint32_t& bs, uint32_t& be) {
        bs = be = 0;
        if (slice_ok && lz < p.lz_last) {
            bs = __ldg(ptr);
            be = __ldg(ptr + 1);
        }
    };
    uint32_t s_next, e_next, s_next2, e_next2;
    bounds(bp, p.lz_first, s_next, e_next);
    bounds(bp + bstride, p.lz_first + 1, s_next2, e_next2);
    bp += 2 * bstride;

#if RG_DUO_PIPE
    // the chunk of entries loaded ahead: gate ids and weight pairs of slots [slot0, slot0 + n), n <= U
    constexpr int CH = RG_DUO_PIPE == 2 ? H : U;
    uint32_t eg[CH];
    float2 ew[CH];
    auto load_chunk = [&](uint32_t slot0, uint32_t n) {
        const uint32_t* base = slots + (size_t)slot0 * 96 + lane;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            if ((uint32_t)j < n) {
                eg[j] = duo_ld_gate(base + j * 96);
                ew[j] = duo_ld_weights(base + j * 96 + 32 + lane);
            }
        }
    };
    load_chunk(s_next >> 1, min((e_next >> 1) - (s_next >> 1), (uint32_t)CH));
#endif

    size_t row = (size_t)p.lz_first * (size_t)p.ncol + (size_t)col - (size_t)p.ncol;
    for (int lz = p.lz_first; lz < p.lz_last; ++lz) {
        const uint32_t s = s_next, e = e_next;
        row += (size_t)p.ncol;
        s_next = s_next2;
        e_next = e_next2;
        bounds(bp, lz + 2, s_next2, e_next2);
        bp += bstride;
        {   // level z+1's slots are contiguous: pull their 128-byte lines (3 per slot) from HBM into L2 now
            const uint32_t l0 = (s_next >> 1) * 3u, l1 = (e_next >> 1) * 3u;
            for (uint32_t l = l0 + (uint32_t)lane; l < l1; l += 32u) prefetch_l2(slots + (size_t)l * 32);
        }
        const uint32_t p0 = s >> 1;
        uint32_t m = (e >> 1) - p0;                            // slots of this level: the same in all lanes
        const bool hv = s & 1u;                                // some row of the slice was summed by heavy_rows_kernel
        float a[HF], b[HF];
#pragma unroll
        for (int k = 0; k < HF; ++k) a[k] = b[k] = 0.f;
        const bool busy = m != 0 || hv;
        if (busy) {
            float a0[F], a1[F], s0[F], s1[F];                  // sum(w*v), sum(w) of column 0 / column 1
#pragma unroll
            for (int f = 0; f < F; ++f) a0[f] = a1[f] = s0[f] = s1[f] = 0.f;
            if (hv) {
                uint32_t cs = 0, ce = 0;
                if (col_ok) {
                    cs = __ldg(p.indptr + row);
                    ce = __ldg(p.indptr + row + 1);
                }
                if (ce - cs > kHeavyRow && (gl & 1) == 0) {    // the first lane of the column adds the chunks' partial sums
#if RG_PDL && !defined(RG_EMU)
                    asm volatile("griddepcontrol.wait;" ::: "memory");   // heavy_rows_kernel complete, its sums visible
#endif
                    const uint32_t r32 = (uint32_t)row;
                    int lo = 0, hi = p.n_heavy - 1;
                    while (lo < hi) {
                        const int mid = (lo + hi) >> 1;
                        if (__ldg(p.heavy_rows + mid) < r32) lo = mid + 1; else hi = mid;
                    }
                    const uint32_t c0 = __ldg(p.heavy_first + lo), c1 = __ldg(p.heavy_first + lo + 1);
                    for (uint32_t c = c0; c < c1; ++c) {
                        const float* hp = p.heavy_part + (size_t)c * (2 * F);
#pragma unroll
                        for (int f = 0; f < F; ++f) {
                            const float hwv = __ldcg(hp + f), hw = __ldcg(hp + F + f);   // L2: written by the heavy kernel
                            if (gl >> 1) { a1[f] += hwv; s1[f] += hw; } else { a0[f] += hwv; s0[f] += hw; }
                        }
                    }
                }
            }

#if RG_DUO_PIPE
            // m slots in chunks of CH: the current chunk's entries are in (eg, ew).  PIPE 1: its gathers are issued, then the
            // loads of the NEXT chunk's entries, then the sums.  PIPE 2: gathers and sums in sub-chunks of U, then the loads
            // of the next chunk -- which only get a head start at the end of a level (they travel during the reduce and the
            // epilogue).  One code path per chunk size: no slot runs that is only padding.
            const uint32_t p0n = s_next >> 1, mn = (e_next >> 1) - p0n;     // the next level's slots
            uint32_t pos = p0;
            auto chunk = [&](auto nn, auto careful_tag) {
                constexpr int N = decltype(nn)::value;
                constexpr bool CAREFUL = decltype(careful_tag)::value;
                m -= (uint32_t)N;
                pos += (uint32_t)N;
#if RG_DUO_PIPE == 1
                float v[N][NV];
                float2 cw[N];
#pragma unroll
                for (int j = 0; j < N; ++j) {
                    load_record<F>(rec, eg[j], v[j]);
                    cw[j] = ew[j];
                }
                if (m > 0) load_chunk(pos, min(m, (uint32_t)CH));
                else load_chunk(p0n, min(mn, (uint32_t)CH));
#pragma unroll
                for (int j = 0; j < N; ++j) duo_accumulate<F, NV, CAREFUL>(cw[j].x, cw[j].y, v[j], a0, a1, s0, s1);
#else
#pragma unroll
                for (int c0 = 0; c0 < N; c0 += U) {
                    float v[U][NV];
#pragma unroll
                    for (int j = 0; j < U; ++j)
                        if (c0 + j < N) load_record<F>(rec, eg[c0 + j], v[j]);
#pragma unroll
                    for (int j = 0; j < U; ++j)
                        if (c0 + j < N) duo_accumulate<F, NV, CAREFUL>(ew[c0 + j].x, ew[c0 + j].y, v[j], a0, a1, s0, s1);
                }
                if (m > 0) load_chunk(pos, min(m, (uint32_t)CH));
                else load_chunk(p0n, min(mn, (uint32_t)CH));
#endif
            };
            auto run = [&](auto careful_tag) {
                while (m >= (uint32_t)CH) chunk(std::integral_constant<int, CH>{}, careful_tag);
                static_assert(CH >= 2 && CH <= 8, "chunk sizes are written out below");
                switch (m) {
                    case 0: break;
                    case 1: chunk(std::integral_constant<int, 1>{}, careful_tag); break;
                    case 2: if constexpr (CH > 2) chunk(std::integral_constant<int, 2>{}, careful_tag); break;
                    case 3: if constexpr (CH > 3) chunk(std::integral_constant<int, 3>{}, careful_tag); break;
                    case 4: if constexpr (CH > 4) chunk(std::integral_constant<int, 4>{}, careful_tag); break;
                    case 5: if constexpr (CH > 5) chunk(std::integral_constant<int, 5>{}, careful_tag); break;
                    case 6: if constexpr (CH > 6) chunk(std::integral_constant<int, 6>{}, careful_tag); break;
                    default: if constexpr (CH > 7) chunk(std::integral_constant<int, 7>{}, careful_tag); break;
                }
            };
            if (m == 0) load_chunk(p0n, min(mn, (uint32_t)CH));   // only a heavy row at this level: nothing to consume, load ahead
            else if (careful) run(std::true_type{});               // rare: an unmasked NaN / inf somewhere in this volume
            else run(std::false_type{});
#else
            // m slots in batches: the H entry loads of a batch are issued together, its gathers in chunks of U; full
            // batches first, then ONE batch of exactly the remaining size, so no slot runs that holds only padding
            const uint32_t* gp = slots + (size_t)p0 * 96 + lane;            // gate ids of slot 0
            const uint32_t* wp = slots + (size_t)p0 * 96 + 32 + 2 * lane;   // weight pairs of slot 0
            auto batch = [&](auto mm, auto careful_tag) {
                constexpr int M = decltype(mm)::value;
                constexpr bool CAREFUL = decltype(careful_tag)::value;
                uint32_t gt[M];
                float2 ww[M];
#pragma unroll
                for (int j = 0; j < M; ++j) {
                    gt[j] = duo_ld_gate(gp + j * 96);
                    ww[j] = duo_ld_weights(wp + j * 96);
                }
#pragma unroll
                for (int c0 = 0; c0 < M; c0 += U) {
                    float v[U][NV];
#pragma unroll
                    for (int j = 0; j < U; ++j)
                        if (c0 + j < M) load_record<F>(rec, gt[c0 + j], v[j]);
#pragma unroll
                    for (int j = 0; j < U; ++j)
                        if (c0 + j < M) duo_accumulate<F, NV, CAREFUL>(ww[c0 + j].x, ww[c0 + j].y, v[j], a0, a1, s0, s1);
                }
            };
            auto run = [&](auto careful_tag) {
                while (m > (uint32_t)H) {
                    batch(std::integral_constant<int, H>{}, careful_tag);
                    gp += H * 96;
                    wp += H * 96;
                    m -= (uint32_t)H;
                }
                static_assert(H >= 1 && H <= 8, "batch sizes are written out below");
                switch (m) {
                    case 0: break;
                    case 1: batch(std::integral_constant<int, 1>{}, careful_tag); break;
                    case 2: if constexpr (H >= 2) batch(std::integral_constant<int, 2>{}, careful_tag); break;
                    case 3: if constexpr (H >= 3) batch(std::integral_constant<int, 3>{}, careful_tag); break;
                    case 4: if constexpr (H >= 4) batch(std::integral_constant<int, 4>{}, careful_tag); break;
                    case 5: if constexpr (H >= 5) batch(std::integral_constant<int, 5>{}, careful_tag); break;
                    case 6: if constexpr (H >= 6) batch(std::integral_constant<int, 6>{}, careful_tag); break;
                    case 7: if constexpr (H >= 7) batch(std::integral_constant<int, 7>{}, careful_tag); break;
                    default: if constexpr (H >= 8) batch(std::integral_constant<int, 8>{}, careful_tag); break;
                }
            };
            if (careful) run(std::true_type{});                // rare: an unmasked NaN / inf somewhere in this volume
            else run(std::false_type{});

#endif

            // reduce-scatter over the group: distance 2 separates the two columns, distance 1 the two halves of the fields
            float t_wv[F], t_w[F];
            {
                const bool up = gl & 2;
#pragma unroll
                for (int f = 0; f < F; ++f) {
                    t_wv[f] = (up ? a1[f] : a0[f]) + __shfl_xor_sync(kFull, up ? a0[f] : a1[f], 2);
                    t_w[f] = (up ? s1[f] : s0[f]) + __shfl_xor_sync(kFull, up ? s0[f] : s1[f], 2);
                }
            }
            {
                const bool odd = gl & 1;
#pragma unroll
                for (int j = 0; j < HF; ++j) {
                    const bool has_hi = HF + j < F;            // resolved at compile time after unrolling
                    const float hi_wv = has_hi ? t_wv[has_hi ? HF + j : 0] : 0.f, hi_w = has_hi ? t_w[has_hi ? HF + j : 0] : 0.f;
                    a[j] = (odd ? hi_wv : t_wv[j]) + __shfl_xor_sync(kFull, odd ? t_wv[j] : hi_wv, 1);
                    b[j] = (odd ? hi_w : t_w[j]) + __shfl_xor_sync(kFull, odd ? t_w[j] : hi_w, 1);
                }
            }
        }

        // EMPTY: the slice has no entries at this level (above the highest sweep, beyond the last gate): every output is
        // the fill value; with a NaN fill the running maximum is not touched either.
        auto finish = [&](auto empty_tag) {
            constexpr bool EMPTY = decltype(empty_tag)::value;
            if constexpr (PROD) {                              // the generic product list: all states of the lane in one pass
                float vv[HF];
                bool on[HF];
#pragma unroll
                for (int k = 0; k < HF; ++k) {
                    on[k] = col_ok && f0 + k < F;
                    vv[k] = !EMPTY && b[k] > 0.f ? fast_div(a[k], b[k]) : p.fill;              // interpolate.py:99-102
                    if (on[k]) {
                        float* const dst = p.grid_out[f0 + k];
                        if (dst != nullptr) __stcs(dst + row, vv[k]);
                    }
                }
                ColumnState::update_words_n<HF>(p.prod, sm_state, HF, p.z_begin + lz, vv, on);
                return;
            }
#pragma unroll
            for (int k = 0; k < HF; ++k) {
                const int f = f0 + k;
                if (col_ok && f < F) {
                    const float v = !EMPTY && b[k] > 0.f ? fast_div(a[k], b[k]) : p.fill;      // interpolate.py:99-102
                    float* const dst = p.grid_out[f];          // from the parameter bank: no register held across levels
                    if (dst != nullptr) __stcs(dst + row, v);
                    if constexpr (PSIG == 2) {
                        const int z = p.z_begin + lz;
                        float* const qs = sm_state + k * kApplyThreads + threadIdx.x;
                        if (!EMPTY || !isnan(p.fill)) {
                            if ((unsigned)(z - p.prod.cmax_z0) < p.prod.cmax_w && !isnan(v)) {
                                const float c = qs[0];
                                qs[0] = isnan(c) ? v : fmaxf(c, v);
                            }
                        }
                        if (z == p.prod.slices[0].z_lo || z == p.prod.slices[0].z_hi) {     // uniform: two levels of the column
                            if (z == p.prod.slices[0].z_lo) qs[HF * kApplyThreads] = v;
                            if (z == p.prod.slices[0].z_hi) qs[2 * HF * kApplyThreads] = v;
                        }
                    }
                }
            }
        };
#if RG_DUO_PIPE
        if (!busy) load_chunk(s_next >> 1, min((e_next >> 1) - (s_next >> 1), (uint32_t)CH));   // empty level: nothing was consumed
#endif
        if (!busy) finish(std::true_type{});
        else finish(std::false_type{});
    }

    if constexpr (PSIG == 2) {
#pragma unroll
        for (int k = 0; k < HF; ++k) {
            const int f = f0 + k;
            if (col_ok && f < F) {
                ColumnState st;
                const float* const qs = sm_state + k * kApplyThreads + threadIdx.x;
                st.cmax = qs[0];
                st.s_lo[0] = qs[HF * kApplyThreads];
                st.s_hi[0] = qs[2 * HF * kApplyThreads];
                st.zz[0] = p.prod.slices[0].z_lo | (p.prod.slices[0].z_hi << 16);   // ownership test of a partial (z-slab) blend
                st.write(p.prod, f, col, p.ncol, 0.f, 0.f);    // only cmax and the LEVEL slice are on: x, y unused
            }
        }
    }
    if constexpr (PROD) {
        if (col_ok) {
            const float x = __ldg(p.prod.x_ax + cx);
            const float y = __ldg(p.prod.y_ax + cy);
#pragma unroll
            for (int k = 0; k < HF; ++k) {
                const int f = f0 + k;
                if (f < F) {
                    ColumnState st;
                    st.load_words(p.prod, sm_state, k, HF);
                    st.load_levels(p.prod, sm_state, HF);
                    st.write(p.prod, f, col, p.ncol, x, y);
                }
            }
        }
    }
}

// ---- launch ----------------------------------------------------------------------------------------------------------
// When heavy_rows_kernel was launched just before, the two are chained by programmatic dependent launch (see
// launch_column_kernel in rg_apply.cu).
template <typename K>
static void launch_duo_kernel(K kernel, unsigned blocks, size_t smem, Context* ctx, const ApplyParams& p)
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

template <int F>
static void launch_duo_f(Context* ctx, const ApplyParams& p)
{
    constexpr int HF = (F + 1) / 2;
    constexpr int kWarps = kApplyThreads / 32;
    const unsigned blocks = (unsigned)(p.duo_qx * ((p.duo_nyp + kWarps - 1) / kWarps));
    const ProductParams& pp = p.prod;
    const bool simple = pp.any && !pp.cmin_on && !pp.cmean_on && pp.n_slices <= 1 &&
                        (pp.n_slices == 0 || pp.slices[0].kind == RG_PROD_LEVEL) && ctx->apply_variant != 3;
    if (!pp.any) {
        launch_duo_kernel(apply_duo_kernel<F, 0>, blocks, 0, ctx, p);
    } else if (simple) {
        launch_duo_kernel(apply_duo_kernel<F, 2>, blocks, (size_t)3 * HF * kApplyThreads * sizeof(float), ctx, p);
    } else {
        const size_t smem = (size_t)(pp.n_state_words * HF + RG_MAX_SLICES) * kApplyThreads * sizeof(float);
        if (smem > 48 * 1024) cudaFuncSetAttribute(apply_duo_kernel<F, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        launch_duo_kernel(apply_duo_kernel<F, 1>, blocks, smem, ctx, p);
    }
}

int launch_duo(Context* ctx, const ApplyParams& p)
{
#ifdef RG_DUO_ONLY_F                                   // register / SASS checks of one field count (never shipped)
    launch_duo_f<RG_DUO_ONLY_F>(ctx, p);
#else
    switch (p.n_fields) {
        case 1: launch_duo_f<1>(ctx, p); break;
        case 2: launch_duo_f<2>(ctx, p); break;
        case 3: launch_duo_f<3>(ctx, p); break;
        case 4: launch_duo_f<4>(ctx, p); break;
        case 5: launch_duo_f<5>(ctx, p); break;
        case 6: launch_duo_f<6>(ctx, p); break;
        case 7: launch_duo_f<7>(ctx, p); break;
        case 8: launch_duo_f<8>(ctx, p); break;
        default: return fail(RG_ERR_INVALID, "n_fields must be 1..8");
    }
#endif
    RG_CUDA(cudaGetLastError());
    return RG_OK;
}

}  // namespace rg
