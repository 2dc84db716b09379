// DEV-ONLY lock-step CUDA emulator (tools/cpu_emu).  NOT a fallback and never loaded by the product:
// it exists so that kernel *logic* (index math, warp collectives, barriers) can be debugged in a container
// without a GPU.  build_emu.py rewrites `k<<<g, b, s, st>>>(args)` into emu_launch(k, g, b, args) and
// compiles the .cu sources with g++ against this header; the result is only ever loaded when
// RADAR_GRID_B200_LIB is pointed at it by hand.  One OS thread per CUDA thread; blocks run one at a time;
// warp collectives and __syncthreads are real barriers, so divergence bugs dead-lock or mis-compare here too.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <barrier>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __grid_constant__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { unsigned x, y; };
struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct uint3 { unsigned x, y, z; };
struct uchar4 { unsigned char x, y, z, w; };
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
inline uint2 make_uint2(unsigned x, unsigned y) { return {x, y}; }
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }

struct EmuWarp {
    std::unique_ptr<std::barrier<>> bar;
    uint64_t xchg[32];
};
struct EmuThread {
    EmuWarp* warp = nullptr;
    std::barrier<>* block_bar = nullptr;
    int lane = 0;
};
inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;
inline thread_local EmuThread emu_self;

// ---- warp collectives -------------------------------------------------------------------------------------
template <typename T>
inline T emu_exchange(T v, int src_lane)
{
    static_assert(sizeof(T) <= 8, "");
    uint64_t raw = 0;
    memcpy(&raw, &v, sizeof(T));
    emu_self.warp->xchg[emu_self.lane] = raw;
    emu_self.warp->bar->arrive_and_wait();
    raw = emu_self.warp->xchg[src_lane & 31];
    emu_self.warp->bar->arrive_and_wait();
    T out;
    memcpy(&out, &raw, sizeof(T));
    return out;
}
template <typename T> inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src); }
template <typename T> inline T __shfl_xor_sync(unsigned, T v, int off) { return emu_exchange(v, emu_self.lane ^ off); }
template <typename T> inline T __shfl_up_sync(unsigned, T v, int off)
{
    const int src = emu_self.lane - off;
    T o = emu_exchange(v, src < 0 ? emu_self.lane : src);
    return src < 0 ? v : o;
}
inline unsigned __ballot_sync(unsigned, bool pred)
{
    emu_self.warp->xchg[emu_self.lane] = pred ? 1 : 0;
    emu_self.warp->bar->arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)(emu_self.warp->xchg[i] & 1) << i;
    emu_self.warp->bar->arrive_and_wait();
    return m;
}
inline bool __any_sync(unsigned m, bool pred) { return __ballot_sync(m, pred) != 0; }
inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
    emu_self.warp->xchg[emu_self.lane] = v;
    emu_self.warp->bar->arrive_and_wait();
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m = m > (unsigned)emu_self.warp->xchg[i] ? m : (unsigned)emu_self.warp->xchg[i];
    emu_self.warp->bar->arrive_and_wait();
    return m;
}
inline void __syncwarp() { emu_self.warp->bar->arrive_and_wait(); }
inline void __syncthreads() { emu_self.block_bar->arrive_and_wait(); }

// ---- scalar intrinsics -------------------------------------------------------------------------------------
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline float __fdiv_rn(float a, float b) { return a / b; }
inline float __fsqrt_rn(float a) { return sqrtf(a); }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __dsqrt_rn(double a) { return sqrt(a); }
inline float __double2float_rn(double a) { return (float)a; }
inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
template <typename T> inline T __ldg(const T* p) { return *p; }
template <typename T> inline T __ldcs(const T* p) { return *p; }
template <typename T> inline T __ldcg(const T* p) { return *p; }
template <typename T> inline void __stcs(T* p, T v) { *p = v; }
template <typename T> inline T max(T a, T b) { return a > b ? a : b; }
template <typename T> inline T min(T a, T b) { return a < b ? a : b; }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicMax(unsigned* p, unsigned v)
{
    unsigned old = __atomic_load_n(p, __ATOMIC_RELAXED);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
    return old;
}

// ---- launch ---------------------------------------------------------------------------------------------
inline void* emu_dynamic_smem = nullptr;

template <typename K, typename... A>
inline void emu_launch(K kernel, dim3 grid, dim3 block, size_t smem_bytes, A... args)
{
    std::vector<char> dyn(smem_bytes + 16);
    emu_dynamic_smem = dyn.data();
    const unsigned nthreads = block.x * block.y * block.z;
    const unsigned nwarps = (nthreads + 31) / 32;
    const uint64_t nblocks = (uint64_t)grid.x * grid.y * grid.z;
    if (nblocks == 0 || nthreads == 0) return;
    std::vector<EmuWarp> warps(nwarps);
    std::unique_ptr<std::barrier<>> block_bar;
    std::barrier<> launch_bar(nthreads);
    auto reset = [&]() {
        for (unsigned w = 0; w < nwarps; ++w) {
            const unsigned in_warp = (w + 1) * 32 <= nthreads ? 32 : nthreads - w * 32;
            warps[w].bar = std::make_unique<std::barrier<>>(in_warp);
        }
        block_bar = std::make_unique<std::barrier<>>(nthreads);
    };
    reset();
    auto worker = [&](unsigned t) {
        threadIdx = {t % block.x, (t / block.x) % block.y, t / (block.x * block.y)};
        blockDim = block;
        gridDim = grid;
        for (uint64_t b = 0; b < nblocks; ++b) {
            blockIdx = {(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((uint64_t)grid.x * grid.y))};
            emu_self.warp = &warps[t / 32];
            emu_self.block_bar = block_bar.get();
            emu_self.lane = t % 32;
            kernel(args...);
            // a finished thread must not hold up collectives of threads still running
            emu_self.warp->bar->arrive_and_drop();
            emu_self.block_bar->arrive_and_drop();
            launch_bar.arrive_and_wait();
            if (t == 0) reset();
            launch_bar.arrive_and_wait();
        }
    };
    std::vector<std::thread> pool;
    pool.reserve(nthreads);
    for (unsigned t = 0; t < nthreads; ++t) pool.emplace_back(worker, t);
    for (auto& th : pool) th.join();
}

// ---- runtime API ------------------------------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2 };
typedef void* cudaStream_t;
typedef void* cudaEvent_t;
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
template <typename K> inline cudaError_t cudaFuncSetAttribute(K, int, int) { return cudaSuccess; }
inline cudaError_t cudaMemGetInfo(size_t* free_b, size_t* total_b) { *free_b = *total_b = (size_t)64 << 30; return cudaSuccess; }
enum { cudaStreamNonBlocking = 1, cudaHostAllocDefault = 0, cudaDevAttrMultiProcessorCount = 16 };
inline const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA error"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
template <typename T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)malloc(n ? n : 1); return *p ? cudaSuccess : cudaErrorMemoryAllocation; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
template <typename T> inline cudaError_t cudaHostAlloc(T** p, size_t n, unsigned) { *p = (T*)malloc(n ? n : 1); return cudaSuccess; }
inline cudaError_t cudaFreeHost(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { if (n) memmove(d, s, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind k, cudaStream_t) { return cudaMemcpy(d, s, n, k); }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t) { if (n) memset(d, v, n); return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)0x1; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventCreate(cudaEvent_t* e) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaEventElapsedTime(float* ms, cudaEvent_t, cudaEvent_t) { *ms = 0.f; return cudaSuccess; }
inline cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, int, int) { *v = 148; return cudaSuccess; }
