// Bare host<->device copy ceiling of one GPU of the box, with nothing of the library in the way:
// cudaMemcpyAsync between a device buffer and host staging memory of several kinds
//     pinned     cudaHostAlloc(Default)               - what rg_host_alloc hands out
//     wc         cudaHostAlloc(WriteCombined)         - H2D source only
//     thp        mmap + madvise(MADV_HUGEPAGE) + cudaHostRegister   - 2 MB pages behind the DMA mappings
//     hugetlb    mmap(MAP_HUGETLB) + cudaHostRegister - only if the box has reserved huge pages
// in three directions (h2d, d2h, both at once on two streams) and two chunkings (one copy of the whole buffer, or
// `pieces` copies as rg_apply issues them).  One process per GPU; tools/pcie_ceiling.py starts N of them at the same
// wall-clock second so that the box-wide ceiling shows up.  Output: one JSON line per (kind, direction).
//
//     nvcc -O2 -o tools/pcie_ceiling tools/pcie_ceiling.cu
//     tools/pcie_ceiling <device> <MiB> <repeats> <start_epoch_seconds or 0> <pieces>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <time.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_));                            \
            exit(2);                                                                            \
        }                                                                                       \
    } while (0)

struct HostBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int how = 0;   // 1 cudaHostAlloc, 2 mmap + register
    std::string kind;
};

static bool alloc_host(const std::string& kind, size_t bytes, HostBuf* out)
{
    out->bytes = bytes;
    out->kind = kind;
    if (kind == "pinned" || kind == "wc") {
        const unsigned flags = kind == "wc" ? cudaHostAllocWriteCombined : cudaHostAllocDefault;
        if (cudaHostAlloc(&out->p, bytes, flags) != cudaSuccess) { cudaGetLastError(); return false; }
        out->how = 1;
        memset(out->p, 1, bytes);
        return true;
    }
    const size_t two_mb = (size_t)2 << 20;
    const size_t len = (bytes + two_mb - 1) & ~(two_mb - 1);
    int flags = MAP_PRIVATE | MAP_ANONYMOUS;
    if (kind == "hugetlb") flags |= MAP_HUGETLB;
    void* p = mmap(nullptr, len + (kind == "thp" ? two_mb : 0), PROT_READ | PROT_WRITE, flags, -1, 0);
    if (p == MAP_FAILED) return false;
    if (kind == "thp") {
        p = (void*)(((uintptr_t)p + two_mb - 1) & ~(uintptr_t)(two_mb - 1));
        madvise(p, len, MADV_HUGEPAGE);
    }
    memset(p, 1, len);                                   // touch: the pages exist (as huge pages when the kernel can) before pinning
    if (cudaHostRegister(p, len, cudaHostRegisterDefault) != cudaSuccess) { cudaGetLastError(); return false; }
    out->p = p;
    out->bytes = bytes;
    out->how = 2;
    return true;
}

static long anon_huge_kb()
{
    FILE* f = fopen("/proc/self/smaps_rollup", "r");
    if (!f) return -1;
    char line[256];
    long kb = -1;
    while (fgets(line, sizeof line, f))
        if (strncmp(line, "AnonHugePages:", 14) == 0) kb = atol(line + 14);
    fclose(f);
    return kb;
}

int main(int argc, char** argv)
{
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    const size_t mib = argc > 2 ? (size_t)atol(argv[2]) : 256;
    const int reps = argc > 3 ? atoi(argv[3]) : 10;
    const long start_at = argc > 4 ? atol(argv[4]) : 0;
    const int pieces = argc > 5 ? atoi(argv[5]) : 1;
    const size_t bytes = mib << 20;
    CK(cudaSetDevice(dev));
    void *d_in, *d_out;
    CK(cudaMalloc(&d_in, bytes));
    CK(cudaMalloc(&d_out, bytes));
    cudaStream_t s_in, s_out;
    CK(cudaStreamCreateWithFlags(&s_in, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&s_out, cudaStreamNonBlocking));

    const char* kinds[] = {"pinned", "wc", "thp", "hugetlb"};
    std::vector<HostBuf> in(4), out(4);
    for (int k = 0; k < 4; ++k) {
        if (!alloc_host(kinds[k], bytes, &in[k])) in[k].p = nullptr;
        if (strcmp(kinds[k], "wc") == 0) { out[k].p = nullptr; continue; }       // reading WC memory from the CPU is the point of not using it for D2H
        if (!alloc_host(kinds[k], bytes, &out[k])) out[k].p = nullptr;
    }
    const long huge_kb = anon_huge_kb();

    auto copy = [&](void* dst, const void* src, cudaMemcpyKind kind, cudaStream_t s) {
        const size_t step = (bytes / pieces + 255) & ~(size_t)255;
        for (size_t o = 0; o < bytes; o += step)
            CK(cudaMemcpyAsync((char*)dst + o, (const char*)src + o, bytes - o < step ? bytes - o : step, kind, s));
    };
    // dir 0 h2d, 1 d2h, 2 both; GB/s summed over the directions in use.  Runs for `seconds` so that N processes
    // started in the same time slot overlap for the whole measurement.
    auto run = [&](int k, int dir, double seconds, int* n_done) -> double {
        auto once = [&]() {
            if (dir != 1) copy(d_in, in[k].p, cudaMemcpyHostToDevice, s_in);
            if (dir != 0) copy(out[k].p, d_out, cudaMemcpyDeviceToHost, s_out);
            CK(cudaStreamSynchronize(s_in));
            CK(cudaStreamSynchronize(s_out));
        };
        once();
        const auto t0 = std::chrono::steady_clock::now();
        int n = 0;
        double s = 0.0;
        do {
            for (int r = 0; r < reps; ++r) once();
            n += reps;
            s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        } while (s < seconds);
        *n_done = n;
        return (double)bytes * n * (dir == 2 ? 2 : 1) / s / 1e9;
    };
    const double slot_s = 2.5, run_s = 1.5;
    auto wait_slot = [&](int idx) {                        // N processes enter every test in the same wall-clock slot
        if (start_at <= 0) return;
        const double target = (double)start_at + idx * slot_s;
        for (;;) {
            struct timespec ts;
            clock_gettime(CLOCK_REALTIME, &ts);
            if ((double)ts.tv_sec + ts.tv_nsec * 1e-9 >= target) break;
            usleep(500);
        }
    };
    const char* dirs[] = {"h2d", "d2h", "both"};
    for (int k = 0; k < 4; ++k) {
        for (int dir = 0; dir < 3; ++dir) {
            const bool ok = (dir == 1 || in[k].p) && (dir == 0 || out[k].p);
            if (!ok) {
                wait_slot(k * 3 + dir);
                printf("{\"device\": %d, \"kind\": \"%s\", \"dir\": \"%s\", \"gbs\": null}\n", dev, kinds[k], dirs[dir]);
                continue;
            }
            wait_slot(k * 3 + dir);
            int n_done = 0;
            const double g = run(k, dir, run_s, &n_done);
            printf("{\"device\": %d, \"kind\": \"%s\", \"dir\": \"%s\", \"gbs\": %.2f, \"mib\": %zu, \"copies\": %d, \"pieces\": %d, "
                   "\"anon_huge_kb\": %ld}\n", dev, kinds[k], dirs[dir], g, mib, n_done, pieces, huge_kb);
            fflush(stdout);
        }
    }
    return 0;
}
