// Runtime plumbing shared by the kernels' host launchers: error capture, the device
// context, a grow-only workspace arena in HBM and launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sonic_b200.h"

namespace sonic {

// ---- per-thread error text -------------------------------------------------------------
inline std::string& last_error_text() {
    static thread_local std::string s;
    return s;
}

inline int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    last_error_text() = buf;
    return code;
}

struct CudaError {
    cudaError_t e;
    const char* what;
    int line;
};

#define SONIC_CUDA(expr)                                                        \
    do {                                                                        \
        cudaError_t _e = (expr);                                                \
        if (_e != cudaSuccess) throw ::sonic::CudaError{_e, #expr, __LINE__};   \
    } while (0)

// ---- context -----------------------------------------------------------------------------
struct Arena {
    // Grow-only bump allocator over one device block, reset at the start of every call.  When a
    // call outgrows the block, extra blocks are chained for the rest of that call and the next
    // reset replaces everything with a single block sized for the high-water mark, so that a
    // steady workload performs no cudaMalloc / cudaFree at all.
    char* base = nullptr;
    size_t cap = 0, off = 0;
    size_t used = 0;             // bytes handed out during the current call (all blocks)
    std::vector<void*> retired;  // blocks outgrown during the current call (freed on reset)

    void reset() {
        for (void* p : retired) cudaFree(p);
        const bool spilled = !retired.empty();
        retired.clear();
        if (spilled || used > cap) {
            if (base) cudaFree(base);
            base = nullptr;
            cap = 0;
            size_t want = used + used / 4 + (size_t(16) << 20);
            if (cudaMalloc((void**)&base, want) == cudaSuccess) cap = want;
            else { base = nullptr; cudaGetLastError(); }
        }
        off = 0;
        used = 0;
    }
    void* alloc(size_t bytes) {
        bytes = (bytes + 255) & ~size_t(255);
        used += bytes;
        if (off + bytes > cap) {
            size_t ncap = cap * 2 > (size_t(64) << 20) ? cap * 2 : (size_t(64) << 20);
            if (ncap < bytes * 2) ncap = bytes * 2;
            if (base) retired.push_back(base);
            base = nullptr;
            cap = 0;
            cudaError_t e = cudaMalloc((void**)&base, ncap);
            if (e != cudaSuccess) {  // retry with exactly what is needed
                cudaGetLastError();
                ncap = bytes;
                SONIC_CUDA(cudaMalloc((void**)&base, ncap));
            }
            cap = ncap;
            off = 0;
        }
        void* p = base + off;
        off += bytes;
        return p;
    }
    template <class T>
    T* get(size_t count) { return (T*)alloc(count * sizeof(T)); }
    void release() {
        for (void* p : retired) cudaFree(p);
        retired.clear();
        if (base) cudaFree(base);
        base = nullptr;
        cap = off = used = 0;
    }
};

constexpr int MAX_DEV = 8;  // the GPUs of one NVSwitch box

struct Ctx {
    bool ready = false;
    int device = 0;                                       // CUDA ordinal
    int slot = 0;                                         // position in the sonic_init device list (= rank of in-library sharding)
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;   // second half of a proof's MSMs (its sort and accumulation overlap the first half's latency-bound tail)
    cudaEvent_t ovl[4] = {};          // fork / accumulate-done / join events of that overlap
    Arena arena;
    std::mutex mu;  // calls may arrive from several OS threads (package.yaml:98-101: -threaded)
    uint64_t launches = 0;
    int opt_window_bits = 0;
    int opt_chunk = 0;
    int opt_aff_fused = 0;                                // affine accumulation: 0 prefix / inverses / add kernels per round, 1 one kernel per round with the inversion inside the block (measured 2x slower)
    int opt_aff_m = 0;                                    // affine accumulation: slots per thread (0 = by round size; 8, 16, 32)
    int opt_aff_tail = 4;                                 // affine accumulation: halvings left to the serial XYZZ tail
    int opt_heavy_mode = 1;                               // 1 one 128-thread block per heavy bucket (default), 0 by quads in two steps (measured: no gain at 8 GPUs, 0.27 -> 0.47 ms on one)
    int opt_overlap = 0;                                  // 1: two MSM batches per proof on two streams (tail of the first under the accumulation of the second); measured slower, off
    int opt_chunk_max = 0;                                // longest chunk the automatic rule may pick (0 = default)
    int opt_acc_mode = 3;                                 // 0 XYZZ, straight-line mixed addition in registers; 1 XYZZ compact (operand file in shared memory); 2 affine with batched inversions; 3 automatic (2 for >= 2^25 entries, else 1)
    bool opt_g2 = false;                                  // SRS.new also generates the G2 h-vectors
    int opt_reduce_mode = 0;                              // 0 automatic (quads of lanes while latency-bound, else thread per K buckets), 1 level by level, 2 thread per K buckets, 3 quads
    int opt_sort_reserve = 1;                             // tiled sort: 1 = tiles reserve their slots of a bucket from a shared cursor (no tile-prefix pass)
    int opt_sort_mode = 1;                                // 0 thread per term + global atomics, 1 tiled counting sort (shared-memory histograms)
    bool sort_smem_set = false;
    int opt_reduce_k = 0;                                 // buckets per thread in the flat reduction (0 = automatic)
    int opt_acc_blocks = 3;                               // resident blocks/SM the accumulate kernel is compiled for (2, 3, 4)
    int opt_precompute = -1;                              // -1 auto, 0 off, else window bits
    uint64_t opt_precompute_budget = uint64_t(8) << 30;   // bytes of HBM the auto mode may spend on tables
    std::map<std::string, double> timing_ms;
    std::map<uint32_t, void*> ntt_cache;  // log2(length) -> resident twiddle tables
    const uint32_t* msm_offsets_total = nullptr;  // device address of the last batch's entry count
    const uint32_t* msm_offsets_total2 = nullptr; // the same for the second half of an overlapped pair
    bool msm_second_half = false;                 // the last MSM ran as two overlapped halves
    cudaEvent_t ev[24] = {};  // 0-3,8,9 msm stages; 4,5 poly / microbench; 6,7 call timer; 10-15 bench marks; 16-21 msm stages of the second half (overlap)
    char* pinned = nullptr;  // staging for small D2H results
    size_t pinned_cap = 0;
};

// One context per device of the sonic_init list.  Kernels are launched on "the current context's"
// stream: the calling thread's by default slot 0; every per-device worker thread of a multi-GPU
// runtime binds its own slot once (ctx_bind), so the launchers below need no device argument.
inline Ctx* ctx_slots() {
    static Ctx c[MAX_DEV];
    return c;
}
inline Ctx*& ctx_current() {
    static thread_local Ctx* p = nullptr;
    return p;
}
inline void ctx_bind(Ctx* c) { ctx_current() = c; }
inline Ctx& ctx() {
    Ctx* p = ctx_current();
    return p ? *p : ctx_slots()[0];
}

#define SONIC_LAUNCH(kernel, grid, block, smem, ...)                          \
    do {                                                                      \
        kernel<<<(grid), (block), (smem), ::sonic::ctx().stream>>>(__VA_ARGS__); \
        ::sonic::ctx().launches++;                                            \
        SONIC_CUDA(cudaGetLastError());                                       \
    } while (0)

inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace sonic
