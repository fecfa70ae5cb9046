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
    // grow-only bump allocator: one cudaMalloc per high-water mark, reset per call
    char* base = nullptr;
    size_t cap = 0, off = 0;
    std::vector<void*> retired;  // blocks outgrown during the current call (freed on reset)

    void reset() {
        off = 0;
        for (void* p : retired) cudaFree(p);
        retired.clear();
    }
    void* alloc(size_t bytes) {
        bytes = (bytes + 255) & ~size_t(255);
        if (off + bytes > cap) {
            // allocate a fresh, larger block; earlier pointers of this call stay valid
            size_t ncap = cap ? cap : (size_t(64) << 20);
            while (ncap < bytes) ncap <<= 1;
            ncap = ncap < bytes * 2 ? bytes * 2 : ncap;
            if (base) retired.push_back(base);
            SONIC_CUDA(cudaMalloc((void**)&base, ncap));
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
        reset();
        if (base) cudaFree(base);
        base = nullptr;
        cap = 0;
    }
};

struct Ctx {
    bool ready = false;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;
    Arena arena;
    std::mutex mu;  // calls may arrive from several OS threads (package.yaml:98-101: -threaded)
    uint64_t launches = 0;
    int opt_window_bits = 0;
    int opt_chunk = 0;
    std::map<std::string, double> timing_ms;
    const uint32_t* msm_offsets_total = nullptr;  // device address of the last batch's entry count
    cudaEvent_t ev[16] = {};  // 0-3,8,9 msm stages; 4,5 poly / microbench; 6,7 call timer; 10-15 bench marks
    char* pinned = nullptr;  // staging for small D2H results
    size_t pinned_cap = 0;
};

inline Ctx& ctx() {
    static Ctx c;
    return c;
}

#define SONIC_LAUNCH(kernel, grid, block, smem, ...)                          \
    do {                                                                      \
        kernel<<<(grid), (block), (smem), ::sonic::ctx().stream>>>(__VA_ARGS__); \
        ::sonic::ctx().launches++;                                            \
        SONIC_CUDA(cudaGetLastError());                                       \
    } while (0)

inline unsigned div_up(uint64_t a, uint64_t b) { return (unsigned)((a + b - 1) / b); }

}  // namespace sonic
