// libsonic_b200: the C ABI declared in include/sonic_b200.h.
// One translation unit; every kernel lives in the .cuh files included here.
#include <algorithm>
#include <cinttypes>
#include <cstdlib>

#include "internal.h"

using namespace sonic;

namespace {

// r, little-endian 64-bit words, for host-side canonical checks of a handful of scalars
const uint64_t FR_MOD64[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};

bool fr_bytes_canonical(const uint8_t* b) {
    uint64_t w[4];
    memcpy(w, b, 32);
    for (int i = 3; i >= 0; --i) {
        if (w[i] < FR_MOD64[i]) return true;
        if (w[i] > FR_MOD64[i]) return false;
    }
    return false;
}

bool fr_bytes_zero(const uint8_t* b) {
    uint64_t w[4];
    memcpy(w, b, 32);
    return (w[0] | w[1] | w[2] | w[3]) == 0;
}

template <class F>
int guarded(F&& body) {
    Ctx& cx = ctx();
    if (!cx.ready) return fail(SONIC_ERR_NOT_INITIALISED, "sonic_init has not been called (no CPU fallback exists)");
    std::lock_guard<std::mutex> lock(cx.mu);
    try {
        SONIC_CUDA(cudaSetDevice(cx.device));
        cx.arena.reset();
        int rc = body(cx);
        return rc;
    } catch (const CudaError& e) {
        cudaGetLastError();
        return fail(SONIC_ERR_CUDA, "CUDA failure: %s (%s) at capi/%d", cudaGetErrorString(e.e), e.what, e.line);
    }
}

// first index in [a, b) with a non-zero scalar (0xffffffff if none)
__global__ void k_first_nonzero(const Fr* __restrict__ s, uint32_t a, uint32_t b, uint32_t* __restrict__ out) {
    const uint32_t i = a + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    if (!s[i].is_zero()) atomicMin(out, i);
}

__global__ void k_check_canonical(const Fr* __restrict__ s, uint32_t n, uint32_t* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const Fr v = s[i];
    uint32_t t = Chain::sub_cc(v.l[0], FrParams::P(0));
#pragma unroll
    for (int k = 1; k < 8; ++k) t = Chain::subc_cc(v.l[k], FrParams::P(k));
    (void)t;
    if (Chain::subc(0, 0) == 0) atomicExch(bad, 1u);
}

__global__ void k_compress_points(const G1Affine* __restrict__ pts, uint64_t n, uint8_t* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) g1_compress(pts[i], out + i * 48);
}

// affine Montgomery -> raw 96 bytes (canonical little-endian x || y; infinity = zeros)
__global__ void k_affine_to_raw(const G1Affine* __restrict__ pts, uint32_t n, Fq* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[2 * i] = fp_from_mont(pts[i].x);
    out[2 * i + 1] = fp_from_mont(pts[i].y);
}

// sum of n raw points -> affine + compressed (single thread: n is the number of ranks)
__global__ void k_sum_raw(const Fq* __restrict__ raw, uint32_t n, uint8_t* __restrict__ out48) {
    if (threadIdx.x || blockIdx.x) return;
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t i = 0; i < n; ++i) {
        G1Affine a;
        a.x = fp_to_mont(raw[2 * i]);
        a.y = fp_to_mont(raw[2 * i + 1]);
        g1_madd(acc, a);
    }
    g1_compress(g1_to_affine(acc), out48);
}

// ---- IMAD microbenchmark: register-only chains of 32x32->64 multiply-accumulates ----------
// Eight independent accumulator chains per thread; every product takes the chain's own low
// word as one factor so that nothing can be hoisted or shared between chains.
//   variant 0: mad.lo.cc / madc.hi.cc pairs (what the field multiplier issues; IMAD.WIDE.U32 + carry)
//   variant 1: mad.wide.u32 on a 64-bit accumulator (IMAD.WIDE.U32, no carry)
//   variant 2: separate mad.lo.u32 and mad.hi.u32 (two IMADs per product)
//   variant 3: four products chained through the carry flag, as one row of the multiplier is
//              (mad.lo.cc / madc.hi.cc / madc.lo.cc / ... : IMAD.WIDE.U32.X with carry in AND out)
template <int VARIANT>
__global__ void __launch_bounds__(256) k_imad_peak(uint32_t* out, uint32_t seed, int iters) {
    uint32_t lo[8], hi[8], m[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { lo[k] = seed + threadIdx.x * 8 + k; hi[k] = seed * 3u + blockIdx.x + k; m[k] = (seed >> 3) + 2 * k + 1; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                if (VARIANT == 0) {
                    uint32_t nl, nh;
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.u32 %1, %2, %3, %5;"
                                 : "=&r"(nl), "=r"(nh) : "r"(lo[k]), "r"(m[k]), "r"(lo[k]), "r"(hi[k]));
                    lo[k] = nl; hi[k] = nh;
                } else if (VARIANT == 1) {
                    uint64_t acc = ((uint64_t)hi[k] << 32) | lo[k];
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc) : "r"(lo[k]), "r"(m[k]));
                    lo[k] = (uint32_t)acc;
                    hi[k] = (uint32_t)(acc >> 32);
                } else if (VARIANT == 2) {
                    uint32_t nl, nh;
                    asm volatile("mad.lo.u32 %0, %2, %3, %4; mad.hi.u32 %1, %2, %3, %5;"
                                 : "=&r"(nl), "=r"(nh) : "r"(lo[k]), "r"(m[k]), "r"(lo[k]), "r"(hi[k]));
                    lo[k] = nl; hi[k] = nh;
                }
            }
            if (VARIANT == 3) {
                // two carry chains of four products each (8 products, like the other variants)
#pragma unroll
                for (int h = 0; h < 8; h += 4) {
                    Chain::mad_wide_cc(lo[h], hi[h], lo[h], m[h], lo[h], hi[h]);
                    Chain::madc_wide_cc(lo[h + 1], hi[h + 1], lo[h + 1], m[h + 1], lo[h + 1], hi[h + 1]);
                    Chain::madc_wide_cc(lo[h + 2], hi[h + 2], lo[h + 2], m[h + 2], lo[h + 2], hi[h + 2]);
                    Chain::madc_wide_cc(lo[h + 3], hi[h + 3], lo[h + 3], m[h + 3], lo[h + 3], hi[h + 3]);
                    hi[h + 3] = Chain::addc(hi[h + 3], 0);
                }
            }
        }
    }
    uint32_t r = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) r ^= lo[k] ^ hi[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

uint8_t* pinned(Ctx& cx, size_t bytes) {
    if (cx.pinned_cap < bytes) {
        if (cx.pinned) cudaFreeHost(cx.pinned);
        size_t cap = std::max(bytes, size_t(1) << 16);
        SONIC_CUDA(cudaMallocHost((void**)&cx.pinned, cap));
        cx.pinned_cap = cap;
    }
    return (uint8_t*)cx.pinned;
}

struct Timer {
    Ctx& cx;
    explicit Timer(Ctx& c) : cx(c) { cx.timing_ms.clear(); SONIC_CUDA(cudaEventRecord(cx.ev[6], cx.stream)); }
    void stop() {
        SONIC_CUDA(cudaEventRecord(cx.ev[7], cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        float ms = 0;
        SONIC_CUDA(cudaEventElapsedTime(&ms, cx.ev[6], cx.ev[7]));
        cx.timing_ms["total"] = ms;
    }
};

// error text of `index` (src/Sonic/CommitmentScheme.hs:70-73) for a shifted exponent
int srs_too_short(bool commit, int64_t e, uint64_t d) {
    if (commit) {
        if (e > 0) return fail(SONIC_ERR_SRS_TOO_SHORT, "commitPoly: gPositiveAlphaX is not long enough: %" PRId64 " >= %" PRIu64, e - 1, d);
        return fail(SONIC_ERR_SRS_TOO_SHORT, "commitPoly: gNegativeAlphaX is not long enough: %" PRId64 " >= %" PRIu64, (e < 0 ? -e : e) - 1, d);
    }
    if (e >= 0) return fail(SONIC_ERR_SRS_TOO_SHORT, "openPoly: gPositiveX is not long enough: %" PRId64 " >= %" PRIu64, e, d + 1);
    return fail(SONIC_ERR_SRS_TOO_SHORT, "openPoly: gNegativeX is not long enough: %" PRId64 " >= %" PRIu64, -e - 1, d);
}

// Checks that no non-zero scalar of a window [lo, lo+len) of exponents falls outside the
// family's range (or on the alpha hole), in the ascending order the reference folds in.
// Returns SONIC_OK or the error; on success clips the window to the valid range.
int check_window(Ctx& cx, const sonic_srs* srs, bool alpha, bool commit_text, const Fr* d_scal,
                 int64_t lo, uint64_t len, int64_t* out_lo, uint64_t* out_len, uint64_t* out_skip) {
    const int64_t d = (int64_t)srs->d;
    const int64_t hi = lo + (int64_t)len;  // exclusive
    // candidate offending index ranges, ascending
    struct Rng { int64_t a, b; } r[3];
    int nr = 0;
    if (lo < -d) r[nr++] = {lo, std::min(hi, -d)};
    if (alpha && lo <= 0 && 0 < hi) r[nr++] = {0, 1};
    if (hi > d + 1) r[nr++] = {std::max(lo, d + 1), hi};
    if (nr) {
        uint32_t* flag = cx.arena.get<uint32_t>(4);
        SONIC_CUDA(cudaMemsetAsync(flag, 0xff, 16, cx.stream));
        for (int i = 0; i < nr; ++i) {
            uint32_t a = (uint32_t)(r[i].a - lo), b = (uint32_t)(r[i].b - lo);
            if (b > a) SONIC_LAUNCH(k_first_nonzero, div_up(b - a, 256), 256, 0, d_scal, a, b, flag + i);
        }
        uint32_t h[4];
        SONIC_CUDA(cudaMemcpyAsync(h, flag, 16, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        for (int i = 0; i < nr; ++i)
            if (h[i] != 0xffffffffu) return srs_too_short(commit_text, lo + (int64_t)h[i], srs->d);
    }
    int64_t clo = std::max(lo, -d), chi = std::min(hi, d + 1);
    if (chi < clo) chi = clo;
    *out_lo = clo;
    *out_len = (uint64_t)(chi - clo);
    *out_skip = (uint64_t)(clo - lo);
    return SONIC_OK;
}

// one MSM over a window of one family; scalars canonical in device memory; result -> host
int msm_window(Ctx& cx, const sonic_srs* srs, int family, bool commit_text, const Fr* d_scal, int64_t lo,
               uint64_t len, uint8_t* out48, uint8_t* out_raw96) {
    int64_t clo;
    uint64_t clen, skip;
    int rc = check_window(cx, srs, family == SONIC_FAMILY_ALPHA, commit_text, d_scal, lo, len, &clo, &clen, &skip);
    if (rc) return rc;
    std::vector<MsmJob> jobs(1);
    jobs[0].point_base = (uint32_t)srs->index(family, clo);
    jobs[0].n = (uint32_t)clen;
    jobs[0].scalar_off = (uint32_t)skip;
    jobs[0].pad = 0;
    G1Affine* d_aff = cx.arena.get<G1Affine>(1);
    uint8_t* d_comp = cx.arena.get<uint8_t>(48);
    msm_run(cx, srs->points, srs->tables, (const uint32_t*)d_scal, jobs, d_aff, d_comp);
    uint8_t* h = pinned(cx, 256);
    if (out48) SONIC_CUDA(cudaMemcpyAsync(h, d_comp, 48, cudaMemcpyDeviceToHost, cx.stream));
    if (out_raw96) {
        Fq* d_raw = cx.arena.get<Fq>(2);
        SONIC_LAUNCH(k_affine_to_raw, 1, 32, 0, d_aff, 1u, d_raw);
        SONIC_CUDA(cudaMemcpyAsync(h + 64, d_raw, 96, cudaMemcpyDeviceToHost, cx.stream));
    }
    SONIC_CUDA(cudaStreamSynchronize(cx.stream));
    msm_collect_timing(cx);
    if (out48) memcpy(out48, h, 48);
    if (out_raw96) memcpy(out_raw96, h + 64, 96);
    return SONIC_OK;
}

int check_canonical_dev(Ctx& cx, const Fr* d_scal, uint64_t n) {
    if (!n) return SONIC_OK;
    uint32_t* bad = cx.arena.get<uint32_t>(1);
    SONIC_CUDA(cudaMemsetAsync(bad, 0, 4, cx.stream));
    SONIC_LAUNCH(k_check_canonical, div_up(n, 256), 256, 0, d_scal, (uint32_t)n, bad);
    uint32_t h = 0;
    SONIC_CUDA(cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, cx.stream));
    SONIC_CUDA(cudaStreamSynchronize(cx.stream));
    if (h) return fail(SONIC_ERR_NONCANONICAL, "an Fr encoding is not a canonical residue (>= r)");
    return SONIC_OK;
}

const Fr* upload_fr(Ctx& cx, const uint8_t* host, uint64_t n) {
    Fr* d = cx.arena.get<Fr>(n ? n : 1);
    if (n) SONIC_CUDA(cudaMemcpyAsync(d, host, n * 32, cudaMemcpyHostToDevice, cx.stream));
    return d;
}

}  // namespace

extern "C" {

int sonic_init(const int* devices, int ndev) {
    Ctx& cx = ctx();
    std::lock_guard<std::mutex> lock(cx.mu);
    if (ndev != 1 && !(ndev == 0 && devices == nullptr))
        return fail(SONIC_ERR_INVALID_ARG, "one process drives one GPU: ndev must be 1");
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(SONIC_ERR_NO_DEVICE, "no CUDA device is usable and there is no CPU fallback");
    }
    int dev = devices ? devices[0] : 0;
    if (dev < 0 || dev >= count) return fail(SONIC_ERR_INVALID_ARG, "device ordinal %d out of range (0..%d)", dev, count - 1);
    if (cx.ready) {
        if (cx.device == dev) return SONIC_OK;
        return fail(SONIC_ERR_INVALID_ARG, "already bound to device %d", cx.device);
    }
    try {
        SONIC_CUDA(cudaSetDevice(dev));
        cudaDeviceProp prop;
        SONIC_CUDA(cudaGetDeviceProperties(&prop, dev));
        if (prop.major < 10)
            return fail(SONIC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", dev, prop.major, prop.minor);
        cx.device = dev;
        cx.sm_count = prop.multiProcessorCount;
        SONIC_CUDA(cudaStreamCreateWithFlags(&cx.stream, cudaStreamNonBlocking));
        for (auto& e : cx.ev) SONIC_CUDA(cudaEventCreate(&e));
        if (const char* e = getenv("SONIC_ACC_BLOCKS")) { int v = atoi(e); if (v >= 2 && v <= 3) cx.opt_acc_blocks = v; }
        if (const char* e = getenv("SONIC_ACC_MODE")) { int v = atoi(e); if (v >= 0 && v <= 1) cx.opt_acc_mode = v; }
        cx.ready = true;
    } catch (const CudaError& e) {
        cudaGetLastError();
        return fail(SONIC_ERR_CUDA, "CUDA failure: %s (%s)", cudaGetErrorString(e.e), e.what);
    }
    return SONIC_OK;
}

void sonic_shutdown(void) {
    Ctx& cx = ctx();
    std::lock_guard<std::mutex> lock(cx.mu);
    if (!cx.ready) return;
    cudaSetDevice(cx.device);
    cudaStreamSynchronize(cx.stream);
    cx.arena.release();
    for (auto& kv : cx.ntt_cache) cudaFree(kv.second);
    cx.ntt_cache.clear();
    if (cx.pinned) cudaFreeHost(cx.pinned);
    cx.pinned = nullptr;
    cx.pinned_cap = 0;
    for (auto& e : cx.ev) { if (e) cudaEventDestroy(e); e = nullptr; }
    cudaStreamDestroy(cx.stream);
    cx.stream = nullptr;
    cx.ready = false;
}

const char* sonic_strerror(int code) {
    switch (code) {
        case SONIC_OK: return "ok";
        case SONIC_ERR_INVALID_ARG: return "invalid argument";
        case SONIC_ERR_SRS_TOO_SHORT: return "SRS is not long enough";
        case SONIC_ERR_D_TOO_SMALL: return "parameter d is not large enough";
        case SONIC_ERR_DIV_BY_ZERO: return "division by zero in Fr";
        case SONIC_ERR_NONCANONICAL: return "non-canonical field element encoding";
        case SONIC_ERR_CUDA: return "CUDA failure";
        case SONIC_ERR_BUFFER_TOO_SMALL: return "output buffer too small";
        case SONIC_ERR_NO_DEVICE: return "no usable CUDA device (no CPU fallback)";
        case SONIC_ERR_NOT_INITIALISED: return "library not initialised";
        default: return "unknown error";
    }
}

size_t sonic_last_error(char* buf, size_t cap) {
    const std::string& s = last_error_text();
    if (buf && cap) {
        size_t n = std::min(cap - 1, s.size());
        memcpy(buf, s.data(), n);
        buf[n] = 0;
    }
    return s.size();
}

int sonic_srs_new(uint64_t d, const uint8_t x[32], const uint8_t alpha[32], sonic_srs** out) {
    if (!out || !x || !alpha) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (d == 0 || d >= (1ull << 28)) return fail(SONIC_ERR_INVALID_ARG, "d out of range");
    if (!fr_bytes_canonical(x) || !fr_bytes_canonical(alpha)) return fail(SONIC_ERR_NONCANONICAL, "x or alpha is not a canonical residue");
    if (fr_bytes_zero(x)) return fail(SONIC_ERR_DIV_BY_ZERO, "SRS.new: recip 0 (x = 0)");
    return guarded([&](Ctx& cx) {
        sonic_srs* s = new sonic_srs;
        s->d = d;
        const uint64_t npts = 2 * s->stride();
        // precomputed window multiples: automatic while they fit the memory budget
        int pre_c = cx.opt_precompute;
        if (pre_c < 0) {
            int lg = 0;
            while ((2ull << lg) <= d) ++lg;          // floor(log2 d)
            // measured on prove() with d = 7n: n = 2^12 c = 13 (5.86 ms; 12: 5.94, 14: 6.11), n = 2^14 c = 15
            // (15.1 ms; 14: 16.1, 16: 15.7), n = 2^16 c = 16 (49.5 ms; 15: 51.2, 17: 51.2)
            pre_c = lg - 1;
            if (pre_c < 4) pre_c = 4;
            if (pre_c > 16) pre_c = 16;
            const uint64_t W = (255 + pre_c - 1) / pre_c;
            if (npts * W * sizeof(G1Affine) > cx.opt_precompute_budget) pre_c = 0;
        }
        uint64_t levels = pre_c > 0 ? (255 + pre_c - 1) / pre_c : 1;
        if (npts * levels >= (1ull << 31)) { pre_c = 0; levels = 1; }
        s->tables.c = pre_c;
        s->tables.W = pre_c > 0 ? (int)levels : 0;
        s->tables.stride = (uint32_t)npts;
        cudaError_t e = cudaMalloc((void**)&s->points, npts * levels * sizeof(G1Affine));
        if (e != cudaSuccess) { delete s; throw CudaError{e, "cudaMalloc(srs)", __LINE__}; }
        if (cx.opt_g2) {
            e = cudaMalloc(&s->g2_points, npts * g2_point_bytes());
            if (e != cudaSuccess) { cudaFree(s->points); delete s; throw CudaError{e, "cudaMalloc(srs g2)", __LINE__}; }
        }
        try {
            Timer tm(cx);
            uint8_t* h = pinned(cx, 64);
            memcpy(h, x, 32);
            memcpy(h + 32, alpha, 32);
            Fr* d_canon = cx.arena.get<Fr>(2);
            SONIC_CUDA(cudaMemcpyAsync(d_canon, h, 64, cudaMemcpyHostToDevice, cx.stream));
            srs_generate(cx, d, d_canon, s->points, pre_c, s->g2_points);
            tm.stop();
        } catch (...) {
            cudaFree(s->points);
            if (s->g2_points) cudaFree(s->g2_points);
            delete s;
            throw;
        }
        *out = s;
        return (int)SONIC_OK;
    });
}

int sonic_srs_g2_range(const sonic_srs* srs, int family, int64_t exponent, uint64_t count, uint8_t* out) {
    if (!srs || !out || (family != SONIC_FAMILY_PLAIN && family != SONIC_FAMILY_ALPHA)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    if (!srs->g2_points) return fail(SONIC_ERR_INVALID_ARG, "this SRS was generated without its G2 vectors (option \"g2\")");
    const int64_t d = (int64_t)srs->d;
    if (count && (exponent < -d || exponent + (int64_t)count - 1 > d))
        return fail(SONIC_ERR_SRS_TOO_SHORT, "pcV: h vector is not long enough: %" PRId64 " >= %" PRIu64, exponent < -d ? -exponent - 1 : exponent + (int64_t)count - 1, srs->d + 1);
    if (!count) return SONIC_OK;
    return guarded([&](Ctx& cx) {
        uint8_t* d_out = cx.arena.get<uint8_t>(count * 96);
        g2_compress_range(cx, srs->g2_points, srs->index(family, exponent), count, d_out);
        SONIC_CUDA(cudaMemcpyAsync(out, d_out, count * 96, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        return (int)SONIC_OK;
    });
}

void sonic_srs_free(sonic_srs* srs) {
    if (!srs) return;
    Ctx& cx = ctx();
    std::lock_guard<std::mutex> lock(cx.mu);
    if (cx.ready) { cudaSetDevice(cx.device); cudaStreamSynchronize(cx.stream); }
    if (srs->points) cudaFree(srs->points);
    if (srs->g2_points) cudaFree(srs->g2_points);
    delete srs;
}

uint64_t sonic_srs_d(const sonic_srs* srs) { return srs ? srs->d : 0; }

// ---- SRS persistence: the resident arrays as one file --------------------
namespace {
struct SrsFileHeader {
    char magic[8];       // "SONICSRS"
    uint32_t version;    // 1
    uint32_t pre_c;      // window bits of the precomputed levels (0 = level 0 only)
    uint64_t d;
    uint64_t levels;
    uint64_t points_per_level;  // 2 * (2d + 1)
};
}  // namespace

int sonic_srs_save(const sonic_srs* srs, const char* path) {
    if (!srs || !path) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    return guarded([&](Ctx& cx) {
        FILE* f = fopen(path, "wb");
        if (!f) return fail(SONIC_ERR_INVALID_ARG, "cannot open %s for writing", path);
        SrsFileHeader h;
        memcpy(h.magic, "SONICSRS", 8);
        h.version = 1;
        h.pre_c = (uint32_t)srs->tables.c;
        h.d = srs->d;
        h.levels = srs->tables.c > 0 ? (uint64_t)srs->tables.W : 1;
        h.points_per_level = 2 * srs->stride();
        bool ok = fwrite(&h, sizeof h, 1, f) == 1;
        const size_t total = (size_t)h.levels * h.points_per_level * sizeof(G1Affine);
        const size_t chunk = size_t(64) << 20;
        uint8_t* stage = pinned(cx, chunk);
        for (size_t off = 0; ok && off < total; off += chunk) {
            const size_t nbytes = std::min(chunk, total - off);
            SONIC_CUDA(cudaMemcpyAsync(stage, (const uint8_t*)srs->points + off, nbytes, cudaMemcpyDeviceToHost, cx.stream));
            SONIC_CUDA(cudaStreamSynchronize(cx.stream));
            ok = fwrite(stage, 1, nbytes, f) == nbytes;
        }
        ok = (fclose(f) == 0) && ok;
        if (!ok) return fail(SONIC_ERR_INVALID_ARG, "short write to %s", path);
        return (int)SONIC_OK;
    });
}

int sonic_srs_load(const char* path, sonic_srs** out) {
    if (!path || !out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    return guarded([&](Ctx& cx) {
        FILE* f = fopen(path, "rb");
        if (!f) return fail(SONIC_ERR_INVALID_ARG, "cannot open %s", path);
        SrsFileHeader h;
        if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, "SONICSRS", 8) != 0 || h.version != 1 || h.d == 0 ||
            h.d >= (1ull << 28) || h.points_per_level != 2 * (2 * h.d + 1) || h.levels == 0 || h.levels > 64 ||
            (h.pre_c == 0 ? h.levels != 1 : h.levels != (255 + h.pre_c - 1) / h.pre_c)) {
            fclose(f);
            return fail(SONIC_ERR_INVALID_ARG, "%s is not an SRS file of this library", path);
        }
        sonic_srs* s = new sonic_srs;
        s->d = h.d;
        s->tables.c = (int)h.pre_c;
        s->tables.W = h.pre_c ? (int)h.levels : 0;
        s->tables.stride = (uint32_t)h.points_per_level;
        const size_t total = (size_t)h.levels * h.points_per_level * sizeof(G1Affine);
        cudaError_t e = cudaMalloc((void**)&s->points, total);
        if (e != cudaSuccess) { fclose(f); delete s; throw CudaError{e, "cudaMalloc(srs)", __LINE__}; }
        const size_t chunk = size_t(64) << 20;
        uint8_t* stage = pinned(cx, chunk);
        bool ok = true;
        for (size_t off = 0; ok && off < total; off += chunk) {
            const size_t nbytes = std::min(chunk, total - off);
            ok = fread(stage, 1, nbytes, f) == nbytes;
            if (ok) {
                cudaError_t ce = cudaMemcpyAsync((uint8_t*)s->points + off, stage, nbytes, cudaMemcpyHostToDevice, cx.stream);
                if (ce == cudaSuccess) ce = cudaStreamSynchronize(cx.stream);
                if (ce != cudaSuccess) { fclose(f); cudaFree(s->points); delete s; throw CudaError{ce, "upload(srs)", __LINE__}; }
            }
        }
        fclose(f);
        if (!ok) { cudaFree(s->points); delete s; return fail(SONIC_ERR_INVALID_ARG, "%s is truncated", path); }
        *out = s;
        return (int)SONIC_OK;
    });
}

int sonic_srs_g1_range(const sonic_srs* srs, int family, int64_t exponent, uint64_t count, uint8_t* out) {
    if (!srs || !out || (family != SONIC_FAMILY_PLAIN && family != SONIC_FAMILY_ALPHA)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    const int64_t d = (int64_t)srs->d;
    const bool alpha = family == SONIC_FAMILY_ALPHA;
    const int64_t last = exponent + (int64_t)count - 1;
    // the record fields of src/Sonic/SRS.hs:13-18 do not hold anything outside [-d, d], nor g^alpha
    if (count && exponent < -d) return srs_too_short(alpha, exponent, srs->d);
    if (count && alpha && exponent <= 0 && 0 <= last) return srs_too_short(true, 0, srs->d);
    if (count && last > d) return srs_too_short(alpha, std::max(exponent, d + 1), srs->d);
    if (!count) return SONIC_OK;
    return guarded([&](Ctx& cx) {
        uint8_t* d_out = cx.arena.get<uint8_t>(count * 48);
        SONIC_LAUNCH(k_compress_points, div_up(count, 128), 128, 0, srs->points + srs->index(family, exponent), count, d_out);
        SONIC_CUDA(cudaMemcpyAsync(out, d_out, count * 48, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        return (int)SONIC_OK;
    });
}

int sonic_srs_g1(const sonic_srs* srs, int family, int64_t exponent, uint8_t out[48]) {
    return sonic_srs_g1_range(srs, family, exponent, 1, out);
}

static int msm_common(const sonic_srs* srs, int family, int64_t lo, uint64_t len, const uint8_t* host_scalars,
                      const void* dev_scalars, uint8_t* out48, uint8_t* out_raw) {
    if (!srs || (family != SONIC_FAMILY_PLAIN && family != SONIC_FAMILY_ALPHA) || (!host_scalars && !dev_scalars && len))
        return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    if (len >= (1ull << 28)) return fail(SONIC_ERR_INVALID_ARG, "len out of range");
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        const Fr* d_scal = dev_scalars ? (const Fr*)dev_scalars : upload_fr(cx, host_scalars, len);
        int rc = check_canonical_dev(cx, d_scal, len);
        if (rc) return rc;
        rc = msm_window(cx, srs, family, family == SONIC_FAMILY_ALPHA, d_scal, lo, len, out48, out_raw);
        tm.stop();
        return rc;
    });
}

int sonic_msm_g1(const sonic_srs* srs, int family, int64_t lo, uint64_t len, const uint8_t* scalars32, uint8_t out[48]) {
    if (!out) return fail(SONIC_ERR_INVALID_ARG, "null output");
    return msm_common(srs, family, lo, len, scalars32, nullptr, out, nullptr);
}

int sonic_msm_g1_partial(const sonic_srs* srs, int family, int64_t lo, uint64_t len, const uint8_t* scalars32, uint8_t out_raw[96]) {
    if (!out_raw) return fail(SONIC_ERR_INVALID_ARG, "null output");
    return msm_common(srs, family, lo, len, scalars32, nullptr, nullptr, out_raw);
}

int sonic_msm_g1_device(const sonic_srs* srs, int family, int64_t lo, uint64_t len, const void* d_scalars32, uint8_t out[48]) {
    if (!out) return fail(SONIC_ERR_INVALID_ARG, "null output");
    return msm_common(srs, family, lo, len, nullptr, d_scalars32, out, nullptr);
}

int sonic_msm_g1_device_partial(const sonic_srs* srs, int family, int64_t lo, uint64_t len, const void* d_scalars32, uint8_t out_raw[96]) {
    if (!out_raw) return fail(SONIC_ERR_INVALID_ARG, "null output");
    return msm_common(srs, family, lo, len, nullptr, d_scalars32, nullptr, out_raw);
}

int sonic_g1_sum(const uint8_t* raw96, uint64_t n, uint8_t out[48]) {
    if ((!raw96 && n) || !out || n > (1u << 20)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    return guarded([&](Ctx& cx) {
        Fq* d_raw = cx.arena.get<Fq>(2 * n + 2);
        if (n) SONIC_CUDA(cudaMemcpyAsync(d_raw, raw96, n * 96, cudaMemcpyHostToDevice, cx.stream));
        uint8_t* d_out = cx.arena.get<uint8_t>(48);
        SONIC_LAUNCH(k_sum_raw, 1, 32, 0, d_raw, (uint32_t)n, d_out);
        SONIC_CUDA(cudaMemcpyAsync(out, d_out, 48, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        return (int)SONIC_OK;
    });
}

int sonic_commit(const sonic_srs* srs, int64_t max, int64_t lo, uint64_t len, const uint8_t* coeffs32, uint8_t out_g1[48]) {
    if (!srs || !out_g1 || (!coeffs32 && len)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    if (len >= (1ull << 28)) return fail(SONIC_ERR_INVALID_ARG, "len out of range");
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        const Fr* d_scal = upload_fr(cx, coeffs32, len);
        int rc = check_canonical_dev(cx, d_scal, len);
        if (rc) return rc;
        // X^(d-max) * f(X): every exponent moves by d - max (src/Sonic/CommitmentScheme.hs:31-33)
        const int64_t shift = (int64_t)srs->d - max;
        rc = msm_window(cx, srs, SONIC_FAMILY_ALPHA, true, d_scal, lo + shift, len, out_g1, nullptr);
        tm.stop();
        return rc;
    });
}

int sonic_open(const sonic_srs* srs, const uint8_t z[32], int64_t lo, uint64_t len, const uint8_t* coeffs32,
               uint8_t out_v[32], uint8_t out_w[48]) {
    if (!srs || !z || !out_v || !out_w || (!coeffs32 && len)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    if (len >= (1ull << 27)) return fail(SONIC_ERR_INVALID_ARG, "len out of range");
    if (!fr_bytes_canonical(z)) return fail(SONIC_ERR_NONCANONICAL, "z is not a canonical residue");
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        // the dense window must contain X^0, where f(z) is subtracted (src/Sonic/CommitmentScheme.hs:44)
        int64_t wlo = std::min<int64_t>(lo, 0);
        int64_t whi = std::max<int64_t>(lo + (int64_t)len, 1);
        const uint64_t wlen = (uint64_t)(whi - wlo);
        Fr* d_canon = cx.arena.get<Fr>(wlen);
        SONIC_CUDA(cudaMemsetAsync(d_canon, 0, wlen * 32, cx.stream));
        if (len) SONIC_CUDA(cudaMemcpyAsync(d_canon + (lo - wlo), coeffs32, len * 32, cudaMemcpyHostToDevice, cx.stream));
        int rc = check_canonical_dev(cx, d_canon, wlen);
        if (rc) return rc;
        const bool z0 = fr_bytes_zero(z);
        if (z0 && wlo < 0) {
            // `eval` multiplies by (recip z)^|lo| when the polynomial has negative powers
            uint32_t* flag = cx.arena.get<uint32_t>(1);
            SONIC_CUDA(cudaMemsetAsync(flag, 0xff, 4, cx.stream));
            SONIC_LAUNCH(k_first_nonzero, div_up((uint64_t)(-wlo), 256), 256, 0, d_canon, 0u, (uint32_t)(-wlo), flag);
            uint32_t h;
            SONIC_CUDA(cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, cx.stream));
            SONIC_CUDA(cudaStreamSynchronize(cx.stream));
            if (h != 0xffffffffu) return fail(SONIC_ERR_DIV_BY_ZERO, "openPoly: recip 0 (z = 0 with negative exponents)");
            // no negative powers after all: drop them from the window
            d_canon += -wlo;
            wlo = 0;
        }
        const uint64_t flen = (uint64_t)(whi - wlo);
        Fr* f = cx.arena.get<Fr>(flen);
        fr_to_mont(cx, d_canon, f, flen, nullptr);
        // tables z^k, z^-k
        Fr* bases = cx.arena.get<Fr>(2);
        uint8_t* h = pinned(cx, 256);
        memcpy(h, z, 32);
        SONIC_CUDA(cudaMemcpyAsync(bases + 1, h, 32, cudaMemcpyHostToDevice, cx.stream));
        fr_to_mont(cx, bases + 1, bases, 1, nullptr);
        fr_inv_few(cx, bases, bases + 1, 1);
        const uint64_t tl = flen + 1;
        Fr* tabs = cx.arena.get<Fr>(2 * tl);
        pow_tables(cx, bases, 2, tabs, tl, tl);
        Fr* q = cx.arena.get<Fr>(flen);
        Fr* val = cx.arena.get<Fr>(1);
        std::vector<OpenJob> jobs(1);
        OpenJob& jb = jobs[0];
        jb.f = f; jb.pz = tabs; jb.pzi = tabs + tl; jb.q_canon = q; jb.value_canon = val;
        jb.len = (uint32_t)flen; jb.lo = (int32_t)wlo; jb.z_is_zero = z0 ? 1u : 0u; jb.pad = 0;
        open_batch(cx, jobs);
        SONIC_CUDA(cudaMemcpyAsync(h + 128, val, 32, cudaMemcpyDeviceToHost, cx.stream));
        rc = msm_window(cx, srs, SONIC_FAMILY_PLAIN, false, q, wlo, flen - 1, out_w, nullptr);
        if (rc) return rc;
        memcpy(out_v, h + 128, 32);
        tm.stop();
        return (int)SONIC_OK;
    });
}


int sonic_circuit_load(uint64_t n, uint64_t Q, const uint8_t* wL, const uint8_t* wR, const uint8_t* wO,
                       const uint8_t* cs, sonic_circuit** out) {
    if (!out || !wL || !wR || !wO || !cs) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    // `sPoly` takes n from `head wL` (src/Sonic/Constraints.hs:53): an empty weight list has no n
    if (n == 0 || Q == 0) return fail(SONIC_ERR_INVALID_ARG, "Empty weights");
    if (n >= (1ull << 24) || Q >= (1ull << 16) || n * Q >= (1ull << 28)) return fail(SONIC_ERR_INVALID_ARG, "circuit too large");
    return guarded([&](Ctx& cx) { return circuit_load(cx, n, Q, wL, wR, wO, cs, out); });
}

int sonic_circuit_load_csr(uint64_t n, uint64_t Q, const uint64_t* rowptr_L, const uint32_t* col_L, const uint8_t* val_L,
                           const uint64_t* rowptr_R, const uint32_t* col_R, const uint8_t* val_R,
                           const uint64_t* rowptr_O, const uint32_t* col_O, const uint8_t* val_O,
                           const uint8_t* cs, sonic_circuit** out) {
    if (!out || !rowptr_L || !rowptr_R || !rowptr_O || !cs) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (n == 0 || Q == 0) return fail(SONIC_ERR_INVALID_ARG, "Empty weights");
    if (n >= (1ull << 24) || Q >= (1ull << 24)) return fail(SONIC_ERR_INVALID_ARG, "circuit too large");
    const uint64_t* rp[3] = {rowptr_L, rowptr_R, rowptr_O};
    const uint32_t* cl[3] = {col_L, col_R, col_O};
    const uint8_t* vl[3] = {val_L, val_R, val_O};
    for (int m = 0; m < 3; ++m)
        if (rp[m][Q] && (!cl[m] || !vl[m])) return fail(SONIC_ERR_INVALID_ARG, "null CSR arrays");
    return guarded([&](Ctx& cx) { return circuit_load_csr(cx, n, Q, rp, cl, vl, cs, out); });
}

void sonic_circuit_free(sonic_circuit* c) {
    if (!c) return;
    Ctx& cx = ctx();
    std::lock_guard<std::mutex> lock(cx.mu);
    if (cx.ready) { cudaSetDevice(cx.device); cudaStreamSynchronize(cx.stream); }
    circuit_free(c);
}

uint64_t sonic_rnd_count(uint64_t Q) { return 2 * Q + 8; }
uint64_t sonic_proof_size(uint64_t Q) { return (4 * Q + 7) * 48 + (2 * Q + 5) * 32; }

int sonic_prove(const sonic_srs* srs, const sonic_circuit* circuit, const uint8_t* aL, const uint8_t* aR,
                const uint8_t* aO, const uint8_t* rnd, uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    if (!srs || !circuit || !aL || !aR || !aO || !rnd || !proof_out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    const uint64_t n = circuit_n(circuit), Q = circuit_Q(circuit);
    // src/Sonic/Protocol.hs:54-55
    if (srs->d < 7 * n)
        return fail(SONIC_ERR_D_TOO_SMALL, "Parameter d is not large enough: %" PRIu64 " should be greater than %" PRIu64, srs->d, 7 * n);
    // every challenge is used as an evaluation point of a polynomial with negative powers
    for (uint64_t i = 4; i < 2 * Q + 8; ++i)
        if (fr_bytes_zero(rnd + 32 * i)) return fail(SONIC_ERR_DIV_BY_ZERO, "prove: recip 0 (challenge %" PRIu64 " is zero)", i);
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        Fr* d_in = cx.arena.get<Fr>(3 * n);
        SONIC_CUDA(cudaMemcpyAsync(d_in, aL, n * 32, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(d_in + n, aR, n * 32, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(d_in + 2 * n, aO, n * 32, cudaMemcpyHostToDevice, cx.stream));
        const Fr* d_rnd = upload_fr(cx, rnd, 2 * Q + 8);
        int rc = prove_run(cx, srs, circuit, d_in, d_rnd, (uint32_t)Q, true, 0, 1, proof_out, cap, written);
        tm.stop();
        return rc;
    });
}

uint64_t sonic_shard_blob_size(uint64_t Q) { return (4 * Q + 7) * 96 + (2 * Q + 5) * 32; }

int sonic_prove_shard(const sonic_srs* srs, const sonic_circuit* circuit, const uint8_t* aL, const uint8_t* aR,
                      const uint8_t* aO, const uint8_t* rnd, uint32_t rank, uint32_t world, uint8_t* blob_out,
                      uint64_t cap, uint64_t* written) {
    if (!srs || !circuit || !aL || !aR || !aO || !rnd || !blob_out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    if (world < 2 || rank >= world || world > 64) return fail(SONIC_ERR_INVALID_ARG, "need 2 <= world <= 64 and rank < world");
    const uint64_t n = circuit_n(circuit), Q = circuit_Q(circuit);
    if (srs->d < 7 * n)
        return fail(SONIC_ERR_D_TOO_SMALL, "Parameter d is not large enough: %" PRIu64 " should be greater than %" PRIu64, srs->d, 7 * n);
    for (uint64_t i = 4; i < 2 * Q + 8; ++i)
        if (fr_bytes_zero(rnd + 32 * i)) return fail(SONIC_ERR_DIV_BY_ZERO, "prove: recip 0 (challenge %" PRIu64 " is zero)", i);
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        Fr* d_in = cx.arena.get<Fr>(3 * n);
        SONIC_CUDA(cudaMemcpyAsync(d_in, aL, n * 32, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(d_in + n, aR, n * 32, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(d_in + 2 * n, aO, n * 32, cudaMemcpyHostToDevice, cx.stream));
        const Fr* d_rnd = upload_fr(cx, rnd, 2 * Q + 8);
        int rc = prove_run(cx, srs, circuit, d_in, d_rnd, (uint32_t)Q, true, rank, world, blob_out, cap, written);
        tm.stop();
        return rc;
    });
}

int sonic_prove_shard_sink(const sonic_srs* srs, const sonic_circuit* circuit, const void* assignment, int assignment_on_device,
                           const void* d_rnd_or_null, const uint8_t* rnd_host, uint32_t rank, uint32_t world,
                           uint8_t* blob_out, uint64_t cap, uint64_t* written, void* d_partials_out) {
    if (!srs || !circuit || !assignment || !rnd_host || !blob_out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    if (world < 2 || rank >= world || world > 64) return fail(SONIC_ERR_INVALID_ARG, "need 2 <= world <= 64 and rank < world");
    const uint64_t n = circuit_n(circuit), Q = circuit_Q(circuit);
    if (srs->d < 7 * n)
        return fail(SONIC_ERR_D_TOO_SMALL, "Parameter d is not large enough: %" PRIu64 " should be greater than %" PRIu64, srs->d, 7 * n);
    for (uint64_t i = 4; i < 2 * Q + 8; ++i)
        if (fr_bytes_zero(rnd_host + 32 * i)) return fail(SONIC_ERR_DIV_BY_ZERO, "prove: recip 0 (challenge %" PRIu64 " is zero)", i);
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        const Fr* d_in = (const Fr*)assignment;
        if (!assignment_on_device) {
            Fr* up = cx.arena.get<Fr>(3 * n);
            SONIC_CUDA(cudaMemcpyAsync(up, assignment, 3 * n * 32, cudaMemcpyHostToDevice, cx.stream));
            d_in = up;
        }
        const Fr* d_rnd = d_rnd_or_null ? (const Fr*)d_rnd_or_null : upload_fr(cx, rnd_host, 2 * Q + 8);
        int rc = prove_run(cx, srs, circuit, d_in, d_rnd, (uint32_t)Q, true, rank, world, blob_out, cap, written, d_partials_out);
        tm.stop();
        return rc;
    });
}

int sonic_prove_combine_device(uint64_t Q, uint32_t world, const void* d_gathered, const uint8_t* own_blob,
                               uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    if (!d_gathered || !own_blob || !proof_out || world < 1 || world > 64 || Q == 0 || Q >= (1u << 16)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    return guarded([&](Ctx& cx) { return prove_combine(cx, (uint32_t)Q, true, world, own_blob, proof_out, cap, written, d_gathered); });
}

int sonic_prove_combine(uint64_t Q, uint32_t world, const uint8_t* blobs, uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    if (!blobs || !proof_out || world < 1 || world > 64 || Q == 0 || Q >= (1u << 16)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    return guarded([&](Ctx& cx) { return prove_combine(cx, (uint32_t)Q, true, world, blobs, proof_out, cap, written); });
}

int sonic_prove_device(const sonic_srs* srs, const sonic_circuit* circuit, const void* d_assignment,
                       const void* d_rnd, const uint8_t* rnd_host, uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    return sonic_prove_shard_device(srs, circuit, d_assignment, d_rnd, rnd_host, 0, 1, proof_out, cap, written);
}

int sonic_prove_shard_device(const sonic_srs* srs, const sonic_circuit* circuit, const void* d_assignment,
                             const void* d_rnd, const uint8_t* rnd_host, uint32_t rank, uint32_t world,
                             uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    if (!srs || !circuit || !d_assignment || !d_rnd || !rnd_host || !proof_out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    if (world < 1 || rank >= world || world > 64) return fail(SONIC_ERR_INVALID_ARG, "need 1 <= world <= 64 and rank < world");
    const uint64_t n = circuit_n(circuit), Q = circuit_Q(circuit);
    if (srs->d < 7 * n)
        return fail(SONIC_ERR_D_TOO_SMALL, "Parameter d is not large enough: %" PRIu64 " should be greater than %" PRIu64, srs->d, 7 * n);
    for (uint64_t i = 4; i < 2 * Q + 8; ++i)
        if (fr_bytes_zero(rnd_host + 32 * i)) return fail(SONIC_ERR_DIV_BY_ZERO, "prove: recip 0 (challenge %" PRIu64 " is zero)", i);
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        int rc = prove_run(cx, srs, circuit, (const Fr*)d_assignment, (const Fr*)d_rnd, (uint32_t)Q, true, rank, world, proof_out, cap, written);
        tm.stop();
        return rc;
    });
}

int sonic_hsc_prove(const sonic_srs* srs, const sonic_circuit* circuit, uint64_t m, const uint8_t* yzs,
                    const uint8_t* uv, uint8_t* out, uint64_t cap, uint64_t* written) {
    if (!srs || !circuit || (!yzs && m) || !uv || !out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    if (m >= (1u << 12)) return fail(SONIC_ERR_INVALID_ARG, "too many (y, z) pairs");
    // lay the draws out where `prove` has them: [6, 6+m) ys, [6+m, 6+2m) zs, then u, v
    std::vector<uint8_t> rnd((2 * m + 8) * 32, 0);
    for (uint64_t j = 0; j < m; ++j) {
        memcpy(&rnd[(6 + j) * 32], yzs + 64 * j, 32);
        memcpy(&rnd[(6 + m + j) * 32], yzs + 64 * j + 32, 32);
    }
    memcpy(&rnd[(6 + 2 * m) * 32], uv, 64);
    for (uint64_t i = 6; i < 2 * m + 8; ++i)
        if (fr_bytes_zero(&rnd[32 * i])) return fail(SONIC_ERR_DIV_BY_ZERO, "hscProve: recip 0 (an evaluation point is zero)");
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        const Fr* d_rnd = upload_fr(cx, rnd.data(), 2 * m + 8);
        int rc = prove_run(cx, srs, circuit, nullptr, d_rnd, (uint32_t)m, false, 0, 1, out, cap, written);
        tm.stop();
        return rc;
    });
}

int sonic_pcv_fold(uint64_t k, const uint8_t* F48, const uint8_t* W48, const uint8_t* v32, const uint8_t* z32,
                   const uint8_t* r32, const uint32_t* group, uint32_t ngroups, uint8_t* out48) {
    if (!F48 || !W48 || !v32 || !z32 || !r32 || !group || !out48 || k == 0 || k > (1u << 16) || ngroups == 0 || ngroups > 64)
        return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    for (uint64_t i = 0; i < k; ++i) {
        if (group[i] >= ngroups) return fail(SONIC_ERR_INVALID_ARG, "group index out of range");
        if (!fr_bytes_canonical(v32 + 32 * i) || !fr_bytes_canonical(z32 + 32 * i) || !fr_bytes_canonical(r32 + 32 * i))
            return fail(SONIC_ERR_NONCANONICAL, "an Fr encoding is not a canonical residue (>= r)");
    }
    return guarded([&](Ctx& cx) { return pcv_fold(cx, (uint32_t)k, F48, W48, v32, z32, r32, group, ngroups, out48); });
}

int sonic_set_option(const char* name, int64_t value) {
    if (!name) return fail(SONIC_ERR_INVALID_ARG, "null option name");
    Ctx& cx = ctx();
    std::lock_guard<std::mutex> lock(cx.mu);
    if (!strcmp(name, "window_bits")) {
        if (value != 0 && (value < 4 || value > 20)) return fail(SONIC_ERR_INVALID_ARG, "window_bits must be 0 or in [4, 20]");
        cx.opt_window_bits = (int)value;
    } else if (!strcmp(name, "g2")) {
        cx.opt_g2 = value != 0;
    } else if (!strcmp(name, "precompute")) {
        if (value != -1 && value != 0 && (value < 4 || value > 20)) return fail(SONIC_ERR_INVALID_ARG, "precompute must be -1 (auto), 0 (off) or window bits in [4, 20]");
        cx.opt_precompute = (int)value;
    } else if (!strcmp(name, "precompute_budget_mb")) {
        if (value < 0) return fail(SONIC_ERR_INVALID_ARG, "budget must be >= 0");
        cx.opt_precompute_budget = (uint64_t)value << 20;
    } else if (!strcmp(name, "reduce_mode")) {
        if (value < 0 || value > 1) return fail(SONIC_ERR_INVALID_ARG, "reduce_mode must be 0 or 1");
        cx.opt_reduce_mode = (int)value;
    } else if (!strcmp(name, "sort_mode")) {
        // 0: thread per term + global atomics; 1: tiled counting sort, automatic tile count; 2..64: tiled, that many tiles per SM
        if (value < 0 || value > 64) return fail(SONIC_ERR_INVALID_ARG, "sort_mode must be in [0, 64]");
        cx.opt_sort_mode = (int)value;
    } else if (!strcmp(name, "reduce_k")) {
        if (value < 0 || value > 256) return fail(SONIC_ERR_INVALID_ARG, "reduce_k must be in [0, 256]");
        cx.opt_reduce_k = (int)value;
    } else if (!strcmp(name, "acc_mode")) {
        if (value < 0 || value > 1) return fail(SONIC_ERR_INVALID_ARG, "acc_mode must be 0 or 1");
        cx.opt_acc_mode = (int)value;
    } else if (!strcmp(name, "acc_blocks")) {
        if (value != 3) return fail(SONIC_ERR_INVALID_ARG, "acc_blocks is fixed at 3 in this build");
        cx.opt_acc_blocks = (int)value;
    } else if (!strcmp(name, "chunk")) {
        if (value < 0 || value > 4096) return fail(SONIC_ERR_INVALID_ARG, "chunk must be in [0, 4096]");
        cx.opt_chunk = (int)value;
    } else {
        return fail(SONIC_ERR_INVALID_ARG, "unknown option %s", name);
    }
    return SONIC_OK;
}

double sonic_last_timing_ms(const char* stage) {
    Ctx& cx = ctx();
    std::lock_guard<std::mutex> lock(cx.mu);
    auto it = cx.timing_ms.find(stage ? stage : "total");
    return it == cx.timing_ms.end() ? 0.0 : it->second;
}

uint64_t sonic_launch_count(void) { return ctx().launches; }

int sonic_bench_mark(int slot) {
    if (slot < 0 || slot > 5) return fail(SONIC_ERR_INVALID_ARG, "bench mark slot must be in [0, 5]");
    Ctx& cx = ctx();
    if (!cx.ready) return fail(SONIC_ERR_NOT_INITIALISED, "sonic_init has not been called");
    std::lock_guard<std::mutex> lock(cx.mu);
    if (cudaEventRecord(cx.ev[10 + slot], cx.stream) != cudaSuccess) return fail(SONIC_ERR_CUDA, "cudaEventRecord failed");
    return SONIC_OK;
}

double sonic_bench_elapsed_ms(int from_slot, int to_slot) {
    if (from_slot < 0 || from_slot > 5 || to_slot < 0 || to_slot > 5) return -1.0;
    Ctx& cx = ctx();
    if (!cx.ready) return -1.0;
    std::lock_guard<std::mutex> lock(cx.mu);
    float ms = -1.0f;
    if (cudaEventSynchronize(cx.ev[10 + to_slot]) != cudaSuccess) return -1.0;
    if (cudaEventElapsedTime(&ms, cx.ev[10 + from_slot], cx.ev[10 + to_slot]) != cudaSuccess) return -1.0;
    return ms;
}

double sonic_imad_peak_lmacs(int variant, int iters) {
    double result = 0;
    int rc = guarded([&](Ctx& cx) {
        const int blocks = cx.sm_count * 8, threads = 256;
        uint32_t* out = cx.arena.get<uint32_t>((size_t)blocks * threads);
        if (iters <= 0) iters = 2000;
        for (int rep = 0; rep < 3; ++rep) {
            SONIC_CUDA(cudaEventRecord(cx.ev[4], cx.stream));
            if (variant == 0) SONIC_LAUNCH(k_imad_peak<0>, blocks, threads, 0, out, 12345u, iters);
            else if (variant == 1) SONIC_LAUNCH(k_imad_peak<1>, blocks, threads, 0, out, 12345u, iters);
            else if (variant == 2) SONIC_LAUNCH(k_imad_peak<2>, blocks, threads, 0, out, 12345u, iters);
            else SONIC_LAUNCH(k_imad_peak<3>, blocks, threads, 0, out, 12345u, iters);
            SONIC_CUDA(cudaEventRecord(cx.ev[5], cx.stream));
            SONIC_CUDA(cudaStreamSynchronize(cx.stream));
            float ms = 0;
            SONIC_CUDA(cudaEventElapsedTime(&ms, cx.ev[4], cx.ev[5]));
            double lmacs = (double)blocks * threads * (double)iters * 64.0 / (ms * 1e-3);
            if (lmacs > result) result = lmacs;
        }
        return (int)SONIC_OK;
    });
    return rc == SONIC_OK ? result : 0.0;
}

int sonic_selftest_field(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, uint32_t n) {
    if (!a || !b || !out || (field != 0 && field != 1)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    return guarded([&](Ctx& cx) { return selftest_field(cx, field, op, a, b, out, n); });
}

int sonic_selftest_g1(int op, const uint32_t* a_xyzz, const uint32_t* b_xyzz, uint32_t* out_affine, uint8_t* out_comp, uint32_t n) {
    if (!a_xyzz || !b_xyzz || !out_affine || !out_comp) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    return guarded([&](Ctx& cx) { return selftest_g1(cx, op, a_xyzz, b_xyzz, out_affine, out_comp, n); });
}

int sonic_dev_alloc(uint64_t bytes, void** out) {
    if (!out) return fail(SONIC_ERR_INVALID_ARG, "null output");
    return guarded([&](Ctx&) {
        SONIC_CUDA(cudaMalloc(out, bytes ? bytes : 1));
        return (int)SONIC_OK;
    });
}

int sonic_dev_free(void* p) {
    return guarded([&](Ctx& cx) {
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        SONIC_CUDA(cudaFree(p));
        return (int)SONIC_OK;
    });
}

int sonic_dev_upload(void* dst, const void* src, uint64_t bytes) {
    return guarded([&](Ctx& cx) {
        SONIC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        return (int)SONIC_OK;
    });
}

int sonic_dev_download(void* dst, const void* src, uint64_t bytes) {
    return guarded([&](Ctx& cx) {
        SONIC_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        return (int)SONIC_OK;
    });
}

}  // extern "C"
