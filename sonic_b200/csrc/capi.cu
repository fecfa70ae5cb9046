// libsonic_b200: the C ABI declared in include/sonic_b200.h.
//
// One process drives the 1, 2, 4 or 8 GPUs named in sonic_init: a context (stream, workspace arena,
// events) per device, one host worker thread per device beyond the first (the calling thread drives
// device 0), one NCCL communicator per device (ncclCommInitAll).  Handles hold one replica per device.
// A call either runs on device 0 alone (small things), or on every device at once:
//   SRS.new        every device generates 1/ndev of every level, then an all-gather over NVLink
//   prove          equal runs of the proof's MSM terms per device (prove.cu), one all-gather of the
//                  ~4 KB exchange records, fold on device 0
//   prove_batch    whole proofs dealt round-robin, no exchange
//   msm / commit   contiguous slices of the window, all-gather of the 96-byte partial sums
#include <nccl.h>

#include <algorithm>
#include <atomic>
#include <cinttypes>
#include <condition_variable>
#include <cstdlib>
#include <functional>
#include <memory>
#include <stdexcept>
#include <string>
#include <thread>

#include <nvtx3/nvToolsExt.h>

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

// ---- the runtime: devices, worker threads, communicators ---------------------------------------------
struct Worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv, cv_done;
    std::function<int()> job;
    int state = 0;  // 0 idle, 1 job posted, 2 done
    bool quit = false;
    int rc = 0;
    std::string err;
    Ctx* cx = nullptr;

    void loop() {
        ctx_bind(cx);
        cudaSetDevice(cx->device);
        std::unique_lock<std::mutex> lk(m);
        for (;;) {
            cv.wait(lk, [&] { return state == 1 || quit; });
            if (quit) return;
            lk.unlock();
            last_error_text().clear();
            int r;
            try {
                r = job();
            } catch (const CudaError& e) {
                cudaGetLastError();
                r = fail(SONIC_ERR_CUDA, "CUDA failure on device %d: %s (%s) at line %d", cx->device, cudaGetErrorString(e.e), e.what, e.line);
            } catch (const std::exception& e) {
                r = fail(SONIC_ERR_CUDA, "device %d: %s", cx->device, e.what());
            }
            lk.lock();
            rc = r;
            err = last_error_text();
            state = 2;
            cv_done.notify_all();
        }
    }
    void post(std::function<int()> f) {
        { std::lock_guard<std::mutex> lk(m); job = std::move(f); state = 1; }
        cv.notify_one();
    }
    int wait() {
        std::unique_lock<std::mutex> lk(m);
        cv_done.wait(lk, [&] { return state == 2; });
        state = 0;
        return rc;
    }
};

// all devices agree whether to enter a collective: a rank that failed before it must not leave the
// others waiting inside NCCL
struct Vote {
    std::atomic<int> arrived{0}, failed{0}, gen{0};
    std::atomic<int> verdict{1};
    int n = 1;
    bool all_ok(bool ok) {
        if (n <= 1) return ok;
        if (!ok) failed.fetch_add(1);
        const int g = gen.load(std::memory_order_acquire);
        if (arrived.fetch_add(1) + 1 == n) {
            verdict.store(failed.load() == 0 ? 1 : 0);
            failed.store(0);
            arrived.store(0);
            gen.store(g + 1, std::memory_order_release);
        } else {
            while (gen.load(std::memory_order_acquire) == g) std::this_thread::yield();
        }
        return verdict.load() != 0;
    }
};

struct Runtime {
    std::mutex mu;  // one API call at a time (calls may arrive from several OS threads: package.yaml:98-101 -threaded)
    bool ready = false;
    int ndev = 0;
    ncclComm_t comm[MAX_DEV] = {};
    bool nccl = false;
    std::unique_ptr<Worker> worker[MAX_DEV];
    Vote vote;
    int64_t opt_shard_min_terms = int64_t(1) << 17;  // per device: below this a standalone MSM stays on device 0
};

Runtime& rt() {
    static Runtime r;
    return r;
}

#define SONIC_NCCL(expr)                                                                          \
    do {                                                                                          \
        ncclResult_t _r = (expr);                                                                 \
        if (_r != ncclSuccess) throw std::runtime_error(std::string("NCCL: ") + ncclGetErrorString(_r) + " (" #expr ")"); \
    } while (0)

// runs body(cx) on device 0 under the API lock
template <class F>
int guarded(F&& body) {
    Runtime& R = rt();
    if (!R.ready) return fail(SONIC_ERR_NOT_INITIALISED, "sonic_init has not been called (no CPU fallback exists)");
    std::lock_guard<std::mutex> lock(R.mu);
    Ctx& cx = ctx_slots()[0];
    ctx_bind(&cx);
    try {
        SONIC_CUDA(cudaSetDevice(cx.device));
        cx.arena.reset();
        return body(cx);
    } catch (const CudaError& e) {
        cudaGetLastError();
        return fail(SONIC_ERR_CUDA, "CUDA failure: %s (%s) at capi/%d", cudaGetErrorString(e.e), e.what, e.line);
    } catch (const std::exception& e) {
        return fail(SONIC_ERR_CUDA, "%s", e.what());
    }
}

// runs body(r, cx_r) on every device at once (device 0 on the calling thread) under the API lock;
// returns the first failure in device order, with its text
template <class F>
int on_all_locked(F&& body) {
    Runtime& R = rt();
    for (int r = 1; r < R.ndev; ++r) {
        Ctx* cx = &ctx_slots()[r];
        R.worker[r]->post([&body, r, cx]() -> int {
            cx->arena.reset();
            return body(r, *cx);
        });
    }
    int rc0;
    std::string err0;
    {
        Ctx& cx = ctx_slots()[0];
        ctx_bind(&cx);
        last_error_text().clear();
        try {
            SONIC_CUDA(cudaSetDevice(cx.device));
            cx.arena.reset();
            rc0 = body(0, cx);
        } catch (const CudaError& e) {
            cudaGetLastError();
            rc0 = fail(SONIC_ERR_CUDA, "CUDA failure: %s (%s) at capi/%d", cudaGetErrorString(e.e), e.what, e.line);
        } catch (const std::exception& e) {
            rc0 = fail(SONIC_ERR_CUDA, "%s", e.what());
        }
        err0 = last_error_text();
    }
    int rc = rc0;
    std::string err = err0;
    for (int r = 1; r < R.ndev; ++r) {
        const int rr = R.worker[r]->wait();
        if (rc == SONIC_OK && rr != SONIC_OK) { rc = rr; err = R.worker[r]->err; }
    }
    last_error_text() = err;
    return rc;
}

template <class F>
int on_all(F&& body) {
    Runtime& R = rt();
    if (!R.ready) return fail(SONIC_ERR_NOT_INITIALISED, "sonic_init has not been called (no CPU fallback exists)");
    std::lock_guard<std::mutex> lock(R.mu);
    return on_all_locked(body);
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
// (out48: compressed; out_raw: the sum as raw 96 bytes again -- either may be null)
__global__ void k_sum_raw(const Fq* __restrict__ raw, uint32_t n, uint8_t* __restrict__ out48, Fq* __restrict__ out_raw) {
    if (threadIdx.x || blockIdx.x) return;
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t i = 0; i < n; ++i) {
        G1Affine a;
        a.x = fp_to_mont(raw[2 * i]);
        a.y = fp_to_mont(raw[2 * i + 1]);
        g1_madd(acc, a);
    }
    const G1Affine s = g1_to_affine(acc);
    if (out48) g1_compress(s, out48);
    if (out_raw) { out_raw[0] = fp_from_mont(s.x); out_raw[1] = fp_from_mont(s.y); }
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

// exceptions -> status (so that a failing device still reaches the next vote)
template <class F>
int attempt(Ctx& cx, F&& f) {
    try {
        return f();
    } catch (const CudaError& e) {
        cudaGetLastError();
        return fail(SONIC_ERR_CUDA, "CUDA failure on device %d: %s (%s) at line %d", cx.device, cudaGetErrorString(e.e), e.what, e.line);
    } catch (const std::exception& e) {
        return fail(SONIC_ERR_CUDA, "device %d: %s", cx.device, e.what());
    }
}

uint8_t* pinned(Ctx& cx, size_t bytes) {
    if (cx.pinned_cap < bytes) {
        if (cx.pinned) cudaFreeHost(cx.pinned);
        cx.pinned = nullptr;
        cx.pinned_cap = 0;
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
int check_window(Ctx& cx, uint64_t srs_d, bool alpha, bool commit_text, const Fr* d_scal,
                 int64_t lo, uint64_t len, int64_t* out_lo, uint64_t* out_len, uint64_t* out_skip) {
    const int64_t d = (int64_t)srs_d;
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
            if (h[i] != 0xffffffffu) return srs_too_short(commit_text, lo + (int64_t)h[i], srs_d);
    }
    int64_t clo = std::max(lo, -d), chi = std::min(hi, d + 1);
    if (chi < clo) chi = clo;
    *out_lo = clo;
    *out_len = (uint64_t)(chi - clo);
    *out_skip = (uint64_t)(clo - lo);
    return SONIC_OK;
}

// Bases for a batch of standalone jobs of one family: the full-range tables if the SRS has them, else
// restricted tables when they hold the whole window, else the plain points.
struct Bases {
    const G1Affine* points;
    MsmTables tables;
    uint32_t base;
};
Bases pick_bases(const SrsRep& srs, int family, int64_t clo, uint64_t clen) {
    if (srs.tables.c == 0 && srs.rt.points && clen) {
        const int64_t slot = srs.rt.find(family, clo, clen);
        if (slot >= 0) return Bases{srs.rt.points, srs.rt.tables, (uint32_t)slot};
    }
    return Bases{srs.points, srs.tables, (uint32_t)srs.index(family, clo)};
}

// one MSM over a window of one family on one device; scalars canonical in device memory.
// The sum is left in device memory: affine (Montgomery) and/or compressed.
int msm_window_dev(Ctx& cx, const SrsRep& srs, int family, bool commit_text, const Fr* d_scal, int64_t lo,
                   uint64_t len, G1Affine* d_aff, uint8_t* d_comp) {
    int64_t clo;
    uint64_t clen, skip;
    int rc = check_window(cx, srs.d, family == SONIC_FAMILY_ALPHA, commit_text, d_scal, lo, len, &clo, &clen, &skip);
    if (rc) return rc;
    const Bases b = pick_bases(srs, family, clo, clen);
    std::vector<MsmJob> jobs(1);
    jobs[0].point_base = b.base;
    jobs[0].n = (uint32_t)clen;
    jobs[0].scalar_off = (uint32_t)skip;
    jobs[0].pad = 0;
    msm_run(cx, b.points, b.tables, (const uint32_t*)d_scal, jobs, d_aff, d_comp);
    return SONIC_OK;
}

// the same, result to the host (48-byte compressed and/or 96-byte raw)
int msm_window(Ctx& cx, const SrsRep& srs, int family, bool commit_text, const Fr* d_scal, int64_t lo,
               uint64_t len, uint8_t* out48, uint8_t* out_raw96) {
    G1Affine* d_aff = cx.arena.get<G1Affine>(1);
    uint8_t* d_comp = cx.arena.get<uint8_t>(48);
    int rc = msm_window_dev(cx, srs, family, commit_text, d_scal, lo, len, d_aff, d_comp);
    if (rc) return rc;
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

// Window tables restricted to what `prove` at this circuit size reads (srs.cu: tables_build), built by
// the first proof of that size on an SRS without full-range tables and kept in the replica.  The hull of
// the MSM shapes of Protocol.hs:63,73,79-81 / Signature.hs:42-63:
//   g^{x^k}:        k in [-4n-8, max(3n, n+Q)]
//   g^{alpha x^k}:  the same range, and [d-3n-4, d] (r'(X,1) shifted by d-n)
void ensure_tables(Ctx& cx, SrsRep& srs, uint64_t n, uint64_t Q, bool has_main) {
    if (srs.tables.c > 0 || cx.opt_precompute == 0) return;
    const int64_t d = (int64_t)srs.d;
    const int64_t lo = std::max<int64_t>(has_main ? -4 * (int64_t)n - 8 : -(int64_t)n, -d);
    const int64_t hi = std::min<int64_t>(std::max<int64_t>(has_main ? 3 * (int64_t)n : 2 * (int64_t)n, (int64_t)(n + Q)), d);
    TableRange rg[3];
    int nr = 0;
    rg[nr++] = TableRange{SONIC_FAMILY_PLAIN, lo, hi, 0};
    const int64_t blo = std::max<int64_t>(d - 3 * (int64_t)n - 4, -d);
    if (has_main && blo > hi + 1) {
        rg[nr++] = TableRange{SONIC_FAMILY_ALPHA, lo, hi, 0};
        rg[nr++] = TableRange{SONIC_FAMILY_ALPHA, blo, d, 0};
    } else {
        rg[nr++] = TableRange{SONIC_FAMILY_ALPHA, lo, has_main ? d : hi, 0};
    }
    // already there?
    if (srs.rt.points && srs.rt.nranges == nr) {
        bool same = true;
        for (int i = 0; i < nr; ++i)
            same = same && srs.rt.range[i].family == rg[i].family && srs.rt.range[i].lo == rg[i].lo && srs.rt.range[i].hi == rg[i].hi;
        if (same) return;
    }
    // window: the automatic rule of SRS.new applied to the effective range 7n (measured best for prove at d = 7n)
    int c = cx.opt_precompute;
    if (c < 0) {
        const uint64_t eff = std::max<uint64_t>(7 * n, 16);
        int lg = 0;
        while ((2ull << lg) <= eff) ++lg;
        c = std::min(16, std::max(4, lg - 1));
    }
    uint64_t size = 0;
    for (int i = 0; i < nr; ++i) size += (uint64_t)(rg[i].hi - rg[i].lo + 1);
    const uint64_t W = (255 + c - 1) / c;
    if (size * W * sizeof(G1Affine) > cx.opt_precompute_budget) { tables_free(&srs.rt); return; }
    tables_build(cx, srs, rg, nr, c, &srs.rt);
}

void free_srs_rep(SrsRep* s) {
    if (!s) return;
    if (s->points) cudaFree(s->points);
    if (s->g2_points) cudaFree(s->g2_points);
    tables_free(&s->rt);
    delete s;
}

// The resident array (all levels, one flat array) is dealt in equal chunks of srs_chunk() points, device r
// generating chunk r; ONE in-place ncclAllGather replicates it (the array is allocated with the padding
// the last chunk may need).
uint64_t srs_chunk(uint64_t total, int ndev) { return (total + (uint64_t)ndev - 1) / (uint64_t)ndev; }

void allgather_points(int r, Ctx& cx, G1Affine* points, uint64_t total) {
    Runtime& R = rt();
    const uint64_t chunk = srs_chunk(total, R.ndev);
    SONIC_NCCL(ncclAllGather(points + (uint64_t)r * chunk, points, chunk * sizeof(G1Affine), ncclUint8, R.comm[r], cx.stream));
}

void shutdown_locked() {
    Runtime& R = rt();
    if (!R.ready) return;
    for (int r = 1; r < R.ndev; ++r) {
        if (!R.worker[r]) continue;
        { std::lock_guard<std::mutex> lk(R.worker[r]->m); R.worker[r]->quit = true; }
        R.worker[r]->cv.notify_one();
        if (R.worker[r]->th.joinable()) R.worker[r]->th.join();
        R.worker[r].reset();
    }
    if (R.nccl) {
        for (int r = 0; r < R.ndev; ++r) if (R.comm[r]) { ncclCommDestroy(R.comm[r]); R.comm[r] = nullptr; }
        R.nccl = false;
    }
    for (int r = 0; r < R.ndev; ++r) {
        Ctx& cx = ctx_slots()[r];
        if (!cx.ready) continue;
        cudaSetDevice(cx.device);
        cudaStreamSynchronize(cx.stream);
        cx.arena.release();
        for (auto& kv : cx.ntt_cache) cudaFree(kv.second);
        cx.ntt_cache.clear();
        if (cx.pinned) cudaFreeHost(cx.pinned);
        cx.pinned = nullptr;
        cx.pinned_cap = 0;
        for (auto& e : cx.ev) { if (e) cudaEventDestroy(e); e = nullptr; }
        for (auto& e : cx.ovl) { if (e) cudaEventDestroy(e); e = nullptr; }
        cudaStreamDestroy(cx.stream);
        cx.stream = nullptr;
        if (cx.stream2) cudaStreamDestroy(cx.stream2);
        cx.stream2 = nullptr;
        cx.ready = false;
    }
    cudaSetDevice(ctx_slots()[0].device);
    R.ndev = 0;
    R.vote.n = 1;
    R.ready = false;
}

}  // namespace

extern "C" {

int sonic_init(const int* devices, int ndev) {
    Runtime& R = rt();
    std::lock_guard<std::mutex> lock(R.mu);
    if (ndev == 0 && devices == nullptr) ndev = 1;
    if (ndev < 1 || ndev > MAX_DEV) return fail(SONIC_ERR_INVALID_ARG, "ndev must be in [1, %d]", MAX_DEV);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) {
        cudaGetLastError();
        return fail(SONIC_ERR_NO_DEVICE, "no CUDA device is usable and there is no CPU fallback");
    }
    int devs[MAX_DEV];
    for (int r = 0; r < ndev; ++r) {
        devs[r] = devices ? devices[r] : r;
        if (devs[r] < 0 || devs[r] >= count) return fail(SONIC_ERR_INVALID_ARG, "device ordinal %d out of range (0..%d)", devs[r], count - 1);
        for (int q = 0; q < r; ++q)
            if (devs[q] == devs[r]) return fail(SONIC_ERR_INVALID_ARG, "device ordinal %d listed twice", devs[r]);
    }
    if (R.ready) {
        bool same = R.ndev == ndev;
        for (int r = 0; same && r < ndev; ++r) same = ctx_slots()[r].device == devs[r];
        if (same) return SONIC_OK;
        return fail(SONIC_ERR_INVALID_ARG, "already bound to %d device(s) starting at %d: call sonic_shutdown first", R.ndev, ctx_slots()[0].device);
    }
    try {
        for (int r = 0; r < ndev; ++r) {
            Ctx& cx = ctx_slots()[r];
            SONIC_CUDA(cudaSetDevice(devs[r]));
            cudaDeviceProp prop;
            SONIC_CUDA(cudaGetDeviceProperties(&prop, devs[r]));
            if (prop.major < 10)
                return fail(SONIC_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", devs[r], prop.major, prop.minor);
            cx.device = devs[r];
            cx.slot = r;
            cx.sm_count = prop.multiProcessorCount;
            // the main stream carries the latency-bound tails and gets the higher priority; the second stream only ever
            // runs the second half of a proof's MSMs under the first half's tail
            int prio_least = 0, prio_greatest = 0;
            SONIC_CUDA(cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest));
            SONIC_CUDA(cudaStreamCreateWithPriority(&cx.stream, cudaStreamNonBlocking, prio_greatest));
            SONIC_CUDA(cudaStreamCreateWithPriority(&cx.stream2, cudaStreamNonBlocking, prio_least));
            for (auto& e : cx.ev) SONIC_CUDA(cudaEventCreate(&e));
            for (auto& e : cx.ovl) SONIC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            cx.launches = 0;
            if (const char* e = getenv("SONIC_ACC_MODE")) { int v = atoi(e); if (v >= 0 && v <= 3) cx.opt_acc_mode = v; }   // test / profiling hook
            cx.ready = true;
        }
        R.ndev = ndev;
        R.vote.n = ndev;
        R.ready = true;
        if (ndev > 1) {
            // peer access over NVLink between every pair: device-resident inputs of one device are read by the others directly
            for (int r = 0; r < ndev; ++r) {
                SONIC_CUDA(cudaSetDevice(devs[r]));
                for (int q = 0; q < ndev; ++q) {
                    if (q == r) continue;
                    int can = 0;
                    if (cudaDeviceCanAccessPeer(&can, devs[r], devs[q]) == cudaSuccess && can) {
                        cudaError_t pe = cudaDeviceEnablePeerAccess(devs[q], 0);
                        if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) throw CudaError{pe, "cudaDeviceEnablePeerAccess", __LINE__};
                        cudaGetLastError();
                    }
                }
            }
            // one communicator per device, all in this process; every collective below is issued by the
            // device's own thread on the device's own stream
            ncclResult_t nr = ncclCommInitAll(R.comm, ndev, devs);
            if (nr != ncclSuccess) {
                shutdown_locked();
                return fail(SONIC_ERR_CUDA, "ncclCommInitAll over %d devices failed: %s", ndev, ncclGetErrorString(nr));
            }
            R.nccl = true;
            for (int r = 1; r < ndev; ++r) {
                R.worker[r].reset(new Worker);
                R.worker[r]->cx = &ctx_slots()[r];
                Worker* w = R.worker[r].get();
                w->th = std::thread([w] { w->loop(); });
            }
        }
        SONIC_CUDA(cudaSetDevice(devs[0]));
        ctx_bind(&ctx_slots()[0]);
        if (ndev > 1) {
            // NCCL connects its channels lazily, at the first collective (~0.3 s on 8 GPUs): pay that here, not inside
            // the first SRS.new or prove
            int rc = on_all_locked([&](int r, Ctx& cx) -> int {
                return attempt(cx, [&]() -> int {
                    uint8_t* buf = cx.arena.get<uint8_t>(256 * (size_t)(ndev + 1));
                    SONIC_CUDA(cudaMemsetAsync(buf, 0, 256 * (size_t)(ndev + 1), cx.stream));
                    SONIC_NCCL(ncclAllGather(buf + 256 * (size_t)ndev, buf, 256, ncclUint8, R.comm[r], cx.stream));
                    SONIC_CUDA(cudaStreamSynchronize(cx.stream));
                    return (int)SONIC_OK;
                });
            });
            if (rc != SONIC_OK) {
                const std::string why = last_error_text();
                shutdown_locked();
                return fail(rc, "NCCL warm-up over %d devices failed: %s", ndev, why.c_str());
            }
        }
        // a process that exits without sonic_shutdown must not hang in the teardown of live NCCL communicators and
        // worker threads: registered after the CUDA runtime came up, so it runs before the runtime's own exit handler
        static bool at_exit_registered = false;
        if (!at_exit_registered) {
            at_exit_registered = true;
            atexit([] { sonic_shutdown(); });
        }
    } catch (const CudaError& e) {
        cudaGetLastError();
        shutdown_locked();
        return fail(SONIC_ERR_CUDA, "CUDA failure: %s (%s)", cudaGetErrorString(e.e), e.what);
    }
    return SONIC_OK;
}

void sonic_shutdown(void) {
    Runtime& R = rt();
    std::lock_guard<std::mutex> lock(R.mu);
    shutdown_locked();
}

int sonic_device_count(void) { return rt().ready ? rt().ndev : 0; }

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
    Runtime& R = rt();
    if (!R.ready) return fail(SONIC_ERR_NOT_INITIALISED, "sonic_init has not been called (no CPU fallback exists)");
    std::lock_guard<std::mutex> lock(R.mu);
    const Ctx& c0 = ctx_slots()[0];
    const uint64_t npts = 2 * (2 * d + 1);
    // precomputed window multiples of the whole range: automatic while they fit the memory budget
    // (else the first proof of a circuit size builds tables for the ranges it reads: ensure_tables)
    int pre_c = c0.opt_precompute;
    if (pre_c < 0) {
        int lg = 0;
        while ((2ull << lg) <= d) ++lg;          // floor(log2 d)
        // measured on prove() with d = 7n: n = 2^12 c = 13 (5.86 ms; 12: 5.94, 14: 6.11), n = 2^14 c = 15
        // (15.1 ms; 14: 16.1, 16: 15.7), n = 2^16 c = 16 (49.5 ms; 15: 51.2, 17: 51.2)
        pre_c = lg - 1;
        if (pre_c < 4) pre_c = 4;
        if (pre_c > 16) pre_c = 16;
        const uint64_t W = (255 + pre_c - 1) / pre_c;
        if (npts * W * sizeof(G1Affine) > c0.opt_precompute_budget) pre_c = 0;
    }
    uint64_t levels = pre_c > 0 ? (255 + pre_c - 1) / pre_c : 1;
    if (npts * levels >= (1ull << 31)) { pre_c = 0; levels = 1; }
    sonic_srs* s = new sonic_srs;
    s->d = d;
    s->ndev = R.ndev;
    const bool want_g2 = c0.opt_g2;
    int rc = on_all_locked([&](int r, Ctx& cx) -> int {
        nvtxRangePushA("sonic.srs_new");
        SrsRep* rep = new SrsRep;
        s->rep[r] = rep;
        rep->d = d;
        rep->tables.c = pre_c;
        rep->tables.W = pre_c > 0 ? (int)levels : 0;
        rep->tables.stride = (uint32_t)npts;
        int rc = attempt(cx, [&]() -> int {
            // (+ ndev points: the last chunk of the all-gather may reach past the end)
            SONIC_CUDA(cudaMalloc((void**)&rep->points, (npts * levels + (uint64_t)R.ndev) * sizeof(G1Affine)));
            if (want_g2) SONIC_CUDA(cudaMalloc(&rep->g2_points, npts * g2_point_bytes()));
            Timer tm(cx);
            uint8_t* h = pinned(cx, 64);
            memcpy(h, x, 32);
            memcpy(h + 32, alpha, 32);
            Fr* d_canon = cx.arena.get<Fr>(2);
            SONIC_CUDA(cudaMemcpyAsync(d_canon, h, 64, cudaMemcpyHostToDevice, cx.stream));
            const uint64_t chunk = srs_chunk(npts * levels, R.ndev);
            srs_generate(cx, d, d_canon, rep->points, pre_c, chunk * (uint64_t)r, chunk, rep->g2_points);
            SONIC_CUDA(cudaEventRecord(cx.ev[4], cx.stream));
            return (int)SONIC_OK;
        });
        if (R.ndev > 1) {
            if (!R.vote.all_ok(rc == SONIC_OK)) { nvtxRangePop(); return rc ? rc : fail(SONIC_ERR_CUDA, "SRS.new failed on another device"); }
            rc = attempt(cx, [&]() -> int {
                allgather_points(r, cx, rep->points, npts * levels);
                return (int)SONIC_OK;
            });
        }
        if (rc == SONIC_OK)
            rc = attempt(cx, [&]() -> int {
                SONIC_CUDA(cudaEventRecord(cx.ev[7], cx.stream));
                SONIC_CUDA(cudaStreamSynchronize(cx.stream));
                float ms = 0, gen = 0;
                SONIC_CUDA(cudaEventElapsedTime(&ms, cx.ev[6], cx.ev[7]));
                SONIC_CUDA(cudaEventElapsedTime(&gen, cx.ev[6], cx.ev[4]));
                cx.timing_ms["total"] = ms;
                cx.timing_ms["srs.generate"] = gen;
                cx.timing_ms["srs.allgather"] = ms - gen;
                return (int)SONIC_OK;
            });
        nvtxRangePop();
        return rc;
    });
    if (rc != SONIC_OK) {
        for (int r = 0; r < s->ndev; ++r) {
            if (s->rep[r]) { cudaSetDevice(ctx_slots()[r].device); free_srs_rep(s->rep[r]); }
        }
        cudaSetDevice(ctx_slots()[0].device);
        delete s;
        return rc;
    }
    *out = s;
    return SONIC_OK;
}

int sonic_srs_g2_range(const sonic_srs* srs, int family, int64_t exponent, uint64_t count, uint8_t* out) {
    if (!srs || !out || (family != SONIC_FAMILY_PLAIN && family != SONIC_FAMILY_ALPHA)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    if (!srs->rep[0]->g2_points) return fail(SONIC_ERR_INVALID_ARG, "this SRS was generated without its G2 vectors (option \"g2\")");
    const int64_t d = (int64_t)srs->d;
    if (count && (exponent < -d || exponent + (int64_t)count - 1 > d))
        return fail(SONIC_ERR_SRS_TOO_SHORT, "pcV: h vector is not long enough: %" PRId64 " >= %" PRIu64, exponent < -d ? -exponent - 1 : exponent + (int64_t)count - 1, srs->d + 1);
    if (!count) return SONIC_OK;
    return guarded([&](Ctx& cx) {
        uint8_t* d_out = cx.arena.get<uint8_t>(count * 96);
        g2_compress_range(cx, srs->rep[0]->g2_points, srs->rep[0]->index(family, exponent), count, d_out);
        SONIC_CUDA(cudaMemcpyAsync(out, d_out, count * 96, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        return (int)SONIC_OK;
    });
}

void sonic_srs_free(sonic_srs* srs) {
    if (!srs) return;
    Runtime& R = rt();
    std::lock_guard<std::mutex> lock(R.mu);
    for (int r = 0; r < srs->ndev; ++r) {
        Ctx& cx = ctx_slots()[r];
        if (cx.ready) { cudaSetDevice(cx.device); cudaStreamSynchronize(cx.stream); }
        free_srs_rep(srs->rep[r]);
    }
    if (ctx_slots()[0].ready) cudaSetDevice(ctx_slots()[0].device);
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
        const SrsRep& rep = *srs->rep[0];
        FILE* f = fopen(path, "wb");
        if (!f) return fail(SONIC_ERR_INVALID_ARG, "cannot open %s for writing", path);
        SrsFileHeader h;
        memcpy(h.magic, "SONICSRS", 8);
        h.version = 1;
        h.pre_c = (uint32_t)rep.tables.c;
        h.d = srs->d;
        h.levels = rep.tables.c > 0 ? (uint64_t)rep.tables.W : 1;
        h.points_per_level = 2 * rep.stride();
        bool ok = fwrite(&h, sizeof h, 1, f) == 1;
        const size_t total = (size_t)h.levels * h.points_per_level * sizeof(G1Affine);
        const size_t chunk = size_t(64) << 20;
        uint8_t* stage = pinned(cx, chunk);
        for (size_t off = 0; ok && off < total; off += chunk) {
            const size_t nbytes = std::min(chunk, total - off);
            SONIC_CUDA(cudaMemcpyAsync(stage, (const uint8_t*)rep.points + off, nbytes, cudaMemcpyDeviceToHost, cx.stream));
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
    Runtime& R = rt();
    return guarded([&](Ctx& cx) {
        FILE* f = fopen(path, "rb");
        if (!f) return fail(SONIC_ERR_INVALID_ARG, "cannot open %s", path);
        SrsFileHeader h;
        const bool got = fread(&h, sizeof h, 1, f) == 1;
        // the invariants sonic_srs_new enforces: window bits, level count, 31-bit point ids (bit 31 of an MSM entry is the sign)
        bool ok = got && memcmp(h.magic, "SONICSRS", 8) == 0 && h.version == 1 && h.d != 0 && h.d < (1ull << 28) &&
                  h.points_per_level == 2 * (2 * h.d + 1) && (h.pre_c == 0 || (h.pre_c >= 4 && h.pre_c <= 20)) &&
                  h.levels == (h.pre_c == 0 ? 1 : (255 + h.pre_c - 1) / h.pre_c) && h.levels * h.points_per_level < (1ull << 31);
        if (ok) {
            // the file must hold exactly the arrays the header announces
            const size_t total = (size_t)h.levels * h.points_per_level * sizeof(G1Affine);
            ok = fseek(f, 0, SEEK_END) == 0 && (uint64_t)ftell(f) == sizeof h + total && fseek(f, sizeof h, SEEK_SET) == 0;
            if (!ok && got) { fclose(f); return fail(SONIC_ERR_INVALID_ARG, "%s is truncated (or longer than its header says)", path); }
        }
        if (!ok) {
            fclose(f);
            return fail(SONIC_ERR_INVALID_ARG, "%s is not an SRS file of this library", path);
        }
        sonic_srs* s = new sonic_srs;
        s->d = h.d;
        s->ndev = R.ndev;
        const size_t total = (size_t)h.levels * h.points_per_level * sizeof(G1Affine);
        auto cleanup = [&]() {
            for (int r = 0; r < s->ndev; ++r) if (s->rep[r]) { cudaSetDevice(ctx_slots()[r].device); free_srs_rep(s->rep[r]); }
            cudaSetDevice(cx.device);
            delete s;
        };
        for (int r = 0; r < R.ndev; ++r) {
            SrsRep* rep = new SrsRep;
            s->rep[r] = rep;
            rep->d = h.d;
            rep->tables.c = (int)h.pre_c;
            rep->tables.W = h.pre_c ? (int)h.levels : 0;
            rep->tables.stride = (uint32_t)h.points_per_level;
            cudaSetDevice(ctx_slots()[r].device);
            cudaError_t e = cudaMalloc((void**)&rep->points, total);
            cudaSetDevice(cx.device);
            if (e != cudaSuccess) { fclose(f); cleanup(); throw CudaError{e, "cudaMalloc(srs)", __LINE__}; }
        }
        const size_t chunk = size_t(64) << 20;
        uint8_t* stage = pinned(cx, chunk);
        bool rd = true;
        for (size_t off = 0; rd && off < total; off += chunk) {
            const size_t nbytes = std::min(chunk, total - off);
            rd = fread(stage, 1, nbytes, f) == nbytes;
            for (int r = 0; rd && r < R.ndev; ++r) {   // pinned memory is visible to every device; cudaMemcpy picks the route
                cudaError_t ce = cudaMemcpy((uint8_t*)s->rep[r]->points + off, stage, nbytes, cudaMemcpyHostToDevice);
                if (ce != cudaSuccess) { fclose(f); cleanup(); throw CudaError{ce, "upload(srs)", __LINE__}; }
            }
        }
        fclose(f);
        if (!rd) { cleanup(); return fail(SONIC_ERR_INVALID_ARG, "%s is truncated", path); }
        *out = s;
        return (int)SONIC_OK;
    });
}

int sonic_srs_g1_range(const sonic_srs* srs, int family, int64_t exponent, uint64_t count, uint8_t* out) {
    if (!srs || !out || (family != SONIC_FAMILY_PLAIN && family != SONIC_FAMILY_ALPHA)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    const int64_t d = (int64_t)srs->d;
    const bool alpha = family == SONIC_FAMILY_ALPHA;
    if (count >= (1ull << 31)) return fail(SONIC_ERR_INVALID_ARG, "count out of range");
    const int64_t last = exponent + (int64_t)count - 1;
    // the record fields of src/Sonic/SRS.hs:13-18 do not hold anything outside [-d, d], nor g^alpha
    if (count && exponent < -d) return srs_too_short(alpha, exponent, srs->d);
    if (count && alpha && exponent <= 0 && 0 <= last) return srs_too_short(true, 0, srs->d);
    if (count && last > d) return srs_too_short(alpha, std::max(exponent, d + 1), srs->d);
    if (!count) return SONIC_OK;
    return guarded([&](Ctx& cx) {
        const SrsRep& rep = *srs->rep[0];
        uint8_t* d_out = cx.arena.get<uint8_t>(count * 48);
        SONIC_LAUNCH(k_compress_points, div_up(count, 128), 128, 0, rep.points + rep.index(family, exponent), count, d_out);
        SONIC_CUDA(cudaMemcpyAsync(out, d_out, count * 48, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        return (int)SONIC_OK;
    });
}

int sonic_srs_g1(const sonic_srs* srs, int family, int64_t exponent, uint8_t out[48]) {
    return sonic_srs_g1_range(srs, family, exponent, 1, out);
}

// ---- standalone MSM / commitPoly ------------------------------------------------------------------------
// A window that is long enough is cut into one contiguous slice per device; every device reduces its
// slice to a partial sum (96 raw bytes), one all-gather, fold on device 0.  Range violations keep the
// reference's ascending order because the slices ascend with the device index.
static int msm_common(const sonic_srs* srs, int family, bool commit_text, int64_t lo, uint64_t len, const uint8_t* host_scalars,
                      const void* dev_scalars, uint8_t* out48, uint8_t* out_raw) {
    if (!srs || (family != SONIC_FAMILY_PLAIN && family != SONIC_FAMILY_ALPHA) || (!host_scalars && !dev_scalars && len))
        return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    if (len >= (1ull << 28)) return fail(SONIC_ERR_INVALID_ARG, "len out of range");
    if (lo < -(int64_t(1) << 29) || lo > (int64_t(1) << 29)) return fail(SONIC_ERR_INVALID_ARG, "exponent out of range");
    Runtime& R = rt();
    const bool shard = R.ready && R.ndev > 1 && host_scalars && (int64_t)(len / R.ndev) >= R.opt_shard_min_terms;
    if (!shard) {
        return guarded([&](Ctx& cx) {
            Timer tm(cx);
            const Fr* d_scal = dev_scalars ? (const Fr*)dev_scalars : upload_fr(cx, host_scalars, len);
            int rc = check_canonical_dev(cx, d_scal, len);
            if (rc) return rc;
            rc = msm_window(cx, *srs->rep[0], family, commit_text, d_scal, lo, len, out48, out_raw);
            tm.stop();
            return rc;
        });
    }
    return on_all([&](int r, Ctx& cx) -> int {
        const uint64_t a = len * r / R.ndev, b = len * (r + 1) / R.ndev;
        Fq* d_part = cx.arena.get<Fq>(2);
        Fq* d_all = cx.arena.get<Fq>(2 * (size_t)R.ndev);
        int rc = attempt(cx, [&]() -> int {
            Timer tm(cx);
            const Fr* d_scal = upload_fr(cx, host_scalars + 32 * a, b - a);
            int rc = check_canonical_dev(cx, d_scal, b - a);
            if (rc) return rc;
            G1Affine* d_aff = cx.arena.get<G1Affine>(1);
            rc = msm_window_dev(cx, *srs->rep[r], family, commit_text, d_scal, lo + (int64_t)a, b - a, d_aff, nullptr);
            if (rc) return rc;
            SONIC_LAUNCH(k_affine_to_raw, 1, 32, 0, d_aff, 1u, d_part);
            return (int)SONIC_OK;
        });
        if (!R.vote.all_ok(rc == SONIC_OK)) return rc;   // rc == 0 here: another device reports the failure
        return attempt(cx, [&]() -> int {
            SONIC_NCCL(ncclAllGather(d_part, d_all, 96, ncclUint8, R.comm[r], cx.stream));
            if (r == 0) {
                uint8_t* d_out = cx.arena.get<uint8_t>(48);
                uint8_t* h = pinned(cx, 256);
                if (out48) {
                    SONIC_LAUNCH(k_sum_raw, 1, 32, 0, d_all, (uint32_t)R.ndev, d_out, (Fq*)nullptr);
                    SONIC_CUDA(cudaMemcpyAsync(h, d_out, 48, cudaMemcpyDeviceToHost, cx.stream));
                }
                if (out_raw) {
                    Fq* d_sum = cx.arena.get<Fq>(2);
                    SONIC_LAUNCH(k_sum_raw, 1, 32, 0, d_all, (uint32_t)R.ndev, (uint8_t*)nullptr, d_sum);
                    SONIC_CUDA(cudaMemcpyAsync(h + 64, d_sum, 96, cudaMemcpyDeviceToHost, cx.stream));
                }
                SONIC_CUDA(cudaEventRecord(cx.ev[7], cx.stream));
                SONIC_CUDA(cudaStreamSynchronize(cx.stream));
                float ms = 0;
                SONIC_CUDA(cudaEventElapsedTime(&ms, cx.ev[6], cx.ev[7]));
                msm_collect_timing(cx);
                cx.timing_ms["total"] = ms;
                if (out48) memcpy(out48, h, 48);
                if (out_raw) memcpy(out_raw, h + 64, 96);
            } else {
                SONIC_CUDA(cudaStreamSynchronize(cx.stream));
                msm_collect_timing(cx);
            }
            return (int)SONIC_OK;
        });
    });
}

int sonic_msm_g1(const sonic_srs* srs, int family, int64_t lo, uint64_t len, const uint8_t* scalars32, uint8_t out[48]) {
    if (!out) return fail(SONIC_ERR_INVALID_ARG, "null output");
    return msm_common(srs, family, family == SONIC_FAMILY_ALPHA, lo, len, scalars32, nullptr, out, nullptr);
}

int sonic_msm_g1_partial(const sonic_srs* srs, int family, int64_t lo, uint64_t len, const uint8_t* scalars32, uint8_t out_raw[96]) {
    if (!out_raw) return fail(SONIC_ERR_INVALID_ARG, "null output");
    return msm_common(srs, family, family == SONIC_FAMILY_ALPHA, lo, len, scalars32, nullptr, nullptr, out_raw);
}

int sonic_msm_g1_device(const sonic_srs* srs, int family, int64_t lo, uint64_t len, const void* d_scalars32, uint8_t out[48]) {
    if (!out) return fail(SONIC_ERR_INVALID_ARG, "null output");
    return msm_common(srs, family, family == SONIC_FAMILY_ALPHA, lo, len, nullptr, d_scalars32, out, nullptr);
}

int sonic_msm_g1_device_partial(const sonic_srs* srs, int family, int64_t lo, uint64_t len, const void* d_scalars32, uint8_t out_raw[96]) {
    if (!out_raw) return fail(SONIC_ERR_INVALID_ARG, "null output");
    return msm_common(srs, family, family == SONIC_FAMILY_ALPHA, lo, len, nullptr, d_scalars32, nullptr, out_raw);
}

int sonic_g1_sum(const uint8_t* raw96, uint64_t n, uint8_t out[48]) {
    if ((!raw96 && n) || !out || n > (1u << 20)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    return guarded([&](Ctx& cx) {
        Fq* d_raw = cx.arena.get<Fq>(2 * n + 2);
        if (n) SONIC_CUDA(cudaMemcpyAsync(d_raw, raw96, n * 96, cudaMemcpyHostToDevice, cx.stream));
        uint8_t* d_out = cx.arena.get<uint8_t>(48);
        SONIC_LAUNCH(k_sum_raw, 1, 32, 0, d_raw, (uint32_t)n, d_out, (Fq*)nullptr);
        SONIC_CUDA(cudaMemcpyAsync(out, d_out, 48, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        return (int)SONIC_OK;
    });
}

// first / last exponent with a non-zero coefficient in a dense host window (len <= 2^28): the sparse
// reference holds nothing else, so only these can index the SRS
static bool nonzero_span(const uint8_t* coeffs32, uint64_t len, uint64_t* first, uint64_t* last) {
    uint64_t a = 0, b = len;
    while (a < len && fr_bytes_zero(coeffs32 + 32 * a)) ++a;
    if (a == len) return false;
    while (b > a && fr_bytes_zero(coeffs32 + 32 * (b - 1))) --b;
    *first = a;
    *last = b - 1;
    return true;
}

int sonic_commit(const sonic_srs* srs, int64_t max, int64_t lo, uint64_t len, const uint8_t* coeffs32, uint8_t out_g1[48]) {
    if (!srs || !out_g1 || (!coeffs32 && len)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    if (len >= (1ull << 28)) return fail(SONIC_ERR_INVALID_ARG, "len out of range");
    // X^(d-max) * f(X): every exponent moves by d - max (src/Sonic/CommitmentScheme.hs:31-33).  The
    // reference folds the terms in ascending order, so the lowest non-zero term is indexed first: if it
    // already lies outside the SRS (however far: no exponent arithmetic below can overflow after this)
    // the panic is its.
    const int64_t d = (int64_t)srs->d;
    uint64_t fz = 0, lz = 0;
    if (!nonzero_span(coeffs32, len, &fz, &lz)) { lo = 0; len = 0; }
    else { coeffs32 += 32 * fz; lo += (int64_t)fz; len = lz - fz + 1; }
    if (len) {
        const __int128 e0 = (__int128)lo + d - max;
        if (e0 < -(__int128)d || e0 == 0 || e0 > (__int128)d) {
            const __int128 lim = (__int128)1 << 62;
            return srs_too_short(true, (int64_t)(e0 > lim ? lim : e0 < -lim ? -lim : e0), srs->d);
        }
    }
    return msm_common(srs, SONIC_FAMILY_ALPHA, true, len ? lo + (d - max) : 0, len, coeffs32, nullptr, out_g1, nullptr);
}

int sonic_open(const sonic_srs* srs, const uint8_t z[32], int64_t lo, uint64_t len, const uint8_t* coeffs32,
               uint8_t out_v[32], uint8_t out_w[48]) {
    if (!srs || !z || !out_v || !out_w || (!coeffs32 && len)) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    if (len >= (1ull << 27)) return fail(SONIC_ERR_INVALID_ARG, "len out of range");
    if (!fr_bytes_canonical(z)) return fail(SONIC_ERR_NONCANONICAL, "z is not a canonical residue");
    // only the non-zero span matters (the sparse reference holds nothing else)
    uint64_t fz = 0, lz = 0;
    if (!nonzero_span(coeffs32, len, &fz, &lz)) { lo = 0; len = 0; }
    else { coeffs32 += 32 * fz; lo += (int64_t)fz; len = lz - fz + 1; }
    const bool z0 = fr_bytes_zero(z);
    // `eval` multiplies by (recip z)^|lo| when the polynomial has negative powers
    if (len && z0 && lo < 0) return fail(SONIC_ERR_DIV_BY_ZERO, "openPoly: recip 0 (z = 0 with negative exponents)");
    // The dense window below spans [min(lo,0), max(hi,1)): a span reaching beyond +-2^28 cannot lie inside any
    // SRS this library holds (d < 2^28) and is settled here instead of allocating |lo| * 32 bytes.  The quotient's
    // lowest coefficient is -f_lo / z != 0, so a far negative end is exactly the reference's first offending
    // term; at a far positive end the first offending exponent is d+1 unless that quotient coefficient
    // happens to vanish (probability 2^-255 for a sampled z).
    const int64_t lim = int64_t(1) << 28;
    if (len && lo < -lim) return fail(SONIC_ERR_SRS_TOO_SHORT, "openPoly: gNegativeX is not long enough: %" PRId64 " >= %" PRIu64, -lo - 1, srs->d);
    if (len && (lo > lim || lo + (int64_t)len > lim))
        return fail(SONIC_ERR_SRS_TOO_SHORT, "openPoly: gPositiveX is not long enough: %" PRIu64 " >= %" PRIu64, srs->d + 1, srs->d + 1);
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        const SrsRep& rep = *srs->rep[0];
        // the dense window must contain X^0, where f(z) is subtracted (src/Sonic/CommitmentScheme.hs:44)
        int64_t wlo = std::min<int64_t>(lo, 0);
        int64_t whi = std::max<int64_t>(lo + (int64_t)len, 1);
        const uint64_t wlen = (uint64_t)(whi - wlo);
        Fr* d_canon = cx.arena.get<Fr>(wlen);
        SONIC_CUDA(cudaMemsetAsync(d_canon, 0, wlen * 32, cx.stream));
        if (len) SONIC_CUDA(cudaMemcpyAsync(d_canon + (lo - wlo), coeffs32, len * 32, cudaMemcpyHostToDevice, cx.stream));
        int rc = check_canonical_dev(cx, d_canon, wlen);
        if (rc) return rc;
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
        rc = msm_window(cx, rep, SONIC_FAMILY_PLAIN, false, q, wlo, flen - 1, out_w, nullptr);
        if (rc) return rc;
        memcpy(out_v, h + 128, 32);
        tm.stop();
        return (int)SONIC_OK;
    });
}

// ---- circuits -----------------------------------------------------------------------------------------------
// Q bounds the y-dimension of several launches (3Q+4 openings, Q+1 evaluations of s, Q dot products):
// gridDim.y <= 65535, so Q < 2^14 keeps every one of them legal and the refusal is up front.
static const uint64_t SONIC_MAX_Q = 1ull << 14;

static void free_circuit(sonic_circuit* c) {
    for (int r = 0; r < c->ndev; ++r) {
        if (!c->rep[r]) continue;
        cudaSetDevice(ctx_slots()[r].device);
        circuit_free(c->rep[r]);
    }
    if (ctx_slots()[0].ready) cudaSetDevice(ctx_slots()[0].device);
    delete c;
}

int sonic_circuit_load(uint64_t n, uint64_t Q, const uint8_t* wL, const uint8_t* wR, const uint8_t* wO,
                       const uint8_t* cs, sonic_circuit** out) {
    if (!out || !wL || !wR || !wO || !cs) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    // `sPoly` takes n from `head wL` (src/Sonic/Constraints.hs:53): an empty weight list has no n
    if (n == 0 || Q == 0) return fail(SONIC_ERR_INVALID_ARG, "Empty weights");
    if (n >= (1ull << 24) || Q >= SONIC_MAX_Q || n * Q >= (1ull << 28)) return fail(SONIC_ERR_INVALID_ARG, "circuit too large (n < 2^24, Q < 2^14, nQ < 2^28)");
    sonic_circuit* c = new sonic_circuit;
    c->n = n;
    c->Q = Q;
    c->ndev = rt().ndev;
    int rc = on_all([&](int r, Ctx& cx) -> int { return circuit_load(cx, n, Q, wL, wR, wO, cs, &c->rep[r]); });
    if (rc) { c->ndev = MAX_DEV; free_circuit(c); return rc; }
    *out = c;
    return SONIC_OK;
}

int sonic_circuit_load_csr(uint64_t n, uint64_t Q, const uint64_t* rowptr_L, const uint32_t* col_L, const uint8_t* val_L,
                           const uint64_t* rowptr_R, const uint32_t* col_R, const uint8_t* val_R,
                           const uint64_t* rowptr_O, const uint32_t* col_O, const uint8_t* val_O,
                           const uint8_t* cs, sonic_circuit** out) {
    if (!out || !rowptr_L || !rowptr_R || !rowptr_O || !cs) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    *out = nullptr;
    if (n == 0 || Q == 0) return fail(SONIC_ERR_INVALID_ARG, "Empty weights");
    if (n >= (1ull << 24) || Q >= SONIC_MAX_Q) return fail(SONIC_ERR_INVALID_ARG, "circuit too large (n < 2^24, Q < 2^14)");
    const uint64_t* rp[3] = {rowptr_L, rowptr_R, rowptr_O};
    const uint32_t* cl[3] = {col_L, col_R, col_O};
    const uint8_t* vl[3] = {val_L, val_R, val_O};
    for (int m = 0; m < 3; ++m)
        if (rp[m][Q] && (!cl[m] || !vl[m])) return fail(SONIC_ERR_INVALID_ARG, "null CSR arrays");
    sonic_circuit* c = new sonic_circuit;
    c->n = n;
    c->Q = Q;
    c->ndev = rt().ndev;
    int rc = on_all([&](int r, Ctx& cx) -> int { return circuit_load_csr(cx, n, Q, rp, cl, vl, cs, &c->rep[r]); });
    if (rc) { c->ndev = MAX_DEV; free_circuit(c); return rc; }
    *out = c;
    return SONIC_OK;
}

void sonic_circuit_free(sonic_circuit* c) {
    if (!c) return;
    Runtime& R = rt();
    std::lock_guard<std::mutex> lock(R.mu);
    for (int r = 0; r < c->ndev; ++r) {
        Ctx& cx = ctx_slots()[r];
        if (cx.ready) { cudaSetDevice(cx.device); cudaStreamSynchronize(cx.stream); }
    }
    free_circuit(c);
}

uint64_t sonic_rnd_count(uint64_t Q) { return 2 * Q + 8; }
uint64_t sonic_proof_size(uint64_t Q) { return (4 * Q + 7) * 48 + (2 * Q + 5) * 32; }

// ---- prove ----------------------------------------------------------------------------------------------------
namespace {

int prove_precheck(const sonic_srs* srs, const sonic_circuit* circuit, const uint8_t* rnd, uint64_t* n_out, uint64_t* Q_out) {
    const uint64_t n = circuit->n, Q = circuit->Q;
    *n_out = n;
    *Q_out = Q;
    // src/Sonic/Protocol.hs:54-55
    if (srs->d < 7 * n)
        return fail(SONIC_ERR_D_TOO_SMALL, "Parameter d is not large enough: %" PRIu64 " should be greater than %" PRIu64, srs->d, 7 * n);
    // every challenge is used as an evaluation point of a polynomial with negative powers
    for (uint64_t i = 4; i < 2 * Q + 8; ++i)
        if (fr_bytes_zero(rnd + 32 * i)) return fail(SONIC_ERR_DIV_BY_ZERO, "prove: recip 0 (challenge %" PRIu64 " is zero)", i);
    return SONIC_OK;
}

// one whole proof on one device (world = 1): enqueue, one copy of the result buffer, finish
int prove_single(Ctx& cx, SrsRep& srs, const CircuitRep& circ, uint64_t n, uint64_t Q, uint32_t M, bool has_main,
                 const void* assignment, bool assignment_on_device, const void* d_rnd_or_null, const uint8_t* rnd_host,
                 uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    const ProveLayout lay(M, has_main);
    Timer tm(cx);
    const Fr* d_in = (const Fr*)assignment;
    if (has_main && !assignment_on_device) {
        Fr* up = cx.arena.get<Fr>(3 * n);
        SONIC_CUDA(cudaMemcpyAsync(up, assignment, 3 * n * 32, cudaMemcpyHostToDevice, cx.stream));
        d_in = up;
    }
    const Fr* d_rnd = d_rnd_or_null ? (const Fr*)d_rnd_or_null : upload_fr(cx, rnd_host, 2 * (uint64_t)M + 8);
    ensure_tables(cx, srs, n, Q, has_main);
    uint8_t* d_out = cx.arena.get<uint8_t>(lay.out_bytes());
    int rc = prove_enqueue(cx, srs, circ, d_in, d_rnd, M, has_main, 0, 1, d_out);
    if (rc) return rc;
    uint8_t* h = pinned(cx, lay.out_bytes());
    SONIC_CUDA(cudaMemcpyAsync(h, d_out, lay.out_bytes(), cudaMemcpyDeviceToHost, cx.stream));
    tm.stop();
    prove_collect_timing(cx);
    return prove_finish(lay, h, rnd_host, proof_out, cap, written);
}

// one proof over all devices of the runtime: equal runs of its MSM terms per device, one all-gather of
// the exchange records, fold on device 0.  `assignment`: aL | aR | aO contiguous on the host, or three
// separate host pointers.
int prove_all_devices(const sonic_srs* srs, const sonic_circuit* circuit, const uint8_t* aL, const uint8_t* aR,
                      const uint8_t* aO, const void* d_assignment, const void* d_rnd_dev, const uint8_t* rnd, uint8_t* proof_out,
                      uint64_t cap, uint64_t* written) {
    Runtime& R = rt();
    const uint64_t n = circuit->n, Q = circuit->Q;
    const ProveLayout lay((uint32_t)Q, true);
    return on_all([&](int r, Ctx& cx) -> int {
        uint8_t* d_rec = cx.arena.get<uint8_t>(lay.rec_bytes());
        uint8_t* d_all = cx.arena.get<uint8_t>(lay.rec_bytes() * (size_t)R.ndev);
        int rc = attempt(cx, [&]() -> int {
            Timer tm(cx);
            const Fr* d_in;
            const Fr* d_rnd;
            if (d_assignment) {
                // inputs resident in the memory of device 0: the other devices pull them over NVLink (peer copy)
                if (r == 0) {
                    d_in = (const Fr*)d_assignment;
                    d_rnd = (const Fr*)d_rnd_dev;
                } else {
                    Fr* in = cx.arena.get<Fr>(3 * n);
                    Fr* rn = cx.arena.get<Fr>(2 * Q + 8);
                    SONIC_CUDA(cudaMemcpyPeerAsync(in, cx.device, d_assignment, ctx_slots()[0].device, 3 * n * 32, cx.stream));
                    SONIC_CUDA(cudaMemcpyPeerAsync(rn, cx.device, d_rnd_dev, ctx_slots()[0].device, (2 * Q + 8) * 32, cx.stream));
                    d_in = in;
                    d_rnd = rn;
                }
            } else {
                Fr* in = cx.arena.get<Fr>(3 * n);
                SONIC_CUDA(cudaMemcpyAsync(in, aL, n * 32, cudaMemcpyHostToDevice, cx.stream));
                SONIC_CUDA(cudaMemcpyAsync(in + n, aR, n * 32, cudaMemcpyHostToDevice, cx.stream));
                SONIC_CUDA(cudaMemcpyAsync(in + 2 * n, aO, n * 32, cudaMemcpyHostToDevice, cx.stream));
                d_in = in;
                d_rnd = upload_fr(cx, rnd, 2 * Q + 8);
            }
            ensure_tables(cx, *srs->rep[r], n, Q, true);
            return prove_enqueue(cx, *srs->rep[r], *circuit->rep[r], d_in, d_rnd, (uint32_t)Q, true, (uint32_t)r, (uint32_t)R.ndev, d_rec);
        });
        if (!R.vote.all_ok(rc == SONIC_OK)) return rc;
        return attempt(cx, [&]() -> int {
            nvtxRangePushA("sonic.prove.exchange");
            SONIC_NCCL(ncclAllGather(d_rec, d_all, lay.rec_bytes(), ncclUint8, R.comm[r], cx.stream));
            int rc = SONIC_OK;
            if (r == 0) {
                uint8_t* d_out = cx.arena.get<uint8_t>(lay.out_bytes());
                prove_fold_enqueue(cx, lay, (uint32_t)R.ndev, d_all, d_out);
                uint8_t* h = pinned(cx, lay.out_bytes());
                SONIC_CUDA(cudaMemcpyAsync(h, d_out, lay.out_bytes(), cudaMemcpyDeviceToHost, cx.stream));
                SONIC_CUDA(cudaEventRecord(cx.ev[7], cx.stream));
                SONIC_CUDA(cudaStreamSynchronize(cx.stream));
                float ms = 0;
                SONIC_CUDA(cudaEventElapsedTime(&ms, cx.ev[6], cx.ev[7]));
                prove_collect_timing(cx);
                cx.timing_ms["total"] = ms;
                rc = prove_finish(lay, h, rnd, proof_out, cap, written);
            } else {
                SONIC_CUDA(cudaEventRecord(cx.ev[7], cx.stream));
                SONIC_CUDA(cudaStreamSynchronize(cx.stream));
                float ms = 0;
                SONIC_CUDA(cudaEventElapsedTime(&ms, cx.ev[6], cx.ev[7]));
                prove_collect_timing(cx);
                cx.timing_ms["total"] = ms;
            }
            nvtxRangePop();
            return rc;
        });
    });
}

}  // namespace

int sonic_prove(const sonic_srs* srs, const sonic_circuit* circuit, const uint8_t* aL, const uint8_t* aR,
                const uint8_t* aO, const uint8_t* rnd, uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    if (!srs || !circuit || !aL || !aR || !aO || !rnd || !proof_out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    uint64_t n, Q;
    int rc = prove_precheck(srs, circuit, rnd, &n, &Q);
    if (rc) return rc;
    if (rt().ready && rt().ndev > 1) return prove_all_devices(srs, circuit, aL, aR, aO, nullptr, nullptr, rnd, proof_out, cap, written);
    return guarded([&](Ctx& cx) {
        Fr* d_in = cx.arena.get<Fr>(3 * n);
        SONIC_CUDA(cudaMemcpyAsync(d_in, aL, n * 32, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(d_in + n, aR, n * 32, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(d_in + 2 * n, aO, n * 32, cudaMemcpyHostToDevice, cx.stream));
        return prove_single(cx, *srs->rep[0], *circuit->rep[0], n, Q, (uint32_t)Q, true, d_in, true, nullptr, rnd, proof_out, cap, written);
    });
}

// `count` independent proofs of one circuit (BASELINE config 5): whole proofs are dealt round-robin to
// the devices, which work through theirs at the same time; nothing is exchanged.
int sonic_prove_batch(const sonic_srs* srs, const sonic_circuit* circuit, uint64_t count, const uint8_t* assignments,
                      const uint8_t* rnds, uint8_t* proofs_out, uint64_t cap, uint64_t* written) {
    if (!srs || !circuit || (count && (!assignments || !rnds || !proofs_out))) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    const uint64_t n = circuit->n, Q = circuit->Q;
    const uint64_t psize = sonic_proof_size(Q), nr = 2 * Q + 8;
    if (written) *written = count * psize;
    if (cap < count * psize) return fail(SONIC_ERR_BUFFER_TOO_SMALL, "%" PRIu64 " proofs need %" PRIu64 " bytes", count, count * psize);
    for (uint64_t i = 0; i < count; ++i) {
        uint64_t n_, Q_;
        int rc = prove_precheck(srs, circuit, rnds + i * nr * 32, &n_, &Q_);
        if (rc) return rc;
    }
    Runtime& R = rt();
    return on_all([&](int r, Ctx& cx) -> int {
        float total = 0;
        for (uint64_t i = (uint64_t)r; i < count; i += (uint64_t)R.ndev) {
            cx.arena.reset();
            uint64_t w = 0;
            int rc = prove_single(cx, *srs->rep[r], *circuit->rep[r], n, Q, (uint32_t)Q, true, assignments + i * 3 * n * 32, false, nullptr,
                                  rnds + i * nr * 32, proofs_out + i * psize, psize, &w);
            if (rc) return rc;
            total += (float)cx.timing_ms["total"];
        }
        cx.timing_ms["batch"] = total;
        return (int)SONIC_OK;
    });
}

// ---- one proof sharded over several PROCESSES (one per GPU), the caller moves the records ---------------------
uint64_t sonic_shard_exchange_size(uint64_t Q) { return ProveLayout((uint32_t)Q, true).rec_bytes(); }
uint64_t sonic_shard_blob_size(uint64_t Q) { return ProveLayout((uint32_t)Q, true).rec_bytes() + 64; }

int sonic_prove_shard_sink(const sonic_srs* srs, const sonic_circuit* circuit, const void* assignment, int assignment_on_device,
                           const void* d_rnd_or_null, const uint8_t* rnd_host, uint32_t rank, uint32_t world,
                           uint8_t* blob_out, uint64_t cap, uint64_t* written, void* d_record_out) {
    if (!srs || !circuit || !assignment || !rnd_host || (!blob_out && !d_record_out)) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    if (world < 2 || rank >= world || world > 64) return fail(SONIC_ERR_INVALID_ARG, "need 2 <= world <= 64 and rank < world");
    uint64_t n, Q;
    int rc = prove_precheck(srs, circuit, rnd_host, &n, &Q);
    if (rc) return rc;
    const ProveLayout lay((uint32_t)Q, true);
    if (written) *written = lay.rec_bytes() + 64;
    if (blob_out && cap < lay.rec_bytes() + 64) return fail(SONIC_ERR_BUFFER_TOO_SMALL, "shard blob needs %llu bytes", (unsigned long long)(lay.rec_bytes() + 64));
    return guarded([&](Ctx& cx) {
        Timer tm(cx);
        const Fr* d_in = (const Fr*)assignment;
        if (!assignment_on_device) {
            Fr* up = cx.arena.get<Fr>(3 * n);
            SONIC_CUDA(cudaMemcpyAsync(up, assignment, 3 * n * 32, cudaMemcpyHostToDevice, cx.stream));
            d_in = up;
        }
        const Fr* d_rnd = d_rnd_or_null ? (const Fr*)d_rnd_or_null : upload_fr(cx, rnd_host, 2 * Q + 8);
        ensure_tables(cx, *srs->rep[0], n, Q, true);
        uint8_t* d_rec = cx.arena.get<uint8_t>(lay.rec_bytes());
        int rc = prove_enqueue(cx, *srs->rep[0], *circuit->rep[0], d_in, d_rnd, (uint32_t)Q, true, rank, world, d_rec);
        if (rc) return rc;
        if (d_record_out) SONIC_CUDA(cudaMemcpyAsync(d_record_out, d_rec, lay.rec_bytes(), cudaMemcpyDeviceToDevice, cx.stream));
        if (blob_out) {
            uint8_t* h = pinned(cx, lay.rec_bytes());
            SONIC_CUDA(cudaMemcpyAsync(h, d_rec, lay.rec_bytes(), cudaMemcpyDeviceToHost, cx.stream));
            tm.stop();
            memcpy(blob_out, h, lay.rec_bytes());
            memcpy(blob_out + lay.rec_bytes(), rnd_host + 32 * (6 + 2 * Q), 64);   // hscU, hscV
        } else {
            tm.stop();
        }
        prove_collect_timing(cx);
        return (int)SONIC_OK;
    });
}

int sonic_prove_shard(const sonic_srs* srs, const sonic_circuit* circuit, const uint8_t* aL, const uint8_t* aR,
                      const uint8_t* aO, const uint8_t* rnd, uint32_t rank, uint32_t world, uint8_t* blob_out,
                      uint64_t cap, uint64_t* written) {
    if (!srs || !circuit || !aL || !aR || !aO || !rnd || !blob_out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    if (world < 2) return fail(SONIC_ERR_INVALID_ARG, "need 2 <= world <= 64 and rank < world");
    const uint64_t n = circuit->n;
    std::vector<uint8_t> in(3 * n * 32);
    memcpy(in.data(), aL, n * 32);
    memcpy(in.data() + n * 32, aR, n * 32);
    memcpy(in.data() + 2 * n * 32, aO, n * 32);
    return sonic_prove_shard_sink(srs, circuit, in.data(), 0, nullptr, rnd, rank, world, blob_out, cap, written, nullptr);
}

int sonic_prove_shard_device(const sonic_srs* srs, const sonic_circuit* circuit, const void* d_assignment,
                             const void* d_rnd, const uint8_t* rnd_host, uint32_t rank, uint32_t world,
                             uint8_t* out, uint64_t cap, uint64_t* written) {
    if (!srs || !circuit || !d_assignment || !d_rnd || !rnd_host || !out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    if (world == 1) return sonic_prove_device(srs, circuit, d_assignment, d_rnd, rnd_host, out, cap, written);
    return sonic_prove_shard_sink(srs, circuit, d_assignment, 1, d_rnd, rnd_host, rank, world, out, cap, written, nullptr);
}

static int combine_common(uint64_t Q, uint32_t world, const uint8_t* host_blobs, const void* d_gathered, const uint8_t* uv,
                          uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    const ProveLayout lay((uint32_t)Q, true);
    return guarded([&](Ctx& cx) {
        const uint8_t* d_recs = (const uint8_t*)d_gathered;
        if (!d_recs) {
            uint8_t* up = cx.arena.get<uint8_t>(lay.rec_bytes() * (size_t)world);
            for (uint32_t r = 0; r < world; ++r)
                SONIC_CUDA(cudaMemcpyAsync(up + r * lay.rec_bytes(), host_blobs + r * (lay.rec_bytes() + 64), lay.rec_bytes(), cudaMemcpyHostToDevice, cx.stream));
            d_recs = up;
        }
        uint8_t* d_out = cx.arena.get<uint8_t>(lay.out_bytes());
        prove_fold_enqueue(cx, lay, world, d_recs, d_out);
        uint8_t* h = pinned(cx, lay.out_bytes());
        SONIC_CUDA(cudaMemcpyAsync(h, d_out, lay.out_bytes(), cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        // prove_finish reads hscU, hscV at their place in the draw order
        std::vector<uint8_t> rnd((2 * Q + 8) * 32, 0);
        memcpy(&rnd[32 * (6 + 2 * Q)], uv, 64);
        return prove_finish(lay, h, rnd.data(), proof_out, cap, written);
    });
}

int sonic_prove_combine_device(uint64_t Q, uint32_t world, const void* d_gathered, const uint8_t* own_blob,
                               uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    if (!d_gathered || !own_blob || !proof_out || world < 1 || world > 64 || Q == 0 || Q >= SONIC_MAX_Q) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    const ProveLayout lay((uint32_t)Q, true);
    return combine_common(Q, world, nullptr, d_gathered, own_blob + lay.rec_bytes(), proof_out, cap, written);
}

int sonic_prove_combine(uint64_t Q, uint32_t world, const uint8_t* blobs, uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    if (!blobs || !proof_out || world < 1 || world > 64 || Q == 0 || Q >= SONIC_MAX_Q) return fail(SONIC_ERR_INVALID_ARG, "bad argument");
    const ProveLayout lay((uint32_t)Q, true);
    return combine_common(Q, world, blobs, nullptr, blobs + lay.rec_bytes(), proof_out, cap, written);
}

int sonic_prove_device(const sonic_srs* srs, const sonic_circuit* circuit, const void* d_assignment,
                       const void* d_rnd, const uint8_t* rnd_host, uint8_t* proof_out, uint64_t cap, uint64_t* written) {
    if (!srs || !circuit || !d_assignment || !d_rnd || !rnd_host || !proof_out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    uint64_t n, Q;
    int rc = prove_precheck(srs, circuit, rnd_host, &n, &Q);
    if (rc) return rc;
    if (rt().ready && rt().ndev > 1) return prove_all_devices(srs, circuit, nullptr, nullptr, nullptr, d_assignment, d_rnd, rnd_host, proof_out, cap, written);
    return guarded([&](Ctx& cx) {
        return prove_single(cx, *srs->rep[0], *circuit->rep[0], n, Q, (uint32_t)Q, true, d_assignment, true, d_rnd, rnd_host, proof_out, cap, written);
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
        return prove_single(cx, *srs->rep[0], *circuit->rep[0], circuit->n, circuit->Q, (uint32_t)m, false, nullptr, true, nullptr, rnd.data(), out, cap, written);
    });
}

int sonic_hsc_prove_terms(const sonic_srs* srs, uint64_t nterms, const int64_t* eX, const int64_t* eY, const uint8_t* coeff32,
                          uint64_t m, const uint8_t* yzs, const uint8_t* uv, uint8_t* out, uint64_t cap, uint64_t* written) {
    if (!srs || (nterms && (!eX || !eY || !coeff32)) || (!yzs && m) || !uv || !out) return fail(SONIC_ERR_INVALID_ARG, "null argument");
    if (m >= (1u << 12) || nterms >= (1ull << 28)) return fail(SONIC_ERR_INVALID_ARG, "too many (y, z) pairs or terms");
    std::vector<uint8_t> rnd((2 * m + 8) * 32, 0);
    for (uint64_t j = 0; j < m; ++j) {
        memcpy(&rnd[(6 + j) * 32], yzs + 64 * j, 32);
        memcpy(&rnd[(6 + m + j) * 32], yzs + 64 * j + 32, 32);
    }
    memcpy(&rnd[(6 + 2 * m) * 32], uv, 64);
    // a zero evaluation point only matters where a negative power is taken (`eval` / `pow` with recip): keep
    // the reference's behaviour by refusing it exactly then
    bool neg_x = false, neg_y = false;
    for (uint64_t t = 0; t < nterms; ++t) {
        if (fr_bytes_zero(coeff32 + 32 * t)) continue;
        neg_x = neg_x || eX[t] < 0;
        neg_y = neg_y || eY[t] < 0;
    }
    for (uint64_t j = 0; j < m; ++j) {
        if (neg_y && fr_bytes_zero(&rnd[(6 + j) * 32])) return fail(SONIC_ERR_DIV_BY_ZERO, "hscProve: recip 0 (y_%" PRIu64 " = 0 with negative powers of Y)", j + 1);
        if (neg_x && fr_bytes_zero(&rnd[(6 + m + j) * 32])) return fail(SONIC_ERR_DIV_BY_ZERO, "hscProve: recip 0 (z_%" PRIu64 " = 0 with negative powers of X)", j + 1);
    }
    if ((neg_x && fr_bytes_zero(&rnd[(6 + 2 * m) * 32])) || (neg_y && fr_bytes_zero(&rnd[(7 + 2 * m) * 32])))
        return fail(SONIC_ERR_DIV_BY_ZERO, "hscProve: recip 0 (u or v is zero with negative powers)");
    for (uint64_t i = 6; i < 2 * m + 8; ++i)
        if (fr_bytes_zero(&rnd[32 * i])) return fail(SONIC_ERR_INVALID_ARG, "hscProve: a zero evaluation point is not supported by this entry");
    return guarded([&](Ctx& cx) {
        const ProveLayout lay((uint32_t)m, false);
        Timer tm(cx);
        const Fr* d_rnd = upload_fr(cx, rnd.data(), 2 * m + 8);
        uint8_t* d_out = cx.arena.get<uint8_t>(lay.out_bytes());
        int rc = hsc_terms_enqueue(cx, *srs->rep[0], nterms, eX, eY, coeff32, (uint32_t)m, d_rnd, d_out);
        if (rc) return rc;
        uint8_t* h = pinned(cx, lay.out_bytes());
        SONIC_CUDA(cudaMemcpyAsync(h, d_out, lay.out_bytes(), cudaMemcpyDeviceToHost, cx.stream));
        tm.stop();
        msm_collect_timing(cx);
        return prove_finish(lay, h, rnd.data(), out, cap, written);
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
    Runtime& R = rt();
    std::lock_guard<std::mutex> lock(R.mu);
    // options are per context; every device of the runtime gets the same value
    auto each = [&](auto&& set) { for (int r = 0; r < MAX_DEV; ++r) set(ctx_slots()[r]); };
    if (!strcmp(name, "window_bits")) {
        if (value != 0 && (value < 4 || value > 20)) return fail(SONIC_ERR_INVALID_ARG, "window_bits must be 0 or in [4, 20]");
        each([&](Ctx& cx) { cx.opt_window_bits = (int)value; });
    } else if (!strcmp(name, "g2")) {
        each([&](Ctx& cx) { cx.opt_g2 = value != 0; });
    } else if (!strcmp(name, "precompute")) {
        if (value != -1 && value != 0 && (value < 4 || value > 20)) return fail(SONIC_ERR_INVALID_ARG, "precompute must be -1 (auto), 0 (off) or window bits in [4, 20]");
        each([&](Ctx& cx) { cx.opt_precompute = (int)value; });
    } else if (!strcmp(name, "precompute_budget_mb")) {
        if (value < 0) return fail(SONIC_ERR_INVALID_ARG, "budget must be >= 0");
        each([&](Ctx& cx) { cx.opt_precompute_budget = (uint64_t)value << 20; });
    } else if (!strcmp(name, "reduce_mode")) {
        if (value < 0 || value > 3) return fail(SONIC_ERR_INVALID_ARG, "reduce_mode must be 0 (automatic), 1 (level by level), 2 (thread per K buckets) or 3 (quads of lanes)");
        each([&](Ctx& cx) { cx.opt_reduce_mode = (int)value; });
    } else if (!strcmp(name, "sort_mode")) {
        // 0: thread per term + global atomics; 1: tiled counting sort, automatic tile count; 2..64: tiled, that many tiles per SM
        if (value < 0 || value > 64) return fail(SONIC_ERR_INVALID_ARG, "sort_mode must be in [0, 64]");
        each([&](Ctx& cx) { cx.opt_sort_mode = (int)value; });
    } else if (!strcmp(name, "sort_reserve")) {
        if (value < 0 || value > 1) return fail(SONIC_ERR_INVALID_ARG, "sort_reserve must be 0 or 1");
        each([&](Ctx& cx) { cx.opt_sort_reserve = (int)value; });
    } else if (!strcmp(name, "reduce_k")) {
        if (value < 0 || value > 256) return fail(SONIC_ERR_INVALID_ARG, "reduce_k must be in [0, 256]");
        each([&](Ctx& cx) { cx.opt_reduce_k = (int)value; });
    } else if (!strcmp(name, "acc_mode")) {
        if (value < 0 || value > 3) return fail(SONIC_ERR_INVALID_ARG, "acc_mode must be 0 (XYZZ in registers), 1 (XYZZ, operand file in shared memory), 2 (affine, batched inversions) or 3 (automatic)");
        each([&](Ctx& cx) { cx.opt_acc_mode = (int)value; });
    } else if (!strcmp(name, "acc_blocks")) {
        if (value != 3) return fail(SONIC_ERR_INVALID_ARG, "acc_blocks is fixed at 3 in this build");
        each([&](Ctx& cx) { cx.opt_acc_blocks = (int)value; });
    } else if (!strcmp(name, "chunk")) {
        if (value < 0 || value > 4096) return fail(SONIC_ERR_INVALID_ARG, "chunk must be in [0, 4096]");
        each([&](Ctx& cx) { cx.opt_chunk = (int)value; });
    } else if (!strcmp(name, "aff_fused")) {
        if (value < 0 || value > 2) return fail(SONIC_ERR_INVALID_ARG, "aff_fused must be 0, 1 or 2");
        each([&](Ctx& cx) { cx.opt_aff_fused = (int)value; });
    } else if (!strcmp(name, "aff_m")) {
        if (value != 0 && value != 8 && value != 16 && value != 32 && value != 64) return fail(SONIC_ERR_INVALID_ARG, "aff_m must be 0, 8, 16, 32 or 64");
        each([&](Ctx& cx) { cx.opt_aff_m = (int)value; });
    } else if (!strcmp(name, "aff_tail")) {
        if (value < 0 || value > 5) return fail(SONIC_ERR_INVALID_ARG, "aff_tail must be in [0, 5]");
        each([&](Ctx& cx) { cx.opt_aff_tail = (int)value; });
    } else if (!strcmp(name, "overlap")) {
        if (value < 0 || value > 1) return fail(SONIC_ERR_INVALID_ARG, "overlap must be 0 or 1");
        each([&](Ctx& cx) { cx.opt_overlap = (int)value; });
    } else if (!strcmp(name, "heavy_mode")) {
        if (value < 0 || value > 1) return fail(SONIC_ERR_INVALID_ARG, "heavy_mode must be 0 or 1");
        each([&](Ctx& cx) { cx.opt_heavy_mode = (int)value; });
    } else if (!strcmp(name, "chunk_max")) {
        if (value < 0 || value > 4096) return fail(SONIC_ERR_INVALID_ARG, "chunk_max must be in [0, 4096]");
        each([&](Ctx& cx) { cx.opt_chunk_max = (int)value; });
    } else if (!strcmp(name, "shard_min_terms")) {
        if (value < 1) return fail(SONIC_ERR_INVALID_ARG, "shard_min_terms must be >= 1");
        R.opt_shard_min_terms = value;
    } else {
        return fail(SONIC_ERR_INVALID_ARG, "unknown option %s", name);
    }
    return SONIC_OK;
}

double sonic_last_timing_ms_dev(int slot, const char* stage) {
    Runtime& R = rt();
    if (slot < 0 || slot >= MAX_DEV) return 0.0;
    std::lock_guard<std::mutex> lock(R.mu);
    Ctx& cx = ctx_slots()[slot];
    auto it = cx.timing_ms.find(stage ? stage : "total");
    return it == cx.timing_ms.end() ? 0.0 : it->second;
}

double sonic_last_timing_ms(const char* stage) { return sonic_last_timing_ms_dev(0, stage); }

uint64_t sonic_launch_count(void) {
    uint64_t t = 0;
    for (int r = 0; r < MAX_DEV; ++r) t += ctx_slots()[r].launches;
    return t;
}

int sonic_bench_mark(int slot) {
    if (slot < 0 || slot > 5) return fail(SONIC_ERR_INVALID_ARG, "bench mark slot must be in [0, 5]");
    Runtime& R = rt();
    if (!R.ready) return fail(SONIC_ERR_NOT_INITIALISED, "sonic_init has not been called");
    std::lock_guard<std::mutex> lock(R.mu);
    Ctx& cx = ctx_slots()[0];
    cudaSetDevice(cx.device);
    if (cudaEventRecord(cx.ev[10 + slot], cx.stream) != cudaSuccess) return fail(SONIC_ERR_CUDA, "cudaEventRecord failed");
    return SONIC_OK;
}

double sonic_bench_elapsed_ms(int from_slot, int to_slot) {
    if (from_slot < 0 || from_slot > 5 || to_slot < 0 || to_slot > 5) return -1.0;
    Runtime& R = rt();
    if (!R.ready) return -1.0;
    std::lock_guard<std::mutex> lock(R.mu);
    Ctx& cx = ctx_slots()[0];
    cudaSetDevice(cx.device);
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

double sonic_selftest_latency_ns(int op, int iters, int blocks, int threads) {
    if (op < 0 || op > 8 || iters < 1 || iters > (1 << 20) || blocks < 1 || blocks > (1 << 16) || threads < 1 || threads > 1024 || (threads & 31)) return -1.0;
    double ns = -1.0;
    guarded([&](Ctx& cx) {
        ns = selftest_latency_ns(cx, op, iters, blocks, threads);
        return (int)SONIC_OK;
    });
    return ns;
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
