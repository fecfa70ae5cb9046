// Host-side interfaces between the translation units of libsonic_b200.
#pragma once
#include <vector>

#include "common.cuh"
#include "g1.cuh"

namespace sonic {

// ---- msm.cu ------------------------------------------------------------------------------
struct MsmJob {
    uint32_t point_base;  // index of the base for term 0 in the unified SRS point array
    uint32_t n;           // number of terms
    uint32_t scalar_off;  // index (in scalars) of term 0 in the scalar array
    uint32_t pad;
};

constexpr int MSM_MAX_JOBS = 64;
constexpr int MSM_HEAVY_PIECES = 24;
constexpr int MSM_RED_THREADS = 128;

struct MsmJobTable {
    MsmJob job[MSM_MAX_JOBS];
    uint32_t prefix[MSM_MAX_JOBS + 1];
    int M;
};


// d_scalars: canonical little-endian scalars, 8 words each.  Results: one affine point
// (Montgomery form) and/or one 48-byte compressed encoding per job, in device memory.
// Precomputed window multiples of the bases: level j of the point array holds 2^(c j) * P for
// every base P, `stride` points per level.  With them all windows of a job share one bucket set
// (no per-window reduction, no Horner tail).  c == 0: no tables, plain windowed Pippenger.
struct MsmTables {
    int c = 0;
    int W = 0;
    uint32_t stride = 0;
};
// `sync` (optional) orders two batches that run on two streams: the bucket accumulation -- the one stage that
// fills the machine -- waits for `wait_before_acc` and signals `signal_after_acc`.
struct MsmSync {
    cudaEvent_t wait_before_acc = nullptr;
    cudaEvent_t signal_after_acc = nullptr;
    bool second = false;   // the second half: its stage events and counters go to their own slots
};
void msm_run(Ctx& cx, const G1Affine* d_points, const MsmTables& tables, const uint32_t* d_scalars,
             const std::vector<MsmJob>& jobs, G1Affine* d_out_aff, uint8_t* d_out_comp, const MsmSync* sync = nullptr);
void msm_collect_timing(Ctx& cx);
void exclusive_scan_u32(Arena& ar, const uint32_t* in, uint32_t* out, uint32_t n);


// ---- poly.cu -----------------------------------------------------------------------------
struct OpenJob {
    const Fr* f;       // coefficients, Montgomery
    const Fr* pz;      // z^k
    const Fr* pzi;     // z^-k  (unused when z == 0)
    Fr* q_canon;       // out: len-1 quotient coefficients for exponents lo .. lo+len-2, canonical (may be null)
    Fr* value_canon;   // out: f(z), canonical
    uint32_t len;
    int32_t lo;
    uint32_t z_is_zero;
    uint32_t pad;
};

struct NttPlan {
    uint32_t logL = 0;
    Fr* tw = nullptr;    // omega^k
    Fr* twi = nullptr;   // omega^-k
    Fr* params = nullptr;
};

void fr_to_mont(Ctx& cx, const Fr* in, Fr* out, uint64_t n, uint32_t* bad_flag);  // flags encodings >= r
void fr_from_mont(Ctx& cx, const Fr* in, Fr* out, uint64_t n);
void fr_inv_few(Ctx& cx, const Fr* in, Fr* out, int n);
void fr_mul_pointwise(Ctx& cx, Fr* a, const Fr* b, uint32_t n);
// tab[t*stride + k] = bases[t]^k, k < len
// with d_sel: only the ntab rows listed there (row indices into bases / tab)
void pow_tables(Ctx& cx, const Fr* bases, int ntab, Fr* tab, uint64_t len, uint64_t stride, const uint32_t* d_sel = nullptr);
// runs the three open passes over a batch of jobs (value + quotient per job)
void open_batch(Ctx& cx, const std::vector<OpenJob>& jobs);
NttPlan ntt_prepare(Ctx& cx, uint32_t logL);
void ntt_forward(const NttPlan& p, Fr* a);
void ntt_inverse(const NttPlan& p, Fr* a);

}  // namespace sonic

namespace sonic {
// ---- selftest.cu -------------------------------------------------------------------------
int selftest_field(Ctx& cx, int which, int op, const void* a, const void* b, void* out, uint32_t n);
int pcv_fold(Ctx& cx, uint32_t k, const uint8_t* F48, const uint8_t* W48, const uint8_t* v32, const uint8_t* z32,
             const uint8_t* r32, const uint32_t* group, uint32_t ngroups, uint8_t* out48);
int selftest_g1(Ctx& cx, int op, const void* a, const void* b, void* out_aff, void* out_comp, uint32_t n);
double selftest_latency_ns(Ctx& cx, int op, int iters, int blocks, int threads);

// ---- resident SRS, one replica per device --------------------------------------------------
// Device layout: one array of affine points indexed by exponent,
//   points[family * (2d+1) + (k + d)],  k in [-d, d],  family 0 = plain, 1 = alpha;
// the alpha slot k = 0 holds the infinity marker (0,0) and is never referenced by a job.
//
// Window tables restricted to the exponent ranges one circuit size touches (tables.cu): level j holds
// 2^(c j) * P for the points of up to four exponent ranges, laid out range after range, `size` points
// per level; level 0 is a copy of the SRS points themselves so that an MSM reads one array.
struct TableRange {
    int family;
    int64_t lo, hi;     // exponents, inclusive
    uint32_t offset;    // first slot of the range inside a level
};
struct RestrictedTables {
    int nranges = 0;
    TableRange range[4];
    uint32_t size = 0;             // points per level
    MsmTables tables;              // c, W, stride = size
    G1Affine* points = nullptr;    // W x size
    // slot of (family, exponent lo .. lo+len-1) if one range holds the whole window, else -1
    int64_t find(int family, int64_t lo, uint64_t len) const {
        for (int i = 0; i < nranges; ++i)
            if (range[i].family == family && lo >= range[i].lo && lo + (int64_t)len - 1 <= range[i].hi)
                return (int64_t)range[i].offset + (lo - range[i].lo);
        return -1;
    }
};
struct SrsRep {
    uint64_t d = 0;
    G1Affine* points = nullptr;  // levels x 2*(2d+1) affine points, Montgomery form; level 0 is the SRS
    MsmTables tables;            // precomputed window multiples of the WHOLE range (c == 0: level 0 only)
    void* g2_points = nullptr;   // optional: 2*(2d+1) affine G2 points (the h-vectors), Montgomery form
    RestrictedTables rt;         // built on demand by the first proof of a circuit size when `tables` is empty
    uint64_t stride() const { return 2 * d + 1; }
    uint64_t index(int family, int64_t k) const { return (uint64_t)family * stride() + (uint64_t)(k + (int64_t)d); }
};

// ---- srs.cu ------------------------------------------------------------------------------
// Generates elements [first, first + count) of the resident point array taken as one flat array of
// levels x 2(2d+1) points (the whole array when count covers it); a multi-GPU runtime gives each device one
// equal slice and all-gathers once.
// d_canon: x, alpha canonical (2 Fr) in device memory.  pre_c > 0 additionally fills levels
// 1..W-1 with the 2^(pre_c j) multiples (W = ceil(255/pre_c)).
// d_g2_points (nullable): also fills the G2 h-vectors, 2*(2d+1) affine G2 points, same exponent indexing, no hole.
void srs_generate(Ctx& cx, uint64_t d, const Fr* d_canon, G1Affine* d_points, int pre_c, uint64_t first, uint64_t count,
                  void* d_g2_points = nullptr);
void srs_generate_g2(Ctx& cx, const Fr* scal_m, uint64_t npts, void* d_g2_points);
void g2_compress_range(Ctx& cx, const void* d_g2_points, uint64_t first, uint64_t count, uint8_t* d_out);
size_t g2_point_bytes();
// ---- srs.cu: restricted window tables -----------------------------------------------------------------------------
// Window tables for `ranges` by repeated doubling of the resident points (no trapdoor needed).
void tables_build(Ctx& cx, const SrsRep& srs, const TableRange* ranges, int nranges, int c, RestrictedTables* out);
void tables_free(RestrictedTables* t);

// ---- prove.cu ----------------------------------------------------------------------------
struct CircuitRep;
int circuit_load(Ctx& cx, uint64_t n, uint64_t Q, const uint8_t* wL, const uint8_t* wR, const uint8_t* wO,
                 const uint8_t* cs, CircuitRep** out);
int circuit_load_csr(Ctx& cx, uint64_t n, uint64_t Q, const uint64_t* const row_ptr[3], const uint32_t* const col[3],
                     const uint8_t* const val[3], const uint8_t* cs, CircuitRep** out);
void circuit_free(CircuitRep* c);

// Sizes of the buffers of one proof (M = number of (y_j, z_j) pairs; M = Q inside `prove`).
struct ProveLayout {
    uint32_t M;
    bool has_main;
    uint32_t nm;   // G1 elements: 4M+7 (prove) or 4M+2 (hscProve alone)
    uint32_t nv;   // field values computed on the device: 2M+3 or 2M
    uint32_t nF;   // field values of the record: nv + hscU, hscV
    ProveLayout(uint32_t M_, bool main_) : M(M_), has_main(main_), nm(main_ ? 4 * M_ + 7 : 4 * M_ + 2),
                                           nv(main_ ? 2 * M_ + 3 : 2 * M_), nF(nv + 2) {}
    // status words (u32): 3 nm range flags | 1 encoding flag | srsD (low, high: the panic texts quote it)
    size_t status_words() const { return 3 * (size_t)nm + 3; }
    // result buffer: nm compressed G1 | nv values | status
    size_t out_vals() const { return (size_t)nm * 48; }
    size_t out_status() const { return out_vals() + (size_t)nv * 32; }
    size_t out_bytes() const { return (out_status() + status_words() * 4 + 31) & ~size_t(31); }
    // exchange record of one rank of a sharded proof: nm raw partial sums (96 B) | nv values | status
    size_t rec_vals() const { return (size_t)nm * 96; }
    size_t rec_status() const { return rec_vals() + (size_t)nv * 32; }
    size_t rec_bytes() const { return (rec_status() + status_words() * 4 + 31) & ~size_t(31); }
    size_t proof_bytes() const { return (size_t)nm * 48 + (size_t)nF * 32; }
};

// Enqueues one proof (or one rank's share of it) on the context's stream.
// d_in: canonical aL|aR|aO (3n Fr); d_rnd: canonical draws in the reference's order, 2M+8 of them.
// world == 1: d_result receives the result buffer (ProveLayout::out_bytes).  world > 1: this rank's run of the
// proof's MSM terms and the field values it owns; d_result receives the exchange record (rec_bytes).
// Nothing is synchronised: the caller gathers / folds / copies and then calls prove_finish.
int prove_enqueue(Ctx& cx, SrsRep& srs, const CircuitRep& circ, const Fr* d_in, const Fr* d_rnd, uint32_t M,
                  bool has_main, uint32_t rank, uint32_t world, uint8_t* d_result);
// [world] exchange records (device) -> result buffer (device): `<>` per commitment, the one
// contribution per field value, the first range violation of every MSM
void prove_fold_enqueue(Ctx& cx, const ProveLayout& lay, uint32_t world, const uint8_t* d_records, uint8_t* d_out);
// Host side of the end of a proof: flags -> the reference's panics, then the proof bytes in record order.
int prove_finish(const ProveLayout& lay, const uint8_t* h_out, const uint8_t* rnd_host, uint8_t* proof, uint64_t cap,
                 uint64_t* written);
void prove_collect_timing(Ctx& cx);
// hscProve for a sparse s(X,Y) given as terms c_t X^(eX_t) Y^(eY_t) (host arrays); result buffer of ProveLayout(M, false)
int hsc_terms_enqueue(Ctx& cx, SrsRep& srs, uint64_t nterms, const int64_t* eX, const int64_t* eY, const uint8_t* coeff32,
                      uint32_t M, const Fr* d_rnd, uint8_t* d_result);
}  // namespace sonic

// ---- the public handles: one replica per device of the sonic_init list -----------------------
struct sonic_srs {
    uint64_t d = 0;
    int ndev = 0;
    sonic::SrsRep* rep[sonic::MAX_DEV] = {};
};
struct sonic_circuit {
    uint64_t n = 0, Q = 0;
    int ndev = 0;
    sonic::CircuitRep* rep[sonic::MAX_DEV] = {};
};
