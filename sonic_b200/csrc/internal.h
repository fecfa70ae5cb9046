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
void msm_run(Ctx& cx, const G1Affine* d_points, const MsmTables& tables, const uint32_t* d_scalars,
             const std::vector<MsmJob>& jobs, G1Affine* d_out_aff, uint8_t* d_out_comp);
void msm_collect_timing(Ctx& cx);
void exclusive_scan_u32(Arena& ar, const uint32_t* in, uint32_t* out, uint32_t n);


// ---- srs.cu ------------------------------------------------------------------------------
// Generates the resident point array.  d_canon: x, alpha canonical (2 Fr) in device memory.
// pre_c > 0 additionally fills levels 1..W-1 with the 2^(pre_c j) multiples (W = ceil(255/pre_c)).
// d_g2_points (nullable): also fills the G2 h-vectors, 2*(2d+1) affine G2 points, same exponent indexing, no hole.
void srs_generate(Ctx& cx, uint64_t d, const Fr* d_canon, G1Affine* d_points, int pre_c, void* d_g2_points = nullptr);
void srs_generate_g2(Ctx& cx, const Fr* scal_m, uint64_t npts, void* d_g2_points);
void g2_compress_range(Ctx& cx, const void* d_g2_points, uint64_t first, uint64_t count, uint8_t* d_out);
size_t g2_point_bytes();

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
void pow_tables(Ctx& cx, const Fr* bases, int ntab, Fr* tab, uint64_t len, uint64_t stride);
// runs the three open passes over a batch of jobs (value + quotient per job)
void open_batch(Ctx& cx, const std::vector<OpenJob>& jobs);
NttPlan ntt_prepare(Ctx& cx, uint32_t logL);
void ntt_forward(const NttPlan& p, Fr* a);
void ntt_inverse(const NttPlan& p, Fr* a);

}  // namespace sonic

struct sonic_circuit;

namespace sonic {
// ---- selftest.cu -------------------------------------------------------------------------
int selftest_field(Ctx& cx, int which, int op, const void* a, const void* b, void* out, uint32_t n);
int pcv_fold(Ctx& cx, uint32_t k, const uint8_t* F48, const uint8_t* W48, const uint8_t* v32, const uint8_t* z32,
             const uint8_t* r32, const uint32_t* group, uint32_t ngroups, uint8_t* out48);
int selftest_g1(Ctx& cx, int op, const void* a, const void* b, void* out_aff, void* out_comp, uint32_t n);
}  // namespace sonic

namespace sonic {
// ---- prove.cu ----------------------------------------------------------------------------
int circuit_load(Ctx& cx, uint64_t n, uint64_t Q, const uint8_t* wL, const uint8_t* wR, const uint8_t* wO,
                 const uint8_t* cs, sonic_circuit** out);
int circuit_load_csr(Ctx& cx, uint64_t n, uint64_t Q, const uint64_t* const row_ptr[3], const uint32_t* const col[3],
                     const uint8_t* const val[3], const uint8_t* cs, sonic_circuit** out);
void circuit_free(sonic_circuit* c);
uint64_t circuit_n(const sonic_circuit* c);
uint64_t circuit_Q(const sonic_circuit* c);
// d_in: canonical aL|aR|aO (3n Fr); d_rnd: canonical draws in the reference's order, 2M+8 of them
// (M = number of (y_j, z_j) pairs; M = Q inside `prove`).
// world == 1: `out` receives the proof.  world > 1: this rank's slice of every MSM; `out` receives a
// shard blob (raw partial sums + field values) for prove_combine.
int prove_run(Ctx& cx, const sonic_srs* srs, const sonic_circuit* circ, const Fr* d_in, const Fr* d_rnd,
              uint32_t M, bool has_main, uint32_t rank, uint32_t world, uint8_t* out, uint64_t cap, uint64_t* written,
              void* d_partials_out = nullptr);
int prove_combine(Ctx& cx, uint32_t M, bool has_main, uint32_t world, const uint8_t* blobs, uint8_t* out,
                  uint64_t cap, uint64_t* written, const void* d_gathered = nullptr);
}  // namespace sonic

// Resident SRS.  Device layout: one array of affine points indexed by exponent,
//   points[family * (2d+1) + (k + d)],  k in [-d, d],  family 0 = plain, 1 = alpha;
// the alpha slot k = 0 holds the infinity marker (0,0) and is never referenced by a job.
struct sonic_srs {
    uint64_t d = 0;
    sonic::G1Affine* points = nullptr;  // levels x 2*(2d+1) affine points, Montgomery form; level 0 is the SRS
    sonic::MsmTables tables;            // precomputed window multiples (c == 0: level 0 only)
    void* g2_points = nullptr;          // optional: 2*(2d+1) affine G2 points (the h-vectors), Montgomery form
    uint64_t stride() const { return 2 * d + 1; }
    uint64_t index(int family, int64_t k) const { return (uint64_t)family * stride() + (uint64_t)(k + (int64_t)d); }
};
