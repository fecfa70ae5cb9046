/* libsonic_b200 -- C ABI of the B200-native Sonic prover hot path.
 *
 * The reference (sdiehl/sonic, pure Haskell) has no FFI of its own; the boundary is
 * defined by the Haskell functions whose bodies this library replaces.  Every entry
 * point names the reference interface it stands in for (paths relative to the
 * reference tree).  INTEGRATION.md shows the `foreign import ccall` side.
 *
 * Conventions
 *   Fr  : 32 bytes, little-endian, canonical residue in [0, r)   (not Montgomery)
 *   G1  : 48 bytes, compressed: big-endian x; bit7 = 1, bit6 = infinity,
 *         bit5 = (y > (q-1)/2); infinity = 0xc0 || 0^47
 *   Ownership: the caller owns all buffers; the library copies in and never keeps a
 *   host pointer.  Handles are library-owned; their contents never change after creation (an SRS
 *   handle may additionally cache window tables derived from its points).
 *   Threading: calls may come from any OS thread; the library runs one call at a time.
 *   Errors: 0 on success, a SONIC_ERR_* code otherwise; nothing is thrown across the
 *   boundary.  sonic_last_error() returns, for the calling thread, the text the
 *   reference would have passed to `panic` where there is one.
 *   There is no CPU fallback: every call fails with SONIC_ERR_NO_DEVICE when no
 *   CUDA device is usable.
 */
#ifndef SONIC_B200_H
#define SONIC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SONIC_FR_BYTES 32
#define SONIC_G1_BYTES 48
#define SONIC_G1_RAW_BYTES 96 /* uncompressed affine: x || y, each 48-byte little-endian canonical; infinity = zeros */

enum {
    SONIC_OK = 0,
    SONIC_ERR_INVALID_ARG = 1,
    SONIC_ERR_SRS_TOO_SHORT = 2,  /* `index` panic, src/Sonic/CommitmentScheme.hs:70-73 */
    SONIC_ERR_D_TOO_SMALL = 3,    /* `prove` guard, src/Sonic/Protocol.hs:54-55 */
    SONIC_ERR_DIV_BY_ZERO = 4,    /* `recip 0`: x = 0 in SRS.new, or z = 0 with negative exponents */
    SONIC_ERR_NONCANONICAL = 5,   /* an Fr encoding >= r */
    SONIC_ERR_CUDA = 6,
    SONIC_ERR_BUFFER_TOO_SMALL = 7,
    SONIC_ERR_NO_DEVICE = 8,
    SONIC_ERR_NOT_INITIALISED = 9
};

/* families of SRS bases, indexed by exponent k in [-d, d] */
enum {
    SONIC_FAMILY_PLAIN = 0, /* g^{x^k}:       gNegativeX / gPositiveX,            src/Sonic/SRS.hs:33-34 */
    SONIC_FAMILY_ALPHA = 1  /* g^{alpha x^k}: gNegativeAlphaX / gPositiveAlphaX,  src/Sonic/SRS.hs:37-39; k = 0 absent */
};

typedef struct sonic_srs sonic_srs;
typedef struct sonic_circuit sonic_circuit;

/* Binds the calling process to `ndev` CUDA devices of one box (1 <= ndev <= 8; `devices` = their CUDA
 * ordinals, NULL = 0..ndev-1).  With ndev > 1 the library itself spreads the work:
 * one context, stream and host worker thread per device, one NCCL communicator per device
 * (ncclCommInitAll); every handle holds a replica per device.  The reference's entry points keep their
 * signatures -- `prove` is still one call in one process (src/Sonic/Protocol.hs:47-53):
 *   sonic_srs_new      each device generates 1/ndev of the exponent range (of every table level),
 *                      one all-gather over NVLink                         (src/Sonic/SRS.hs:27-43)
 *   sonic_prove        equal runs of the proof's MSM terms per device, one NCCL all-gather of the
 *                      ~4 KB exchange records, `<>` per commitment on device 0
 *   sonic_prove_batch  whole proofs dealt round-robin, no exchange
 *   sonic_msm_g1, sonic_commit   contiguous slices of a long window, all-gather of 96-byte partial sums
 * Everything else runs on devices[0].  (One process per GPU is still possible: sonic_prove_shard.) */
int sonic_init(const int* devices, int ndev);
void sonic_shutdown(void);
/* number of devices this process is bound to (0 before sonic_init) */
int sonic_device_count(void);

const char* sonic_strerror(int code);
/* copies the calling thread's last error text (NUL-terminated) into buf; returns its length */
size_t sonic_last_error(char* buf, size_t cap);

/* SRS.new :: Int -> Fr -> Fr -> SRS   (src/Sonic/SRS.hs:27-43), G1 vectors only.
 * Generates all 4d+1 G1 elements on the device by fixed-base batch multiplication and
 * keeps them resident in HBM until sonic_srs_free. */
int sonic_srs_new(uint64_t d, const uint8_t x[32], const uint8_t alpha[32], sonic_srs** out);
void sonic_srs_free(sonic_srs* srs);
uint64_t sonic_srs_d(const sonic_srs* srs); /* srsD, src/Sonic/SRS.hs:12 */

/* SRS persistence (the reference has no on-disk format: SRS.hs:11-22 derives nothing):
 * the resident G1 arrays, full-range window tables included, as one file, so that a large setup is
 * paid once per machine rather than once per process.  The trapdoor is never stored; neither are the
 * optional G2 vectors nor tables restricted to a circuit size (those are rebuilt from the points). */
int sonic_srs_save(const sonic_srs* srs, const char* path);
int sonic_srs_load(const char* path, sonic_srs** out);

/* Element of gNegativeX/gPositiveX/gNegativeAlphaX/gPositiveAlphaX by exponent
 * (record fields, src/Sonic/SRS.hs:13-18).  family ALPHA, exponent 0 -> SONIC_ERR_SRS_TOO_SHORT. */
int sonic_srs_g1(const sonic_srs* srs, int family, int64_t exponent, uint8_t out[48]);
/* `count` consecutive elements starting at `exponent`, 48 bytes each */
int sonic_srs_g1_range(const sonic_srs* srs, int family, int64_t exponent, uint64_t count, uint8_t* out);

/* The G2 half of the record (hNegativeX / hPositiveX / hNegativeAlphaX / hPositiveAlphaX,
 * src/Sonic/SRS.hs:35-36,40-41), generated by SRS.new when option "g2" is
 * set (off by default: only pcV reads G2, four elements of it).  96-byte compressed G2 each
 * (x.c1 || x.c0 big-endian, G1's flag bits); family ALPHA includes exponent 0 (SRS.hs:41). */
int sonic_srs_g2_range(const sonic_srs* srs, int family, int64_t exponent, uint64_t count, uint8_t* out);

/* commitPoly :: SRS -> Int -> VLaurent Fr -> G1   (src/Sonic/CommitmentScheme.hs:20-33)
 * f is given as a dense window: coeffs32[i] is the coefficient of X^(lo+i), i < len.
 * Zero coefficients are what the sparse reference does not hold: they never index the SRS. */
int sonic_commit(const sonic_srs* srs, int64_t max, int64_t lo, uint64_t len,
                 const uint8_t* coeffs32, uint8_t out_g1[48]);

/* openPoly :: SRS -> Fr -> VLaurent Fr -> (Fr, G1)   (src/Sonic/CommitmentScheme.hs:36-48) */
int sonic_open(const sonic_srs* srs, const uint8_t z[32], int64_t lo, uint64_t len,
               const uint8_t* coeffs32, uint8_t out_v[32], uint8_t out_w[48]);

/* The fold inside commitPoly/openPoly as a standalone multi-scalar multiplication
 * (src/Sonic/CommitmentScheme.hs:26-29,45-48): sum_i scalars[i] * base[family][lo+i]. */
int sonic_msm_g1(const sonic_srs* srs, int family, int64_t lo, uint64_t len,
                 const uint8_t* scalars32, uint8_t out[48]);
/* Same, for one slice of a sharded MSM: the partial sum leaves as 96 raw bytes so that
 * ranks can gather the partials (NCCL) and fold them with sonic_g1_sum. */
int sonic_msm_g1_partial(const sonic_srs* srs, int family, int64_t lo, uint64_t len,
                         const uint8_t* scalars32, uint8_t out_raw[96]);
/* `<>` over n raw partial sums -> compressed G1 (src/Sonic/CommitmentScheme.hs:26,45) */
int sonic_g1_sum(const uint8_t* raw96, uint64_t n, uint8_t out[48]);

/* Scalars already resident in device memory (canonical little-endian, 32 bytes each);
 * `d_scalars32` is a CUDA device pointer.  Used when the coefficient vectors are produced
 * on the device, and by bench.py to time the path without host copies. */
int sonic_msm_g1_device(const sonic_srs* srs, int family, int64_t lo, uint64_t len,
                        const void* d_scalars32, uint8_t out[48]);

int sonic_msm_g1_device_partial(const sonic_srs* srs, int family, int64_t lo, uint64_t len,
                                const void* d_scalars32, uint8_t out_raw[96]);

/* ArithCircuit{weights = GateWeights{wL,wR,wO}, cs}  (src/Sonic/Protocol.hs:53; layout of
 * src/Sonic/Constraints.hs:38-53): three dense Q x n row-major matrices of Fr and Q constants,
 * kept resident across proofs.  Limits: n < 2^24, Q < 2^14, n Q < 2^28 (dense). */
int sonic_circuit_load(uint64_t n, uint64_t Q, const uint8_t* wL, const uint8_t* wR,
                       const uint8_t* wO, const uint8_t* cs, sonic_circuit** out);
/* Same circuit from sparse weights: each matrix in CSR -- rowptr (Q+1
 * entries, rowptr[0] = 0), col (gate index, 0-based) and val (32-byte Fr per non-zero); repeated
 * (row, col) pairs add up, as terms of a sparse polynomial do.  Replaces the Q x n dense transfer
 * and the Q*n work of `sPoly`'s list walk (src/Sonic/Constraints.hs:48-49) by nnz. */
int sonic_circuit_load_csr(uint64_t n, uint64_t Q, const uint64_t* rowptr_L, const uint32_t* col_L,
                           const uint8_t* val_L, const uint64_t* rowptr_R, const uint32_t* col_R,
                           const uint8_t* val_R, const uint64_t* rowptr_O, const uint32_t* col_O,
                           const uint8_t* val_O, const uint8_t* cs, sonic_circuit** out);
void sonic_circuit_free(sonic_circuit* c);

/* number of Fr values `prove` draws from MonadRandom: 2Q + 8, in the order
 * c_{n+1..n+4}, y, z, ys[Q], zs[Q], u, v  (src/Sonic/Protocol.hs:58,66,76,84-85; Signature.hs:48,60) */
uint64_t sonic_rnd_count(uint64_t Q);
/* bytes of an encoded proof: (4Q+7) G1 + (2Q+5) Fr */
uint64_t sonic_proof_size(uint64_t Q);

/* prove :: SRS -> Assignment Fr -> ArithCircuit Fr -> m (Proof, RndOracle)
 * (src/Sonic/Protocol.hs:47-109, with hscProve, src/Sonic/Signature.hs:32-72, inside).
 * aL/aR/aO: n Fr each.  rnd: sonic_rnd_count(Q) Fr.  The proof is written in the field
 * order of `Proof` / `HscProof` (Protocol.hs:28-38, Signature.hs:22-29):
 *   prR prT prA prWa prB prWb prWt prS | Q x (S_j s_j W_j) | Q x (s'_j W'_j Q_j) | hscQv hscC hscU hscV */
int sonic_prove(const sonic_srs* srs, const sonic_circuit* circuit, const uint8_t* aL,
                const uint8_t* aR, const uint8_t* aO, const uint8_t* rnd, uint8_t* proof_out,
                uint64_t cap, uint64_t* written);

/* `count` independent proofs of one circuit in one call (BASELINE config 5: 64 proofs at n = 2^14).
 * assignments: count x (aL | aR | aO), 3n Fr each; rnds: count x sonic_rnd_count(Q) Fr; proofs_out:
 * count x sonic_proof_size(Q) bytes.  Proof i is `prove` on (assignment i, draws i): with several
 * devices, proof i runs on device i mod ndev, all devices at the same time, nothing exchanged. */
int sonic_prove_batch(const sonic_srs* srs, const sonic_circuit* circuit, uint64_t count, const uint8_t* assignments,
                      const uint8_t* rnds, uint8_t* proofs_out, uint64_t cap, uint64_t* written);

/* One proof sharded over `world` PROCESSES (one per GPU; the multi-process twin of sonic_init with
 * ndev > 1).  Every rank calls sonic_prove_shard with the same inputs.  The exponent windows of the
 * proof's 4Q+7 MSMs, concatenated in record order, are cut into `world` runs of terms (equal, except
 * that the ranks that also build t(X,y) get up to n/2 terms less (n/2 * min(world-2, 6)/6)): a rank sums a few whole MSMs plus at
 * most two partial ones (the identity for the rest) and builds only the polynomials, power tables and
 * openings those MSMs need; a field value of the proof is computed by the lowest rank that opens the
 * polynomial it belongs to.  The result is an exchange record of sonic_shard_exchange_size(Q) bytes:
 *   (4Q+7) raw partial sums (96 B) | (2Q+3) field values (zeros where another rank leads) |
 *   status words: per MSM the first non-zero coefficient outside the SRS, the encoding flag, srsD
 * and the shard blob = record | hscU | hscV.  The ranks exchange the records (one NCCL all-gather);
 * sonic_prove_combine folds them: `<>` per commitment, the one contribution per field value, the
 * minimum of the violation flags -- so every rank that folds reaches the same verdict, success or the
 * reference's panic text, whichever rank saw the offending coefficient. */
uint64_t sonic_shard_exchange_size(uint64_t Q);
uint64_t sonic_shard_blob_size(uint64_t Q);
int sonic_prove_shard(const sonic_srs* srs, const sonic_circuit* circuit, const uint8_t* aL,
                      const uint8_t* aR, const uint8_t* aO, const uint8_t* rnd, uint32_t rank,
                      uint32_t world, uint8_t* blob_out, uint64_t cap, uint64_t* written);
int sonic_prove_combine(uint64_t Q, uint32_t world, const uint8_t* blobs, uint8_t* proof_out,
                        uint64_t cap, uint64_t* written);

/* Exchange on the device: sonic_prove_shard_sink leaves the exchange record at `d_record_out`, a device
 * buffer of the caller (sonic_shard_exchange_size(Q) bytes) -- typically the input of an NCCL
 * all-gather -- and sonic_prove_combine_device folds the gathered device buffer ([world] records)
 * directly; `own_blob` only supplies hscU, hscV.  blob_out may be NULL when d_record_out is given.
 * `assignment`: aL | aR | aO contiguous (3n Fr), host or device as flagged; `d_rnd_or_null`: the
 * draws in device memory if already there. */
int sonic_prove_shard_sink(const sonic_srs* srs, const sonic_circuit* circuit, const void* assignment,
                           int assignment_on_device, const void* d_rnd_or_null, const uint8_t* rnd_host,
                           uint32_t rank, uint32_t world, uint8_t* blob_out, uint64_t cap,
                           uint64_t* written, void* d_record_out);
int sonic_prove_combine_device(uint64_t Q, uint32_t world, const void* d_gathered, const uint8_t* own_blob,
                               uint8_t* proof_out, uint64_t cap, uint64_t* written);

/* Same proof, with the assignment (aL | aR | aO, 3n Fr, canonical) and the draws already
 * resident in device memory (of devices[0]; the other devices of a multi-GPU runtime pull them over
 * NVLink); `rnd_host` is the host copy of the draws (zero checks, hscU/hscV).
 * Used by bench.py to time the path without the host->device copy of the inputs. */
int sonic_prove_device(const sonic_srs* srs, const sonic_circuit* circuit, const void* d_assignment,
                       const void* d_rnd, const uint8_t* rnd_host, uint8_t* proof_out, uint64_t cap,
                       uint64_t* written);

int sonic_prove_shard_device(const sonic_srs* srs, const sonic_circuit* circuit, const void* d_assignment,
                             const void* d_rnd, const uint8_t* rnd_host, uint32_t rank, uint32_t world,
                             uint8_t* out, uint64_t cap, uint64_t* written);

/* hscProve :: SRS -> BiVLaurent Fr -> [(Fr,Fr)] -> m HscProof   (src/Sonic/Signature.hs:32-72)
 * s(X,Y) is the circuit's; yzs = m pairs (y_j, z_j) interleaved; uv = the two draws u, v.
 * Output: Q x (S_j s_j W_j) | Q x (s'_j W'_j Q_j) | hscQv hscC hscU hscV. */
int sonic_hsc_prove(const sonic_srs* srs, const sonic_circuit* circuit, uint64_t m,
                    const uint8_t* yzs, const uint8_t* uv, uint8_t* out, uint64_t cap,
                    uint64_t* written);

/* The same hscProve for ANY sparse `BiVLaurent Fr`, as the reference's signature allows and its test
 * builds it (`sPoly weights`, test/Test/Signature.hs:30-36): s(X,Y) = sum_t coeff_t X^(eX[t]) Y^(eY[t]),
 * nterms terms in any order (outer variable X, inner Y: src/Sonic/Utils.hs:15); terms with a zero
 * coefficient are ignored and repeated monomials add up, as in the sparse normal form.  |exponent| <= 2^26.
 * A shim walks `GHC.Exts.toList sXY` (and `toList` of each inner polynomial) to produce the three arrays. */
int sonic_hsc_prove_terms(const sonic_srs* srs, uint64_t nterms, const int64_t* eX, const int64_t* eY,
                          const uint8_t* coeff32, uint64_t m, const uint8_t* yzs, const uint8_t* uv,
                          uint8_t* out, uint64_t cap, uint64_t* written);

/* Verifier-side G1 folding (the pairings stay on the host library).
 * pcV (src/Sonic/CommitmentScheme.hs:51-68) for k checks (F_i, z_i, (v_i, W_i)) with caller-chosen
 * random weights r_i reduces to ONE multi-pairing with the G1 inputs written to out48:
 *   out[0] = A = sum r_i W_i                  to pair with hPositiveAlphaX[1]
 *   out[1] = B = sum r_i (v_i g - z_i W_i)    to pair with hPositiveAlphaX[0]
 *   out[2+m] = C_m = sum_{group_i = m} r_i F_i to pair with h^{x^{-d+max_m}}
 * so that verify is:  e(A, h^{alpha x}) e(B, h^alpha) == prod_m e(C_m, hxi_m). */
int sonic_pcv_fold(uint64_t k, const uint8_t* F48, const uint8_t* W48, const uint8_t* v32,
                   const uint8_t* z32, const uint8_t* r32, const uint32_t* group, uint32_t ngroups,
                   uint8_t* out48);

/* ---- tuning and measurement hooks (not part of the reference surface) ---- */
/* option names: "window_bits" (0 = automatic), "chunk" (0 = automatic),
 * "precompute" (-1 = automatic, 0 = off, c = window bits): SRS.new also stores the multiples
 * 2^(c j) * base of every SRS element so that all windows of an MSM share one bucket set; when those
 * full-range tables exceed the budget, the first proof of a circuit size builds tables for the 17n+23
 * exponents that size reads (by doubling the resident points; kept in the SRS handle);
 * "precompute_budget_mb": HBM the automatic mode may spend on those tables (default 8192);
 * "shard_min_terms": a standalone MSM is cut across the devices when it has at least this many terms per device (default 2^17);
 * "g2" (0/1): SRS.new also generates the G2 vectors; "sort_mode" (0: thread per term with global
 * atomics, 1: tiled counting sort with shared-memory histograms [default], 2..64: tiled with that
 * many tiles per SM); "sort_reserve" (1 default: a tile takes its slots of a bucket from the bucket's cursor by an atomic
 * add -- the order of tiles inside a bucket is the order of arrival; 0: a prefix pass over the tile histograms fixes it);
 * "acc_mode": the bucket stage -- 3 (default) automatic: pairwise rounds in affine coordinates with batched inversions
 * when the batch has at least 5*2^22 bucket entries, the XYZZ chunk kernel below; 2 affine always; 1 XYZZ, operands in
 * shared memory; 0 XYZZ, operands in registers.  Every mode returns the same bytes.  "aff_tail" (0..5, default 4): the
 * last halvings of a bucket left to a serial tail; "aff_m" (0 automatic | 8 | 16 | 32 | 64): output slots per thread
 * of a round; "aff_fused" (0 | 1 | 2): one kernel per round with the inversion inside the block (measured slower);
 * "acc_blocks", "chunk_max", "reduce_mode" (0 automatic, 1 level by level, 2 thread per K buckets, 3 quads of lanes),
 * "reduce_k", "heavy_mode", "overlap": kernel tuning knobs */
int sonic_set_option(const char* name, int64_t value);
/* device time in milliseconds of the kernels of the last call, by stage name; returns 0
 * if unknown.  Stages: "msm", "msm.sort", "msm.accumulate", "msm.reduce", "msm.accumulate_kernel", "poly", "total";
 * the same call also reports counters of the last MSM batch: "msm.window_bits", "msm.windows",
 * "msm.terms", "msm.entries", "msm.jobs", "msm.chunk", "msm.buckets", "msm.sort_tiles", "msm.affine" (1 when the
 * bucket stage ran in affine coordinates) */
double sonic_last_timing_ms(const char* stage);
/* the same for device `slot` of the sonic_init list (each device times its own share of a call) */
double sonic_last_timing_ms_dev(int slot, const char* stage);
/* number of kernel launches issued by this library since sonic_init */
uint64_t sonic_launch_count(void);
/* CUDA events on the library's own stream, for harnesses that time several calls as one
 * region (events of another runtime only see that runtime's stream): record into slot 0..5, read the time
 * between two recorded slots (synchronises on the later one). */
int sonic_bench_mark(int slot);
double sonic_bench_elapsed_ms(int from_slot, int to_slot);
/* Register-only integer multiply-add microbenchmark: returns measured 32x32->64
 * multiply-accumulates per second on the bound device (the roofline denominator). */
double sonic_imad_peak_lmacs(int variant, int iters);
/* Device self-tests of the arithmetic layer (raw Montgomery-form limbs, little-endian u32):
 * field 0 = Fq (12 limbs), 1 = Fr (8 limbs); op 0 mul, 1 add, 2 sub, 3 to_mont, 4 from_mont,
 * 5 inv, 6 sqr, 7 neg, 8 inv by binary Euclid (the MSM's final to-affine).  G1: points as XYZZ (48 limbs); op 0 acc+affine(b.x,b.y), 1 acc+b, 2 2*acc,
 * 3 2*affine(a.x,a.y), 4 acc+b and 5 2*acc by a quad of lanes; outputs affine (24 limbs) and the compressed encoding. */
int sonic_selftest_field(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, uint32_t n);
int sonic_selftest_g1(int op, const uint32_t* a_xyzz, const uint32_t* b_xyzz, uint32_t* out_affine,
                      uint8_t* out_comp, uint32_t n);
/* Latency probe: ns per operation of a chain of `iters` dependent operations run by every thread of
 * blocks x threads (one warp per SM = what a lone warp of the MSM's tail stages sees).
 * op 0 Fq multiplication, 1 full addition, 2 doubling, 3 mixed addition, 4 / 5 addition / doubling shared by a
 * quad of lanes (csrc/g1coop.cuh).  sonic_selftest_g1 ops 4 / 5 check those against ops 1 / 2. */
double sonic_selftest_latency_ns(int op, int iters, int blocks, int threads);
/* device memory helpers for harnesses that have no CUDA binding of their own */
int sonic_dev_alloc(uint64_t bytes, void** out);
int sonic_dev_free(void* p);
int sonic_dev_upload(void* dst, const void* src, uint64_t bytes);
int sonic_dev_download(void* dst, const void* src, uint64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* SONIC_B200_H */
