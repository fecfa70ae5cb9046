/* A C consumer of include/sonic_b200.h: no Python, no torch -- what a `foreign import ccall` shim links
 * against.  Compiled with `gcc -std=c11 -Wall -Wextra -Werror` (tests/c/Makefile), so it is also the
 * compile-time check of every prototype in the header.
 *
 *   abi_multi <ndev> [log_n]
 *
 * 1. binds ONE device, generates an SRS, loads a circuit, proves (sonic_prove), proves a batch;
 * 2. shuts down, binds <ndev> devices IN THE SAME PROCESS (sonic_init(devices, ndev)) and repeats:
 *    SRS.new is then sharded over the devices and all-gathered, sonic_prove deals the proof's MSM terms
 *    to the devices and folds the NCCL-gathered records, sonic_prove_batch deals whole proofs;
 * 3. everything must be byte-identical between the two runs: sampled SRS elements of both families,
 *    the proof, the batch, a commitment / opening pair, a standalone MSM, and the panic text of an
 *    unsatisfied assignment.
 * Exit status 0 = identical. */
#include <inttypes.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "sonic_b200.h"

/* ---- every prototype of the header, spelled out once: a drifting declaration fails to compile ---- */
typedef struct {
    int (*init)(const int*, int);
    void (*shutdown)(void);
    int (*device_count)(void);
    const char* (*strerror_)(int);
    size_t (*last_error)(char*, size_t);
    int (*srs_new)(uint64_t, const uint8_t*, const uint8_t*, sonic_srs**);
    void (*srs_free)(sonic_srs*);
    uint64_t (*srs_d)(const sonic_srs*);
    int (*srs_save)(const sonic_srs*, const char*);
    int (*srs_load)(const char*, sonic_srs**);
    int (*srs_g1)(const sonic_srs*, int, int64_t, uint8_t*);
    int (*srs_g1_range)(const sonic_srs*, int, int64_t, uint64_t, uint8_t*);
    int (*srs_g2_range)(const sonic_srs*, int, int64_t, uint64_t, uint8_t*);
    int (*commit)(const sonic_srs*, int64_t, int64_t, uint64_t, const uint8_t*, uint8_t*);
    int (*open)(const sonic_srs*, const uint8_t*, int64_t, uint64_t, const uint8_t*, uint8_t*, uint8_t*);
    int (*msm_g1)(const sonic_srs*, int, int64_t, uint64_t, const uint8_t*, uint8_t*);
    int (*msm_g1_partial)(const sonic_srs*, int, int64_t, uint64_t, const uint8_t*, uint8_t*);
    int (*g1_sum)(const uint8_t*, uint64_t, uint8_t*);
    int (*msm_g1_device)(const sonic_srs*, int, int64_t, uint64_t, const void*, uint8_t*);
    int (*msm_g1_device_partial)(const sonic_srs*, int, int64_t, uint64_t, const void*, uint8_t*);
    int (*circuit_load)(uint64_t, uint64_t, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, sonic_circuit**);
    int (*circuit_load_csr)(uint64_t, uint64_t, const uint64_t*, const uint32_t*, const uint8_t*, const uint64_t*, const uint32_t*,
                            const uint8_t*, const uint64_t*, const uint32_t*, const uint8_t*, const uint8_t*, sonic_circuit**);
    void (*circuit_free)(sonic_circuit*);
    uint64_t (*rnd_count)(uint64_t);
    uint64_t (*proof_size)(uint64_t);
    int (*prove)(const sonic_srs*, const sonic_circuit*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, uint8_t*, uint64_t, uint64_t*);
    int (*prove_batch)(const sonic_srs*, const sonic_circuit*, uint64_t, const uint8_t*, const uint8_t*, uint8_t*, uint64_t, uint64_t*);
    uint64_t (*shard_exchange_size)(uint64_t);
    uint64_t (*shard_blob_size)(uint64_t);
    int (*prove_shard)(const sonic_srs*, const sonic_circuit*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, uint32_t, uint32_t,
                       uint8_t*, uint64_t, uint64_t*);
    int (*prove_combine)(uint64_t, uint32_t, const uint8_t*, uint8_t*, uint64_t, uint64_t*);
    int (*prove_shard_sink)(const sonic_srs*, const sonic_circuit*, const void*, int, const void*, const uint8_t*, uint32_t, uint32_t, uint8_t*,
                            uint64_t, uint64_t*, void*);
    int (*prove_combine_device)(uint64_t, uint32_t, const void*, const uint8_t*, uint8_t*, uint64_t, uint64_t*);
    int (*prove_device)(const sonic_srs*, const sonic_circuit*, const void*, const void*, const uint8_t*, uint8_t*, uint64_t, uint64_t*);
    int (*prove_shard_device)(const sonic_srs*, const sonic_circuit*, const void*, const void*, const uint8_t*, uint32_t, uint32_t, uint8_t*,
                              uint64_t, uint64_t*);
    int (*hsc_prove)(const sonic_srs*, const sonic_circuit*, uint64_t, const uint8_t*, const uint8_t*, uint8_t*, uint64_t, uint64_t*);
    int (*hsc_prove_terms)(const sonic_srs*, uint64_t, const int64_t*, const int64_t*, const uint8_t*, uint64_t, const uint8_t*, const uint8_t*,
                           uint8_t*, uint64_t, uint64_t*);
    int (*pcv_fold)(uint64_t, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint8_t*, const uint32_t*, uint32_t, uint8_t*);
    int (*set_option)(const char*, int64_t);
    double (*last_timing_ms)(const char*);
    double (*last_timing_ms_dev)(int, const char*);
    uint64_t (*launch_count)(void);
    int (*bench_mark)(int);
    double (*bench_elapsed_ms)(int, int);
    double (*imad_peak_lmacs)(int, int);
    int (*selftest_field)(int, int, const uint32_t*, const uint32_t*, uint32_t*, uint32_t);
    int (*selftest_g1)(int, const uint32_t*, const uint32_t*, uint32_t*, uint8_t*, uint32_t);
    double (*selftest_latency_ns)(int, int, int, int);
    int (*dev_alloc)(uint64_t, void**);
    int (*dev_free)(void*);
    int (*dev_upload)(void*, const void*, uint64_t);
    int (*dev_download)(void*, const void*, uint64_t);
} sonic_abi;

static const sonic_abi ABI = {
    sonic_init, sonic_shutdown, sonic_device_count, sonic_strerror, sonic_last_error, sonic_srs_new, sonic_srs_free, sonic_srs_d,
    sonic_srs_save, sonic_srs_load, sonic_srs_g1, sonic_srs_g1_range, sonic_srs_g2_range, sonic_commit, sonic_open, sonic_msm_g1,
    sonic_msm_g1_partial, sonic_g1_sum, sonic_msm_g1_device, sonic_msm_g1_device_partial, sonic_circuit_load, sonic_circuit_load_csr,
    sonic_circuit_free, sonic_rnd_count, sonic_proof_size, sonic_prove, sonic_prove_batch, sonic_shard_exchange_size,
    sonic_shard_blob_size, sonic_prove_shard, sonic_prove_combine, sonic_prove_shard_sink, sonic_prove_combine_device,
    sonic_prove_device, sonic_prove_shard_device, sonic_hsc_prove, sonic_hsc_prove_terms, sonic_pcv_fold, sonic_set_option,
    sonic_last_timing_ms, sonic_last_timing_ms_dev, sonic_launch_count, sonic_bench_mark, sonic_bench_elapsed_ms,
    sonic_imad_peak_lmacs, sonic_selftest_field, sonic_selftest_g1, sonic_selftest_latency_ns, sonic_dev_alloc, sonic_dev_free, sonic_dev_upload, sonic_dev_download,
};

/* ---- deterministic inputs -------------------------------------------------------------------------- */
static uint64_t sm_state;
static uint64_t splitmix64(void) {
    uint64_t z = (sm_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* a canonical Fr: four words, the top one masked to 62 bits (r > 2^254) */
static void fr_random(uint8_t out[32]) {
    uint64_t w[4] = {splitmix64(), splitmix64(), splitmix64(), splitmix64() & ((1ull << 62) - 1)};
    w[0] |= 1; /* never zero: the draws are evaluation points */
    memcpy(out, w, 32);
}
static void fr_from_u128(uint8_t out[32], unsigned __int128 v) {
    memset(out, 0, 32);
    uint64_t lo = (uint64_t)v, hi = (uint64_t)(v >> 64);
    memcpy(out, &lo, 8);
    memcpy(out + 8, &hi, 8);
}

#define CHECK(call)                                                                         \
    do {                                                                                    \
        int rc_ = (call);                                                                   \
        if (rc_ != SONIC_OK) {                                                              \
            char msg_[512];                                                                 \
            sonic_last_error(msg_, sizeof msg_);                                            \
            fprintf(stderr, "%s:%d: %s -> %d (%s): %s\n", __FILE__, __LINE__, #call, rc_, sonic_strerror(rc_), msg_); \
            exit(2);                                                                        \
        }                                                                                   \
    } while (0)

static double now_ms(void) {
    struct timespec t;
    clock_gettime(CLOCK_MONOTONIC, &t);
    return t.tv_sec * 1e3 + t.tv_nsec * 1e-6;
}

typedef struct {
    uint64_t n, Q, d, batch;
    uint8_t *wL, *wR, *wO, *cs, *assign /* batch x (aL|aR|aO) */, *rnd /* batch x (2Q+8) */, *x, *alpha;
} workload;

typedef struct {
    uint8_t* srs_sample;   /* 4 ranges x 64 elements x 48 B */
    uint8_t* proof;        /* one proof */
    uint8_t* batch;        /* `batch` proofs */
    uint8_t commit[48], open_v[32], open_w[48], msm[48];
    char panic_text[512];
    int panic_code;
    double srs_ms, prove_ms, batch_ms;
} results;

static void make_workload(workload* w, unsigned log_n) {
    w->n = 1ull << log_n;
    w->Q = 3;
    w->d = 7 * w->n + 5;
    w->batch = 8;
    const uint64_t n = w->n, Q = w->Q;
    w->wL = calloc(Q * n, 32);
    w->wR = calloc(Q * n, 32);
    w->wO = calloc(Q * n, 32);
    w->cs = calloc(Q, 32);
    w->assign = calloc(w->batch * 3 * n, 32);
    w->rnd = calloc(w->batch * (2 * Q + 8), 32);
    w->x = calloc(1, 32);
    w->alpha = calloc(1, 32);
    sm_state = 0x534F4E4943ull;
    fr_random(w->x);
    fr_random(w->alpha);
    /* one all-ones row per matrix (test/Test/Reference.hs:141-149): row 0 of wL, 1 of wR, 2 of wO */
    for (uint64_t i = 0; i < n; ++i) {
        w->wL[32 * (0 * n + i)] = 1;
        w->wR[32 * (1 * n + i)] = 1;
        w->wO[32 * (2 * n + i)] = 1;
    }
    /* assignment 0 fixes cs; the other assignments of the batch are permutations-in-value of it: the
     * same multiset of (aL, aR) pairs in rotated order keeps every sum, hence satisfies the same cs */
    unsigned __int128 sL = 0, sR = 0, sO = 0;
    uint32_t* a = malloc(2 * n * sizeof(uint32_t));
    for (uint64_t i = 0; i < n; ++i) {
        a[2 * i] = (uint32_t)splitmix64();
        a[2 * i + 1] = (uint32_t)splitmix64();
        sL += a[2 * i];
        sR += a[2 * i + 1];
        sO += (unsigned __int128)a[2 * i] * a[2 * i + 1];
    }
    fr_from_u128(w->cs + 0, sL);
    fr_from_u128(w->cs + 32, sR);
    fr_from_u128(w->cs + 64, sO);
    for (uint64_t b = 0; b < w->batch; ++b) {
        uint8_t* base = w->assign + b * 3 * n * 32;
        for (uint64_t i = 0; i < n; ++i) {
            const uint64_t k = (i + 17 * b) % n;
            fr_from_u128(base + 32 * i, a[2 * k]);
            fr_from_u128(base + 32 * (n + i), a[2 * k + 1]);
            fr_from_u128(base + 32 * (2 * n + i), (unsigned __int128)a[2 * k] * a[2 * k + 1]);
        }
        for (uint64_t k = 0; k < 2 * Q + 8; ++k) fr_random(w->rnd + (b * (2 * Q + 8) + k) * 32);
    }
    free(a);
}

static void run(const workload* w, int ndev, results* r) {
    int devices[8];
    for (int i = 0; i < ndev; ++i) devices[i] = i;
    CHECK(sonic_init(devices, ndev));
    if (sonic_device_count() != ndev) { fprintf(stderr, "sonic_device_count() = %d, expected %d\n", sonic_device_count(), ndev); exit(2); }
    const uint64_t n = w->n, Q = w->Q, d = w->d;
    const uint64_t psize = sonic_proof_size(Q);
    sonic_srs* srs = NULL;
    double t0 = now_ms();
    CHECK(sonic_srs_new(d, w->x, w->alpha, &srs));
    r->srs_ms = now_ms() - t0;
    /* SRS elements from both ends and the middle of both families (the hole g^alpha excluded) */
    r->srs_sample = calloc(4 * 64, 48);
    CHECK(sonic_srs_g1_range(srs, SONIC_FAMILY_PLAIN, -(int64_t)d, 64, r->srs_sample));
    CHECK(sonic_srs_g1_range(srs, SONIC_FAMILY_PLAIN, (int64_t)d - 63, 64, r->srs_sample + 64 * 48));
    CHECK(sonic_srs_g1_range(srs, SONIC_FAMILY_ALPHA, -64, 64, r->srs_sample + 128 * 48));
    CHECK(sonic_srs_g1_range(srs, SONIC_FAMILY_ALPHA, 1, 64, r->srs_sample + 192 * 48));
    sonic_circuit* circ = NULL;
    CHECK(sonic_circuit_load(n, Q, w->wL, w->wR, w->wO, w->cs, &circ));
    r->proof = calloc(1, psize);
    uint64_t written = 0;
    const uint8_t* a0 = w->assign;
    CHECK(sonic_prove(srs, circ, a0, a0 + 32 * n, a0 + 64 * n, w->rnd, r->proof, psize, &written));   /* warm-up (arena, tables) */
    t0 = now_ms();
    CHECK(sonic_prove(srs, circ, a0, a0 + 32 * n, a0 + 64 * n, w->rnd, r->proof, psize, &written));
    r->prove_ms = now_ms() - t0;
    if (written != psize) { fprintf(stderr, "written %" PRIu64 " != %" PRIu64 "\n", written, psize); exit(2); }
    r->batch = calloc(w->batch, psize);
    CHECK(sonic_prove_batch(srs, circ, w->batch, w->assign, w->rnd, r->batch, w->batch * psize, &written));
    t0 = now_ms();
    CHECK(sonic_prove_batch(srs, circ, w->batch, w->assign, w->rnd, r->batch, w->batch * psize, &written));
    r->batch_ms = now_ms() - t0;
    if (memcmp(r->batch, r->proof, psize) != 0) { fprintf(stderr, "ndev=%d: batch proof 0 differs from sonic_prove\n", ndev); exit(1); }
    /* commitPoly / openPoly / MSM on aL as a polynomial over X^1..X^n */
    CHECK(sonic_commit(srs, (int64_t)d, 1, n, a0, r->commit));
    CHECK(sonic_open(srs, w->rnd + 32 * 5, 1, n, a0, r->open_v, r->open_w));
    CHECK(sonic_set_option("shard_min_terms", 256));   /* let the small MSM below be cut across the devices too */
    CHECK(sonic_msm_g1(srs, SONIC_FAMILY_PLAIN, -(int64_t)n, 3 * n, w->assign, r->msm));
    CHECK(sonic_set_option("shard_min_terms", 1 << 17));
    /* an assignment that does not satisfy the circuit: the X^0 term of t(X,y) indexes g^alpha (CommitmentScheme.hs:70-73) */
    uint8_t* bad = malloc(3 * n * 32);
    memcpy(bad, a0, 3 * n * 32);
    bad[64 * n] ^= 1;
    uint8_t* scratch = malloc(psize);
    r->panic_code = sonic_prove(srs, circ, bad, bad + 32 * n, bad + 64 * n, w->rnd, scratch, psize, &written);
    sonic_last_error(r->panic_text, sizeof r->panic_text);
    free(bad);
    free(scratch);
    sonic_circuit_free(circ);
    sonic_srs_free(srs);
    sonic_shutdown();
    if (sonic_device_count() != 0) { fprintf(stderr, "device count after shutdown\n"); exit(2); }
}

static uint64_t fnv(const uint8_t* p, size_t n) {
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 1099511628211ull;
    return h;
}

int main(int argc, char** argv) {
    (void)ABI;
    const int ndev = argc > 1 ? atoi(argv[1]) : 2;
    const unsigned log_n = argc > 2 ? (unsigned)atoi(argv[2]) : 12;
    if (ndev < 1 || ndev > 8 || log_n < 1 || log_n > 18) { fprintf(stderr, "usage: abi_multi <ndev 1..8> [log_n 1..18]\n"); return 2; }
    workload w;
    make_workload(&w, log_n);
    const uint64_t psize = sonic_proof_size(w.Q);
    results one, many;
    memset(&one, 0, sizeof one);
    memset(&many, 0, sizeof many);
    run(&w, 1, &one);
    run(&w, ndev, &many);
    int bad = 0;
#define SAME(field, bytes, what)                                                                          \
    if (memcmp(one.field, many.field, (bytes)) != 0) { fprintf(stderr, "MISMATCH ndev=%d vs 1: %s\n", ndev, what); bad = 1; }
    SAME(srs_sample, 4 * 64 * 48, "SRS elements (sharded SRS.new + all-gather)");
    SAME(proof, psize, "proof bytes (sonic_prove over all devices)");
    SAME(batch, w.batch * psize, "batch proofs");
    SAME(commit, 48, "commitPoly");
    SAME(open_v, 32, "openPoly value");
    SAME(open_w, 48, "openPoly witness");
    SAME(msm, 48, "standalone MSM (sliced across devices)");
    if (one.panic_code != SONIC_ERR_SRS_TOO_SHORT || many.panic_code != one.panic_code || strcmp(one.panic_text, many.panic_text) != 0) {
        fprintf(stderr, "MISMATCH panic: 1 device -> %d '%s', %d devices -> %d '%s'\n", one.panic_code, one.panic_text, ndev, many.panic_code, many.panic_text);
        bad = 1;
    }
    printf("{\"ndev\": %d, \"n\": %" PRIu64 ", \"Q\": %" PRIu64 ", \"d\": %" PRIu64 ", \"proof_fnv64\": \"%016" PRIx64 "\", \"identical\": %s, "
           "\"panic\": \"%s\", \"ms\": {\"srs_new\": [%.2f, %.2f], \"prove\": [%.3f, %.3f], \"batch_of_%" PRIu64 "\": [%.3f, %.3f]}}\n",
           ndev, w.n, w.Q, w.d, fnv(one.proof, psize), bad ? "false" : "true", one.panic_text, one.srs_ms, many.srs_ms, one.prove_ms, many.prove_ms,
           w.batch, one.batch_ms, many.batch_ms);
    return bad;
}
