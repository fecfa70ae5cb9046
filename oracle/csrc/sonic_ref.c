/* CPU restatement of the Sonic prover hot path in plain C (gcc, unsigned __int128).
 *
 * TEST INFRASTRUCTURE ONLY.  Linked by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs; never by the product library.
 * PARITY UNPINNED: the reference (sdiehl/sonic) is Haskell, cannot be built in this
 * environment (no GHC) and holds no golden vectors; this file follows its ALGORITHM
 * (citations below, paths relative to the reference tree) and is checked against the
 * Python big-int oracle and the fixtures in tests/golden/.
 *
 *   SRS.new            src/Sonic/SRS.hs:27-43             per element: pow, (* alpha), mul gen
 *   commitPoly         src/Sonic/CommitmentScheme.hs:20-33 one double-and-add `mul` + one `<>` per term
 *   openPoly           src/Sonic/CommitmentScheme.hs:36-48 eval, divide by (X - z), same fold
 *   prove / hscProve   src/Sonic/Protocol.hs:47-109, src/Sonic/Signature.hs:32-72
 *
 * The arithmetic packages (galois-field, elliptic-curve, poly) are not on disk; what is
 * restated is their published mathematics.  Where the reference is asymptotically
 * infeasible (bivariate sparse product for t(X,Y), Constraints.hs:61; list indexing in
 * sPoly, :48-49) the univariate equivalent is used so that n = 2^12 finishes; the MSM
 * loop, which dominates, is the reference's own: a full scalar multiplication per term.
 *
 * Encodings at this file's boundary are those of include/sonic_b200.h.
 */
#define _POSIX_C_SOURCE 200809L
#include <pthread.h>
#include <time.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t u64;

/* ---------------------------------------------------------------- prime fields ---- */
#define NQ 6
#define NR 4
static const u64 QM[NQ] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                           0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};
static const u64 RM[NR] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
static const u64 Q_INV = 0x89f3fffcfffcfffdull; /* -q^-1 mod 2^64 */
static const u64 R_INV = 0xfffffffeffffffffull; /* -r^-1 mod 2^64 */
static const u64 Q_R2[NQ] = {0xf4df1f341c341746ull, 0x0a76e6a609d104f1ull, 0x8de5476c4c95b6d5ull,
                             0x67eb88a9939d83c0ull, 0x9a793e85b519952dull, 0x11988fe592cae3aaull};
static const u64 R_R2[NR] = {0xc999e990f3f29c6dull, 0x2b6cedcb87925c23ull, 0x05d314967254398full, 0x0748d9d99f59ff11ull};
static const u64 Q_ONE[NQ] = {0x760900000002fffdull, 0xebf4000bc40c0002ull, 0x5f48985753c758baull,
                              0x77ce585370525745ull, 0x5c071a97a256ec6dull, 0x15f65ec3fa80e493ull};
static const u64 R_ONE[NR] = {0x00000001fffffffeull, 0x5884b7fa00034802ull, 0x998c4fefecbc4ff5ull, 0x1824b159acc5056full};
/* generator of G1, Montgomery form */
static const u64 GX[NQ] = {0x5cb38790fd530c16ull, 0x7817fc679976fff5ull, 0x154f95c7143ba1c1ull,
                           0xf0ae6acdf3d0e747ull, 0xedce6ecc21dbf440ull, 0x120177419e0bfb75ull};
static const u64 GY[NQ] = {0xbaac93d50ce72271ull, 0x8c22631a7918fd8eull, 0xdd595f13570725ceull,
                           0x51ac582950405194ull, 0x0e1c8c3fad0059c0ull, 0x0bbc3efc5008a26aull};

static int ge_n(const u64* a, const u64* m, int n) {
    for (int i = n - 1; i >= 0; --i) {
        if (a[i] > m[i]) return 1;
        if (a[i] < m[i]) return 0;
    }
    return 1;
}
static void sub_n(u64* r, const u64* a, const u64* b, int n) {
    u64 borrow = 0;
    for (int i = 0; i < n; ++i) {
        u128 t = (u128)a[i] - b[i] - borrow;
        r[i] = (u64)t;
        borrow = (u64)(t >> 64) & 1;
    }
}
static u64 add_n(u64* r, const u64* a, const u64* b, int n) {
    u64 carry = 0;
    for (int i = 0; i < n; ++i) {
        u128 t = (u128)a[i] + b[i] + carry;
        r[i] = (u64)t;
        carry = (u64)(t >> 64);
    }
    return carry;
}
/* Montgomery product (CIOS), n limbs */
static inline __attribute__((always_inline)) void mont_mul(u64* r, const u64* a, const u64* b, const u64* m, u64 inv, const int n) {
    u64 t[NQ + 2];
    memset(t, 0, sizeof t);
    for (int i = 0; i < n; ++i) {
        u64 c = 0;
        for (int j = 0; j < n; ++j) {
            u128 p = (u128)a[j] * b[i] + t[j] + c;
            t[j] = (u64)p;
            c = (u64)(p >> 64);
        }
        u128 s = (u128)t[n] + c;
        t[n] = (u64)s;
        t[n + 1] = (u64)(s >> 64);
        u64 mm = t[0] * inv;
        u128 p = (u128)mm * m[0] + t[0];
        c = (u64)(p >> 64);
        for (int j = 1; j < n; ++j) {
            p = (u128)mm * m[j] + t[j] + c;
            t[j - 1] = (u64)p;
            c = (u64)(p >> 64);
        }
        s = (u128)t[n] + c;
        t[n - 1] = (u64)s;
        t[n] = t[n + 1] + (u64)(s >> 64);
    }
    if (t[n] || ge_n(t, m, n)) sub_n(r, t, m, n);
    else memcpy(r, t, 8 * n);
}
static void mod_add(u64* r, const u64* a, const u64* b, const u64* m, int n) {
    u64 t[NQ];
    u64 c = add_n(t, a, b, n);
    if (c || ge_n(t, m, n)) sub_n(r, t, m, n);
    else memcpy(r, t, 8 * n);
}
static void mod_sub(u64* r, const u64* a, const u64* b, const u64* m, int n) {
    u64 t[NQ];
    if (ge_n(a, b, n)) { sub_n(r, a, b, n); return; }
    sub_n(t, b, a, n);
    sub_n(r, m, t, n);
}
static int is_zero_n(const u64* a, int n) {
    u64 t = 0;
    for (int i = 0; i < n; ++i) t |= a[i];
    return t == 0;
}

typedef struct { u64 l[NQ]; } fq;
typedef struct { u64 l[NR]; } fr;
static fq fq_mul(fq a, fq b) { fq r; mont_mul(r.l, a.l, b.l, QM, Q_INV, NQ); return r; }
static fq fq_add(fq a, fq b) { fq r; mod_add(r.l, a.l, b.l, QM, NQ); return r; }
static fq fq_sub(fq a, fq b) { fq r; mod_sub(r.l, a.l, b.l, QM, NQ); return r; }
static fr fr_mul(fr a, fr b) { fr r; mont_mul(r.l, a.l, b.l, RM, R_INV, NR); return r; }
static fr fr_add(fr a, fr b) { fr r; mod_add(r.l, a.l, b.l, RM, NR); return r; }
static fr fr_sub(fr a, fr b) { fr r; mod_sub(r.l, a.l, b.l, RM, NR); return r; }
static fq fq_one(void) { fq r; memcpy(r.l, Q_ONE, sizeof r.l); return r; }
static fr fr_one(void) { fr r; memcpy(r.l, R_ONE, sizeof r.l); return r; }
static fr fr_zero(void) { fr r; memset(r.l, 0, sizeof r.l); return r; }
static fq fq_zero(void) { fq r; memset(r.l, 0, sizeof r.l); return r; }
static fq fq_to_mont(fq a) { fq r2; memcpy(r2.l, Q_R2, sizeof r2.l); return fq_mul(a, r2); }
static fq fq_from_mont(fq a) { fq o = fq_zero(); o.l[0] = 1; return fq_mul(a, o); }
static fr fr_to_mont(fr a) { fr r2; memcpy(r2.l, R_R2, sizeof r2.l); return fr_mul(a, r2); }
static fr fr_from_mont(fr a) { fr o = fr_zero(); o.l[0] = 1; return fr_mul(a, o); }

/* `pow` of Data.Field.Galois by square-and-multiply, as SRS.hs:33-41 calls it per element */
static fr fr_pow_u64(fr a, u64 e) {
    fr r = fr_one();
    for (int b = 63; b >= 0; --b) {
        r = fr_mul(r, r);
        if ((e >> b) & 1) r = fr_mul(r, a);
    }
    return r;
}
static fr fr_pow_limbs(fr a, const u64* e, int n) {
    fr r = fr_one();
    for (int i = n - 1; i >= 0; --i)
        for (int b = 63; b >= 0; --b) {
            r = fr_mul(r, r);
            if ((e[i] >> b) & 1) r = fr_mul(r, a);
        }
    return r;
}
static fr fr_inv(fr a) { /* recip = a^(r-2) */
    u64 e[NR];
    memcpy(e, RM, sizeof e);
    e[0] -= 2;
    return fr_pow_limbs(a, e, NR);
}
static fq fq_inv(fq a) {
    u64 e[NQ];
    memcpy(e, QM, sizeof e);
    e[0] -= 2;
    fq r = fq_one();
    for (int i = NQ - 1; i >= 0; --i)
        for (int b = 63; b >= 0; --b) {
            r = fq_mul(r, r);
            if ((e[i] >> b) & 1) r = fq_mul(r, a);
        }
    return r;
}

/* ---------------------------------------------------------------- G1 (Jacobian) ---- */
typedef struct { fq x, y, z; } g1; /* infinity <=> z == 0 */
static g1 g1_inf(void) { g1 p; p.x = fq_one(); p.y = fq_one(); p.z = fq_zero(); return p; }
static int g1_is_inf(const g1* p) { return is_zero_n(p->z.l, NQ); }
static g1 g1_gen(void) { g1 p; memcpy(p.x.l, GX, sizeof GX); memcpy(p.y.l, GY, sizeof GY); p.z = fq_one(); return p; }

static g1 g1_dbl(g1 p) {
    if (g1_is_inf(&p)) return p;
    fq A = fq_mul(p.x, p.x), B = fq_mul(p.y, p.y), C = fq_mul(B, B);
    fq xb = fq_add(p.x, B);
    fq D = fq_sub(fq_sub(fq_mul(xb, xb), A), C);
    D = fq_add(D, D);
    fq E = fq_add(fq_add(A, A), A), F = fq_mul(E, E);
    g1 r;
    r.x = fq_sub(F, fq_add(D, D));
    fq c8 = fq_add(C, C); c8 = fq_add(c8, c8); c8 = fq_add(c8, c8);
    r.y = fq_sub(fq_mul(E, fq_sub(D, r.x)), c8);
    fq yz = fq_mul(p.y, p.z);
    r.z = fq_add(yz, yz);
    return r;
}
/* `<>` */
static g1 g1_add(g1 p, g1 q) {
    if (g1_is_inf(&p)) return q;
    if (g1_is_inf(&q)) return p;
    fq z1z1 = fq_mul(p.z, p.z), z2z2 = fq_mul(q.z, q.z);
    fq u1 = fq_mul(p.x, z2z2), u2 = fq_mul(q.x, z1z1);
    fq s1 = fq_mul(fq_mul(p.y, q.z), z2z2), s2 = fq_mul(fq_mul(q.y, p.z), z1z1);
    fq h = fq_sub(u2, u1), rr = fq_sub(s2, s1);
    if (is_zero_n(h.l, NQ)) {
        if (is_zero_n(rr.l, NQ)) return g1_dbl(p);
        return g1_inf();
    }
    fq hh = fq_mul(h, h), hhh = fq_mul(h, hh), v = fq_mul(u1, hh);
    g1 r;
    r.x = fq_sub(fq_sub(fq_mul(rr, rr), hhh), fq_add(v, v));
    r.y = fq_sub(fq_mul(rr, fq_sub(v, r.x)), fq_mul(s1, hhh));
    r.z = fq_mul(fq_mul(p.z, q.z), h);
    return r;
}
/* `mul`: left-to-right double-and-add on the canonical scalar (one per MSM term in the reference) */
static g1 g1_mul(g1 p, const u64 k[NR]) {
    g1 acc = g1_inf();
    int started = 0;
    for (int i = NR - 1; i >= 0; --i)
        for (int b = 63; b >= 0; --b) {
            if (started) acc = g1_dbl(acc);
            if ((k[i] >> b) & 1) { acc = g1_add(acc, p); started = 1; }
        }
    return acc;
}
static void g1_affine(const g1* p, fq* x, fq* y, int* inf) {
    if (g1_is_inf(p)) { *inf = 1; *x = fq_zero(); *y = fq_zero(); return; }
    *inf = 0;
    fq zi = fq_inv(p->z), zi2 = fq_mul(zi, zi);
    *x = fq_mul(p->x, zi2);
    *y = fq_mul(p->y, fq_mul(zi2, zi));
}
static void g1_compress(const g1* p, uint8_t out[48]) {
    fq x, y;
    int inf;
    g1_affine(p, &x, &y, &inf);
    if (inf) { memset(out, 0, 48); out[0] = 0xC0; return; }
    fq xc = fq_from_mont(x), yc = fq_from_mont(y);
    for (int i = 0; i < NQ; ++i)
        for (int b = 0; b < 8; ++b) out[8 * i + b] = (uint8_t)(xc.l[NQ - 1 - i] >> (56 - 8 * b));
    out[0] |= 0x80;
    /* y > (q-1)/2  <=>  2y > q - 1  <=>  2y >= q  (q odd) */
    u64 t[NQ];
    u64 c = add_n(t, yc.l, yc.l, NQ);
    if (c || ge_n(t, QM, NQ)) out[0] |= 0x20;
}
/* raw 96-byte affine form (canonical little-endian x || y, infinity = zeros) */
static void g1_to_raw(const g1* p, uint8_t out[96]) {
    fq x, y;
    int inf;
    g1_affine(p, &x, &y, &inf);
    if (inf) { memset(out, 0, 96); return; }
    fq xc = fq_from_mont(x), yc = fq_from_mont(y);
    memcpy(out, xc.l, 48);
    memcpy(out + 48, yc.l, 48);
}
static g1 g1_from_raw(const uint8_t in[96]) {
    fq x, y;
    memcpy(x.l, in, 48);
    memcpy(y.l, in + 48, 48);
    if (is_zero_n(x.l, NQ) && is_zero_n(y.l, NQ)) return g1_inf();
    g1 p;
    p.x = fq_to_mont(x);
    p.y = fq_to_mont(y);
    p.z = fq_one();
    return p;
}

/* ---------------------------------------------------------------- threads ---- */
typedef void (*range_fn)(void* ctx, uint64_t lo, uint64_t hi, int tid);
typedef struct { range_fn fn; void* ctx; uint64_t lo, hi; int tid; } task;
static void* task_main(void* a) { task* t = (task*)a; t->fn(t->ctx, t->lo, t->hi, t->tid); return NULL; }
static void parallel_for(uint64_t n, int threads, range_fn fn, void* ctx) {
    if (threads < 1) threads = 1;
    if ((uint64_t)threads > n) threads = n ? (int)n : 1;
    if (threads == 1) { fn(ctx, 0, n, 0); return; }
    pthread_t* th = malloc(sizeof(pthread_t) * threads);
    task* ts = malloc(sizeof(task) * threads);
    for (int i = 0; i < threads; ++i) {
        ts[i].fn = fn; ts[i].ctx = ctx; ts[i].tid = i;
        ts[i].lo = n * i / threads; ts[i].hi = n * (i + 1) / threads;
        pthread_create(&th[i], NULL, task_main, &ts[i]);
    }
    for (int i = 0; i < threads; ++i) pthread_join(th[i], NULL);
    free(th); free(ts);
}

/* ---------------------------------------------------------------- MSM, reference algorithm ---- */
typedef struct { const uint8_t* pts; const uint8_t* scal; g1* partial; } msm_ctx;
static void msm_range(void* c, uint64_t lo, uint64_t hi, int tid) {
    msm_ctx* m = (msm_ctx*)c;
    g1 acc = g1_inf();
    for (uint64_t i = lo; i < hi; ++i) {
        u64 k[NR];
        memcpy(k, m->scal + 32 * i, 32);
        if (is_zero_n(k, NR)) continue; /* the sparse form holds no zero coefficient */
        g1 p = g1_from_raw(m->pts + 96 * i);
        acc = g1_add(acc, g1_mul(p, k)); /* acc <> (base `mul` v), CommitmentScheme.hs:26-29 */
    }
    m->partial[tid] = acc;
}
/* points: n raw affine (96 B), scalars: n canonical Fr (32 B).  out48 and/or out_raw96. */
void ref_msm_naive(const uint8_t* points_raw, const uint8_t* scalars, uint64_t n, int threads,
                   uint8_t* out48, uint8_t* out_raw96) {
    if (threads < 1) threads = 1;
    g1* partial = malloc(sizeof(g1) * threads);
    for (int i = 0; i < threads; ++i) partial[i] = g1_inf();
    msm_ctx c = {points_raw, scalars, partial};
    parallel_for(n, threads, msm_range, &c);
    g1 acc = g1_inf();
    for (int i = 0; i < threads; ++i) acc = g1_add(acc, partial[i]);
    free(partial);
    if (out48) g1_compress(&acc, out48);
    if (out_raw96) g1_to_raw(&acc, out_raw96);
}

/* ---------------------------------------------------------------- MSM, bucket method (context only) ---- */
/* NOT the reference's algorithm: a plain windowed Pippenger on the CPU so that the GPU speed-up
 * is not quoted against the naive fold alone ("B-pip").  One window per
 * task; unsigned c-bit digits; threads split the windows. */
typedef struct { const uint8_t* pts; const uint8_t* scal; uint64_t n; int c; int nwin; g1* win; } pip_ctx;
static void pip_range(void* cv, uint64_t lo, uint64_t hi, int tid) {
    (void)tid;
    pip_ctx* p = (pip_ctx*)cv;
    const uint64_t nb = ((uint64_t)1 << p->c) - 1;
    g1* buckets = malloc(sizeof(g1) * nb);
    for (uint64_t w = lo; w < hi; ++w) {
        for (uint64_t b = 0; b < nb; ++b) buckets[b] = g1_inf();
        const int bit = (int)w * p->c;
        for (uint64_t i = 0; i < p->n; ++i) {
            u64 k[NR + 1];
            memcpy(k, p->scal + 32 * i, 32);
            k[NR] = 0;
            const int limb = bit >> 6, sh = bit & 63;
            u64 dgt = k[limb] >> sh;
            if (sh && limb + 1 <= NR) dgt |= k[limb + 1] << (64 - sh);
            dgt &= nb;
            if (!dgt) continue;
            g1 pt = g1_from_raw(p->pts + 96 * i);
            buckets[dgt - 1] = g1_add(buckets[dgt - 1], pt);
        }
        g1 run = g1_inf(), acc = g1_inf();
        for (uint64_t b = nb; b-- > 0;) {
            run = g1_add(run, buckets[b]);
            acc = g1_add(acc, run);
        }
        p->win[w] = acc;
    }
    free(buckets);
}
void ref_msm_pippenger(const uint8_t* points_raw, const uint8_t* scalars, uint64_t n, int c, int threads, uint8_t* out48) {
    pip_ctx p;
    p.pts = points_raw; p.scal = scalars; p.n = n; p.c = c;
    p.nwin = (255 + c - 1) / c;
    p.win = malloc(sizeof(g1) * p.nwin);
    parallel_for((uint64_t)p.nwin, threads, pip_range, &p);
    g1 acc = g1_inf();
    for (int w = p.nwin - 1; w >= 0; --w) {
        for (int i = 0; i < c; ++i) acc = g1_dbl(acc);
        acc = g1_add(acc, p.win[w]);
    }
    free(p.win);
    g1_compress(&acc, out48);
}

/* ---------------------------------------------------------------- SRS.new ---- */
typedef struct { uint64_t d; fr x, xinv, alpha; uint8_t* out; } srs_ctx;
/* element index e in [0, 2*(2d+1)): family = e / (2d+1), exponent k = e % (2d+1) - d */
static void srs_range(void* c, uint64_t lo, uint64_t hi, int tid) {
    (void)tid;
    srs_ctx* s = (srs_ctx*)c;
    const uint64_t stride = 2 * s->d + 1;
    g1 G = g1_gen();
    for (uint64_t e = lo; e < hi; ++e) {
        const int family = (int)(e / stride);
        const int64_t k = (int64_t)(e % stride) - (int64_t)s->d;
        uint8_t* o = s->out + 96 * e;
        if (family == 1 && k == 0) { memset(o, 0, 96); continue; } /* g^alpha is not shared, SRS.hs:38 */
        fr v = k >= 0 ? fr_pow_u64(s->x, (u64)k) : fr_pow_u64(s->xinv, (u64)(-k)); /* pow x i / pow xInv i */
        if (family == 1) v = fr_mul(v, s->alpha);                                    /* (*) alpha */
        fr vc = fr_from_mont(v);
        g1 p = g1_mul(G, vc.l);                                                      /* mul gen */
        g1_to_raw(&p, o);
    }
}
/* out: 2*(2d+1) raw points, [family][k + d]; the alpha slot k = 0 is zeros.  returns 0, or 4 if x = 0 */
int ref_srs_new(uint64_t d, const uint8_t x[32], const uint8_t alpha[32], int threads, uint8_t* out) {
    srs_ctx s;
    s.d = d;
    fr xc, ac;
    memcpy(xc.l, x, 32);
    memcpy(ac.l, alpha, 32);
    if (is_zero_n(xc.l, NR)) return 4;
    s.x = fr_to_mont(xc);
    s.alpha = fr_to_mont(ac);
    s.xinv = fr_inv(s.x);
    s.out = out;
    parallel_for(2 * (2 * d + 1), threads, srs_range, &s);
    return 0;
}

/* ---------------------------------------------------------------- polynomials (dense windows) ---- */
static fr fr_load(const uint8_t* b) { fr v; memcpy(v.l, b, 32); return fr_to_mont(v); }
static void fr_store(uint8_t* b, fr v) { fr c = fr_from_mont(v); memcpy(b, c.l, 32); }
static fr fr_pow_i64(fr x, fr xinv, int64_t e) { return e >= 0 ? fr_pow_u64(x, (u64)e) : fr_pow_u64(xinv, (u64)(-e)); }

/* commitPoly on a dense window f[lo..lo+len) against the raw SRS table of ref_srs_new.
 * returns 0 ok, 2 = `index` panic (err_e = the offending shifted exponent) */
int ref_commit(const uint8_t* srs_raw, uint64_t d, int64_t maxm, int64_t lo, uint64_t len,
               const uint8_t* coeffs, int threads, uint8_t out48[48], int64_t* err_e) {
    const int64_t shift = (int64_t)d - maxm;
    const uint64_t stride = 2 * d + 1;
    /* ascending fold: the first offending term panics */
    for (uint64_t i = 0; i < len; ++i) {
        u64 k[NR];
        memcpy(k, coeffs + 32 * i, 32);
        if (is_zero_n(k, NR)) continue;
        int64_t e = lo + (int64_t)i + shift;
        if (e == 0 || e < -(int64_t)d || e > (int64_t)d) { if (err_e) *err_e = e; return 2; }
    }
    int64_t e0 = lo + shift;
    /* clip to the table (only zero coefficients are clipped) */
    uint64_t skip = 0;
    if (e0 < -(int64_t)d) skip = (uint64_t)(-(int64_t)d - e0);
    if (skip > len) skip = len;
    uint64_t n = len - skip;
    int64_t first = e0 + (int64_t)skip;
    if (first + (int64_t)n - 1 > (int64_t)d) n = (uint64_t)((int64_t)d - first + 1);
    if ((int64_t)n < 0) n = 0;
    ref_msm_naive(srs_raw + 96 * (stride + (uint64_t)(first + (int64_t)d)), coeffs + 32 * skip, n, threads, out48, NULL);
    return 0;
}

/* openPoly on a dense window containing X^0.  returns 0 ok, 2 = `index` panic, 4 = recip 0 */
int ref_open(const uint8_t* srs_raw, uint64_t d, const uint8_t z32[32], int64_t lo, uint64_t len,
             const uint8_t* coeffs, int threads, uint8_t out_v[32], uint8_t out_w[48], int64_t* err_e) {
    fr zc;
    memcpy(zc.l, z32, 32);
    const int z0 = is_zero_n(zc.l, NR);
    fr z = fr_to_mont(zc);
    /* the window must contain exponent 0 (callers pad with zeros) */
    if (!(lo <= 0 && 0 < lo + (int64_t)len)) return 1;
    fr* g = malloc(sizeof(fr) * len);
    int any_neg = 0;
    for (uint64_t i = 0; i < len; ++i) {
        g[i] = fr_load(coeffs + 32 * i);
        if ((int64_t)i + lo < 0 && !is_zero_n(g[i].l, NR)) any_neg = 1;
    }
    if (z0 && any_neg) { free(g); return 4; }
    /* fz = eval f z  (Horner on X^lo * g(X)) */
    fr acc = fr_zero();
    for (uint64_t i = len; i-- > 0;) acc = fr_add(fr_mul(acc, z), g[i]);
    fr fz = acc;
    if (lo < 0 && !z0) fz = fr_mul(acc, fr_pow_u64(fr_inv(z), (u64)(-lo)));
    if (z0) fz = g[-lo];
    fr_store(out_v, fz);
    g[-lo] = fr_sub(g[-lo], fz);
    /* quotient by (X - z): synthetic division from the top */
    uint8_t* q = malloc(32 * (len ? len : 1));
    fr carry = fr_zero();
    for (uint64_t k = len - 1; k >= 1; --k) {
        carry = fr_add(g[k], fr_mul(z, carry));
        fr_store(q + 32 * (k - 1), carry);
    }
    free(g);
    const uint64_t qn = len - 1;
    const uint64_t stride = 2 * d + 1;
    (void)stride;
    for (uint64_t i = 0; i < qn; ++i) {
        u64 k[NR];
        memcpy(k, q + 32 * i, 32);
        if (is_zero_n(k, NR)) continue;
        int64_t e = lo + (int64_t)i;
        if (e < -(int64_t)d || e > (int64_t)d) { if (err_e) *err_e = e; free(q); return 2; }
    }
    uint64_t skip = 0;
    if (lo < -(int64_t)d) skip = (uint64_t)(-(int64_t)d - lo);
    if (skip > qn) skip = qn;
    uint64_t n = qn - skip;
    int64_t first = lo + (int64_t)skip;
    if (n && first + (int64_t)n - 1 > (int64_t)d) n = (uint64_t)((int64_t)d - first + 1);
    ref_msm_naive(srs_raw + 96 * (uint64_t)(first + (int64_t)d), q + 32 * skip, n, threads, out_w, NULL);
    free(q);
    return 0;
}

/* ---------------------------------------------------------------- prove ---- */
typedef struct { const fr* a; uint64_t na; const fr* b; uint64_t nb; fr* out; } conv_ctx;
static void conv_range(void* c, uint64_t lo, uint64_t hi, int tid) {
    (void)tid;
    conv_ctx* v = (conv_ctx*)c;
    for (uint64_t k = lo; k < hi; ++k) {
        fr acc = fr_zero();
        uint64_t i0 = k >= v->nb ? k - v->nb + 1 : 0;
        uint64_t i1 = k < v->na ? k : v->na - 1;
        for (uint64_t i = i0; i <= i1; ++i) acc = fr_add(acc, fr_mul(v->a[i], v->b[k - i]));
        v->out[k] = acc;
    }
}

static void dump(uint8_t* dst, const fr* v, uint64_t n) { for (uint64_t i = 0; i < n; ++i) fr_store(dst + 32 * i, v[i]); }

/* s(X,y): out over X^{-n..2n}  (Constraints.hs:34-53 evaluated at Y = y, Utils.hs:20-21) */
static void build_sxy(uint64_t n, uint64_t Q, const fr* wL, const fr* wR, const fr* wO, fr y, fr yinv, fr* out) {
    fr* yp = malloc(sizeof(fr) * (Q + 1));
    yp[0] = fr_pow_u64(y, n);
    for (uint64_t q = 1; q <= Q; ++q) yp[q] = fr_mul(yp[q - 1], y);
    fr yi = fr_one(), yni = fr_one();
    out[n] = fr_zero();
    for (uint64_t i = 1; i <= n; ++i) {
        yi = fr_mul(yi, y);
        yni = fr_mul(yni, yinv);
        fr su = fr_zero(), sv = fr_zero(), sw = fr_zero();
        for (uint64_t q = 1; q <= Q; ++q) {
            su = fr_add(su, fr_mul(yp[q], wL[(q - 1) * n + i - 1]));
            sv = fr_add(sv, fr_mul(yp[q], wR[(q - 1) * n + i - 1]));
            sw = fr_add(sw, fr_mul(yp[q], wO[(q - 1) * n + i - 1]));
        }
        out[n - i] = su;
        out[n + i] = sv;
        out[2 * n + i] = fr_sub(fr_sub(sw, yi), yni);
    }
    free(yp);
}

/* Whole prover; same contract as sonic_prove (include/sonic_b200.h).  srs_raw from ref_srs_new.
 * returns 0 ok, 2 index panic (err_e), 3 d < 7n, 4 recip 0 */
int ref_prove(const uint8_t* srs_raw, uint64_t d, uint64_t n, uint64_t Q, const uint8_t* wL8, const uint8_t* wR8,
              const uint8_t* wO8, const uint8_t* cs8, const uint8_t* aL8, const uint8_t* aR8, const uint8_t* aO8,
              const uint8_t* rnd8, int threads, uint8_t* out, int64_t* err_e) {
    if (d < 7 * n) return 3; /* Protocol.hs:54-55 */
    const uint64_t nr = 2 * Q + 8;
    fr* rnd = malloc(sizeof(fr) * nr);
    for (uint64_t i = 0; i < nr; ++i) rnd[i] = fr_load(rnd8 + 32 * i);
    for (uint64_t i = 4; i < nr; ++i) if (is_zero_n(rnd[i].l, NR)) { free(rnd); return 4; }
    fr* wL = malloc(sizeof(fr) * Q * n), *wR = malloc(sizeof(fr) * Q * n), *wO = malloc(sizeof(fr) * Q * n);
    for (uint64_t i = 0; i < Q * n; ++i) { wL[i] = fr_load(wL8 + 32 * i); wR[i] = fr_load(wR8 + 32 * i); wO[i] = fr_load(wO8 + 32 * i); }
    int rc = 0;
    uint8_t* o = out;
    const fr y = rnd[4], z = rnd[5];
    const fr yinv = fr_inv(y);
    /* r'(X,1) over X^{-2n-4..n}: Constraints.hs:23-31, Protocol.hs:58-62 */
    const uint64_t rl = 3 * n + 5;
    const int64_t rlo = -2 * (int64_t)n - 4;
    fr* rx1 = calloc(rl, sizeof(fr));
    for (uint64_t i = 1; i <= n; ++i) {
        rx1[(int64_t)i - rlo] = fr_load(aL8 + 32 * (i - 1));
        rx1[-(int64_t)i - rlo] = fr_load(aR8 + 32 * (i - 1));
        rx1[-(int64_t)i - (int64_t)n - rlo] = fr_load(aO8 + 32 * (i - 1));
    }
    for (uint64_t i = 1; i <= 4; ++i) rx1[-2 * (int64_t)n - (int64_t)i - rlo] = rnd[i - 1];
    uint8_t* rx1b = malloc(32 * rl);
    dump(rx1b, rx1, rl);
    /* s(X,y), r'(X,y) + s(X,y), t(X,y) */
    const uint64_t sl = 3 * n + 1;
    fr* sxy = malloc(sizeof(fr) * sl);
    build_sxy(n, Q, wL, wR, wO, y, yinv, sxy);
    const uint64_t rsl = 4 * n + 5;
    fr* rs = calloc(rsl, sizeof(fr));
    for (uint64_t k = 0; k < rl; ++k) rs[k] = fr_mul(rx1[k], fr_pow_i64(y, yinv, rlo + (int64_t)k)); /* r(X,y) = r(Xy,1) */
    for (uint64_t k = 0; k < sl; ++k) { uint64_t idx = (uint64_t)(-(int64_t)n + (int64_t)k - rlo); rs[idx] = fr_add(rs[idx], sxy[k]); }
    const uint64_t tl = 7 * n + 9;
    const int64_t tlo = -4 * (int64_t)n - 8;
    fr* t = malloc(sizeof(fr) * tl);
    conv_ctx cc = {rx1, rl, rs, rsl, t};
    parallel_for(tl, threads, conv_range, &cc);
    fr ky = fr_zero();
    for (uint64_t q = 0; q < Q; ++q) ky = fr_add(ky, fr_mul(fr_load(cs8 + 32 * q), fr_pow_u64(y, n + 1 + q))); /* kPoly, :67-68 */
    t[-tlo] = fr_sub(t[-tlo], ky);
    uint8_t* tb = malloc(32 * tl);
    dump(tb, t, tl);
    uint8_t* sb = malloc(32 * sl);
    uint8_t zb[32], yzb[32], vb[32];
    fr_store(zb, z);
    fr_store(yzb, fr_mul(y, z));
#define TRY(expr) do { rc = (expr); if (rc) goto done; } while (0)
    TRY(ref_commit(srs_raw, d, (int64_t)n, rlo, rl, rx1b, threads, o, err_e)); o += 48;            /* prR  Protocol.hs:63 */
    TRY(ref_commit(srs_raw, d, (int64_t)d, tlo, tl, tb, threads, o, err_e)); o += 48;              /* prT  :73 */
    TRY(ref_open(srs_raw, d, zb, rlo, rl, rx1b, threads, o, o + 32, err_e)); o += 80;              /* prA, prWa  :79 */
    TRY(ref_open(srs_raw, d, yzb, rlo, rl, rx1b, threads, o, o + 32, err_e)); o += 80;             /* prB, prWb  :80 */
    TRY(ref_open(srs_raw, d, zb, tlo, tl, tb, threads, vb, o, err_e)); o += 48;                    /* prWt  :81 */
    {   /* prS = s(z, y)  :83 */
        fr acc = fr_zero();
        for (uint64_t i = sl; i-- > 0;) acc = fr_add(fr_mul(acc, z), sxy[i]);
        acc = fr_mul(acc, fr_pow_u64(fr_inv(z), n));
        fr_store(o, acc); o += 32;
    }
    {
        const fr* ys = rnd + 6, *zs = rnd + 6 + Q;
        const fr u = rnd[6 + 2 * Q], v = rnd[7 + 2 * Q];
        const fr uinv = fr_inv(u);
        /* s(u,Y) over Y^{-n..n+Q}: Utils.hs:17-18 on Constraints.hs:34-53 */
        const uint64_t ul = 2 * n + Q + 1;
        fr* suy = calloc(ul, sizeof(fr));
        for (uint64_t i = 1; i <= n; ++i) {
            fr m = fr_sub(fr_zero(), fr_pow_u64(u, i + n));
            suy[n + i] = m;
            suy[n - i] = m;
        }
        for (uint64_t q = 1; q <= Q; ++q) {
            fr acc = fr_zero();
            for (uint64_t i = 1; i <= n; ++i) {
                acc = fr_add(acc, fr_mul(fr_pow_u64(uinv, i), wL[(q - 1) * n + i - 1]));
                acc = fr_add(acc, fr_mul(fr_pow_u64(u, i), wR[(q - 1) * n + i - 1]));
                acc = fr_add(acc, fr_mul(fr_pow_u64(u, i + n), wO[(q - 1) * n + i - 1]));
            }
            suy[2 * n + q] = acc;
        }
        uint8_t* ub = malloc(32 * ul);
        dump(ub, suy, ul);
        uint8_t* sj_all = malloc(32 * sl * (Q ? Q : 1));
        uint8_t pt[32];
        for (uint64_t j = 0; j < Q; ++j) {                                                          /* hscS  Signature.hs:40-45 */
            build_sxy(n, Q, wL, wR, wO, ys[j], fr_inv(ys[j]), sxy);
            uint8_t* sjb = sj_all + 32 * sl * j;
            dump(sjb, sxy, sl);
            TRY(ref_commit(srs_raw, d, (int64_t)d, -(int64_t)n, sl, sjb, threads, o, err_e)); o += 48;
            fr_store(pt, zs[j]);
            TRY(ref_open(srs_raw, d, pt, -(int64_t)n, sl, sjb, threads, o, o + 32, err_e)); o += 80;
        }
        uint8_t ubuf[32];
        fr_store(ubuf, u);
        for (uint64_t j = 0; j < Q; ++j) {                                                          /* hscW  Signature.hs:53-57 */
            uint8_t tmpv[32];
            TRY(ref_open(srs_raw, d, ubuf, -(int64_t)n, sl, sj_all + 32 * sl * j, threads, tmpv, o + 32, err_e));
            fr_store(pt, ys[j]);
            TRY(ref_open(srs_raw, d, pt, -(int64_t)n, ul, ub, threads, o, o + 80, err_e));
            o += 128;
        }
        fr_store(pt, v);
        TRY(ref_open(srs_raw, d, pt, -(int64_t)n, ul, ub, threads, vb, o, err_e)); o += 48;           /* hscQv  :63 */
        TRY(ref_commit(srs_raw, d, (int64_t)d, -(int64_t)n, ul, ub, threads, o, err_e)); o += 48;     /* hscC  :52 */
        fr_store(o, u); o += 32;
        fr_store(o, v); o += 32;
        free(suy); free(ub); free(sj_all);
    }
done:
    (void)sb;
    free(sb); free(tb); free(t); free(rs); free(sxy); free(rx1b); free(rx1); free(wL); free(wR); free(wO); free(rnd);
    return rc;
}

/* single-thread primitive timing for the roofline discussion: ns per Fq multiplication */
double ref_time_fq_mul(uint64_t iters) {
    fq a = fq_one(), b;
    memcpy(b.l, GX, sizeof GX);
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (uint64_t i = 0; i < iters; ++i) a = fq_mul(a, b);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    volatile u64 sink = a.l[0];
    (void)sink;
    return ((t1.tv_sec - t0.tv_sec) * 1e9 + (t1.tv_nsec - t0.tv_nsec)) / (double)iters;
}
