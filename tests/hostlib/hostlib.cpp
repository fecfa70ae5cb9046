// Host build of the device arithmetic headers (portable carry-chain emulation) so that
// the exact limb algorithms can be checked against the big-int oracle without a GPU.
// TEST INFRASTRUCTURE ONLY: the product library never links this file.
#include <cstring>
#include "../../sonic_b200/csrc/g1.cuh"
#include "../../sonic_b200/csrc/scalar.cuh"
#include "../../sonic_b200/csrc/g2.cuh"

using namespace sonic;

template <class F> static F ld(const uint32_t* p) { F r; memcpy(r.l, p, sizeof(r.l)); return r; }
template <class F> static void st(uint32_t* p, const F& v) { memcpy(p, v.l, sizeof(v.l)); }

extern "C" {
// op: 0 mul, 1 add, 2 sub, 3 to_mont(a), 4 from_mont(a), 5 inv(a), 6 sqr(a), 7 neg(a), 8 inv(a) by binary Euclid
void ht_fq_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fq x = ld<Fq>(a), y = ld<Fq>(b), r;
    switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_to_mont(x); break;
        case 4: r = fp_from_mont(x); break;
        case 5: r = fp_inv(x); break;
        case 6: r = fp_sqr(x); break;
        case 8: r = fp_inv_euclid(x); break;
        default: r = fp_neg(x); break;
    }
    st(out, r);
}
// op: 0 unrolled CIOS, 1 rolled CIOS, 2 Karatsuba + wide reduction, 3 a*b + c*d (one reduction), 4 a*b - c*d
void ht_fq_mulvar(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out) {
    Fq x = ld<Fq>(a), y = ld<Fq>(b), z = ld<Fq>(c), w = ld<Fq>(d), r;
    switch (op) {
        case 0: r = fp_mul_unrolled(x, y); break;
        case 1: r = fp_mul_rolled<FqParams, 2>(x, y); break;
        case 2: r = fp_mul_karatsuba(x, y); break;
        case 3: r = fp_mul_add2(x, y, z, w); break;
        case 5: r = fp_mul_rolled<FqParams, 4>(x, y); break;
        case 6: r = fp_mul_rolled<FqParams, 6>(x, y); break;
        default: r = fp_mul_sub2(x, y, z, w); break;
    }
    st(out, r);
}
void ht_fq_mul2(const uint32_t* a1, const uint32_t* b1, const uint32_t* a2, const uint32_t* b2, uint32_t* r1, uint32_t* r2) {
    Fq x, y;
    fp_mul2(ld<Fq>(a1), ld<Fq>(b1), ld<Fq>(a2), ld<Fq>(b2), x, y);
    st(r1, x);
    st(r2, y);
}
void ht_fr_mulvar(int op, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out) {
    Fr x = ld<Fr>(a), y = ld<Fr>(b), z = ld<Fr>(c), w = ld<Fr>(d), r;
    switch (op) {
        case 0: r = fp_mul_unrolled(x, y); break;
        case 1: r = fp_mul_rolled(x, y); break;
        case 2: r = fp_mul_karatsuba(x, y); break;
        case 3: r = fp_mul_add2(x, y, z, w); break;
        default: r = fp_mul_sub2(x, y, z, w); break;
    }
    st(out, r);
}
// raw wide product of two K-limb numbers (K = 4 or 6) and of two 2K-limb numbers by Karatsuba
void ht_wide_sqr(int N, const uint32_t* a, uint32_t* out) {
    if (N == 12) wide_sqr<12>(out, a);
    else if (N == 8) wide_sqr<8>(out, a);
    else if (N == 6) wide_sqr<6>(out, a);
    else wide_sqr<4>(out, a);
}
void ht_wide_mul(int K, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    if (K == 6) wide_mul<6>(out, a, b);
    else if (K == 4) wide_mul<4>(out, a, b);
    else if (K == 12) wide_mul_karatsuba<12>(out, a, b);
    else if (K == 112) wide_mul<12>(out, a, b);
    else if (K == 108) wide_mul<8>(out, a, b);
    else wide_mul_karatsuba<8>(out, a, b);
}
void ht_fr_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    Fr x = ld<Fr>(a), y = ld<Fr>(b), r;
    switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_to_mont(x); break;
        case 4: r = fp_from_mont(x); break;
        case 5: r = fp_inv(x); break;
        case 6: r = fp_sqr(x); break;
        case 8: r = fp_inv_euclid(x); break;
        default: r = fp_neg(x); break;
    }
    st(out, r);
}
static G1XYZZ ldx(const uint32_t* p) { G1XYZZ r; r.x = ld<Fq>(p); r.y = ld<Fq>(p + 12); r.zz = ld<Fq>(p + 24); r.zzz = ld<Fq>(p + 36); return r; }
static void stx(uint32_t* p, const G1XYZZ& r) { st(p, r.x); st(p + 12, r.y); st(p + 24, r.zz); st(p + 36, r.zzz); }
static G1Affine lda(const uint32_t* p) { G1Affine r; r.x = ld<Fq>(p); r.y = ld<Fq>(p + 12); return r; }
static void sta(uint32_t* p, const G1Affine& r) { st(p, r.x); st(p + 12, r.y); }

// all points in Montgomery form. op: 0 madd(acc, aff b), 1 add(acc, xyzz b), 2 dbl(acc), 3 mdbl(aff a -> xyzz)
void ht_g1_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
    G1XYZZ r;
    switch (op) {
        case 0: r = ldx(a); g1_madd(r, lda(b)); break;
        case 1: r = ldx(a); g1_add(r, ldx(b)); break;
        case 2: r = g1_dbl(ldx(a)); break;
        default: r = g1_mdbl(lda(a)); break;
    }
    stx(out, r);
}
void ht_g1_to_affine(const uint32_t* a, uint32_t* out) { sta(out, g1_to_affine(ldx(a))); }
void ht_g1_compress(const uint32_t* aff, uint8_t* out48) { g1_compress(lda(aff), out48); }
int ht_g1_decompress(const uint8_t* in48, uint32_t* out_aff) {
    G1Affine a;
    if (!g1_decompress(in48, a)) return 0;
    sta(out_aff, a);
    return 1;
}
void ht_g1_mul_scalar(const uint32_t* aff, const uint32_t* k8, uint32_t* out_xyzz) {
    stx(out_xyzz, g1_mul_scalar(lda(aff), ld<Fr>(k8)));
}
void ht_g1_gen(uint32_t* out) { sta(out, G1Affine::gen()); }

// k * G2 generator by double-and-add with the device formulas -> 96-byte compressed
void ht_g2_mul_gen(const uint32_t* k8, uint8_t* out96) {
    Fr k = ld<Fr>(k8);
    G2XYZZ r = G2XYZZ::inf();
    const G2Affine g = G2Affine::gen();
    bool started = false;
    for (int i = 7; i >= 0; --i)
        for (int b = 31; b >= 0; --b) {
            if (started) r = g2_dbl(r);
            if ((k.l[i] >> b) & 1) { g2_madd(r, g); started = true; }
        }
    g2_compress(g2_to_affine(r), out96);
}

// signed-digit recoding of a canonical scalar: returns W digits
int ht_recode(const uint32_t* scalar8, int c, int32_t* digits) {
    ScalarDigits sd(scalar8, c);
    int W = msm_num_windows(c);
    for (int j = 0; j < W; ++j) digits[j] = sd.next();
    return W;
}
}
