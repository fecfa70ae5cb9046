// BLS12-381 G1 group law in extended Jacobian ("XYZZ") coordinates.
//
// Replaces (by value) the elliptic-curve package's `<>`, `mul`, `mempty` on
// `G1 BLS12381` used at src/Sonic/CommitmentScheme.hs:26-29,45-48 and
// src/Sonic/SRS.hs:33-39 of the reference.  Results are only ever compared /
// exported in canonical affine form, so the internal representation is free.
//
// XYZZ: x = X/ZZ, y = Y/ZZZ with ZZ^3 = ZZZ^2; infinity <=> ZZ == 0.
// Affine table entries use (0,0) for infinity ((0,0) is not on y^2 = x^3 + 4).
// Formulas: EFD madd-2008-s (8M+2S), add-2008-s (12M+2S), dbl-2008-s-1, mdbl-2008-s-1.
#pragma once
#include "field.cuh"

// 1: the mixed addition issues its independent products in interleaved pairs (fp_mul2)
#ifndef SONIC_MADD_PAIRED
#define SONIC_MADD_PAIRED 0
#endif

namespace sonic {

struct G1Affine {
    Fq x, y;
    SONIC_HD bool is_inf() const {
        uint32_t t = 0;
        for (int i = 0; i < Fq::N; ++i) t |= x.l[i] | y.l[i];
        return t == 0;
    }
    static SONIC_HD G1Affine inf() { G1Affine p; p.x = Fq::zero(); p.y = Fq::zero(); return p; }
    static SONIC_HD G1Affine gen() {
        G1Affine p;
        for (int i = 0; i < Fq::N; ++i) { p.x.l[i] = FqParams::GX_M(i); p.y.l[i] = FqParams::GY_M(i); }
        return p;
    }
};

struct G1XYZZ {
    Fq x, y, zz, zzz;
    SONIC_HD bool is_inf() const { return zz.is_zero(); }
    static SONIC_HD G1XYZZ inf() {
        G1XYZZ p; p.x = Fq::one(); p.y = Fq::one(); p.zz = Fq::zero(); p.zzz = Fq::zero(); return p;
    }
    static SONIC_HD G1XYZZ from_affine(const G1Affine& a) {
        if (a.is_inf()) return inf();
        G1XYZZ p; p.x = a.x; p.y = a.y; p.zz = Fq::one(); p.zzz = Fq::one(); return p;
    }
};

SONIC_HD G1Affine g1_neg(const G1Affine& a) {
    G1Affine r; r.x = a.x; r.y = a.y.is_zero() ? a.y : fp_neg(a.y); return r;
}

// 2*(affine point)
SONIC_HD G1XYZZ g1_mdbl(const G1Affine& a) {
    if (a.is_inf()) return G1XYZZ::inf();
    Fq U = fp_dbl(a.y);
    Fq V = fp_sqr(U);
    Fq W = fp_mul(U, V);
    Fq S = fp_mul(a.x, V);
    Fq X2 = fp_sqr(a.x);
    Fq M = fp_add(fp_dbl(X2), X2);
    G1XYZZ r;
    r.x = fp_sub(fp_sqr(M), fp_dbl(S));
    r.y = fp_mul_sub2(M, fp_sub(S, r.x), W, a.y);
    r.zz = V;
    r.zzz = W;
    return r;
}

SONIC_HD G1XYZZ g1_dbl(const G1XYZZ& p) {
    // infinity stays infinity: ZZ3 = V*ZZ1 = 0
    Fq U = fp_dbl(p.y);
    Fq V = fp_sqr(U);
    Fq W = fp_mul(U, V);
    Fq S = fp_mul(p.x, V);
    Fq X2 = fp_sqr(p.x);
    Fq M = fp_add(fp_dbl(X2), X2);
    G1XYZZ r;
    r.x = fp_sub(fp_sqr(M), fp_dbl(S));
    r.y = fp_mul_sub2(M, fp_sub(S, r.x), W, p.y);
    r.zz = fp_mul(V, p.zz);
    r.zzz = fp_mul(W, p.zzz);
    return r;
}

// acc += a   (mixed addition; all exceptional cases handled, reference bench uses x = 1
// so P + P and P + (-P) really occur: bench/Main.hs:23)
SONIC_HD void g1_madd(G1XYZZ& acc, const G1Affine& a) {
    if (a.is_inf()) return;
    if (acc.is_inf()) { acc = G1XYZZ::from_affine(a); return; }
#if SONIC_MADD_PAIRED
    // independent products issued in pairs with interleaved rows (fp_mul2)
    Fq U2, S2;
    fp_mul2(a.x, acc.zz, a.y, acc.zzz, U2, S2);
    Fq P = fp_sub(U2, acc.x);
    Fq R = fp_sub(S2, acc.y);
    if (P.is_zero()) {
        if (R.is_zero()) acc = g1_mdbl(a);
        else acc = G1XYZZ::inf();
        return;
    }
    Fq PP, RR;
    fp_mul2(P, P, R, R, PP, RR);
    Fq PPP, Q;
    fp_mul2(P, PP, acc.x, PP, PPP, Q);
    Fq X3 = fp_sub(fp_sub(RR, PPP), fp_dbl(Q));
    Fq Y3 = fp_mul_sub2(R, fp_sub(Q, X3), acc.y, PPP);
    acc.x = X3;
    acc.y = Y3;
    Fq zz3, zzz3;
    fp_mul2(acc.zz, PP, acc.zzz, PPP, zz3, zzz3);
    acc.zz = zz3;
    acc.zzz = zzz3;
#else
    Fq U2 = fp_mul(a.x, acc.zz);
    Fq S2 = fp_mul(a.y, acc.zzz);
    Fq P = fp_sub(U2, acc.x);
    Fq R = fp_sub(S2, acc.y);
    if (P.is_zero()) {
        if (R.is_zero()) acc = g1_mdbl(a);
        else acc = G1XYZZ::inf();
        return;
    }
    Fq PP = fp_sqr(P);
    Fq PPP = fp_mul(P, PP);
    Fq Q = fp_mul(acc.x, PP);
    Fq X3 = fp_sub(fp_sub(fp_sqr(R), PPP), fp_dbl(Q));
    Fq Y3 = fp_mul_sub2(R, fp_sub(Q, X3), acc.y, PPP);  // R*(Q-X3) - Y1*PPP, one reduction
    acc.x = X3;
    acc.y = Y3;
    acc.zz = fp_mul(acc.zz, PP);
    acc.zzz = fp_mul(acc.zzz, PPP);
#endif
}

// acc += b   (full addition)
SONIC_HD void g1_add(G1XYZZ& acc, const G1XYZZ& b) {
    if (b.is_inf()) return;
    if (acc.is_inf()) { acc = b; return; }
    Fq U1 = fp_mul(acc.x, b.zz);
    Fq U2 = fp_mul(b.x, acc.zz);
    Fq S1 = fp_mul(acc.y, b.zzz);
    Fq S2 = fp_mul(b.y, acc.zzz);
    Fq P = fp_sub(U2, U1);
    Fq R = fp_sub(S2, S1);
    if (P.is_zero()) {
        if (R.is_zero()) acc = g1_dbl(acc);
        else acc = G1XYZZ::inf();
        return;
    }
    Fq PP = fp_sqr(P);
    Fq PPP = fp_mul(P, PP);
    Fq Q = fp_mul(U1, PP);
    Fq X3 = fp_sub(fp_sub(fp_sqr(R), PPP), fp_dbl(Q));
    Fq Y3 = fp_mul_sub2(R, fp_sub(Q, X3), S1, PPP);  // one reduction for the two products
    acc.x = X3;
    acc.y = Y3;
    acc.zz = fp_mul(fp_mul(acc.zz, b.zz), PP);
    acc.zzz = fp_mul(fp_mul(acc.zzz, b.zzz), PPP);
}

SONIC_HD G1XYZZ g1_neg_xyzz(const G1XYZZ& p) {
    G1XYZZ r = p; r.y = p.y.is_zero() ? p.y : fp_neg(p.y); return r;
}

// canonical affine form (one field inversion)
SONIC_HD G1Affine g1_to_affine(const G1XYZZ& p) {
    if (p.is_inf()) return G1Affine::inf();
    Fq zi = fp_inv(p.zzz);                 // 1/ZZZ
    Fq t = fp_mul(p.zz, zi);               // ZZ/ZZZ = 1/Z
    Fq zzi = fp_sqr(t);                    // 1/ZZ
    G1Affine r;
    r.x = fp_mul(p.x, zzi);
    r.y = fp_mul(p.y, zi);
    return r;
}

// the same for one point on one thread (the end of an MSM): binary-Euclid inverse
SONIC_HD G1Affine g1_to_affine_single(const G1XYZZ& p) {
    if (p.is_inf()) return G1Affine::inf();
    Fq zi = fp_inv_euclid(p.zzz);
    Fq t = fp_mul(p.zz, zi);
    Fq zzi = fp_sqr(t);
    G1Affine r;
    r.x = fp_mul(p.x, zzi);
    r.y = fp_mul(p.y, zi);
    return r;
}

// k * p for a small unsigned multiplier (double-and-add, MSB first)
SONIC_HD G1XYZZ g1_mul_small(const G1XYZZ& p, uint32_t k) {
    G1XYZZ r = G1XYZZ::inf();
    for (int b = 31; b >= 0; --b) {
        r = g1_dbl(r);
        if ((k >> b) & 1) g1_add(r, p);
    }
    return r;
}

// k * p for a full-width canonical scalar (double-and-add, MSB first); for the handful of
// arbitrary-base multiplications outside the MSM (verifier-side folding)
SONIC_HD G1XYZZ g1_mul_scalar(const G1Affine& p, const Fr& k_canonical) {
    G1XYZZ r = G1XYZZ::inf();
    bool started = false;
    for (int i = Fr::N - 1; i >= 0; --i)
        for (int b = 31; b >= 0; --b) {
            if (started) r = g1_dbl(r);
            if ((k_canonical.l[i] >> b) & 1) { g1_madd(r, p); started = true; }
        }
    return r;
}

// Inverse of g1_compress.  Returns false for a malformed encoding (flags, x >= q, x not on the curve).
SONIC_HD bool g1_decompress(const uint8_t in[48], G1Affine& out) {
    if (!(in[0] & 0x80)) return false;
    if (in[0] & 0x40) {
        uint32_t rest = in[0] & 0x3F;
        for (int i = 1; i < 48; ++i) rest |= in[i];
        if (rest) return false;
        out = G1Affine::inf();
        return true;
    }
    Fq xc;
    for (int i = 0; i < 12; ++i) {
        uint32_t w = ((uint32_t)in[4 * i] << 24) | ((uint32_t)in[4 * i + 1] << 16) | ((uint32_t)in[4 * i + 2] << 8) | in[4 * i + 3];
        if (i == 0) w &= 0x1FFFFFFFu;
        xc.l[11 - i] = w;
    }
    // x < q ?
    uint32_t t = Chain::sub_cc(xc.l[0], FqParams::P(0));
    (void)t;
#pragma unroll
    for (int i = 1; i < 12; ++i) t = Chain::subc_cc(xc.l[i], FqParams::P(i));
    if (Chain::subc(0, 0) == 0) return false;
    Fq x = fp_to_mont(xc);
    Fq b;
    for (int i = 0; i < 12; ++i) b.l[i] = FqParams::B_M(i);
    Fq rhs = fp_add(fp_mul(fp_sqr(x), x), b);
    uint32_t e[12];
    for (int i = 0; i < 12; ++i) e[i] = FqParams::SQRT_EXP(i);
    Fq y = fp_pow_limbs(rhs, e, 12);
    if (fp_sqr(y) != rhs) return false;
    const bool want_big = (in[0] & 0x20) != 0;
    if (fp_canonical_gt_half(fp_from_mont(y)) != want_big) y = fp_neg(y);
    out.x = x;
    out.y = y;
    return true;
}

// 48-byte compressed boundary encoding: big-endian x, flag bits
// 0x80 compressed, 0x40 infinity, 0x20 y > (q-1)/2.  Input is canonical affine in Montgomery form.
SONIC_HD void g1_compress(const G1Affine& a, uint8_t out[48]) {
    if (a.is_inf()) {
        out[0] = 0xC0;
        for (int i = 1; i < 48; ++i) out[i] = 0;
        return;
    }
    Fq xc = fp_from_mont(a.x);
    Fq yc = fp_from_mont(a.y);
    for (int i = 0; i < 12; ++i) {
        uint32_t w = xc.l[11 - i];
        out[4 * i + 0] = (uint8_t)(w >> 24);
        out[4 * i + 1] = (uint8_t)(w >> 16);
        out[4 * i + 2] = (uint8_t)(w >> 8);
        out[4 * i + 3] = (uint8_t)(w);
    }
    out[0] |= 0x80;
    if (fp_canonical_gt_half(yc)) out[0] |= 0x20;
}

}  // namespace sonic
