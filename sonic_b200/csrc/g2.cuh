// Fq2 = Fq[u]/(u^2+1) and the G2 group law (E': y^2 = x^3 + 4(1+u)) in XYZZ coordinates.
//
// Not on the prover's path: G2 only appears in the h-vectors of the SRS
// (src/Sonic/SRS.hs:35-36,40-41), of which pcV reads four elements
// (src/Sonic/CommitmentScheme.hs:62-68).  This is the G2 fixed-base batch that
// SURVEY.md section 8f lists as a "next" item, so that SRS.new can be fully device-side.
// An Fq2 product is two fused sums of two Fq products (one Montgomery reduction each).
#pragma once
#include "field.cuh"

namespace sonic {

struct Fq2 {
    Fq c0, c1;
    static SONIC_HD Fq2 zero() { Fq2 r; r.c0 = Fq::zero(); r.c1 = Fq::zero(); return r; }
    static SONIC_HD Fq2 one() { Fq2 r; r.c0 = Fq::one(); r.c1 = Fq::zero(); return r; }
    SONIC_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    SONIC_HD bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
    SONIC_HD bool operator!=(const Fq2& o) const { return !(*this == o); }
};

SONIC_HD Fq2 f2_add(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fp_add(a.c0, b.c0); r.c1 = fp_add(a.c1, b.c1); return r; }
SONIC_HD Fq2 f2_sub(const Fq2& a, const Fq2& b) { Fq2 r; r.c0 = fp_sub(a.c0, b.c0); r.c1 = fp_sub(a.c1, b.c1); return r; }
SONIC_HD Fq2 f2_dbl(const Fq2& a) { return f2_add(a, a); }
SONIC_HD Fq2 f2_neg(const Fq2& a) { Fq2 r; r.c0 = fp_neg(a.c0); r.c1 = fp_neg(a.c1); return r; }
#if defined(__CUDACC__)
#define SONIC_G2_FN __host__ __device__ __noinline__
#else
#define SONIC_G2_FN inline
#endif
// out of line on the device: every G2 formula calls these a dozen times and the code is cold
SONIC_G2_FN Fq2 f2_mul(const Fq2& a, const Fq2& b) {
    Fq2 r;
    r.c0 = fp_mul_sub2(a.c0, b.c0, a.c1, b.c1);
    r.c1 = fp_mul_add2(a.c0, b.c1, a.c1, b.c0);
    return r;
}
SONIC_G2_FN Fq2 f2_sqr(const Fq2& a) {
    Fq2 r;
    r.c0 = fp_mul(fp_add(a.c0, a.c1), fp_sub(a.c0, a.c1));
    r.c1 = fp_dbl(fp_mul(a.c0, a.c1));
    return r;
}
SONIC_HD Fq2 f2_inv(const Fq2& a) {
    Fq n = fp_inv(fp_add(fp_sqr(a.c0), fp_sqr(a.c1)));
    Fq2 r;
    r.c0 = fp_mul(a.c0, n);
    r.c1 = fp_neg(fp_mul(a.c1, n));
    return r;
}

struct G2Affine {
    Fq2 x, y;
    SONIC_HD bool is_inf() const { return x.is_zero() && y.is_zero(); }  // (0,0) is not on E'
    static SONIC_HD G2Affine inf() { G2Affine p; p.x = Fq2::zero(); p.y = Fq2::zero(); return p; }
    static SONIC_HD G2Affine gen() {
        G2Affine p;
        for (int i = 0; i < Fq::N; ++i) {
            p.x.c0.l[i] = FqParams::G2X0_M(i); p.x.c1.l[i] = FqParams::G2X1_M(i);
            p.y.c0.l[i] = FqParams::G2Y0_M(i); p.y.c1.l[i] = FqParams::G2Y1_M(i);
        }
        return p;
    }
};

struct G2XYZZ {
    Fq2 x, y, zz, zzz;
    SONIC_HD bool is_inf() const { return zz.is_zero(); }
    static SONIC_HD G2XYZZ inf() { G2XYZZ p; p.x = Fq2::one(); p.y = Fq2::one(); p.zz = Fq2::zero(); p.zzz = Fq2::zero(); return p; }
    static SONIC_HD G2XYZZ from_affine(const G2Affine& a) {
        if (a.is_inf()) return inf();
        G2XYZZ p; p.x = a.x; p.y = a.y; p.zz = Fq2::one(); p.zzz = Fq2::one(); return p;
    }
};

SONIC_HD G2XYZZ g2_mdbl(const G2Affine& a) {
    if (a.is_inf()) return G2XYZZ::inf();
    Fq2 U = f2_dbl(a.y), V = f2_sqr(U), W = f2_mul(U, V), S = f2_mul(a.x, V), X2 = f2_sqr(a.x);
    Fq2 M = f2_add(f2_dbl(X2), X2);
    G2XYZZ r;
    r.x = f2_sub(f2_sqr(M), f2_dbl(S));
    r.y = f2_sub(f2_mul(M, f2_sub(S, r.x)), f2_mul(W, a.y));
    r.zz = V;
    r.zzz = W;
    return r;
}

SONIC_HD G2XYZZ g2_dbl(const G2XYZZ& p) {
    Fq2 U = f2_dbl(p.y), V = f2_sqr(U), W = f2_mul(U, V), S = f2_mul(p.x, V), X2 = f2_sqr(p.x);
    Fq2 M = f2_add(f2_dbl(X2), X2);
    G2XYZZ r;
    r.x = f2_sub(f2_sqr(M), f2_dbl(S));
    r.y = f2_sub(f2_mul(M, f2_sub(S, r.x)), f2_mul(W, p.y));
    r.zz = f2_mul(V, p.zz);
    r.zzz = f2_mul(W, p.zzz);
    return r;
}

SONIC_HD void g2_madd(G2XYZZ& acc, const G2Affine& a) {
    if (a.is_inf()) return;
    if (acc.is_inf()) { acc = G2XYZZ::from_affine(a); return; }
    Fq2 U2 = f2_mul(a.x, acc.zz), S2 = f2_mul(a.y, acc.zzz);
    Fq2 P = f2_sub(U2, acc.x), R = f2_sub(S2, acc.y);
    if (P.is_zero()) {
        if (R.is_zero()) acc = g2_mdbl(a);
        else acc = G2XYZZ::inf();
        return;
    }
    Fq2 PP = f2_sqr(P), PPP = f2_mul(P, PP), Q = f2_mul(acc.x, PP);
    Fq2 X3 = f2_sub(f2_sub(f2_sqr(R), PPP), f2_dbl(Q));
    Fq2 Y3 = f2_sub(f2_mul(R, f2_sub(Q, X3)), f2_mul(acc.y, PPP));
    acc.x = X3;
    acc.y = Y3;
    acc.zz = f2_mul(acc.zz, PP);
    acc.zzz = f2_mul(acc.zzz, PPP);
}

SONIC_HD G2Affine g2_to_affine(const G2XYZZ& p) {
    if (p.is_inf()) return G2Affine::inf();
    Fq2 zi = f2_inv(p.zzz), t = f2_mul(p.zz, zi), zzi = f2_sqr(t);
    G2Affine r;
    r.x = f2_mul(p.x, zzi);
    r.y = f2_mul(p.y, zi);
    return r;
}

// 96-byte compressed encoding: x.c1 || x.c0 big-endian, flags as for G1, sign = y lexicographically
// larger than -y (c1 compared first, c0 if c1 == 0)
SONIC_HD void g2_compress(const G2Affine& a, uint8_t out[96]) {
    if (a.is_inf()) {
        out[0] = 0xC0;
        for (int i = 1; i < 96; ++i) out[i] = 0;
        return;
    }
    const Fq x1 = fp_from_mont(a.x.c1), x0 = fp_from_mont(a.x.c0);
    const Fq y1 = fp_from_mont(a.y.c1), y0 = fp_from_mont(a.y.c0);
    for (int i = 0; i < 12; ++i) {
        const uint32_t w1 = x1.l[11 - i], w0 = x0.l[11 - i];
        out[4 * i + 0] = (uint8_t)(w1 >> 24); out[4 * i + 1] = (uint8_t)(w1 >> 16); out[4 * i + 2] = (uint8_t)(w1 >> 8); out[4 * i + 3] = (uint8_t)w1;
        out[48 + 4 * i + 0] = (uint8_t)(w0 >> 24); out[48 + 4 * i + 1] = (uint8_t)(w0 >> 16); out[48 + 4 * i + 2] = (uint8_t)(w0 >> 8); out[48 + 4 * i + 3] = (uint8_t)w0;
    }
    out[0] |= 0x80;
    const bool big = y1.is_zero() ? fp_canonical_gt_half(y0) : fp_canonical_gt_half(y1);
    if (big) out[0] |= 0x20;
}

}  // namespace sonic
