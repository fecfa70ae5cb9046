// Point operations shared by a "quad" of four adjacent lanes.
//
// The stages after the bucket accumulation (fix-up, bucket reduction, final folds) are chains of
// DEPENDENT full additions: a thread alone spends 14 field multiplications, one after the other,
// on each, and a lone warp cannot hide the carry-chain latency of the multiplier (measured: a
// dependent g1_add takes ~20 us, see profiles/r02*_latency.md).  The multiplications inside one
// addition are not all dependent, though: add-2008-s has depth 4.  Here the four lanes of a quad
// hold the same operands, each computes ONE of the independent products of a stage -- the same
// instruction stream on different data, so the warp never diverges -- and the products are
// exchanged with quad-wide shuffles:
//
//   add:  [X1 ZZ2 | X2 ZZ1 | Y1 ZZZ2 | Y2 ZZZ1] -> [P^2 | R^2 | ZZ1 ZZ2 | ZZZ1 ZZZ2]
//         -> [P PP | U1 PP | ZZ12 PP | S1 P] -> [R (Q - X3) - (S1 P) PP | ZZZ12 PPP | - | -]
//   dbl:  [U^2 | X^2 | - | -] -> [U V | X V | V ZZ | M^2] -> [M (S - X3) - W Y | W ZZZ | - | -]
//
// 3 products + one fused two-product step instead of 14 (add), 2 + 1 instead of 7.5 (dbl).
// Every lane of the quad ends with the same result; exceptional cases (infinity, P = +-Q) are
// decided on identical data by all four lanes and take the generic formulas redundantly.
#pragma once
#include "g1.cuh"

namespace sonic {

struct Quad {
    unsigned mask;  // the four lanes of this quad within the warp
    int lane;       // 0..3
    SONIC_D Quad() {
        const int l = threadIdx.x & 31;
        lane = l & 3;
        mask = 0xFu << (l & ~3);
    }
    // value `v` of quad lane `src`, for all four lanes
    SONIC_D Fq get(const Fq& v, int src) const {
        Fq r;
#pragma unroll
        for (int k = 0; k < Fq::N; ++k) r.l[k] = __shfl_sync(mask, v.l[k], src, 4);
        return r;
    }
};

// branch-free 4-way select on the quad lane (all lanes run the same instructions)
SONIC_D Fq quad_pick(int lane, const Fq& a0, const Fq& a1, const Fq& a2, const Fq& a3) {
    Fq r;
#pragma unroll
    for (int k = 0; k < Fq::N; ++k) {
        const uint32_t lo = lane & 1 ? a1.l[k] : a0.l[k];
        const uint32_t hi = lane & 1 ? a3.l[k] : a2.l[k];
        r.l[k] = lane & 2 ? hi : lo;
    }
    return r;
}

// acc += b, both XYZZ, by the four lanes of a quad (same inputs and same result on every lane)
SONIC_D void g1_add_quad(const Quad& q, G1XYZZ& acc, const G1XYZZ& b) {
    if (b.is_inf()) return;
    if (acc.is_inf()) { acc = b; return; }
    // stage 1
    Fq t = fp_mul(quad_pick(q.lane, acc.x, b.x, acc.y, b.y), quad_pick(q.lane, b.zz, acc.zz, b.zzz, acc.zzz));
    const Fq U1 = q.get(t, 0), U2 = q.get(t, 1), S1 = q.get(t, 2), S2 = q.get(t, 3);
    const Fq P = fp_sub(U2, U1), R = fp_sub(S2, S1);
    if (P.is_zero()) {
        if (R.is_zero()) acc = g1_dbl(acc);
        else acc = G1XYZZ::inf();
        return;
    }
    // stage 2
    t = fp_mul(quad_pick(q.lane, P, R, acc.zz, acc.zzz), quad_pick(q.lane, P, R, b.zz, b.zzz));
    const Fq PP = q.get(t, 0), RR = q.get(t, 1), ZZ12 = q.get(t, 2), ZZZ12 = q.get(t, 3);
    // stage 3
    t = fp_mul(quad_pick(q.lane, P, U1, ZZ12, S1), quad_pick(q.lane, PP, PP, PP, P));
    const Fq PPP = q.get(t, 0), Q = q.get(t, 1), ZZ3 = q.get(t, 2), S1P = q.get(t, 3);
    const Fq X3 = fp_sub(fp_sub(RR, PPP), fp_dbl(Q));
    // stage 4: lane 0: R (Q - X3) - (S1 P) PP ; lane 1: ZZZ12 PPP - 0 ; lanes 2, 3 repeat lane 0's work
    const Fq zero = Fq::zero();
    const Fq a = q.lane == 1 ? ZZZ12 : R, bq = q.lane == 1 ? PPP : fp_sub(Q, X3);
    const Fq c = q.lane == 1 ? zero : S1P, dq = q.lane == 1 ? zero : PP;
    t = fp_mul_sub2(a, bq, c, dq);
    acc.x = X3;
    acc.y = q.get(t, 0);
    acc.zz = ZZ3;
    acc.zzz = q.get(t, 1);
}

// 2 * p by the four lanes of a quad
SONIC_D G1XYZZ g1_dbl_quad(const Quad& q, const G1XYZZ& p) {
    if (p.is_inf()) return G1XYZZ::inf();
    const Fq U = fp_dbl(p.y);
    // stage 1: [U^2 | X^2 | - | -]
    Fq t = fp_mul(q.lane & 1 ? p.x : U, q.lane & 1 ? p.x : U);
    const Fq V = q.get(t, 0), X2 = q.get(t, 1);
    const Fq M = fp_add(fp_dbl(X2), X2);
    // stage 2: [U V | X V | V ZZ | M^2]
    t = fp_mul(quad_pick(q.lane, U, p.x, V, M), quad_pick(q.lane, V, V, p.zz, M));
    const Fq W = q.get(t, 0), S = q.get(t, 1), ZZ3 = q.get(t, 2), MM = q.get(t, 3);
    const Fq X3 = fp_sub(MM, fp_dbl(S));
    // stage 3: lane 0: M (S - X3) - W Y ; lane 1: W ZZZ - 0
    const Fq zero = Fq::zero();
    const Fq a = q.lane == 1 ? W : M, bq = q.lane == 1 ? p.zzz : fp_sub(S, X3);
    const Fq c = q.lane == 1 ? zero : W, dq = q.lane == 1 ? zero : p.y;
    t = fp_mul_sub2(a, bq, c, dq);
    G1XYZZ r;
    r.x = X3;
    r.y = q.get(t, 0);
    r.zz = ZZ3;
    r.zzz = q.get(t, 1);
    return r;
}
}  // namespace sonic
