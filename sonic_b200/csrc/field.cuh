// Montgomery arithmetic for the two BLS12-381 prime fields on 32-bit limbs.
//
// Replaces (by value, not by algorithm) the galois-field `Prime` arithmetic the
// reference reaches through `Fr`/`Fq` (reference call sites: src/Sonic/SRS.hs:29-41,
// src/Sonic/Utils.hs:18, src/Sonic/CommitmentScheme.hs:43-44).
//
// Layout: little-endian 32-bit limbs, Fq = 12 limbs (R = 2^384), Fr = 8 limbs
// (R = 2^256); values are kept fully reduced in [0, p) in Montgomery form.
//
// The multiplier is a row-wise CIOS that keeps two interleaved accumulators:
// an "even" one aligned at limb 0 and an "odd" one aligned at limb 1.  Every
// 32x32->64 product of a row lands limb-aligned in one of the two, so a whole
// row is two carry chains of `mad.lo.cc / madc.hi.cc` pairs, which ptxas fuses
// into one IMAD.WIDE.U32(.X) per product: N*N products for a*b, N*N for m*p and
// N for the m's = 2N^2+N IMADs (300 for Fq, 136 for Fr) -- the LMAC count the
// roofline is quoted in.  After a row is reduced the
// value is shifted down one limb, which swaps the roles of the two accumulators.
//
// The same source compiles for the host (portable 64-bit emulation of the carry
// chain) so that tests can exercise the exact limb algorithm without a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SONIC_HD __host__ __device__ __forceinline__
#define SONIC_D __device__ __forceinline__
#else
#define SONIC_HD inline
#define SONIC_D inline
#endif

#define SONIC_CONST_ARRAY(name, n, ...)                                  \
    static SONIC_HD constexpr uint32_t name(int i) {                     \
        constexpr uint32_t t[n] = {__VA_ARGS__};                         \
        return t[i];                                                     \
    }

#include "constants.cuh"

namespace sonic {

// --------------------------------------------------------------------------------------
// carry-chain primitives
// --------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
struct Chain {
    // r = a + b, sets CF
    static SONIC_D uint32_t add_cc(uint32_t a, uint32_t b) {
        uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
    }
    static SONIC_D uint32_t addc_cc(uint32_t a, uint32_t b) {
        uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
    }
    static SONIC_D uint32_t addc(uint32_t a, uint32_t b) {
        uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
    }
    static SONIC_D uint32_t sub_cc(uint32_t a, uint32_t b) {
        uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
    }
    static SONIC_D uint32_t subc_cc(uint32_t a, uint32_t b) {
        uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
    }
    static SONIC_D uint32_t subc(uint32_t a, uint32_t b) {
        uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
    }
    // (hi:lo) = a*b
    static SONIC_D void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
        asm volatile("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b));
    }
    // (hi:lo) = a*b + (chi:clo), sets CF  (no carry-in)
    static SONIC_D void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                     : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
    }
    // (hi:lo) = a*b + (chi:clo) + CF, sets CF
    static SONIC_D void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
                     : "=&r"(lo), "=r"(hi) : "r"(a), "r"(b), "r"(clo), "r"(chi));
    }
};
#else
// Host emulation of the PTX carry flag (tests only; never the product path).
struct Chain {
    static inline uint32_t& cf() { static thread_local uint32_t f = 0; return f; }
    static inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
    static inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cf(); cf() = (uint32_t)(t >> 32); return (uint32_t)t; }
    static inline uint32_t addc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a + b + cf()); }
    static inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; cf() = (uint32_t)(t >> 63); return (uint32_t)t; }
    static inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cf(); cf() = (uint32_t)(t >> 63); return (uint32_t)t; }
    static inline uint32_t subc(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a - b - cf()); }
    static inline void mul_wide(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a * b; lo = (uint32_t)t; hi = (uint32_t)(t >> 32); }
    static inline void mad_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
        unsigned __int128 t = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)chi << 32) | clo);
        lo = (uint32_t)t; hi = (uint32_t)(t >> 32); cf() = (uint32_t)(t >> 64);
    }
    static inline void madc_wide_cc(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b, uint32_t clo, uint32_t chi) {
        unsigned __int128 t = (unsigned __int128)((uint64_t)a * b) + (((uint64_t)chi << 32) | clo) + cf();
        lo = (uint32_t)t; hi = (uint32_t)(t >> 32); cf() = (uint32_t)(t >> 64);
    }
};
#endif

// --------------------------------------------------------------------------------------
// Field element
// --------------------------------------------------------------------------------------
template <class P>
struct alignas(16) Fp {
    using Params = P;
    static constexpr int N = P::N;
    uint32_t l[N];

    static SONIC_HD Fp zero() { Fp r; for (int i = 0; i < N; ++i) r.l[i] = 0; return r; }
    static SONIC_HD Fp one() { Fp r; for (int i = 0; i < N; ++i) r.l[i] = P::ONE(i); return r; }
    static SONIC_HD Fp r2() { Fp r; for (int i = 0; i < N; ++i) r.l[i] = P::R2(i); return r; }

    SONIC_HD bool is_zero() const { uint32_t t = 0; for (int i = 0; i < N; ++i) t |= l[i]; return t == 0; }
    SONIC_HD bool operator==(const Fp& o) const { uint32_t t = 0; for (int i = 0; i < N; ++i) t |= l[i] ^ o.l[i]; return t == 0; }
    SONIC_HD bool operator!=(const Fp& o) const { return !(*this == o); }
};

// r = a - p if a >= p  (a < 2p)
template <class P>
SONIC_HD void fp_reduce_once(Fp<P>& a) {
    constexpr int N = P::N;
    uint32_t t[N];
    t[0] = Chain::sub_cc(a.l[0], P::P(0));
#pragma unroll
    for (int i = 1; i < N; ++i) t[i] = Chain::subc_cc(a.l[i], P::P(i));
    uint32_t borrow = Chain::subc(0, 0);  // 0xffffffff if a < p
#pragma unroll
    for (int i = 0; i < N; ++i) a.l[i] = borrow ? a.l[i] : t[i];
}

template <class P>
SONIC_HD Fp<P> fp_add(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.l[0] = Chain::add_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; ++i) r.l[i] = Chain::addc_cc(a.l[i], b.l[i]);
    // p < 2^(32N-1) for both fields, so a+b < 2p never carries out of limb N-1
    fp_reduce_once(r);
    return r;
}

template <class P>
SONIC_HD Fp<P> fp_sub(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    Fp<P> r;
    r.l[0] = Chain::sub_cc(a.l[0], b.l[0]);
#pragma unroll
    for (int i = 1; i < N; ++i) r.l[i] = Chain::subc_cc(a.l[i], b.l[i]);
    uint32_t mask = Chain::subc(0, 0);  // all ones on borrow
    r.l[0] = Chain::add_cc(r.l[0], P::P(0) & mask);
#pragma unroll
    for (int i = 1; i < N - 1; ++i) r.l[i] = Chain::addc_cc(r.l[i], P::P(i) & mask);
    r.l[N - 1] = Chain::addc(r.l[N - 1], P::P(N - 1) & mask);
    return r;
}

template <class P>
SONIC_HD Fp<P> fp_neg(const Fp<P>& a) {
    return fp_sub(Fp<P>::zero(), a);
}

template <class P>
SONIC_HD Fp<P> fp_dbl(const Fp<P>& a) { return fp_add(a, a); }

// ---- Montgomery multiplication ---------------------------------------------------------
// Reduction half of a row: `ev` is the even-aligned accumulator, `od` the odd-aligned one.
// Adds m*p so that ev[0] becomes 0.
template <class P>
SONIC_HD void mont_reduce_row(uint32_t* ev, uint32_t* od) {
    constexpr int N = P::N;
    const uint32_t m = ev[0] * P::INV;
    // odd limbs of p into the odd accumulator
    Chain::mad_wide_cc(od[0], od[1], P::P(1), m, od[0], od[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) Chain::madc_wide_cc(od[j], od[j + 1], P::P(j + 1), m, od[j], od[j + 1]);
    // even limbs of p into the even accumulator; its carry-out has the weight of od[N-1]
    Chain::mad_wide_cc(ev[0], ev[1], P::P(0), m, ev[0], ev[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) Chain::madc_wide_cc(ev[j], ev[j + 1], P::P(j), m, ev[j], ev[j + 1]);
    od[N - 1] = Chain::addc(od[N - 1], 0);
}

// First row: accumulators start from zero.  On exit `ev` (even-aligned, ev[0]==0) and `od`.
template <class P>
SONIC_HD void mont_row_first(uint32_t* ev, uint32_t* od, const uint32_t* a, uint32_t bi) {
    constexpr int N = P::N;
#pragma unroll
    for (int j = 0; j < N; j += 2) {
        Chain::mul_wide(ev[j], ev[j + 1], a[j], bi);
        Chain::mul_wide(od[j], od[j + 1], a[j + 1], bi);
    }
    mont_reduce_row<P>(ev, od);
}

// Next row.  On entry `pe` is the previous even-aligned accumulator (pe[0]==0) and `po` the
// previous odd-aligned one.  Dividing by 2^32 turns `po` into the new even-aligned
// accumulator and (pe >> 64) into the new odd-aligned one, with pe[1] folded into po[0].
// On exit the new even-aligned value lives in `po` and the new odd-aligned one in `pe`.
template <class P>
SONIC_HD void mont_row_next(uint32_t* pe, uint32_t* po, const uint32_t* a, uint32_t bi) {
    constexpr int N = P::N;
    po[0] = Chain::add_cc(po[0], pe[1]);  // carry has the weight of the new odd limb 0
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) Chain::madc_wide_cc(pe[j], pe[j + 1], a[j + 1], bi, pe[j + 2], pe[j + 3]);
    Chain::madc_wide_cc(pe[N - 2], pe[N - 1], a[N - 1], bi, 0, 0);
    Chain::mad_wide_cc(po[0], po[1], a[0], bi, po[0], po[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) Chain::madc_wide_cc(po[j], po[j + 1], a[j], bi, po[j], po[j + 1]);
    pe[N - 1] = Chain::addc(pe[N - 1], 0);
    mont_reduce_row<P>(po, pe);
}

// Fully unrolled variant: 2N^2+N IMADs in one straight line (about 420 SASS instructions for Fq).
template <class P>
SONIC_HD Fp<P> fp_mul_unrolled(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    uint32_t ev[N], od[N];
    mont_row_first<P>(ev, od, a.l, b.l[0]);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_row_next<P>(ev, od, a.l, b.l[i]);          // even-aligned now in od
        if (i + 1 < N) mont_row_next<P>(od, ev, a.l, b.l[i + 1]);  // back in ev
    }
    // N is even: after the loop the even-aligned accumulator is `od`, the odd-aligned `ev`.
    Fp<P> r;
    r.l[0] = Chain::add_cc(od[1], ev[0]);
#pragma unroll
    for (int k = 1; k < N - 1; ++k) r.l[k] = Chain::addc_cc(od[k + 1], ev[k]);
    r.l[N - 1] = Chain::addc(ev[N - 1], 0);
    fp_reduce_once(r);
    return r;
}

// Two independent products with their rows interleaved: four carry chains in flight instead of
// two, so that a scheduler with few resident warps still finds a ready IMAD while the previous
// one's carry is in flight (ncu: the accumulate kernel's top stall is `wait` on IMAD.WIDE.U32.X).
template <class P>
SONIC_HD void fp_mul2(const Fp<P>& a1, const Fp<P>& b1, const Fp<P>& a2, const Fp<P>& b2, Fp<P>& r1, Fp<P>& r2) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    uint32_t e1[N], o1[N], e2[N], o2[N];
    mont_row_first<P>(e1, o1, a1.l, b1.l[0]);
    mont_row_first<P>(e2, o2, a2.l, b2.l[0]);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_row_next<P>(e1, o1, a1.l, b1.l[i]);
        mont_row_next<P>(e2, o2, a2.l, b2.l[i]);
        if (i + 1 < N) {
            mont_row_next<P>(o1, e1, a1.l, b1.l[i + 1]);
            mont_row_next<P>(o2, e2, a2.l, b2.l[i + 1]);
        }
    }
    r1.l[0] = Chain::add_cc(o1[1], e1[0]);
#pragma unroll
    for (int k = 1; k < N - 1; ++k) r1.l[k] = Chain::addc_cc(o1[k + 1], e1[k]);
    r1.l[N - 1] = Chain::addc(e1[N - 1], 0);
    r2.l[0] = Chain::add_cc(o2[1], e2[0]);
#pragma unroll
    for (int k = 1; k < N - 1; ++k) r2.l[k] = Chain::addc_cc(o2[k + 1], e2[k]);
    r2.l[N - 1] = Chain::addc(e2[N - 1], 0);
    fp_reduce_once(r1);
    fp_reduce_once(r2);
}

// Rolled variant: the same rows, two per loop iteration (so the accumulator roles return to
// where they started), the multiplier limbs rotated through registers.  Same IMAD count, one
// sixth of the code: a point addition then fits the instruction cache, which the fully
// unrolled form (67 KB per mixed addition) does not.
template <class P, int ROWS_PER_ITER = 2>
SONIC_HD Fp<P> fp_mul_rolled(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0 && ROWS_PER_ITER % 2 == 0 && N % ROWS_PER_ITER == 0, "row grouping");
    uint32_t ev[N], od[N], bb[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { ev[k] = 0; od[k] = 0; bb[k] = b.l[k]; }
#pragma unroll 1
    for (int i = 0; i < N; i += ROWS_PER_ITER) {
#pragma unroll
        for (int r = 0; r < ROWS_PER_ITER; r += 2) {
            mont_row_next<P>(ev, od, a.l, bb[r]);      // even-aligned now in od
            mont_row_next<P>(od, ev, a.l, bb[r + 1]);  // and back in ev
        }
#pragma unroll
        for (int k = 0; k < N - ROWS_PER_ITER; ++k) bb[k] = bb[k + ROWS_PER_ITER];
    }
    Fp<P> r;
    r.l[0] = Chain::add_cc(ev[1], od[0]);
#pragma unroll
    for (int k = 1; k < N - 1; ++k) r.l[k] = Chain::addc_cc(ev[k + 1], od[k]);
    r.l[N - 1] = Chain::addc(od[N - 1], 0);
    fp_reduce_once(r);
    return r;
}

// ---- wide (unreduced) products and a separate Montgomery reduction -----------------------
// Splitting "multiply" from "reduce" lets the multiplier spend fewer IMADs than the interleaved
// CIOS above: (1) one level of Karatsuba on the 2K x 2K limb product (3 K x K products instead
// of 4: 108 instead of 144 IMADs for Fq), paid for with additions that run on the ALU pipe, which
// idles while the IMAD pipe is saturated; (2) sums of two products are reduced once (fp_mul_add2).

// K x K schoolbook product with the same even/odd accumulator trick (K even): out[2K] = a * b.
template <int K>
SONIC_HD void wide_mul(uint32_t* out, const uint32_t* a, const uint32_t* b) {
    static_assert(K % 2 == 0, "even limb count");
    uint32_t E[2 * K + 2], O[2 * K + 2];  // O[k] holds limb k+1
#pragma unroll
    for (int k = 0; k < 2 * K + 2; ++k) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int j = 0; j < K; j += 2) {
        Chain::mul_wide(E[j], E[j + 1], a[j], b[0]);
        Chain::mul_wide(O[j], O[j + 1], a[j + 1], b[0]);
    }
#pragma unroll
    for (int i = 1; i < K; ++i) {
        if (i & 1) {
            // odd row: a[even j] lands on odd positions i+j -> O[i+j-1]; a[odd j] on even positions -> E[i+j]
            Chain::mad_wide_cc(O[i - 1], O[i], a[0], b[i], O[i - 1], O[i]);
#pragma unroll
            for (int j = 2; j < K; j += 2) Chain::madc_wide_cc(O[i + j - 1], O[i + j], a[j], b[i], O[i + j - 1], O[i + j]);
            O[i + K - 1] = Chain::addc(O[i + K - 1], 0);
            Chain::mad_wide_cc(E[i + 1], E[i + 2], a[1], b[i], E[i + 1], E[i + 2]);
#pragma unroll
            for (int j = 3; j < K; j += 2) Chain::madc_wide_cc(E[i + j], E[i + j + 1], a[j], b[i], E[i + j], E[i + j + 1]);
            E[i + K + 1] = Chain::addc(E[i + K + 1], 0);
        } else {
            // even row: a[even j] on even positions -> E[i+j]; a[odd j] on odd positions -> O[i+j-1]
            Chain::mad_wide_cc(E[i], E[i + 1], a[0], b[i], E[i], E[i + 1]);
#pragma unroll
            for (int j = 2; j < K; j += 2) Chain::madc_wide_cc(E[i + j], E[i + j + 1], a[j], b[i], E[i + j], E[i + j + 1]);
            E[i + K] = Chain::addc(E[i + K], 0);
            Chain::mad_wide_cc(O[i], O[i + 1], a[1], b[i], O[i], O[i + 1]);
#pragma unroll
            for (int j = 3; j < K; j += 2) Chain::madc_wide_cc(O[i + j - 1], O[i + j], a[j], b[i], O[i + j - 1], O[i + j]);
            O[i + K] = Chain::addc(O[i + K], 0);
        }
    }
    out[0] = E[0];
    out[1] = Chain::add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 2 * K - 1; ++k) out[k] = Chain::addc_cc(E[k], O[k - 1]);
    out[2 * K - 1] = Chain::addc(E[2 * K - 1], O[2 * K - 2]);
}

// out[2N] = a^2: the N(N-1)/2 off-diagonal products once, doubled, plus the N squares on the
// diagonal -- N(N+1)/2 IMADs (78 for Fq) instead of N^2.
template <int N>
SONIC_HD void wide_sqr(uint32_t* out, const uint32_t* a) {
    static_assert(N % 2 == 0, "even limb count");
    uint32_t E[2 * N + 2], O[2 * N + 2];  // O[k] holds limb k+1
#pragma unroll
    for (int k = 0; k < 2 * N + 2; ++k) { E[k] = 0; O[k] = 0; }
#pragma unroll
    for (int i = 0; i < N - 1; ++i) {
        // a[i] * a[j], j > i: position i+j; j = i+1, i+3, .. are odd positions -> O[i+j-1]
        Chain::mad_wide_cc(O[2 * i], O[2 * i + 1], a[i + 1], a[i], O[2 * i], O[2 * i + 1]);
        int top = 2 * i;
#pragma unroll
        for (int j = i + 3; j < N; j += 2) {
            Chain::madc_wide_cc(O[i + j - 1], O[i + j], a[j], a[i], O[i + j - 1], O[i + j]);
            top = i + j - 1;
        }
        O[top + 2] = Chain::addc(O[top + 2], 0);
        // j = i+2, i+4, .. are even positions -> E[i+j]
        if (i + 2 < N) {
            Chain::mad_wide_cc(E[2 * i + 2], E[2 * i + 3], a[i + 2], a[i], E[2 * i + 2], E[2 * i + 3]);
            int etop = 2 * i + 2;
#pragma unroll
            for (int j = i + 4; j < N; j += 2) {
                Chain::madc_wide_cc(E[i + j], E[i + j + 1], a[j], a[i], E[i + j], E[i + j + 1]);
                etop = i + j;
            }
            E[etop + 2] = Chain::addc(E[etop + 2], 0);
        }
    }
    // S = E + (O << 32): the off-diagonal half
    uint32_t S[2 * N];
    S[0] = E[0];
    S[1] = Chain::add_cc(E[1], O[0]);
#pragma unroll
    for (int k = 2; k < 2 * N - 1; ++k) S[k] = Chain::addc_cc(E[k], O[k - 1]);
    S[2 * N - 1] = Chain::addc(E[2 * N - 1], O[2 * N - 2]);
    // out = 2*S + sum_i a_i^2 << 64 i
    uint32_t lo, hi;
    Chain::mul_wide(lo, hi, a[0], a[0]);
    out[0] = Chain::add_cc(lo, S[0] << 1);
    out[1] = Chain::addc_cc(hi, (S[1] << 1) | (S[0] >> 31));
#pragma unroll
    for (int i = 1; i < N; ++i) {
        Chain::mul_wide(lo, hi, a[i], a[i]);
        out[2 * i] = Chain::addc_cc(lo, (S[2 * i] << 1) | (S[2 * i - 1] >> 31));
        out[2 * i + 1] = Chain::addc_cc(hi, (S[2 * i + 1] << 1) | (S[2 * i] >> 31));
    }
}

// out[2N] = a * b by one level of Karatsuba over halves of K = N/2 limbs
template <int N>
SONIC_HD void wide_mul_karatsuba(uint32_t* out, const uint32_t* a, const uint32_t* b) {
    constexpr int K = N / 2;
    static_assert(N % 4 == 0, "halves must have an even limb count");
    uint32_t z2[2 * K], sa[K], sb[K], zm[2 * K + 1];
    wide_mul<K>(out, a, b);            // z0 -> out[0 .. 2K)
    wide_mul<K>(z2, a + K, b + K);     // z2
    // sa = a0 + a1, sb = b0 + b1 with carry bits ca, cb
    sa[0] = Chain::add_cc(a[0], a[K]);
#pragma unroll
    for (int k = 1; k < K; ++k) sa[k] = Chain::addc_cc(a[k], a[K + k]);
    const uint32_t ca = Chain::addc(0, 0);
    sb[0] = Chain::add_cc(b[0], b[K]);
#pragma unroll
    for (int k = 1; k < K; ++k) sb[k] = Chain::addc_cc(b[k], b[K + k]);
    const uint32_t cb = Chain::addc(0, 0);
    wide_mul<K>(zm, sa, sb);
    // zm += (ca ? sb : 0) << 32K  +  (cb ? sa : 0) << 32K  +  (ca & cb) << 64K
    const uint32_t ma = 0u - ca, mb = 0u - cb;
    zm[K] = Chain::add_cc(zm[K], sb[0] & ma);
#pragma unroll
    for (int k = 1; k < K; ++k) zm[K + k] = Chain::addc_cc(zm[K + k], sb[k] & ma);
    zm[2 * K] = Chain::addc(ca & cb, 0);
    zm[K] = Chain::add_cc(zm[K], sa[0] & mb);
#pragma unroll
    for (int k = 1; k < K; ++k) zm[K + k] = Chain::addc_cc(zm[K + k], sa[k] & mb);
    zm[2 * K] = Chain::addc(zm[2 * K], 0);
    // z1 = zm - z0 - z2  (non-negative, 2K+1 limbs)
    zm[0] = Chain::sub_cc(zm[0], out[0]);
#pragma unroll
    for (int k = 1; k < 2 * K; ++k) zm[k] = Chain::subc_cc(zm[k], out[k]);
    zm[2 * K] = Chain::subc(zm[2 * K], 0);
    zm[0] = Chain::sub_cc(zm[0], z2[0]);
#pragma unroll
    for (int k = 1; k < 2 * K; ++k) zm[k] = Chain::subc_cc(zm[k], z2[k]);
    zm[2 * K] = Chain::subc(zm[2 * K], 0);
    // out = z0 + z1 << 32K + z2 << 64K
#pragma unroll
    for (int k = 0; k < 2 * K; ++k) out[2 * K + k] = z2[k];
    out[K] = Chain::add_cc(out[K], zm[0]);
#pragma unroll
    for (int k = 1; k <= 2 * K; ++k) out[K + k] = Chain::addc_cc(out[K + k], zm[k]);
#pragma unroll
    for (int k = 3 * K + 1; k < 4 * K - 1; ++k) out[k] = Chain::addc_cc(out[k], 0);
    out[4 * K - 1] = Chain::addc(out[4 * K - 1], 0);
}

// One reduction row on a sliding window: on entry `pe` is the even-aligned accumulator with
// pe[0] == 0 and `po` the odd-aligned one; the window moves down one limb, `t` (the next limb of
// the wide operand) enters at the top, and m*p is added so that the new low limb vanishes.
// On exit the even-aligned accumulator is `po`, the odd-aligned one `pe`.
template <class P>
SONIC_HD void mont_reduce_next(uint32_t* pe, uint32_t* po, uint32_t t) {
    constexpr int N = P::N;
    po[0] = Chain::add_cc(po[0], pe[1]);  // carry has the weight of the new odd limb 0
    const uint32_t m = po[0] * P::INV;
#pragma unroll
    for (int j = 0; j < N - 2; j += 2) Chain::madc_wide_cc(pe[j], pe[j + 1], P::P(j + 1), m, pe[j + 2], pe[j + 3]);
    Chain::madc_wide_cc(pe[N - 2], pe[N - 1], P::P(N - 1), m, t, 0);  // the top odd slot is free: t goes there
    Chain::mad_wide_cc(po[0], po[1], P::P(0), m, po[0], po[1]);
#pragma unroll
    for (int j = 2; j < N; j += 2) Chain::madc_wide_cc(po[j], po[j + 1], P::P(j), m, po[j], po[j + 1]);
    pe[N - 1] = Chain::addc(pe[N - 1], 0);
}

// Montgomery reduction of a 2N-limb value T < p * 2^(32N) * (small): returns T / R mod p, < 2p
// before the final conditional subtraction (callers with larger T pass extra_sub = true).
template <class P>
SONIC_HD Fp<P> mont_reduce_wide(const uint32_t* T, bool extra_sub = false) {
    constexpr int N = P::N;
    uint32_t ev[N], od[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { ev[k] = T[k]; od[k] = 0; }
    mont_reduce_row<P>(ev, od);
#pragma unroll
    for (int i = 1; i < N; i += 2) {
        mont_reduce_next<P>(ev, od, T[N + i - 1]);              // even-aligned now in od
        if (i + 1 < N) mont_reduce_next<P>(od, ev, T[N + i]);  // back in ev
    }
    // N even: even-aligned accumulator is `od` (od[0] == 0), odd-aligned `ev`; the last limb enters at the top
    Fp<P> r;
    r.l[0] = Chain::add_cc(od[1], ev[0]);
#pragma unroll
    for (int k = 1; k < N - 1; ++k) r.l[k] = Chain::addc_cc(od[k + 1], ev[k]);
    r.l[N - 1] = Chain::addc(ev[N - 1], T[2 * N - 1]);
    fp_reduce_once(r);
    if (extra_sub) fp_reduce_once(r);
    return r;
}

template <class P>
SONIC_HD Fp<P> fp_mul_karatsuba(const Fp<P>& a, const Fp<P>& b) {
    uint32_t T[2 * P::N];
    wide_mul_karatsuba<P::N>(T, a.l, b.l);
    return mont_reduce_wide<P>(T);
}

// a*b + c*d with a single reduction.  Inputs reduced; 2p^2 + R p < 2 p R for both fields
// (p / R < 1/2), so the result is < 2p before the conditional subtraction.
#ifndef SONIC_WIDE_KARATSUBA
#define SONIC_WIDE_KARATSUBA 0
#endif
template <class P>
SONIC_HD Fp<P> fp_mul_add2(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    constexpr int N = P::N;
    uint32_t T[2 * N], U[2 * N];
#if SONIC_WIDE_KARATSUBA
    wide_mul_karatsuba<N>(T, a.l, b.l);
    wide_mul_karatsuba<N>(U, c.l, d.l);
#else
    wide_mul<N>(T, a.l, b.l);
    wide_mul<N>(U, c.l, d.l);
#endif
    T[0] = Chain::add_cc(T[0], U[0]);
#pragma unroll
    for (int k = 1; k < 2 * N - 1; ++k) T[k] = Chain::addc_cc(T[k], U[k]);
    T[2 * N - 1] = Chain::addc(T[2 * N - 1], U[2 * N - 1]);
    return mont_reduce_wide<P>(T);
}

// a*b - c*d
template <class P>
SONIC_HD Fp<P> fp_mul_sub2(const Fp<P>& a, const Fp<P>& b, const Fp<P>& c, const Fp<P>& d) {
    return fp_mul_add2(a, b, fp_neg(c), d);
}

// which multiplier `fp_mul` is: 0 interleaved CIOS (unrolled), 1 the same rolled, 2 Karatsuba + wide reduction
#ifndef SONIC_MUL_VARIANT
#define SONIC_MUL_VARIANT 0
#endif

template <class P>
SONIC_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#if SONIC_MUL_VARIANT == 1
    return fp_mul_rolled<P, (P::N % 4 == 0 ? 4 : 2)>(a, b);
#elif SONIC_MUL_VARIANT == 3
    return fp_mul_rolled<P, P::N / 2>(a, b);
#elif SONIC_MUL_VARIANT == 2
    return fp_mul_karatsuba(a, b);
#else
    return fp_mul_unrolled(a, b);
#endif
}

// 0: squaring is a multiplication; 1: half-product squaring + wide reduction
#ifndef SONIC_SQR_WIDE
#define SONIC_SQR_WIDE 1
#endif
template <class P>
SONIC_HD Fp<P> fp_sqr(const Fp<P>& a) {
#if SONIC_SQR_WIDE
    uint32_t T[2 * P::N];
    wide_sqr<P::N>(T, a.l);
    return mont_reduce_wide<P>(T);
#else
    return fp_mul(a, a);
#endif
}

// The same product as fp_mul through the wide path: the N^2 partial products first (rows independent of the
// reduction, so more of them are in flight), then one Montgomery reduction.  Same IMAD count as the interleaved
// CIOS; used where a lone warp waits on the latency of a multiplication rather than on the pipe (g1coop.cuh).
template <class P>
SONIC_HD Fp<P> fp_mul_wide(const Fp<P>& a, const Fp<P>& b) {
    uint32_t T[2 * P::N];
    wide_mul<P::N>(T, a.l, b.l);
    return mont_reduce_wide<P>(T);
}

// canonical <-> Montgomery
template <class P>
SONIC_HD Fp<P> fp_to_mont(const Fp<P>& a) { return fp_mul(a, Fp<P>::r2()); }

template <class P>
SONIC_HD Fp<P> fp_from_mont(const Fp<P>& a) {
    Fp<P> o = Fp<P>::zero();
    o.l[0] = 1;
    return fp_mul(a, o);
}

// a^e for a public exponent given as N limbs (square-and-multiply, MSB first)
template <class P>
SONIC_HD Fp<P> fp_pow_limbs(const Fp<P>& a, const uint32_t* e, int nlimbs) {
    Fp<P> r = Fp<P>::one();
    bool started = false;
    for (int i = nlimbs - 1; i >= 0; --i) {
        for (int b = 31; b >= 0; --b) {
            if (started) r = fp_sqr(r);
            if ((e[i] >> b) & 1) {
                r = started ? fp_mul(r, a) : a;
                started = true;
            }
        }
    }
    return r;
}

// Fermat inverse; inv(0) = 0 (callers that must reject 0 test before calling)
template <class P>
SONIC_HD Fp<P> fp_inv(const Fp<P>& a) {
    uint32_t e[P::N];
    for (int i = 0; i < P::N; ++i) e[i] = P::PM2(i);
    return fp_pow_limbs(a, e, P::N);
}

// Inverse of ONE element on ONE thread without the Fermat ladder: Kaliski's almost-inverse (binary
// extended Euclid on the limbs -- shifts, additions, subtractions; k <= 2*bits iterations) yields
// a^-1 * 2^k, then four Montgomery products remove the power of two and restore the Montgomery
// factor.  The dependent chain is about 8x shorter than 2*bits field multiplications, which is what
// the end of an MSM pays for (to-affine of a single point on a single thread).  Data-dependent
// branches: not for warps of independent inversions (use fp_inv there).  inv(0) = 0.
template <class P>
SONIC_HD Fp<P> fp_inv_euclid(const Fp<P>& a) {
    constexpr int N = P::N;
    if (a.is_zero()) return a;
    uint32_t u[N], v[N], r[N], s[N], t[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { u[i] = P::P(i); v[i] = a.l[i]; r[i] = 0; s[i] = 0; }
    s[0] = 1;
    int k = 0;
    for (;;) {
        uint32_t vz = 0;
#pragma unroll
        for (int i = 0; i < N; ++i) vz |= v[i];
        if (vz == 0) break;
        if ((u[0] & 1u) == 0) {
#pragma unroll
            for (int i = 0; i < N - 1; ++i) u[i] = (u[i] >> 1) | (u[i + 1] << 31);
            u[N - 1] >>= 1;
#pragma unroll
            for (int i = N - 1; i > 0; --i) s[i] = (s[i] << 1) | (s[i - 1] >> 31);
            s[0] <<= 1;
        } else if ((v[0] & 1u) == 0) {
#pragma unroll
            for (int i = 0; i < N - 1; ++i) v[i] = (v[i] >> 1) | (v[i + 1] << 31);
            v[N - 1] >>= 1;
#pragma unroll
            for (int i = N - 1; i > 0; --i) r[i] = (r[i] << 1) | (r[i - 1] >> 31);
            r[0] <<= 1;
        } else {
            t[0] = Chain::sub_cc(u[0], v[0]);
#pragma unroll
            for (int i = 1; i < N; ++i) t[i] = Chain::subc_cc(u[i], v[i]);
            const uint32_t borrow = Chain::subc(0, 0);  // 0xffffffff if u < v
            uint32_t tz = 0;
#pragma unroll
            for (int i = 0; i < N; ++i) tz |= t[i];
            if (borrow == 0 && tz != 0) {
                // u > v: u = (u - v) / 2, r += s, s *= 2
#pragma unroll
                for (int i = 0; i < N - 1; ++i) u[i] = (t[i] >> 1) | (t[i + 1] << 31);
                u[N - 1] = t[N - 1] >> 1;
                r[0] = Chain::add_cc(r[0], s[0]);
#pragma unroll
                for (int i = 1; i < N; ++i) r[i] = Chain::addc_cc(r[i], s[i]);
#pragma unroll
                for (int i = N - 1; i > 0; --i) s[i] = (s[i] << 1) | (s[i - 1] >> 31);
                s[0] <<= 1;
            } else {
                // v >= u: v = (v - u) / 2, s += r, r *= 2
                t[0] = Chain::sub_cc(v[0], u[0]);
#pragma unroll
                for (int i = 1; i < N; ++i) t[i] = Chain::subc_cc(v[i], u[i]);
#pragma unroll
                for (int i = 0; i < N - 1; ++i) v[i] = (t[i] >> 1) | (t[i + 1] << 31);
                v[N - 1] = t[N - 1] >> 1;
                s[0] = Chain::add_cc(s[0], r[0]);
#pragma unroll
                for (int i = 1; i < N; ++i) s[i] = Chain::addc_cc(s[i], r[i]);
#pragma unroll
                for (int i = N - 1; i > 0; --i) r[i] = (r[i] << 1) | (r[i - 1] >> 31);
                r[0] <<= 1;
            }
        }
        ++k;
    }
    // r < 2p holds a^-1 * 2^k up to sign: x = p - (r mod p)
    Fp<P> x;
#pragma unroll
    for (int i = 0; i < N; ++i) x.l[i] = r[i];
    fp_reduce_once(x);
    x = fp_neg(x);
    // input a = A*R (Montgomery form), so x = A^-1 R^-1 2^k; wanted A^-1 R = x * R^2 / 2^k.
    // Two products by R^2 add R each; a product by 2^j divides by 2^(32N - j): two of them remove 2^k.
    x = fp_mul(fp_mul(x, Fp<P>::r2()), Fp<P>::r2());
    const int d1 = k >> 1, d2 = k - d1;
    Fp<P> w = Fp<P>::zero();
    w.l[(32 * N - d1) >> 5] = 1u << ((32 * N - d1) & 31);
    x = fp_mul(x, w);
    w = Fp<P>::zero();
    w.l[(32 * N - d2) >> 5] = 1u << ((32 * N - d2) & 31);
    return fp_mul(x, w);
}

// a^k for a 64-bit non-negative exponent
template <class P>
SONIC_HD Fp<P> fp_pow_u64(const Fp<P>& a, uint64_t k) {
    uint32_t e[2] = {(uint32_t)k, (uint32_t)(k >> 32)};
    return fp_pow_limbs(a, e, 2);
}

// canonical (non-Montgomery) comparison helper: a > (p-1)/2 ?
template <class P>
SONIC_HD bool fp_canonical_gt_half(const Fp<P>& c) {
    // c - HALF - 1 >= 0  <=>  no borrow from HALF - c
    uint32_t t = Chain::sub_cc(P::HALF(0), c.l[0]);
    (void)t;
#pragma unroll
    for (int i = 1; i < P::N; ++i) t = Chain::subc_cc(P::HALF(i), c.l[i]);
    return Chain::subc(0, 0) != 0;  // borrow => c > HALF
}

using Fq = Fp<FqParams>;
using Fr = Fp<FrParams>;

}  // namespace sonic
