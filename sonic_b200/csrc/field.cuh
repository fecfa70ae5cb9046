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

// Rolled variant: the same rows, two per loop iteration (so the accumulator roles return to
// where they started), the multiplier limbs rotated through registers.  Same IMAD count, one
// sixth of the code: a point addition then fits the instruction cache, which the fully
// unrolled form (67 KB per mixed addition) does not.
template <class P>
SONIC_HD Fp<P> fp_mul_rolled(const Fp<P>& a, const Fp<P>& b) {
    constexpr int N = P::N;
    static_assert(N % 2 == 0, "even limb count");
    uint32_t ev[N], od[N], bb[N];
#pragma unroll
    for (int k = 0; k < N; ++k) { ev[k] = 0; od[k] = 0; bb[k] = b.l[k]; }
#pragma unroll 1
    for (int i = 0; i < N; i += 2) {
        mont_row_next<P>(ev, od, a.l, bb[0]);  // even-aligned now in od
        mont_row_next<P>(od, ev, a.l, bb[1]);  // and back in ev
#pragma unroll
        for (int k = 0; k < N - 2; ++k) bb[k] = bb[k + 2];
    }
    Fp<P> r;
    r.l[0] = Chain::add_cc(ev[1], od[0]);
#pragma unroll
    for (int k = 1; k < N - 1; ++k) r.l[k] = Chain::addc_cc(ev[k + 1], od[k]);
    r.l[N - 1] = Chain::addc(od[N - 1], 0);
    fp_reduce_once(r);
    return r;
}

#ifndef SONIC_MUL_ROLLED
#define SONIC_MUL_ROLLED 0
#endif

template <class P>
SONIC_HD Fp<P> fp_mul(const Fp<P>& a, const Fp<P>& b) {
#if SONIC_MUL_ROLLED
    return fp_mul_rolled(a, b);
#else
    return fp_mul_unrolled(a, b);
#endif
}

template <class P>
SONIC_HD Fp<P> fp_sqr(const Fp<P>& a) { return fp_mul(a, a); }

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
