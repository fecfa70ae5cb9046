// Signed-digit (window c) recoding of a canonical Fr scalar for the bucket method.
//
// A scalar s in [0, r) is first folded to |s'| <= (r-1)/2 < 2^254 by s' = s - r when
// s > (r-1)/2 (the point's sign absorbs it), then cut into W = ceil(255/c) digits in
// [-2^(c-1), 2^(c-1)].  Because |s'| < 2^254 and c*W >= 255 the top window never carries out.
#pragma once
#include "field.cuh"

namespace sonic {

SONIC_HD constexpr int msm_num_windows(int c) { return (255 + c - 1) / c; }

struct ScalarDigits {
    uint32_t s[9];   // folded magnitude, one zero limb of padding for the bit extractor
    uint32_t carry;
    int c, pos;
    bool neg;

    SONIC_HD ScalarDigits(const uint32_t* canonical8, int c_) : carry(0), c(c_), pos(0) {
        Fr v;
        for (int i = 0; i < 8; ++i) v.l[i] = canonical8[i];
        neg = fp_canonical_gt_half<FrParams>(v);
        if (neg) {
            // r - v
            s[0] = Chain::sub_cc(FrParams::P(0), v.l[0]);
#pragma unroll
            for (int i = 1; i < 8; ++i) s[i] = Chain::subc_cc(FrParams::P(i), v.l[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) s[i] = v.l[i];
        }
        s[8] = 0;
    }

    // digits are produced from the least significant window up
    SONIC_HD int32_t next() {
        const int limb = pos >> 5, sh = pos & 31;
        uint64_t two = ((uint64_t)s[limb + 1 > 8 ? 8 : limb + 1] << 32) | s[limb > 8 ? 8 : limb];
        if (limb >= 8) two = 0;
        uint32_t raw = (uint32_t)((two >> sh) & ((1u << c) - 1)) + carry;
        pos += c;
        int32_t d;
        if (raw > (1u << (c - 1))) { d = (int32_t)raw - (int32_t)(1u << c); carry = 1; }
        else { d = (int32_t)raw; carry = 0; }
        return neg ? -d : d;
    }
};

}  // namespace sonic
