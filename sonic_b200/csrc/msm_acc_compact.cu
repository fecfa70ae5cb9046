// Bucket accumulation, compact variant (`acc_mode=1`, the default).  See msm.cu for the pipeline.
#include "msm_acc.cuh"

namespace sonic {

// ---- stage 4, compact variant: the same chunked accumulation with the operands of the mixed
// addition in a shared-memory operand file and the field operations in a small loop --------------------
// The straight-line mixed addition above is ~4 000 instructions (64 KB): it does not fit the
// instruction cache, and `no_instruction` is its second stall reason (profiles/r01c).  Here the ten
// multiplications of an addition run through ONE multiplier, ONE squarer and ONE fused a*b-c*d in a
// loop (about 1 800 instructions), with the accumulator and temporaries held per thread in shared
// memory (9 slots x 12 limbs, thread-minor so that lanes never conflict).  Exceptional cases
// (P = +-Q) fall back to the generic formula, which then is cold code.
// Nine slots (54 KB per block, four blocks per SM at 128 registers): X2, Y2 and P are dead by the
// time PP, PPP and Q are produced, so those share their slots.  Measured on a n = 2^16 proof:
// 12 slots / 3 blocks 42.0 ms, 9 slots / 4 blocks 40.9 ms, 7 slots (R in registers) / 5 blocks 41.0 ms.
constexpr int ACC_SLOTS = 9;
constexpr int ACC_COMPACT_BLOCKS = 4;
enum { SL_X1 = 0, SL_Y1, SL_ZZ1, SL_ZZZ1, SL_X2, SL_Y2, SL_P, SL_R, SL_RR, SL_PP = SL_X2, SL_PPP = SL_Y2, SL_Q = SL_P };
enum { OP_MULSUB = 0, OP_MUL = 1, OP_SQR = 2 };

SONIC_D Fq opf_load(const uint32_t* __restrict__ f, int slot) {
    Fq r;
#pragma unroll
    for (int l = 0; l < 12; ++l) r.l[l] = f[(slot * 12 + l) * 128];
    return r;
}
SONIC_D void opf_store(uint32_t* __restrict__ f, int slot, const Fq& v) {
#pragma unroll
    for (int l = 0; l < 12; ++l) f[(slot * 12 + l) * 128] = v.l[l];
}

__global__ void __launch_bounds__(128, ACC_COMPACT_BLOCKS)
k_msm_accumulate_compact(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets, uint32_t GB,
                         uint32_t L, const G1Affine* __restrict__ points,
                         G1XYZZ* __restrict__ buckets, G1XYZZ* __restrict__ head, G1XYZZ* __restrict__ tail) {
    extern __shared__ uint32_t opf_all[];
    uint32_t* f = opf_all + threadIdx.x;  // this thread's column of the operand file
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = offsets[GB];
    const uint64_t start64 = (uint64_t)t * L;
    if (start64 >= total) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (total - start < L) ? total : start + L;
    uint32_t lo = 0, hi = GB;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t gb = lo;
    uint32_t bend = offsets[gb + 1];
    while (bend <= start) { ++gb; bend = offsets[gb + 1]; }
    bool cont = offsets[gb] < start;
    bool fresh = true;
    bool acc_inf = true;
    // the micro-program of one mixed addition: {op, dst, a, b, c}
    //   P = X2*ZZ1 - X1; R = Y2*ZZZ1 - Y1; PP = P^2; PPP = P*PP; Q = X1*PP; ZZ1 = ZZ1*PP; ZZZ1 = ZZZ1*PPP; RR = R^2
    constexpr uint32_t PROG[8] = {
        OP_MULSUB | SL_P << 4 | SL_X2 << 8 | SL_ZZ1 << 12 | SL_X1 << 16,
        OP_MULSUB | SL_R << 4 | SL_Y2 << 8 | SL_ZZZ1 << 12 | SL_Y1 << 16,
        OP_SQR | SL_PP << 4 | SL_P << 8 | SL_P << 12,
        OP_MUL | SL_PPP << 4 | SL_P << 8 | SL_PP << 12,
        OP_MUL | SL_Q << 4 | SL_X1 << 8 | SL_PP << 12,
        OP_MUL | SL_ZZ1 << 4 | SL_ZZ1 << 8 | SL_PP << 12,
        OP_MUL | SL_ZZZ1 << 4 | SL_ZZZ1 << 8 | SL_PPP << 12,
        OP_SQR | SL_RR << 4 | SL_R << 8 | SL_R << 12,
    };
    for (uint32_t p = start; p < end; ++p) {
        const G1Affine pt = fetch_entry(points, entries[p]);
        if (fresh) {
            acc_inf = pt.is_inf();
            opf_store(f, SL_X1, pt.x);
            opf_store(f, SL_Y1, pt.y);
            opf_store(f, SL_ZZ1, Fq::one());
            opf_store(f, SL_ZZZ1, Fq::one());
            fresh = false;
        } else if (!pt.is_inf()) {
            if (acc_inf) {
                acc_inf = false;
                opf_store(f, SL_X1, pt.x);
                opf_store(f, SL_Y1, pt.y);
                opf_store(f, SL_ZZ1, Fq::one());
                opf_store(f, SL_ZZZ1, Fq::one());
            } else {
                opf_store(f, SL_X2, pt.x);
                opf_store(f, SL_Y2, pt.y);
                bool exceptional = false;
#pragma unroll 1
                for (int step = 0; step < 8; ++step) {
                    const uint32_t code = PROG[step];
                    const int op = code & 15, dst = (code >> 4) & 15;
                    const Fq a = opf_load(f, (code >> 8) & 15);
                    Fq r;
                    if (op == OP_SQR) {
                        r = fp_sqr(a);
                    } else {
                        r = fp_mul(a, opf_load(f, (code >> 12) & 15));
                        if (op == OP_MULSUB) r = fp_sub(r, opf_load(f, (code >> 16) & 15));
                    }
                    opf_store(f, dst, r);
                    if (step == 1 && opf_load(f, SL_P).is_zero()) { exceptional = true; break; }
                }
                if (exceptional) {
                    // P = +-Q: doubling or cancellation, through the generic formula (cold)
                    G1XYZZ acc;
                    acc.x = opf_load(f, SL_X1); acc.y = opf_load(f, SL_Y1); acc.zz = opf_load(f, SL_ZZ1); acc.zzz = opf_load(f, SL_ZZZ1);
                    g1_madd(acc, pt);
                    acc_inf = acc.is_inf();
                    opf_store(f, SL_X1, acc.x); opf_store(f, SL_Y1, acc.y); opf_store(f, SL_ZZ1, acc.zz); opf_store(f, SL_ZZZ1, acc.zzz);
                } else {
                    // X3 = R^2 - PPP - 2Q ; Y3 = R*(Q - X3) - Y1*PPP
                    const Fq ppp = opf_load(f, SL_PPP), q = opf_load(f, SL_Q);
                    const Fq x3 = fp_sub(fp_sub(opf_load(f, SL_RR), ppp), fp_dbl(q));
                    const Fq y3 = fp_mul_sub2(opf_load(f, SL_R), fp_sub(q, x3), opf_load(f, SL_Y1), ppp);
                    opf_store(f, SL_X1, x3);
                    opf_store(f, SL_Y1, y3);
                }
            }
        }
        if (p + 1 == bend || p + 1 == end) {
            G1XYZZ acc;
            if (acc_inf) {
                acc = G1XYZZ::inf();
            } else {
                acc.x = opf_load(f, SL_X1); acc.y = opf_load(f, SL_Y1); acc.zz = opf_load(f, SL_ZZ1); acc.zzz = opf_load(f, SL_ZZZ1);
            }
            if (cont) store_xyzz(head + t, acc);
            else if (bend <= end) store_xyzz(buckets + gb, acc);
            else store_xyzz(tail + t, acc);
            if (p + 1 < end) {
                ++gb;
                bend = offsets[gb + 1];
                if (bend <= p + 1) {
                    // a run of empty buckets (a sliver of a job leaves most of its 2^(c-1) buckets empty: walking them
                    // one load at a time cost a millisecond): last bucket whose offset is <= p + 1, by bisection
                    uint32_t blo = gb, bhi = GB;
                    while (bhi - blo > 1) {
                        const uint32_t mid = (blo + bhi) >> 1;
                        if (offsets[mid] <= p + 1) blo = mid; else bhi = mid;
                    }
                    gb = blo;
                    bend = offsets[gb + 1];
                }
                cont = false;
                fresh = true;
            }
        }
    }
}

void launch_accumulate_compact(Ctx& cx, uint32_t chunks, const uint32_t* entries, const uint32_t* offsets, uint32_t GB, uint32_t L,
                               const G1Affine* points, G1XYZZ* buckets, G1XYZZ* head, G1XYZZ* tail) {
    (void)cx;
    const size_t smem = (size_t)ACC_SLOTS * 12 * 128 * sizeof(uint32_t);
    // opt-in above 48 KB of dynamic shared memory: a host-side attribute of the current device's
    // context, set per launch so that it survives sonic_shutdown / sonic_init on another device
    SONIC_CUDA(cudaFuncSetAttribute(k_msm_accumulate_compact, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    SONIC_LAUNCH(k_msm_accumulate_compact, div_up(chunks, 128), 128, smem, entries, offsets, GB, L, points, buckets, head, tail);
}

}  // namespace sonic
