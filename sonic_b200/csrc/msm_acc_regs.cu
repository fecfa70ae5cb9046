// Bucket accumulation, register-resident variant (`acc_mode=0`): the mixed addition as one
// straight line of ~4 000 instructions with the accumulator in registers.  See msm.cu for the pipeline.
#include "msm_acc.cuh"

namespace sonic {

__global__ void __launch_bounds__(128, 3)
k_msm_accumulate(const uint32_t* __restrict__ entries, const uint32_t* __restrict__ offsets, uint32_t GB,
                 uint32_t L, const G1Affine* __restrict__ points,
                 G1XYZZ* __restrict__ buckets, G1XYZZ* __restrict__ head, G1XYZZ* __restrict__ tail) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t total = offsets[GB];  // number of non-zero digits; the grid is sized for the upper bound
    const uint64_t start64 = (uint64_t)t * L;
    if (start64 >= total) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (total - start < L) ? total : start + L;
    // bucket that holds entry `start`: last gb with offsets[gb] <= start (non-empty by construction)
    uint32_t lo = 0, hi = GB;  // invariant: offsets[lo] <= start < offsets[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t gb = lo;
    uint32_t bend = offsets[gb + 1];
    while (bend <= start) { ++gb; bend = offsets[gb + 1]; }  // (defensive; lo already satisfies it)
    bool cont = offsets[gb] < start;  // bucket began in an earlier chunk
    bool fresh = true;
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t p = start; p < end; ++p) {
        G1Affine pt = fetch_entry(points, entries[p]);
        if (fresh) { acc = G1XYZZ::from_affine(pt); fresh = false; }
        else g1_madd(acc, pt);
        if (p + 1 == bend || p + 1 == end) {
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

void launch_accumulate_regs(Ctx& cx, uint32_t chunks, const uint32_t* entries, const uint32_t* offsets, uint32_t GB, uint32_t L,
                            const G1Affine* points, G1XYZZ* buckets, G1XYZZ* head, G1XYZZ* tail) {
    (void)cx;
    SONIC_LAUNCH(k_msm_accumulate, div_up(chunks, 128), 128, 0, entries, offsets, GB, L, points, buckets, head, tail);
}

}  // namespace sonic
