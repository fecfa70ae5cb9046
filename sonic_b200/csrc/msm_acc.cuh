// Shared pieces of the MSM translation units (sort + orchestration in msm.cu, the two bucket-accumulation
// kernels and the reduction stages each in their own, so that ptxas works on the large kernels in parallel).
#pragma once
#include "internal.h"
#include "g1io.cuh"

namespace sonic {

// ---- stage 4: chunked bucket accumulation (the hot kernel) -------------------------------
SONIC_D G1Affine fetch_entry(const G1Affine* __restrict__ points, uint32_t e) {
    G1Affine p = load_affine(points + (e & 0x7fffffffu));
    if (e & 0x80000000u) p.y = fp_neg(p.y);  // entries never reference the point at infinity's y: neg(0)=0
    return p;
}

struct MsmPlan {
    int c, W, sets;
    uint32_t B, GB, L, n_tot;
};

// launchers (msm_acc_regs.cu, msm_acc_compact.cu)
void launch_accumulate_regs(Ctx& cx, uint32_t chunks, const uint32_t* entries, const uint32_t* offsets, uint32_t GB, uint32_t L,
                            const G1Affine* points, G1XYZZ* buckets, G1XYZZ* head, G1XYZZ* tail);
void launch_accumulate_compact(Ctx& cx, uint32_t chunks, const uint32_t* entries, const uint32_t* offsets, uint32_t GB, uint32_t L,
                               const G1Affine* points, G1XYZZ* buckets, G1XYZZ* head, G1XYZZ* tail);

// acc_mode 2 (msm_acc_affine.cu): pairwise tree reduction of the buckets in affine coordinates with batched
// inversions; leaves every bucket final (no head / tail pieces, no fix-up)
void launch_accumulate_affine(Ctx& cx, uint64_t entries_max, const uint32_t* entries, const uint32_t* offsets, uint32_t GB,
                              const G1Affine* points, G1XYZZ* buckets, double max_mean_bucket);

// stages 5-7 (msm_reduce.cu)
void msm_reduce_stage(Ctx& cx, const MsmPlan& p, int M, const uint32_t* offsets, uint32_t chunks, G1XYZZ* buckets,
                      const G1XYZZ* head, const G1XYZZ* tail, G1Affine* d_out_aff, uint8_t* d_out_comp, cudaEvent_t fixup_done, bool buckets_final = false);

}  // namespace sonic
