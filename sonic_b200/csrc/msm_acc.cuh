// Shared pieces of the two bucket-accumulation kernels (each lives in its own translation unit so
// that ptxas works on them in parallel: they are the largest kernels of the library).
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

// launchers (msm_acc_regs.cu, msm_acc_compact.cu)
void launch_accumulate_regs(Ctx& cx, uint32_t chunks, const uint32_t* entries, const uint32_t* offsets, uint32_t GB, uint32_t L,
                            const G1Affine* points, G1XYZZ* buckets, G1XYZZ* head, G1XYZZ* tail);
void launch_accumulate_compact(Ctx& cx, uint32_t chunks, const uint32_t* entries, const uint32_t* offsets, uint32_t GB, uint32_t L,
                               const G1Affine* points, G1XYZZ* buckets, G1XYZZ* head, G1XYZZ* tail);

}  // namespace sonic
