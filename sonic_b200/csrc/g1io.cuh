// 128-bit vector loads / stores of curve points (coalesced 16-byte accesses).
#pragma once
#include "g1.cuh"

namespace sonic {

// ---- vector loads / stores ------------------------------------------------------------
SONIC_D G1Affine load_affine(const G1Affine* p) {
    G1Affine r;
#if defined(__CUDA_ARCH__)
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 v[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) v[i] = __ldg(q + i);
    uint32_t* w = r.x.l;
#pragma unroll
    for (int i = 0; i < 3; ++i) { w[4 * i] = v[i].x; w[4 * i + 1] = v[i].y; w[4 * i + 2] = v[i].z; w[4 * i + 3] = v[i].w; }
    w = r.y.l;
#pragma unroll
    for (int i = 0; i < 3; ++i) { w[4 * i] = v[3 + i].x; w[4 * i + 1] = v[3 + i].y; w[4 * i + 2] = v[3 + i].z; w[4 * i + 3] = v[3 + i].w; }
#else
    r = *p;
#endif
    return r;
}

SONIC_D void store_xyzz(G1XYZZ* p, const G1XYZZ& v) {
#if defined(__CUDA_ARCH__)
    uint4* q = reinterpret_cast<uint4*>(p);
    const uint32_t* w = v.x.l;
    const Fq* f[4] = {&v.x, &v.y, &v.zz, &v.zzz};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        w = f[k]->l;
#pragma unroll
        for (int i = 0; i < 3; ++i) q[3 * k + i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
#else
    *p = v;
#endif
}

SONIC_D G1XYZZ load_xyzz(const G1XYZZ* p) {
    G1XYZZ r;
#if defined(__CUDA_ARCH__)
    const uint4* q = reinterpret_cast<const uint4*>(p);
    Fq* f[4] = {&r.x, &r.y, &r.zz, &r.zzz};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            uint4 v = q[3 * k + i];
            f[k]->l[4 * i] = v.x; f[k]->l[4 * i + 1] = v.y; f[k]->l[4 * i + 2] = v.z; f[k]->l[4 * i + 3] = v.w;
        }
    }
#else
    r = *p;
#endif
    return r;
}

}  // namespace sonic
