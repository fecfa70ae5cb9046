// Device self-test hooks: run the field / curve primitives over arrays so that the GPU
// parity tests can localise a failure below the MSM (tests/test_gpu_arith.py).
#include "internal.h"
#include "g1io.cuh"

namespace sonic {

template <class F>
__global__ void k_selftest_field(int op, const F* __restrict__ a, const F* __restrict__ b, F* __restrict__ out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b[i], r;
    switch (op) {
        case 0: r = fp_mul(x, y); break;
        case 1: r = fp_add(x, y); break;
        case 2: r = fp_sub(x, y); break;
        case 3: r = fp_to_mont(x); break;
        case 4: r = fp_from_mont(x); break;
        case 5: r = fp_inv(x); break;
        case 6: r = fp_sqr(x); break;
        default: r = fp_neg(x); break;
    }
    out[i] = r;
}

// op: 0 madd(acc, affine b), 1 add(acc, xyzz b), 2 dbl(acc), 3 mdbl(affine a); then to_affine
__global__ void k_selftest_g1(int op, const G1XYZZ* __restrict__ a, const G1XYZZ* __restrict__ b, G1Affine* __restrict__ out,
                              uint8_t* __restrict__ comp, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1XYZZ r = load_xyzz(a + i);
    G1Affine bf; bf.x = b[i].x; bf.y = b[i].y;
    G1Affine af; af.x = a[i].x; af.y = a[i].y;
    switch (op) {
        case 0: g1_madd(r, bf); break;
        case 1: g1_add(r, load_xyzz(b + i)); break;
        case 2: r = g1_dbl(r); break;
        default: r = g1_mdbl(af); break;
    }
    G1Affine o = g1_to_affine(r);
    out[i] = o;
    g1_compress(o, comp + (size_t)i * 48);
}

int selftest_field(Ctx& cx, int which, int op, const void* a, const void* b, void* out, uint32_t n) {
    const size_t esz = which == 0 ? sizeof(Fq) : sizeof(Fr);
    char* da = cx.arena.get<char>(esz * n);
    char* db = cx.arena.get<char>(esz * n);
    char* dout = cx.arena.get<char>(esz * n);
    SONIC_CUDA(cudaMemcpyAsync(da, a, esz * n, cudaMemcpyHostToDevice, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(db, b, esz * n, cudaMemcpyHostToDevice, cx.stream));
    if (which == 0) SONIC_LAUNCH(k_selftest_field<Fq>, div_up(n, 128), 128, 0, op, (const Fq*)da, (const Fq*)db, (Fq*)dout, n);
    else SONIC_LAUNCH(k_selftest_field<Fr>, div_up(n, 128), 128, 0, op, (const Fr*)da, (const Fr*)db, (Fr*)dout, n);
    SONIC_CUDA(cudaMemcpyAsync(out, dout, esz * n, cudaMemcpyDeviceToHost, cx.stream));
    SONIC_CUDA(cudaStreamSynchronize(cx.stream));
    return SONIC_OK;
}

int selftest_g1(Ctx& cx, int op, const void* a, const void* b, void* out_aff, void* out_comp, uint32_t n) {
    G1XYZZ* da = cx.arena.get<G1XYZZ>(n);
    G1XYZZ* db = cx.arena.get<G1XYZZ>(n);
    G1Affine* dout = cx.arena.get<G1Affine>(n);
    uint8_t* dcomp = cx.arena.get<uint8_t>((size_t)n * 48);
    SONIC_CUDA(cudaMemcpyAsync(da, a, sizeof(G1XYZZ) * n, cudaMemcpyHostToDevice, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(db, b, sizeof(G1XYZZ) * n, cudaMemcpyHostToDevice, cx.stream));
    SONIC_LAUNCH(k_selftest_g1, div_up(n, 64), 64, 0, op, da, db, dout, dcomp, n);
    SONIC_CUDA(cudaMemcpyAsync(out_aff, dout, sizeof(G1Affine) * n, cudaMemcpyDeviceToHost, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(out_comp, dcomp, (size_t)n * 48, cudaMemcpyDeviceToHost, cx.stream));
    SONIC_CUDA(cudaStreamSynchronize(cx.stream));
    return SONIC_OK;
}

}  // namespace sonic
