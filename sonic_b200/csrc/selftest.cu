// Device self-test hooks: run the field / curve primitives over arrays so that the GPU
// parity tests can localise a failure below the MSM (tests/test_gpu_arith.py).
#include "internal.h"
#include "g1io.cuh"
#include "g1coop.cuh"

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
        case 8: r = fp_inv_euclid(x); break;
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

// the same through the quad-cooperative formulas (g1coop.cuh): four lanes per element.  op 4 add, 5 dbl
__global__ void k_selftest_g1_quad(int op, const G1XYZZ* __restrict__ a, const G1XYZZ* __restrict__ b, G1Affine* __restrict__ out,
                                   uint8_t* __restrict__ comp, uint32_t n) {
    const uint32_t i = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (i >= n) return;   // n is padded to a multiple of 8 elements by the launcher: whole warps stay together
    const Quad q;
    G1XYZZ r = load_xyzz(a + i);
    if (op == 4) g1_add_quad(q, r, load_xyzz(b + i));
    else r = g1_dbl_quad(q, r);
    if (q.lane == 0) {
        G1Affine o = g1_to_affine(r);
        out[i] = o;
        g1_compress(o, comp + (size_t)i * 48);
    }
}

// Latency probe: every thread runs a chain of `iters` DEPENDENT operations; with one warp per block and
// one block per SM the time per operation is the latency a lone warp sees (the regime of the MSM's tail).
//   0 fp_mul (Fq)   1 g1_add   2 g1_dbl   3 g1_madd   4 g1_add_quad   5 g1_dbl_quad   6 fp_mul_wide   7 fp_sqr   8 fp_mul_sub2
__global__ void k_latency(int op, int iters, const G1XYZZ* __restrict__ seed, G1XYZZ* __restrict__ sink) {
    G1XYZZ acc = load_xyzz(seed), b = load_xyzz(seed + 1);
    G1Affine ba;
    ba.x = b.x;
    ba.y = b.y;
    const Quad q;
    for (int i = 0; i < iters; ++i) {
        switch (op) {
            case 0: acc.x = fp_mul(acc.x, b.x); break;
            case 1: g1_add(acc, b); break;
            case 2: acc = g1_dbl(acc); break;
            case 3: g1_madd(acc, ba); break;
            case 4: g1_add_quad(q, acc, b); break;
            case 5: acc = g1_dbl_quad(q, acc); break;
            case 6: acc.x = fp_mul_wide(acc.x, b.x); break;
            case 7: acc.x = fp_sqr(acc.x); break;
            default: acc.x = fp_mul_sub2(acc.x, b.x, acc.x, b.y); break;
        }
    }
    if (threadIdx.x == 0) store_xyzz(sink + blockIdx.x, acc);
}

double selftest_latency_ns(Ctx& cx, int op, int iters, int blocks, int threads) {
    G1XYZZ h[2];
    h[0] = G1XYZZ::from_affine(G1Affine::gen());
    h[1] = g1_dbl(g1_dbl(h[0]));   // 4G in XYZZ with ZZ != 1 (host build of the same headers)
    G1XYZZ* seed = cx.arena.get<G1XYZZ>(2);
    G1XYZZ* sink = cx.arena.get<G1XYZZ>(blocks);
    SONIC_CUDA(cudaMemcpyAsync(seed, h, sizeof h, cudaMemcpyHostToDevice, cx.stream));
    double best = 0;
    for (int rep = 0; rep < 3; ++rep) {
        SONIC_CUDA(cudaEventRecord(cx.ev[4], cx.stream));
        SONIC_LAUNCH(k_latency, blocks, threads, 0, op, iters, seed, sink);
        SONIC_CUDA(cudaEventRecord(cx.ev[5], cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        float ms = 0;
        SONIC_CUDA(cudaEventElapsedTime(&ms, cx.ev[4], cx.ev[5]));
        const double ns = (double)ms * 1e6 / iters;
        if (rep == 0 || ns < best) best = ns;
    }
    return best;
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
    if (op >= 4) SONIC_LAUNCH(k_selftest_g1_quad, div_up((uint64_t)n * 4, 64), 64, 0, op, da, db, dout, dcomp, n);
    else SONIC_LAUNCH(k_selftest_g1, div_up(n, 64), 64, 0, op, da, db, dout, dcomp, n);
    SONIC_CUDA(cudaMemcpyAsync(out_aff, dout, sizeof(G1Affine) * n, cudaMemcpyDeviceToHost, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(out_comp, dcomp, (size_t)n * 48, cudaMemcpyDeviceToHost, cx.stream));
    SONIC_CUDA(cudaStreamSynchronize(cx.stream));
    return SONIC_OK;
}

// ---- verifier-side G1 folding -------------------------------------------
// pcV (src/Sonic/CommitmentScheme.hs:51-68) checks  e(W, h^{alpha x}) e(g^v W^{-z}, h^alpha) == e(F, h^{x^{-d+max}}).
// k such checks with random weights r_i collapse into ONE multi-pairing whose G1 inputs are
//   A   = sum_i r_i W_i                         (against h^{alpha x})
//   B   = sum_i r_i (v_i G - z_i W_i)           (against h^alpha)
//   C_m = sum_{i : group_i = m} r_i F_i         (against h^{x^{-d+max_m}}, one per distinct `max`)
// The pairings themselves stay on the host library.  One thread per check does the three scalar
// multiplications (k is 3Q+4, a few dozen), one thread per output folds.
__global__ void __launch_bounds__(64) k_pcv_terms(uint32_t k, const uint8_t* __restrict__ F48, const uint8_t* __restrict__ W48,
                                                  const Fr* __restrict__ v, const Fr* __restrict__ z, const Fr* __restrict__ r,
                                                  G1XYZZ* __restrict__ tA, G1XYZZ* __restrict__ tB, G1XYZZ* __restrict__ tC,
                                                  uint32_t* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= k) return;
    G1Affine F, W;
    if (!g1_decompress(F48 + (size_t)i * 48, F) || !g1_decompress(W48 + (size_t)i * 48, W)) { atomicExch(bad, 1u); return; }
    const Fr rm = fp_to_mont(r[i]);
    const Fr rv = fp_from_mont(fp_mul(rm, fp_to_mont(v[i])));
    const Fr rz = fp_from_mont(fp_mul(rm, fp_to_mont(z[i])));
    store_xyzz(tA + i, g1_mul_scalar(W, r[i]));
    G1XYZZ b = g1_mul_scalar(G1Affine::gen(), rv);
    g1_add(b, g1_neg_xyzz(g1_mul_scalar(W, rz)));
    store_xyzz(tB + i, b);
    store_xyzz(tC + i, g1_mul_scalar(F, r[i]));
}

__global__ void k_pcv_fold(uint32_t k, const G1XYZZ* __restrict__ tA, const G1XYZZ* __restrict__ tB, const G1XYZZ* __restrict__ tC,
                           const uint32_t* __restrict__ group, uint32_t ngroups, uint8_t* __restrict__ out48) {
    const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;  // 0 = A, 1 = B, 2.. = C_m
    if (o >= 2 + ngroups) return;
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t i = 0; i < k; ++i) {
        if (o == 0) g1_add(acc, load_xyzz(tA + i));
        else if (o == 1) g1_add(acc, load_xyzz(tB + i));
        else if (group[i] == o - 2) g1_add(acc, load_xyzz(tC + i));
    }
    g1_compress(g1_to_affine(acc), out48 + (size_t)o * 48);
}

int pcv_fold(Ctx& cx, uint32_t k, const uint8_t* F48, const uint8_t* W48, const uint8_t* v32, const uint8_t* z32,
             const uint8_t* r32, const uint32_t* group, uint32_t ngroups, uint8_t* out48) {
    uint8_t* dF = cx.arena.get<uint8_t>((size_t)k * 48);
    uint8_t* dW = cx.arena.get<uint8_t>((size_t)k * 48);
    Fr* dv = cx.arena.get<Fr>(3 * (size_t)k);
    uint32_t* dg = cx.arena.get<uint32_t>(k);
    uint32_t* bad = cx.arena.get<uint32_t>(1);
    G1XYZZ* t = cx.arena.get<G1XYZZ>(3 * (size_t)k);
    uint8_t* dout = cx.arena.get<uint8_t>((size_t)(2 + ngroups) * 48);
    SONIC_CUDA(cudaMemcpyAsync(dF, F48, (size_t)k * 48, cudaMemcpyHostToDevice, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(dW, W48, (size_t)k * 48, cudaMemcpyHostToDevice, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(dv, v32, (size_t)k * 32, cudaMemcpyHostToDevice, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(dv + k, z32, (size_t)k * 32, cudaMemcpyHostToDevice, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(dv + 2 * (size_t)k, r32, (size_t)k * 32, cudaMemcpyHostToDevice, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(dg, group, (size_t)k * 4, cudaMemcpyHostToDevice, cx.stream));
    SONIC_CUDA(cudaMemsetAsync(bad, 0, 4, cx.stream));
    SONIC_LAUNCH(k_pcv_terms, div_up(k, 64), 64, 0, k, dF, dW, dv, dv + k, dv + 2 * (size_t)k, t, t + k, t + 2 * (size_t)k, bad);
    SONIC_LAUNCH(k_pcv_fold, div_up(2 + ngroups, 32), 32, 0, k, t, t + k, t + 2 * (size_t)k, dg, ngroups, dout);
    uint32_t hbad = 0;
    SONIC_CUDA(cudaMemcpyAsync(&hbad, bad, 4, cudaMemcpyDeviceToHost, cx.stream));
    SONIC_CUDA(cudaMemcpyAsync(out48, dout, (size_t)(2 + ngroups) * 48, cudaMemcpyDeviceToHost, cx.stream));
    SONIC_CUDA(cudaStreamSynchronize(cx.stream));
    if (hbad) return fail(SONIC_ERR_INVALID_ARG, "a G1 encoding is malformed or not on the curve");
    return SONIC_OK;
}

}  // namespace sonic
