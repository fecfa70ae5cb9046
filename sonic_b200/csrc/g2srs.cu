// G2 fixed-base batch for the h-vectors of the SRS.
// Own translation unit: the Fq2 curve code is large and off the hot path, so it is compiled with
// out-of-line field multiplications and a lower ptxas optimisation level to keep build times sane.
#include "internal.h"
#include "g2.cuh"

namespace sonic {

// ---- G2 fixed-base batch (not on the prover's path) -----------------------
// 8-bit windows: T2[j][v] = v * 2^(8j) * H, v < 256, affine.  One thread per window builds its row
// (a few hundred sequential point operations: a one-time cost), with one Fq2 inversion per row.
constexpr int G2_W = 8;
constexpr int G2_WT = 32;  // ceil(255 / 8)

__global__ void k_g2_table(G2Affine* __restrict__ T, G2XYZZ* __restrict__ scratch, Fq2* __restrict__ pre) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= G2_WT) return;
    G2XYZZ b = G2XYZZ::from_affine(G2Affine::gen());
    for (int i = 0; i < G2_W * j; ++i) b = g2_dbl(b);
    const G2Affine base = g2_to_affine(b);
    G2Affine* row = T + ((size_t)j << G2_W);
    G2XYZZ* sx = scratch + ((size_t)j << G2_W);
    Fq2* pp = pre + ((size_t)j << G2_W);
    row[0] = G2Affine::inf();
    G2XYZZ acc = G2XYZZ::inf();
    Fq2 run = Fq2::one();
    for (int v = 1; v < (1 << G2_W); ++v) {
        g2_madd(acc, base);
        sx[v] = acc;
        run = f2_mul(run, acc.zzz);  // v * base is never infinity: the order of H is the 255-bit prime r
        pp[v] = run;
    }
    Fq2 inv = f2_inv(run);
    for (int v = (1 << G2_W) - 1; v >= 1; --v) {
        const G2XYZZ p = sx[v];
        const Fq2 zi = v > 1 ? f2_mul(inv, pp[v - 1]) : inv;
        inv = f2_mul(inv, p.zzz);
        const Fq2 tz = f2_mul(p.zz, zi);
        G2Affine a;
        a.x = f2_mul(p.x, f2_sqr(tz));
        a.y = f2_mul(p.y, zi);
        row[v] = a;
    }
}

constexpr int G2_AFF_BATCH = 8;
__global__ void __launch_bounds__(64) k_g2_fixed_base(const Fr* __restrict__ scal_m, const G2Affine* __restrict__ T, uint64_t n,
                                                     G2Affine* __restrict__ out) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t first = t * G2_AFF_BATCH;
    if (first >= n) return;
    const int cnt = (n - first < (uint64_t)G2_AFF_BATCH) ? (int)(n - first) : G2_AFF_BATCH;
    G2XYZZ pts[G2_AFF_BATCH];
    Fq2 pre[G2_AFF_BATCH];
    Fq2 run = Fq2::one();
    for (int m = 0; m < cnt; ++m) {
        const Fr s = fp_from_mont(scal_m[first + m]);
        G2XYZZ acc = G2XYZZ::inf();
        for (int j = 0; j < G2_WT; ++j) {
            const uint32_t dgt = (s.l[j >> 2] >> ((j & 3) * 8)) & 0xffu;
            if (dgt) g2_madd(acc, T[((size_t)j << G2_W) + dgt]);
        }
        pts[m] = acc;
        run = f2_mul(run, acc.is_inf() ? Fq2::one() : acc.zzz);
        pre[m] = run;
    }
    Fq2 inv = f2_inv(run);
    for (int m = cnt - 1; m >= 0; --m) {
        const G2XYZZ p = pts[m];
        G2Affine a = G2Affine::inf();
        if (!p.is_inf()) {
            const Fq2 zi = m ? f2_mul(inv, pre[m - 1]) : inv;
            inv = f2_mul(inv, p.zzz);
            const Fq2 tz = f2_mul(p.zz, zi);
            a.x = f2_mul(p.x, f2_sqr(tz));
            a.y = f2_mul(p.y, zi);
        }
        out[first + m] = a;
    }
}

__global__ void k_g2_compress(const G2Affine* __restrict__ pts, uint64_t n, uint8_t* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) g2_compress(pts[i], out + i * 96);
}

void g2_compress_range(Ctx& cx, const void* d_g2_points, uint64_t first, uint64_t count, uint8_t* d_out) {
    (void)cx;
    SONIC_LAUNCH(k_g2_compress, div_up(count, 64), 64, 0, (const G2Affine*)d_g2_points + first, count, d_out);
}

size_t g2_point_bytes() { return sizeof(G2Affine); }


// the h-vectors: same scalars x^k, alpha x^k against the G2 generator (alpha x^0 included: SRS.hs:41)
void srs_generate_g2(Ctx& cx, const Fr* scal_m, uint64_t npts, void* d_g2_points) {
    Arena& ar = cx.arena;
    const size_t tsz = (size_t)G2_WT << G2_W;
    G2Affine* T2 = ar.get<G2Affine>(tsz);
    G2XYZZ* sx = ar.get<G2XYZZ>(tsz);
    Fq2* pre = ar.get<Fq2>(tsz);
    SONIC_LAUNCH(k_g2_table, 1, 32, 0, T2, sx, pre);
    SONIC_LAUNCH(k_g2_fixed_base, div_up(div_up(npts, G2_AFF_BATCH), 64), 64, 0, scal_m, T2, npts, (G2Affine*)d_g2_points);
}

}  // namespace sonic
