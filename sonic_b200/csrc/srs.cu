// SRS generation: the fixed-base batch scalar multiplication behind `SRS.new`.
//
// Reference: src/Sonic/SRS.hs:27-43 builds, per element and independently,
//   gNegativeX[i-1]      = gen * x^-i        i = 1..d     (:33)
//   gPositiveX[i]        = gen * x^i         i = 0..d     (:34)
//   gNegativeAlphaX[i-1] = gen * alpha x^-i  i = 1..d     (:37)
//   gPositiveAlphaX[i-1] = gen * alpha x^i   i = 1..d     (:38-39, g^alpha deliberately absent)
// with a square-and-multiply `pow` and a double-and-add `mul gen` for every element.
//
// Here: the scalars x^k come from per-thread runs of a geometric progression; the points
// from a windowed fixed-base table T[j][v] = v * 2^(w j) * G (built on the device, level by
// level) so that one SRS element costs ceil(255/w) mixed additions and no doubling; a batched
// (Montgomery trick) inversion brings everything to affine for the MSM's 96-byte gathers.
//
// Device layout: one array of affine points indexed by exponent,
//   points[family * (2d+1) + (k + d)],  k in [-d, d],  family 0 = plain, 1 = alpha;
// the alpha slot k = 0 holds the infinity marker (0,0) and is never referenced by a job.
#include <algorithm>

#include "internal.h"
#include "g1io.cuh"

namespace sonic {

constexpr int SRS_RUN = 32;

// params[0] = x, params[1] = alpha (canonical) -> mont[0] = x, mont[1] = 1/x, mont[2] = alpha (Montgomery)
__global__ void k_srs_params(const Fr* __restrict__ canon, Fr* __restrict__ mont) {
    if (threadIdx.x == 0) {
        Fr x = fp_to_mont(canon[0]);
        mont[0] = x;
        mont[1] = fp_inv(x);
        mont[2] = fp_to_mont(canon[1]);
    }
}

// scal[0][k+d] = x^k, scal[1][k+d] = alpha x^k, Montgomery form, k in [-d, d]
__global__ void __launch_bounds__(128) k_srs_scalars(const Fr* __restrict__ mont, uint64_t d, Fr* __restrict__ scal) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const bool negative = blockIdx.y == 1;
    const uint64_t e0 = t * SRS_RUN + (negative ? 1 : 0);
    if (e0 > d) return;
    const Fr base = mont[negative ? 1 : 0];
    const Fr alpha = mont[2];
    Fr v = fp_pow_u64(base, e0);
    const uint64_t stride = 2 * d + 1;
    for (int i = 0; i < SRS_RUN; ++i) {
        const uint64_t e = e0 + i;
        if (e > d) break;
        const uint64_t idx = negative ? d - e : d + e;
        scal[idx] = v;
        scal[stride + idx] = fp_mul(v, alpha);
        v = fp_mul(v, base);
    }
}

// consts[j] = 2^(c j) (Montgomery), j < W
__global__ void k_level_consts(int c, int W, Fr* __restrict__ consts) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= W) return;
    Fr v = Fr::one();
    for (int i = 0; i < c * j; ++i) v = fp_dbl(v);
    consts[j] = v;
}

// out[i] = canonical(scal[i] * factor)
__global__ void __launch_bounds__(256) k_level_scalars(const Fr* __restrict__ scal, const Fr* __restrict__ factor, uint64_t n, Fr* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = fp_from_mont(fp_mul(scal[i], *factor));
}

// T[j][0] = inf, T[j][1] = 2^(w j) G
__global__ void k_tbl_bases(G1XYZZ* __restrict__ T, int w, int Wt) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= Wt) return;
    G1XYZZ b = G1XYZZ::from_affine(G1Affine::gen());
    for (int i = 0; i < w * j; ++i) b = g1_dbl(b);
    G1XYZZ* row = T + ((size_t)j << w);
    store_xyzz(row, G1XYZZ::inf());
    store_xyzz(row + 1, b);
}

// level l >= 1: T[j][2^l + i] = T[j][i] + 2*T[j][2^(l-1)],  i in [0, 2^l)
__global__ void __launch_bounds__(128) k_tbl_level(G1XYZZ* __restrict__ T, int w, int Wt, int level) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const int j = blockIdx.y;
    if (i >= (1u << level)) return;
    G1XYZZ* row = T + ((size_t)j << w);
    G1XYZZ top = g1_dbl(load_xyzz(row + (1u << (level - 1))));
    if (i) g1_add(top, load_xyzz(row + i));
    store_xyzz(row + (1u << level) + i, top);
}

// XYZZ -> affine with one inversion per BATCH points (Montgomery's trick)
constexpr int AFF_BATCH = 16;
__global__ void __launch_bounds__(128) k_batch_affine(const G1XYZZ* __restrict__ in, G1Affine* __restrict__ out, uint64_t n) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    const uint64_t first = t * AFF_BATCH;
    if (first >= n) return;
    const int cnt = (n - first < (uint64_t)AFF_BATCH) ? (int)(n - first) : AFF_BATCH;
    Fq pre[AFF_BATCH];
    Fq run = Fq::one();
    for (int m = 0; m < cnt; ++m) {
        Fq z = in[first + m].zzz;
        if (z.is_zero()) z = Fq::one();  // infinity: keep the product invertible
        run = fp_mul(run, z);
        pre[m] = run;
    }
    Fq inv = fp_inv(run);
    for (int m = cnt - 1; m >= 0; --m) {
        G1XYZZ p = load_xyzz(in + first + m);
        G1Affine a;
        if (p.is_inf()) {
            a = G1Affine::inf();
        } else {
            Fq zi = m ? fp_mul(inv, pre[m - 1]) : inv;  // 1/ZZZ
            inv = fp_mul(inv, p.zzz);
            Fq tz = fp_mul(p.zz, zi);                   // 1/Z
            Fq zzi = fp_sqr(tz);                        // 1/ZZ
            a.x = fp_mul(p.x, zzi);
            a.y = fp_mul(p.y, zi);
        }
        out[first + m] = a;
    }
}

// one SRS element per thread: sum_j T[j][digit_j(s)]
__global__ void __launch_bounds__(128, 3)
k_fixed_base(const Fr* __restrict__ scal, const G1Affine* __restrict__ T, int w, int Wt, uint64_t n,
             uint64_t hole, G1XYZZ* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1XYZZ acc = G1XYZZ::inf();
    if (i != hole) {
        uint32_t s[9];
        const Fr v = scal[i];
#pragma unroll
        for (int k = 0; k < 8; ++k) s[k] = v.l[k];
        s[8] = 0;
        for (int j = 0; j < Wt; ++j) {
            const int pos = j * w, limb = pos >> 5, sh = pos & 31;
            const uint64_t two = ((uint64_t)s[limb + 1] << 32) | s[limb];
            const uint32_t dgt = (uint32_t)(two >> sh) & ((1u << w) - 1);
            if (dgt) g1_madd(acc, load_affine(T + ((size_t)j << w) + dgt));
        }
    }
    store_xyzz(out + i, acc);
}

static int srs_table_bits(uint64_t npoints) {
    int best = 4;
    double bc = 1e300;
    for (int w = 4; w <= 16; ++w) {
        double Wt = (255 + w - 1) / w;
        double cost = double(npoints) * Wt + 2.6 * Wt * double(1u << w);
        if (cost < bc) { bc = cost; best = w; }
    }
    return best;
}

// Generates elements [first, first + count) of the resident point array taken as ONE flat array of
// levels x npts points (flat index = level * npts + family * (2d+1) + k + d); everything when first = 0 and
// count covers it.  A multi-GPU runtime gives every device one equal flat slice -- a slice may span two or
// three levels -- and all-gathers once (capi.cu).  d_canon: x, alpha canonical (2 Fr) in device memory.
// Level 0 is the SRS itself; with pre_c > 0, level j holds the same points times 2^(pre_c j), obtained from the
// same fixed-base table with the scalar multiplied by 2^(pre_c j) mod r.  The scalars x^k, alpha x^k are
// cheap (two Fr products per element against 16 x 3000 LMAC for the point) and are computed in full on
// every device.
void srs_generate(Ctx& cx, uint64_t d, const Fr* d_canon, G1Affine* d_points, int pre_c, uint64_t first, uint64_t count,
                  void* d_g2_points) {
    Arena& ar = cx.arena;
    const uint64_t stride = 2 * d + 1, npts = 2 * stride;
    const int levels = pre_c > 0 ? (255 + pre_c - 1) / pre_c : 1;
    const uint64_t total = npts * (uint64_t)levels;
    if (first > total) first = total;
    if (count > total - first) count = total - first;
    Fr* mont = ar.get<Fr>(3);
    SONIC_LAUNCH(k_srs_params, 1, 32, 0, d_canon, mont);
    Fr* scal_m = ar.get<Fr>(npts);
    SONIC_LAUNCH(k_srs_scalars, dim3(div_up(d / SRS_RUN + 1, 128), 2), 128, 0, mont, d, scal_m);
    if (count) {
        Fr* consts = ar.get<Fr>(levels);
        SONIC_LAUNCH(k_level_consts, div_up(levels, 32), 32, 0, pre_c > 0 ? pre_c : 1, levels, consts);
        const int w = srs_table_bits(count);
        const int Wt = (255 + w - 1) / w;
        const size_t tsize = (size_t)Wt << w;
        G1XYZZ* Tx = ar.get<G1XYZZ>(tsize);
        G1Affine* Ta = ar.get<G1Affine>(tsize);
        SONIC_LAUNCH(k_tbl_bases, div_up(Wt, 32), 32, 0, Tx, w, Wt);
        for (int l = 1; l < w; ++l)
            SONIC_LAUNCH(k_tbl_level, dim3(div_up(1u << l, 128), (unsigned)Wt), 128, 0, Tx, w, Wt, l);
        SONIC_LAUNCH(k_batch_affine, div_up(div_up(tsize, AFF_BATCH), 128), 128, 0, Tx, Ta, (uint64_t)tsize);
        const uint64_t piece_max = std::min<uint64_t>(count, npts);
        Fr* scal = ar.get<Fr>(piece_max);
        G1XYZZ* px = ar.get<G1XYZZ>(piece_max);
        const uint64_t hole_abs = stride + d;   // alpha family, exponent 0: g^alpha is not part of the SRS
        for (int j = 0; j < levels; ++j) {
            // the part of level j inside the flat slice
            const uint64_t lo = std::max<uint64_t>(first, (uint64_t)j * npts), hi = std::min<uint64_t>(first + count, (uint64_t)(j + 1) * npts);
            if (hi <= lo) continue;
            const uint64_t a = lo - (uint64_t)j * npts, n = hi - lo;
            const uint64_t hole = hole_abs >= a && hole_abs < a + n ? hole_abs - a : ~0ull;
            SONIC_LAUNCH(k_level_scalars, div_up(n, 256), 256, 0, scal_m + a, consts + j, n, scal);
            SONIC_LAUNCH(k_fixed_base, div_up(n, 128), 128, 0, scal, Ta, w, Wt, n, hole, px);
            SONIC_LAUNCH(k_batch_affine, div_up(div_up(n, AFF_BATCH), 128), 128, 0, px, d_points + lo, n);
        }
    }
    if (d_g2_points) srs_generate_g2(cx, scal_m, npts, d_g2_points);
}

// ---- window tables restricted to the exponent ranges a circuit size touches -----------------------------
// `prove` at n gates reads g^{x^k} for k in [-4n-8, 3n] and g^{alpha x^k} for k in [-4n-8, 3n] and
// [d-3n-4, d] only (the shapes of Protocol.hs:63,73,79-81 and Signature.hs:42-63): 17n points out of
// 4d+1.  Their multiples 2^(c j) P come from c doublings of the level below -- no trapdoor, so a loaded
// SRS can have them too -- and cost about what the fixed-base path pays per point (c x 7 products).
__global__ void __launch_bounds__(128) k_table_level(const G1Affine* __restrict__ prev, uint32_t n, int c, G1XYZZ* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    G1XYZZ p = g1_mdbl(load_affine(prev + i));
    for (int k = 1; k < c; ++k) p = g1_dbl(p);
    store_xyzz(out + i, p);
}

void tables_build(Ctx& cx, const SrsRep& srs, const TableRange* ranges, int nranges, int c, RestrictedTables* out) {
    tables_free(out);
    uint64_t size = 0;
    for (int i = 0; i < nranges; ++i) {
        out->range[i] = ranges[i];
        out->range[i].offset = (uint32_t)size;
        size += (uint64_t)(ranges[i].hi - ranges[i].lo + 1);
    }
    const int W = (255 + c - 1) / c;
    if (size == 0 || size * (uint64_t)W >= (1ull << 31)) return;
    SONIC_CUDA(cudaMalloc((void**)&out->points, size * (size_t)W * sizeof(G1Affine)));
    out->nranges = nranges;
    out->size = (uint32_t)size;
    out->tables.c = c;
    out->tables.W = W;
    out->tables.stride = (uint32_t)size;
    for (int i = 0; i < nranges; ++i)
        SONIC_CUDA(cudaMemcpyAsync(out->points + out->range[i].offset, srs.points + srs.index(ranges[i].family, ranges[i].lo),
                                   (size_t)(ranges[i].hi - ranges[i].lo + 1) * sizeof(G1Affine), cudaMemcpyDeviceToDevice, cx.stream));
    G1XYZZ* px = cx.arena.get<G1XYZZ>(size);
    for (int j = 1; j < W; ++j) {
        SONIC_LAUNCH(k_table_level, div_up(size, 128), 128, 0, out->points + (size_t)(j - 1) * size, (uint32_t)size, c, px);
        SONIC_LAUNCH(k_batch_affine, div_up(div_up(size, AFF_BATCH), 128), 128, 0, px, out->points + (size_t)j * size, size);
    }
}

void tables_free(RestrictedTables* t) {
    if (t->points) cudaFree(t->points);
    *t = RestrictedTables();
}

}  // namespace sonic
