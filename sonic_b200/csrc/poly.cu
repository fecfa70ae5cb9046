// Fr-side kernels: the Laurent-polynomial arithmetic that feeds the commitments.
//
// Dense vectors stand in for the reference's sparse `VLaurent Fr`: a vector plus the
// exponent `lo` of its first slot.  Zero slots are exactly the terms the sparse form does
// not hold.  All values on the device are in Montgomery form unless a name says "canon".
//
//   power tables      base^k                      (Utils.hs:18 `pow`, CommitmentScheme.hs:43 `eval`)
//   open              f(z) and (f(X)-f(z))/(X-z)  (CommitmentScheme.hs:43-44)
//   NTT               radix-2 over Fr, 2-adicity 32, for t(X,y) (Constraints.hs:61)
#include <vector>

#include "internal.h"

namespace sonic {

// ---- conversions ------------------------------------------------------------------------
__global__ void k_fr_to_mont(const Fr* __restrict__ in, Fr* __restrict__ out, uint64_t n, uint32_t* __restrict__ bad) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fr v = in[i];
    // canonical check: v < r
    uint32_t t = Chain::sub_cc(v.l[0], FrParams::P(0));
#pragma unroll
    for (int k = 1; k < 8; ++k) t = Chain::subc_cc(v.l[k], FrParams::P(k));
    if (Chain::subc(0, 0) == 0 && bad) atomicExch(bad, 1u);  // no borrow: v >= r
    (void)t;
    out[i] = fp_to_mont(v);
}

__global__ void k_fr_from_mont(const Fr* __restrict__ in, Fr* __restrict__ out, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = fp_from_mont(in[i]);
}

// out[i] = 1/in[i]  (0 -> 0), a few elements only
__global__ void k_fr_inv(const Fr* __restrict__ in, Fr* __restrict__ out, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = fp_inv(in[i]);
}

// ---- power tables: tab[t][k] = base[t]^k, k in [0, len) -------------------------------------
// a run of 64 powers per thread: the 64-bit-exponent start (about 27 products at k ~ 2^18) is then a
// third of the run instead of twice it; the tables are compute-bound (9 M entries per proof)
constexpr int POW_RUN = 64;
__global__ void __launch_bounds__(128) k_pow_tables(const Fr* __restrict__ bases, Fr* __restrict__ tab, uint64_t len, uint64_t stride,
                                                    const uint32_t* __restrict__ sel) {
    const uint64_t k0 = (blockIdx.x * (uint64_t)blockDim.x + threadIdx.x) * POW_RUN;
    if (k0 >= len) return;
    const uint32_t row = sel ? sel[blockIdx.y] : blockIdx.y;
    const Fr b = bases[row];
    Fr v = fp_pow_u64(b, k0);
    Fr* o = tab + (size_t)row * stride + k0;
    for (int i = 0; i < POW_RUN && k0 + i < len; ++i) {
        o[i] = v;
        v = fp_mul(v, b);
    }
}

// ---- warp / block reductions over Fr -------------------------------------------------------
SONIC_D Fr fr_shfl_down(const Fr& v, int o) {
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = __shfl_down_sync(0xffffffffu, v.l[k], o);
    return r;
}

// sum over the block; result valid in thread 0.  smem: one Fr per warp.
template <int THREADS>
SONIC_D Fr block_sum_fr(Fr v, Fr* smem) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Fr y = fr_shfl_down(v, o);
        v = fp_add(v, y);  // lanes beyond the edge add garbage that is never read by lane 0
    }
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < THREADS / 32; ++w) v = fp_add(v, smem[w]);
    }
    return v;
}

// ---- openPoly -----------------------------------------------------------------------------
// One "open job": f over exponents [lo, lo+len) (lo <= 0 < lo+len), evaluation point z given
// through its tables pz[k] = z^k, pzi[k] = z^-k.  With h_k = f_k z^k and H = sum_k h_k:
//   f(z) = z^lo * H,
//   quotient q_(k-1) = z^-k * sum_(m>=k) h'_m,   h' = h with H subtracted at slot -lo
// which is the synthetic division (f(X) - f(z)) / (X - z) written as a suffix sum.

constexpr int OPEN_THREADS = 256;
constexpr int OPEN_ITEMS = 4;
constexpr int OPEN_TILE = OPEN_THREADS * OPEN_ITEMS;

// pass 1: per-tile sums of h
__global__ void __launch_bounds__(OPEN_THREADS) k_open_partial(const OpenJob* __restrict__ jobs, Fr* __restrict__ partial, uint32_t tiles_stride) {
    __shared__ Fr smem[OPEN_THREADS / 32];
    const OpenJob jb = jobs[blockIdx.y];
    if (blockIdx.x * OPEN_TILE >= jb.len) return;
    Fr s = Fr::zero();
    // consecutive lanes read consecutive coefficients (the order of a sum is free)
#pragma unroll
    for (int i = 0; i < OPEN_ITEMS; ++i) {
        const uint32_t k = blockIdx.x * OPEN_TILE + i * OPEN_THREADS + threadIdx.x;
        if (k < jb.len) s = fp_add(s, fp_mul(jb.f[k], jb.pz[k]));
    }
    s = block_sum_fr<OPEN_THREADS>(s, smem);
    if (threadIdx.x == 0) partial[(size_t)blockIdx.y * tiles_stride + blockIdx.x] = s;
}

// pass 2 (one warp per job): H, f(z), and the exclusive suffix sums of the tile sums (a warp-wide
// scan per 32 tiles from the top down: a 7n+9-long opening has 448 tiles, and one thread walking them
// took 0.19 ms of a proof)
SONIC_D Fr fr_shfl_idx(const Fr& v, int src) {
    Fr r;
#pragma unroll
    for (int k = 0; k < 8; ++k) r.l[k] = __shfl_sync(0xffffffffu, v.l[k], src);
    return r;
}

__global__ void __launch_bounds__(32) k_open_tiles(const OpenJob* __restrict__ jobs, Fr* __restrict__ partial, uint32_t tiles_stride, Fr* __restrict__ Hout) {
    const OpenJob jb = jobs[blockIdx.x];
    const int lane = threadIdx.x;
    Fr* p = partial + (size_t)blockIdx.x * tiles_stride;
    const uint32_t tiles = (jb.len + OPEN_TILE - 1) / OPEN_TILE;
    Fr h = Fr::zero();
    for (uint32_t t = lane; t < tiles; t += 32) h = fp_add(h, p[t]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h = fp_add(h, fr_shfl_down(h, o));  // lane 0 ends with the sum of all lanes
    const Fr H = fr_shfl_idx(h, 0);
    if (lane == 0) {
        Hout[blockIdx.x] = H;
        Fr fz;
        if (jb.z_is_zero) fz = jb.f[-jb.lo];           // only reached with lo == 0 (host rejects lo < 0 at z = 0)
        else fz = jb.lo < 0 ? fp_mul(H, jb.pzi[-jb.lo]) : fp_mul(H, jb.pz[jb.lo]);
        *jb.value_canon = fp_from_mont(fz);
    }
    if (tiles == 0) return;
    // exclusive suffix over tiles of h' (H removed from the tile that holds slot -lo)
    const uint32_t ct = (uint32_t)(-jb.lo) / OPEN_TILE;
    Fr run = Fr::zero();   // sum of the tiles above the current group of 32
    for (int64_t base = (int64_t)((tiles - 1) / 32) * 32; base >= 0; base -= 32) {
        const uint32_t t = (uint32_t)base + lane;
        Fr v = t < tiles ? p[t] : Fr::zero();
        if (t == ct) v = fp_sub(v, H);
        Fr s = v;   // inclusive suffix inside the group
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const Fr y = fr_shfl_down(s, o);
            if (lane + o < 32) s = fp_add(s, y);
        }
        if (t < tiles) p[t] = fp_add(fp_sub(s, v), run);
        run = fp_add(run, fr_shfl_idx(s, 0));
    }
}

// pass 3: in-tile suffix scan, scale by z^-k, emit canonical quotient coefficients
__global__ void __launch_bounds__(OPEN_THREADS) k_open_quotient(const OpenJob* __restrict__ jobs, const Fr* __restrict__ partial, uint32_t tiles_stride, const Fr* __restrict__ Hin) {
    __shared__ Fr wsum[OPEN_THREADS / 32];
    const OpenJob jb = jobs[blockIdx.y];
    if (jb.q_canon == nullptr || blockIdx.x * OPEN_TILE >= jb.len) return;
    const uint32_t base = blockIdx.x * OPEN_TILE + threadIdx.x * OPEN_ITEMS;
    if (jb.z_is_zero) {
        // X - 0 divides f - f(0): the quotient is a shift
#pragma unroll
        for (int i = 0; i < OPEN_ITEMS; ++i) {
            const uint32_t k = base + i;
            if (k >= 1 && k < jb.len) jb.q_canon[k - 1] = fp_from_mont(jb.f[k]);
        }
        return;
    }
    const Fr H = Hin[blockIdx.y];
    const uint32_t c = (uint32_t)(-jb.lo);
    // the products h_k with consecutive lanes on consecutive coefficients (whole sectors per warp; a thread reading its
    // four consecutive 32-byte values itself made every load touch 32 lines), handed over through shared memory to the
    // thread that owns four consecutive slots of the scan; the scaled sums go back the same way for the stores
    __shared__ Fr stage[OPEN_TILE];
    const uint32_t tile0 = blockIdx.x * OPEN_TILE;
#pragma unroll
    for (int i = 0; i < OPEN_ITEMS; ++i) {
        const uint32_t j = i * OPEN_THREADS + threadIdx.x, k = tile0 + j;
        Fr v = Fr::zero();
        if (k < jb.len) {
            v = fp_mul(jb.f[k], jb.pz[k]);
            if (k == c) v = fp_sub(v, H);
        }
        stage[j] = v;
    }
    __syncthreads();
    Fr h[OPEN_ITEMS];
#pragma unroll
    for (int i = 0; i < OPEN_ITEMS; ++i) h[i] = stage[threadIdx.x * OPEN_ITEMS + i];
    // thread-local inclusive suffix
#pragma unroll
    for (int i = OPEN_ITEMS - 2; i >= 0; --i) h[i] = fp_add(h[i], h[i + 1]);
    // exclusive suffix across the block of the thread totals h[0]
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    Fr inc = h[0];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        Fr y = fr_shfl_down(inc, o);
        if (lane + o < 32) inc = fp_add(inc, y);
    }
    if (lane == 0) wsum[wid] = inc;
    __syncthreads();
    Fr off = partial[(size_t)blockIdx.y * tiles_stride + blockIdx.x];
    for (int w = OPEN_THREADS / 32 - 1; w > wid; --w) off = fp_add(off, wsum[w]);
    off = fp_add(off, fp_sub(inc, h[0]));  // lanes above in this warp
    __syncthreads();   // every thread has taken its products out of the staging buffer
#pragma unroll
    for (int i = 0; i < OPEN_ITEMS; ++i) stage[threadIdx.x * OPEN_ITEMS + i] = fp_add(h[i], off);
    __syncthreads();
#pragma unroll
    for (int i = 0; i < OPEN_ITEMS; ++i) {
        const uint32_t j = i * OPEN_THREADS + threadIdx.x, k = tile0 + j;
        if (k >= 1 && k < jb.len) jb.q_canon[k - 1] = fp_from_mont(fp_mul(stage[j], jb.pzi[k]));
    }
}

// ---- NTT ------------------------------------------------------------------------------------
// Radix-2 over Fr (2-adicity 32), in place.  Forward = decimation in frequency (natural in,
// bit-reversed out); inverse = decimation in time on the bit-reversed data (natural out), so a
// cyclic convolution needs no permutation.  tw[k] = omega^k, k < L/2 (omega^-k for the inverse).
//
// Butterflies are staged through shared memory: one launch runs `nb` consecutive stages on
// tiles of 2^nb x C elements (<= 1024 elements = 32 KB), where the 2^nb rows are the index
// bits those stages touch and the C columns are adjacent low-order indices, so that every
// global access is a run of C consecutive 32-byte elements.  L = 2^19 takes 3 launches per
// transform instead of 19 (stages 6 + 6 + 7 forward; 10 + 6 + 3 inverse).
constexpr int NTT_TILE_LOG = 10;
constexpr int NTT_THREADS = 512;

template <bool INVERSE>
__global__ void __launch_bounds__(NTT_THREADS) k_ntt_pass(Fr* __restrict__ a, const Fr* __restrict__ tw, uint32_t logL,
                                                          uint32_t lowbit, uint32_t nb, uint32_t logC,
                                                          const Fr* __restrict__ scale) {
    __shared__ Fr tile[1 << NTT_TILE_LOG];
    const uint32_t C = 1u << logC;
    const uint32_t tile_elems = 1u << (nb + logC);
    // block -> (hi, lg): bits above the staged ones, and the group of C low-order columns
    const uint32_t lg_count_log = lowbit - logC;
    const uint32_t hi = blockIdx.x >> lg_count_log;
    const uint32_t lg = blockIdx.x & ((1u << lg_count_log) - 1);
    const uint32_t gbase = (hi << (lowbit + nb)) | (lg << logC);
    for (uint32_t t = threadIdx.x; t < tile_elems; t += NTT_THREADS) {
        const uint32_t k = t >> logC, c = t & (C - 1);
        tile[t] = a[gbase | (k << lowbit) | c];
    }
    __syncthreads();
    const uint32_t nbf = tile_elems >> 1;
    for (uint32_t ls = 0; ls < nb; ++ls) {
        // forward: large strides first (row bit nb-1-ls, global stage s = logL-lowbit-nb+ls, half = L >> (s+1));
        // inverse: small strides first (row bit ls, global stage s = lowbit+ls, half = 1 << s)
        const uint32_t kbit = INVERSE ? ls : nb - 1 - ls;
        const uint32_t halfk = 1u << kbit;
        const uint32_t s = INVERSE ? lowbit + ls : logL - lowbit - nb + ls;
        const uint32_t ghalf_log = INVERSE ? s : logL - 1 - s;
        const uint32_t tw_shift = INVERSE ? logL - 1 - s : s;
        for (uint32_t bf = threadIdx.x; bf < nbf; bf += NTT_THREADS) {
            const uint32_t c = bf & (C - 1), kp = bf >> logC;
            const uint32_t jk = kp & (halfk - 1);
            const uint32_t klo = ((kp - jk) << 1) + jk;
            const uint32_t tlo = (klo << logC) | c, thi = ((klo + halfk) << logC) | c;
            const uint32_t glo = gbase | (klo << lowbit) | c;
            const uint32_t j = glo & ((1u << ghalf_log) - 1);
            const Fr w = tw[(size_t)j << tw_shift];
            const Fr u = tile[tlo];
            if (INVERSE) {
                const Fr v = fp_mul(tile[thi], w);
                tile[tlo] = fp_add(u, v);
                tile[thi] = fp_sub(u, v);
            } else {
                const Fr v = tile[thi];
                tile[tlo] = fp_add(u, v);
                tile[thi] = fp_mul(fp_sub(u, v), w);
            }
        }
        __syncthreads();
    }
    for (uint32_t t = threadIdx.x; t < tile_elems; t += NTT_THREADS) {
        const uint32_t k = t >> logC, c = t & (C - 1);
        Fr v = tile[t];
        if (scale) v = fp_mul(v, *scale);
        a[gbase | (k << lowbit) | c] = v;
    }
}

__global__ void __launch_bounds__(256) k_fr_mul_pointwise(Fr* __restrict__ a, const Fr* __restrict__ b, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = fp_mul(a[i], b[i]);
}

__global__ void __launch_bounds__(256) k_fr_scale(Fr* __restrict__ a, const Fr* __restrict__ s, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = fp_mul(a[i], *s);
}

// omega_L = ROOT32^(2^(32-logL)); params[0] = omega, [1] = omega^-1, [2] = 1/L
__global__ void k_ntt_params(uint32_t logL, Fr* __restrict__ params) {
    if (threadIdx.x != 0) return;
    Fr w, wi;
    for (int k = 0; k < 8; ++k) { w.l[k] = FrParams::ROOT32_M(k); wi.l[k] = FrParams::ROOT32_INV_M(k); }
    for (uint32_t s = logL; s < 32; ++s) { w = fp_sqr(w); wi = fp_sqr(wi); }
    params[0] = w;
    params[1] = wi;
    Fr L = Fr::zero();
    L.l[0] = 1u << logL;  // logL <= 31
    params[2] = fp_inv(fp_to_mont(L));
}


// Twiddle tables depend only on the length: they are built once per length and stay resident
// (outside the per-call arena) for the life of the context.
NttPlan ntt_prepare(Ctx& cx, uint32_t logL) {
    auto it = cx.ntt_cache.find(logL);
    if (it != cx.ntt_cache.end()) {
        NttPlan p;
        p.logL = logL;
        p.params = (Fr*)it->second;
        p.tw = p.params + 4;
        p.twi = p.tw + (logL ? (1ull << (logL - 1)) : 1);
        return p;
    }
    NttPlan p;
    p.logL = logL;
    const uint64_t half = logL ? (1ull << (logL - 1)) : 1;
    Fr* mem = nullptr;
    SONIC_CUDA(cudaMalloc((void**)&mem, (4 + 2 * half) * sizeof(Fr)));
    cx.ntt_cache[logL] = mem;
    p.params = mem;
    p.tw = mem + 4;
    p.twi = p.tw + half;
    SONIC_LAUNCH(k_ntt_params, 1, 32, 0, logL, p.params);
    SONIC_LAUNCH(k_pow_tables, dim3(div_up(div_up(half, POW_RUN), 128), 2), 128, 0, p.params, p.tw, half, half, (const uint32_t*)nullptr);
    return p;
}

// Splits the logL stages into passes: one contiguous pass of up to NTT_TILE_LOG stages on the
// low bits, the remaining (strided) stages in groups of up to 6 with C = 1024 >> nb columns.
struct NttPass { uint32_t lowbit, nb, logC; };
static std::vector<NttPass> ntt_passes(uint32_t logL) {
    std::vector<NttPass> v;  // ordered from the low bits up
    const uint32_t nlow = logL < (uint32_t)NTT_TILE_LOG ? logL : (uint32_t)NTT_TILE_LOG;
    v.push_back({0u, nlow, 0u});
    uint32_t bit = nlow;
    while (bit < logL) {
        uint32_t nb = logL - bit < 6 ? logL - bit : 6;
        uint32_t logC = NTT_TILE_LOG - nb;
        if (logC > bit) logC = bit;
        v.push_back({bit, nb, logC});
        bit += nb;
    }
    return v;
}

void ntt_forward(const NttPlan& p, Fr* a) {
    if (p.logL == 0) return;
    std::vector<NttPass> ps = ntt_passes(p.logL);
    for (size_t i = ps.size(); i-- > 0;) {  // decimation in frequency: high bits first
        const NttPass& q = ps[i];
        const uint32_t blocks = 1u << (p.logL - q.nb - q.logC);
        SONIC_LAUNCH(k_ntt_pass<false>, blocks, NTT_THREADS, 0, a, p.tw, p.logL, q.lowbit, q.nb, q.logC, (const Fr*)nullptr);
    }
}

void ntt_inverse(const NttPlan& p, Fr* a) {
    if (p.logL == 0) return;
    std::vector<NttPass> ps = ntt_passes(p.logL);
    for (size_t i = 0; i < ps.size(); ++i) {  // decimation in time: low bits first; 1/L folded into the last store
        const NttPass& q = ps[i];
        const uint32_t blocks = 1u << (p.logL - q.nb - q.logC);
        SONIC_LAUNCH(k_ntt_pass<true>, blocks, NTT_THREADS, 0, a, p.twi, p.logL, q.lowbit, q.nb, q.logC,
                     i + 1 == ps.size() ? (const Fr*)(p.params + 2) : (const Fr*)nullptr);
    }
}

// ---- host launchers ---------------------------------------------------------------------------
void fr_to_mont(Ctx& cx, const Fr* in, Fr* out, uint64_t n, uint32_t* bad_flag) {
    (void)cx;
    if (n) SONIC_LAUNCH(k_fr_to_mont, div_up(n, 256), 256, 0, in, out, n, bad_flag);
}

void fr_from_mont(Ctx& cx, const Fr* in, Fr* out, uint64_t n) {
    (void)cx;
    if (n) SONIC_LAUNCH(k_fr_from_mont, div_up(n, 256), 256, 0, in, out, n);
}

void fr_inv_few(Ctx& cx, const Fr* in, Fr* out, int n) {
    (void)cx;
    if (n) SONIC_LAUNCH(k_fr_inv, div_up(n, 32), 32, 0, in, out, n);
}

void fr_mul_pointwise(Ctx& cx, Fr* a, const Fr* b, uint32_t n) {
    (void)cx;
    if (n) SONIC_LAUNCH(k_fr_mul_pointwise, div_up(n, 256), 256, 0, a, b, n);
}

void pow_tables(Ctx& cx, const Fr* bases, int ntab, Fr* tab, uint64_t len, uint64_t stride, const uint32_t* d_sel) {
    (void)cx;
    if (ntab && len) SONIC_LAUNCH(k_pow_tables, dim3(div_up(div_up(len, POW_RUN), 128), (unsigned)ntab), 128, 0, bases, tab, len, stride, d_sel);
}

void open_batch(Ctx& cx, const std::vector<OpenJob>& jobs) {
    const int nj = (int)jobs.size();
    if (!nj) return;
    uint32_t max_len = 0;
    for (const OpenJob& j : jobs) max_len = j.len > max_len ? j.len : max_len;
    const uint32_t tiles = div_up(max_len ? max_len : 1, OPEN_TILE);
    OpenJob* d_jobs = cx.arena.get<OpenJob>(nj);
    SONIC_CUDA(cudaMemcpyAsync(d_jobs, jobs.data(), sizeof(OpenJob) * nj, cudaMemcpyHostToDevice, cx.stream));
    Fr* partial = cx.arena.get<Fr>((size_t)tiles * nj);
    Fr* H = cx.arena.get<Fr>(nj);
    SONIC_LAUNCH(k_open_partial, dim3(tiles, (unsigned)nj), OPEN_THREADS, 0, d_jobs, partial, tiles);
    SONIC_LAUNCH(k_open_tiles, nj, 32, 0, d_jobs, partial, tiles, H);
    SONIC_LAUNCH(k_open_quotient, dim3(tiles, (unsigned)nj), OPEN_THREADS, 0, d_jobs, partial, tiles, H);
    // the job table is read by kernels still in flight: the caller keeps `jobs` alive until it synchronises
}

}  // namespace sonic
