// Bucket accumulation in affine coordinates with batched inversions (`acc_mode` 2; chosen by the automatic mode 3 for
// batches of at least 5 * 2^22 entries -- msm.cu).
//
// The XYZZ kernels spend ~9 multiplications on every insertion (8M + 2S; 2 712 IMAD.WIDE executed).  An
// affine addition needs 1/(x2 - x1) and then only 3 (lambda, lambda^2, y3); with Montgomery's trick the
// inversion of MANY independent denominators costs 3 multiplications each plus a few real inversions for the
// lot.  Independent additions are what a pairwise tree reduction of the buckets offers: round r adds the
// entries of every bucket two by two (all pairs of all buckets at once), halving every bucket.  5.8
// multiplications (1 794 IMAD.WIDE) per insertion, no chunk head/tail pieces, no fix-up stage -- for ~0.7 KB of
// HBM traffic per insertion (the intermediate points and the running products live in global memory; 87 GB per
// n = 2^16 proof, 2.3 TB/s on average: not the bound).
//
// Per round: A (k_aff_prefix) denominators and their per-thread running products, one product per block;
// B (k_aff_inverses) the inverses of the block products, 256 per block of B and one Euclid inverse each;
// C (k_aff_add) the block's product tree walked back down, 1/d per slot, the additions.
// Hybrid: the last `aff_tail` (4) halvings are not rounds -- a thread sums the <= ~16 points its bucket still has with
// mixed XYZZ additions (k_aff_left_list), a block the few buckets with more than 32 left (k_aff_left_sum).
//
// Measured on a B200 (profiles/r02q_affine_stage.md, r02t_bench_n1.json): prove() at n = 2^16 42.2 ms against 48.3 ms
// with the XYZZ kernel, a 2^24-point MSM 74.4 against 81.6 ms; a rank of eight (15 M entries) 8.3 against 8.0 ms, hence
// the threshold.  The additions run at 68-76 % of the multiplier pipe, the first round's denominator pass at 46 %
// (bound by its gathers).
#include <algorithm>
#include <cmath>

#include "msm_acc.cuh"

namespace sonic {

constexpr int AF_T = 128;               // threads per block of the A / C kernels
constexpr int AF_M_MAX = 32;            // output slots per thread: 32 / 16 / 8, by the size of the round (template parameter M)
constexpr int AF_MAX_ROUNDS = 12;
constexpr int AF_BT = 256;              // threads per block of kernel B
constexpr int AF_LEFT_SERIAL = 32;      // a bucket left with at most this many points after the last round is summed by one thread

// ---- round bookkeeping -------------------------------------------------------------------------------------
// cnt[b] points of bucket b start at base[b] in the round's input; the round leaves n = ceil(cnt / 2) of them.
// A bucket with n == 1 is finished by this round (its point goes to the bucket array) and has nothing in the next.
__global__ void __launch_bounds__(256) k_aff_prepare_first(const uint32_t* __restrict__ offsets, uint32_t GB, uint32_t* __restrict__ base,
                                                           uint32_t* __restrict__ cnt, uint32_t* __restrict__ n_out, G1XYZZ* __restrict__ buckets) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > GB) return;
    if (b == GB) { n_out[GB] = 0; return; }
    const uint32_t a = offsets[b], k = offsets[b + 1] - a;
    base[b] = a;
    cnt[b] = k;
    n_out[b] = (k + 1) / 2;
    if (k == 0) store_xyzz(buckets + b, G1XYZZ::inf());
}

__global__ void __launch_bounds__(256) k_aff_prepare_next(const uint32_t* __restrict__ cnt_prev, const uint32_t* __restrict__ wo_prev, uint32_t GB,
                                                          uint32_t* __restrict__ base, uint32_t* __restrict__ cnt, uint32_t* __restrict__ n_out) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > GB) return;
    if (b == GB) { n_out[GB] = 0; return; }
    const uint32_t n_prev = (cnt_prev[b] + 1) / 2;
    const uint32_t k = n_prev > 1 ? n_prev : 0;
    base[b] = wo_prev[b];
    cnt[b] = k;
    n_out[b] = (k + 1) / 2;
}

// last bucket whose first work slot is <= s   (wo[0] = 0 <= s < wo[GB])
SONIC_D uint32_t aff_find_bucket(const uint32_t* __restrict__ wo, uint32_t GB, uint32_t s) {
    uint32_t lo = 0, hi = GB;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (wo[mid] <= s) lo = mid; else hi = mid;
    }
    return lo;
}

// the inputs of one output slot
struct AffSlot {
    uint32_t b;      // bucket
    uint32_t i0;     // index of the first input
    bool has2;       // a partner exists at i0 + 1
    bool last;       // the bucket is down to one point after this round
};

SONIC_D AffSlot aff_slot(uint32_t s, uint32_t& b, const uint32_t* __restrict__ wo, const uint32_t* __restrict__ base,
                         const uint32_t* __restrict__ cnt, uint32_t GB) {
    if (s >= wo[b + 1]) {
        ++b;
        if (s >= wo[b + 1]) b = aff_find_bucket(wo, GB, s);   // a run of empty buckets
    }
    AffSlot r;
    r.b = b;
    const uint32_t local = s - wo[b], k = cnt[b];
    r.i0 = base[b] + 2 * local;
    r.has2 = 2 * local + 1 < k;
    r.last = k <= 2;
    return r;
}

// x of input i (the first 48 bytes of the point); FIRST: through the sorted entry list into the SRS tables
template <bool FIRST>
SONIC_D Fq aff_load_x(const uint32_t* __restrict__ entries, const G1Affine* __restrict__ pts, uint32_t i) {
    const G1Affine* p = FIRST ? pts + (entries[i] & 0x7fffffffu) : pts + i;
    const uint4* q = reinterpret_cast<const uint4*>(p);
    const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
    Fq x;
    x.l[0] = v0.x; x.l[1] = v0.y; x.l[2] = v0.z; x.l[3] = v0.w;
    x.l[4] = v1.x; x.l[5] = v1.y; x.l[6] = v1.z; x.l[7] = v1.w;
    x.l[8] = v2.x; x.l[9] = v2.y; x.l[10] = v2.z; x.l[11] = v2.w;
    return x;
}
template <bool FIRST>
SONIC_D Fq aff_load_y(const uint32_t* __restrict__ entries, const G1Affine* __restrict__ pts, uint32_t i) {
    uint32_t e = 0;
    const G1Affine* p;
    if (FIRST) { e = entries[i]; p = pts + (e & 0x7fffffffu); } else { p = pts + i; }
    const uint4* q = reinterpret_cast<const uint4*>(p) + 3;
    const uint4 v0 = __ldg(q), v1 = __ldg(q + 1), v2 = __ldg(q + 2);
    Fq y;
    y.l[0] = v0.x; y.l[1] = v0.y; y.l[2] = v0.z; y.l[3] = v0.w;
    y.l[4] = v1.x; y.l[5] = v1.y; y.l[6] = v1.z; y.l[7] = v1.w;
    y.l[8] = v2.x; y.l[9] = v2.y; y.l[10] = v2.z; y.l[11] = v2.w;
    if (FIRST && (e & 0x80000000u)) y = fp_neg(y);
    return y;
}

// (No software prefetch: `prefetch.global.L2` of a slot's gathers, whether in a pass in front of the additions or one to
// eight slots ahead inside the loop, cost 0.8-1.1 ms per n = 2^16 proof -- the stage is bound by its products, not by
// the latency of its loads.)

// What a slot does, and the denominator it contributes to the batch (1 when nothing is inverted).  A point of G1
// with x = 0 does not exist ((0, +-2) has order 3), so x == 0 identifies the infinity marker (0,0).
enum { AK_COPY = 0, AK_TAKE2, AK_INF, AK_DBL, AK_ADD };
template <bool FIRST>
SONIC_D int aff_kind_x(const uint32_t* __restrict__ entries, const G1Affine* __restrict__ pts, const AffSlot& sl, const Fq& x1, const Fq& x2, Fq& d) {
    d = Fq::one();
    if (!sl.has2) return AK_COPY;
    if (x2.is_zero()) return AK_COPY;     // P1 + inf
    if (x1.is_zero()) return AK_TAKE2;    // inf + P2
    if (x1 == x2) {
        const Fq y1 = aff_load_y<FIRST>(entries, pts, sl.i0), y2 = aff_load_y<FIRST>(entries, pts, sl.i0 + 1);
        if (y1 == y2 && !y1.is_zero()) { d = fp_dbl(y1); return AK_DBL; }
        return AK_INF;                    // P - P
    }
    d = fp_sub(x2, x1);
    return AK_ADD;
}
template <bool FIRST>
SONIC_D int aff_kind(const uint32_t* __restrict__ entries, const G1Affine* __restrict__ pts, const AffSlot& sl, const Fq& x1, Fq& d) {
    d = Fq::one();
    if (!sl.has2) return AK_COPY;
    const Fq x2 = aff_load_x<FIRST>(entries, pts, sl.i0 + 1);
    if (x2.is_zero()) return AK_COPY;     // P1 + inf
    if (x1.is_zero()) return AK_TAKE2;    // inf + P2
    if (x1 == x2) {
        const Fq y1 = aff_load_y<FIRST>(entries, pts, sl.i0), y2 = aff_load_y<FIRST>(entries, pts, sl.i0 + 1);
        if (y1 == y2 && !y1.is_zero()) { d = fp_dbl(y1); return AK_DBL; }
        return AK_INF;                    // P - P
    }
    d = fp_sub(x2, x1);
    return AK_ADD;
}

// ---- A: denominators and their running products -------------------------------------------------------------------
template <bool FIRST, int AF_M>
__global__ void __launch_bounds__(AF_T, 4)
k_aff_prefix(const uint32_t* __restrict__ entries, const G1Affine* __restrict__ pts, const uint32_t* __restrict__ wo,
             const uint32_t* __restrict__ base, const uint32_t* __restrict__ cnt, uint32_t GB, Fq* __restrict__ pre,
             Fq* __restrict__ block_tot) {
    __shared__ Fq tree[AF_T];
    const uint32_t total = wo[GB];
    if (blockIdx.x * (uint32_t)(AF_T * AF_M) >= total) return;   // the whole block is beyond the round's slots (uniform)
    const uint32_t s0 = (blockIdx.x * AF_T + threadIdx.x) * AF_M;
    Fq run = Fq::one();
    if (s0 < total) {
        uint32_t b = aff_find_bucket(wo, GB, s0);
        // the x coordinates of slot j + 1 are loaded before the product of slot j: one thread walks its slots alone, and
        // without this its loads and its products alternate
        AffSlot sl = aff_slot(s0, b, wo, base, cnt, GB);
        Fq x1 = aff_load_x<FIRST>(entries, pts, sl.i0);
        Fq x2 = sl.has2 ? aff_load_x<FIRST>(entries, pts, sl.i0 + 1) : Fq::zero();
        for (int j = 0; j < AF_M; ++j) {
            const uint32_t s = s0 + j;
            if (s >= total) break;
            AffSlot nsl = sl;
            Fq nx1 = x1, nx2 = x2;
            if (j + 1 < AF_M && s + 1 < total) {
                nsl = aff_slot(s + 1, b, wo, base, cnt, GB);
                nx1 = aff_load_x<FIRST>(entries, pts, nsl.i0);
                nx2 = nsl.has2 ? aff_load_x<FIRST>(entries, pts, nsl.i0 + 1) : Fq::zero();
            }
            Fq d;
            aff_kind_x<FIRST>(entries, pts, sl, x1, x2, d);
            run = fp_mul(run, d);
            pre[s] = run;
            sl = nsl;
            x1 = nx1;
            x2 = nx2;
        }
    }
    // product of the block's thread totals
    tree[threadIdx.x] = run;
    __syncthreads();
    for (int s = AF_T / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) tree[threadIdx.x] = fp_mul(tree[threadIdx.x], tree[threadIdx.x + s]);
        __syncthreads();
    }
    if (threadIdx.x == 0) block_tot[blockIdx.x] = tree[0];
}

// ---- B: the inverses of all block totals ----------------------------------------------------------------------------------------
// Every block of AF_BT threads takes an equal share of the totals and pays one Euclid inversion for it: thread t owns `per`
// consecutive totals: running products (kept in global scratch), a product tree over the thread totals in shared memory,
// the inverse of the root, the tree walked back down, back-substitution.  (One block for everything spent 0.13-0.33 ms
// per round on a single SM -- fifteen serial products per thread in the first round; the launcher now gives a block at
// most AF_BT totals, so `per` is 1.)
__global__ void __launch_bounds__(AF_BT)
k_aff_inverses(const Fq* __restrict__ tot, const uint32_t* __restrict__ wo, uint32_t GB, uint32_t slots_per_block, Fq* __restrict__ scratch,
               Fq* __restrict__ inv) {
    __shared__ Fq tree[2 * AF_BT];   // node 1 = root, leaves at AF_BT + t
    const uint32_t total = wo[GB];
    const uint32_t nb = (total + slots_per_block - 1) / slots_per_block;
    const uint32_t share = (nb + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * share, hi = lo + share < nb ? lo + share : nb;
    if (lo >= hi) return;
    const uint32_t per = (hi - lo + AF_BT - 1) / AF_BT;
    const uint32_t i0 = lo + threadIdx.x * per < hi ? lo + threadIdx.x * per : hi, i1 = i0 + per < hi ? i0 + per : hi;
    Fq run = Fq::one();
    for (uint32_t i = i0; i < i1; ++i) {
        run = fp_mul(run, tot[i]);
        scratch[i] = run;
    }
    tree[AF_BT + threadIdx.x] = run;
    __syncthreads();
    for (int w = AF_BT / 2; w >= 1; w >>= 1) {   // nodes w .. 2w-1
        if ((int)threadIdx.x < w) {
            const int i = w + threadIdx.x;
            tree[i] = fp_mul(tree[2 * i], tree[2 * i + 1]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) tree[1] = fp_inv_euclid(tree[1]);
    __syncthreads();
    for (int w = 1; w <= AF_BT / 2; w <<= 1) {    // children of nodes w .. 2w-1
        if ((int)threadIdx.x < w) {
            const int i = w + threadIdx.x;
            const Fq l = tree[2 * i], r = tree[2 * i + 1], iv = tree[i];
            tree[2 * i] = fp_mul(iv, r);
            tree[2 * i + 1] = fp_mul(iv, l);
        }
        __syncthreads();
    }
    Fq inv_run = tree[AF_BT + threadIdx.x];   // 1 / (product of this thread's totals)
    for (uint32_t i = i1; i-- > i0;) {
        const Fq prev = i > i0 ? scratch[i - 1] : Fq::one();
        inv[i] = fp_mul(inv_run, prev);
        inv_run = fp_mul(inv_run, tot[i]);
    }
}

// ---- C: 1/d per slot, the additions ---------------------------------------------------------------------------------------
template <bool FIRST, int AF_M>
__global__ void __launch_bounds__(AF_T, 3)
k_aff_add(const uint32_t* __restrict__ entries, const G1Affine* __restrict__ pts, const uint32_t* __restrict__ wo,
          const uint32_t* __restrict__ base, const uint32_t* __restrict__ cnt, uint32_t GB, const Fq* __restrict__ pre,
          const Fq* __restrict__ inv_tot, G1Affine* __restrict__ out, G1XYZZ* __restrict__ buckets) {
    __shared__ Fq tree[2 * AF_T];   // node 1 = root, leaves at AF_T + t
    const uint32_t total = wo[GB];
    if (blockIdx.x * (uint32_t)(AF_T * AF_M) >= total) return;
    const uint32_t s0 = (blockIdx.x * AF_T + threadIdx.x) * AF_M;
    const uint32_t s1 = s0 >= total ? s0 : (total - s0 < (uint32_t)AF_M ? total : s0 + AF_M);   // this thread's slots [s0, s1)
    // 1 / (this thread's product): the block's product tree, walked down from the inverse of its root
    tree[AF_T + threadIdx.x] = s1 > s0 ? pre[s1 - 1] : Fq::one();
    __syncthreads();
    for (int w = AF_T / 2; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) {
            const int i = w + threadIdx.x;
            tree[i] = fp_mul(tree[2 * i], tree[2 * i + 1]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) tree[1] = inv_tot[blockIdx.x];
    __syncthreads();
    for (int w = 1; w <= AF_T / 2; w <<= 1) {
        if ((int)threadIdx.x < w) {
            const int i = w + threadIdx.x;
            const Fq l = tree[2 * i], r = tree[2 * i + 1], iv = tree[i];
            tree[2 * i] = fp_mul(iv, r);
            tree[2 * i + 1] = fp_mul(iv, l);
        }
        __syncthreads();
    }
    if (s1 <= s0) return;
    Fq inv_run = tree[AF_T + threadIdx.x];
    // the slots' buckets, forwards (the walk needs ascending slots); the additions then run backwards
    uint32_t bs[AF_M];
    {
        uint32_t b = aff_find_bucket(wo, GB, s0);
        for (int j = 0; j < AF_M; ++j) {
            const uint32_t s = s0 + j;
            if (s >= s1) break;
            if (s >= wo[b + 1]) {
                ++b;
                if (s >= wo[b + 1]) b = aff_find_bucket(wo, GB, s);
            }
            bs[j] = b;
        }
    }
    for (int j = AF_M - 1; j >= 0; --j) {
        const uint32_t s = s0 + j;
        if (s >= s1) continue;
        uint32_t b = bs[j];
        const AffSlot sl = aff_slot(s, b, wo, base, cnt, GB);
        const Fq x1 = aff_load_x<FIRST>(entries, pts, sl.i0);
        Fq d;
        const int kind = aff_kind<FIRST>(entries, pts, sl, x1, d);
        const Fq prev = j > 0 ? pre[s - 1] : Fq::one();
        const Fq inv_d = fp_mul(inv_run, prev);
        inv_run = fp_mul(inv_run, d);
        G1Affine r;
        if (kind == AK_COPY) {
            r.x = x1;
            r.y = aff_load_y<FIRST>(entries, pts, sl.i0);
        } else if (kind == AK_TAKE2) {
            r.x = aff_load_x<FIRST>(entries, pts, sl.i0 + 1);
            r.y = aff_load_y<FIRST>(entries, pts, sl.i0 + 1);
        } else if (kind == AK_INF) {
            r = G1Affine::inf();
        } else {
            const Fq y1 = aff_load_y<FIRST>(entries, pts, sl.i0);
            Fq lambda, x2;
            if (kind == AK_DBL) {
                const Fq xx = fp_sqr(x1);
                lambda = fp_mul(fp_add(fp_dbl(xx), xx), inv_d);   // 3 x^2 / 2y
                x2 = x1;
            } else {
                x2 = aff_load_x<FIRST>(entries, pts, sl.i0 + 1);
                const Fq y2 = aff_load_y<FIRST>(entries, pts, sl.i0 + 1);
                lambda = fp_mul(fp_sub(y2, y1), inv_d);
            }
            r.x = fp_sub(fp_sub(fp_sqr(lambda), x1), x2);
            r.y = fp_sub(fp_mul(lambda, fp_sub(x1, r.x)), y1);
        }
        if (sl.last) store_xyzz(buckets + sl.b, G1XYZZ::from_affine(r));
        else out[s] = r;
    }
}

// ---- A + B + C in one kernel: the inversion stays inside the block ----------------------------------------------------------
// (Option `aff_fused` = 1; NOT the default: measured 88 ms against 44 ms for the split form on a n = 2^16 proof.)  The
// split form reads every input twice from global memory (x for the denominators, then the whole points), stores and
// re-reads a 48-byte running product per slot and needs kernel B between the two halves.  Here a block of AF_T threads
// x AF_FM slots keeps the running products in shared memory (thread-minor columns, conflict-free), builds its product
// tree once, lets ONE thread invert the root (binary Euclid) while the block's other warps wait at the barrier, and
// goes straight on to the additions.  With 61 KB of shared memory only three blocks fit an SM, and three are not
// enough to cover one another's inversions: the multiplier pipe idles half of the time.
constexpr int AF_FM = 8;
constexpr int AF_FSLOTS = AF_T * AF_FM;
constexpr size_t AF_FSMEM = (size_t)AF_FM * 12 * AF_T * 4 + 2 * (size_t)AF_T * sizeof(Fq);

SONIC_D void aff_col_store(uint32_t* __restrict__ col, int j, const Fq& v) {
#pragma unroll
    for (int l = 0; l < 12; ++l) col[(j * 12 + l) * AF_T] = v.l[l];
}
SONIC_D Fq aff_col_load(const uint32_t* __restrict__ col, int j) {
    Fq v;
#pragma unroll
    for (int l = 0; l < 12; ++l) v.l[l] = col[(j * 12 + l) * AF_T];
    return v;
}

template <bool FIRST>
__global__ void __launch_bounds__(AF_T, 3)
k_aff_fused(const uint32_t* __restrict__ entries, const G1Affine* __restrict__ pts, const uint32_t* __restrict__ wo,
            const uint32_t* __restrict__ base, const uint32_t* __restrict__ cnt, uint32_t GB, G1Affine* __restrict__ out,
            G1XYZZ* __restrict__ buckets) {
    extern __shared__ uint32_t af_dyn[];
    uint32_t* col = af_dyn + threadIdx.x;                                                   // this thread's column of running products
    Fq* tree = reinterpret_cast<Fq*>(af_dyn + (size_t)AF_FM * 12 * AF_T);                  // node 1 = root, leaves at AF_T + t
    const uint32_t total = wo[GB];
    if (blockIdx.x * (uint32_t)AF_FSLOTS >= total) return;
    const uint32_t s0 = (blockIdx.x * AF_T + threadIdx.x) * AF_FM;
    const uint32_t s1 = s0 >= total ? s0 : (total - s0 < (uint32_t)AF_FM ? total : s0 + AF_FM);
    uint32_t bs[AF_FM];
    Fq run = Fq::one();
    if (s1 > s0) {
        uint32_t b = aff_find_bucket(wo, GB, s0);
#pragma unroll
        for (int j = 0; j < AF_FM; ++j) {
            const uint32_t s = s0 + j;
            if (s < s1) {
                const AffSlot sl = aff_slot(s, b, wo, base, cnt, GB);
                bs[j] = b;
                const Fq x1 = aff_load_x<FIRST>(entries, pts, sl.i0);
                Fq d;
                aff_kind<FIRST>(entries, pts, sl, x1, d);
                run = fp_mul(run, d);
                aff_col_store(col, j, run);
            }
        }
    }
    tree[AF_T + threadIdx.x] = run;
    __syncthreads();
    for (int w = AF_T / 2; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) {
            const int i = w + threadIdx.x;
            tree[i] = fp_mul(tree[2 * i], tree[2 * i + 1]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) tree[1] = fp_inv_euclid(tree[1]);
    __syncthreads();
    for (int w = 1; w <= AF_T / 2; w <<= 1) {
        if ((int)threadIdx.x < w) {
            const int i = w + threadIdx.x;
            const Fq l = tree[2 * i], r = tree[2 * i + 1], iv = tree[i];
            tree[2 * i] = fp_mul(iv, r);
            tree[2 * i + 1] = fp_mul(iv, l);
        }
        __syncthreads();
    }
    if (s1 <= s0) return;
    Fq inv_run = tree[AF_T + threadIdx.x];
#pragma unroll
    for (int j = AF_FM - 1; j >= 0; --j) {
        const uint32_t s = s0 + j;
        if (s >= s1) continue;
        uint32_t b = bs[j];
        const AffSlot sl = aff_slot(s, b, wo, base, cnt, GB);
        const Fq x1 = aff_load_x<FIRST>(entries, pts, sl.i0);
        Fq d;
        const int kind = aff_kind<FIRST>(entries, pts, sl, x1, d);
        const Fq prev = j > 0 ? aff_col_load(col, j > 0 ? j - 1 : 0) : Fq::one();
        const Fq inv_d = fp_mul(inv_run, prev);
        inv_run = fp_mul(inv_run, d);
        G1Affine r;
        if (kind == AK_COPY) {
            r.x = x1;
            r.y = aff_load_y<FIRST>(entries, pts, sl.i0);
        } else if (kind == AK_TAKE2) {
            r.x = aff_load_x<FIRST>(entries, pts, sl.i0 + 1);
            r.y = aff_load_y<FIRST>(entries, pts, sl.i0 + 1);
        } else if (kind == AK_INF) {
            r = G1Affine::inf();
        } else {
            const Fq y1 = aff_load_y<FIRST>(entries, pts, sl.i0);
            Fq lambda, x2;
            if (kind == AK_DBL) {
                const Fq xx = fp_sqr(x1);
                lambda = fp_mul(fp_add(fp_dbl(xx), xx), inv_d);   // 3 x^2 / 2y
                x2 = x1;
            } else {
                x2 = aff_load_x<FIRST>(entries, pts, sl.i0 + 1);
                const Fq y2 = aff_load_y<FIRST>(entries, pts, sl.i0 + 1);
                lambda = fp_mul(fp_sub(y2, y1), inv_d);
            }
            r.x = fp_sub(fp_sub(fp_sqr(lambda), x1), x2);
            r.y = fp_sub(fp_mul(lambda, fp_sub(x1, r.x)), y1);
        }
        if (sl.last) store_xyzz(buckets + sl.b, G1XYZZ::from_affine(r));
        else out[s] = r;
    }
}

// ---- A + B + C in one kernel, running products in global memory (option `aff_fused` = 2) ------------------------------------
// The shared-memory form above can hold 8 slots per thread, and a block then waits on its barriers (fourteen tree levels
// and the Euclid inverse of the root, ~60 us) for a third of its life.  With the running products in `pre` (written and
// re-read by the same thread a few hundred microseconds apart) a thread takes M = 16 or 32 slots, the wait is a tenth of
// the block's life and the other three blocks of the SM cover it; the separate denominator pass over the inputs, kernel
// B and its single-SM serial section are gone.
template <bool FIRST, int AF_M>
__global__ void __launch_bounds__(AF_T, 4)
k_aff_round(const uint32_t* __restrict__ entries, const G1Affine* __restrict__ pts, const uint32_t* __restrict__ wo,
            const uint32_t* __restrict__ base, const uint32_t* __restrict__ cnt, uint32_t GB, Fq* __restrict__ pre,
            G1Affine* __restrict__ out, G1XYZZ* __restrict__ buckets) {
    __shared__ Fq tree[2 * AF_T];   // node 1 = root, leaves at AF_T + t
    const uint32_t total = wo[GB];
    if (blockIdx.x * (uint32_t)(AF_T * AF_M) >= total) return;
    const uint32_t s0 = (blockIdx.x * AF_T + threadIdx.x) * AF_M;
    const uint32_t s1 = s0 >= total ? s0 : (total - s0 < (uint32_t)AF_M ? total : s0 + AF_M);   // this thread's slots [s0, s1)
    uint32_t bs[AF_M];
    {
        Fq run = Fq::one();
        if (s1 > s0) {
            uint32_t b = aff_find_bucket(wo, GB, s0);
            for (int j = 0; j < AF_M; ++j) {
                const uint32_t s = s0 + j;
                if (s >= s1) break;
                const AffSlot sl = aff_slot(s, b, wo, base, cnt, GB);
                bs[j] = b;
                const Fq x1 = aff_load_x<FIRST>(entries, pts, sl.i0);
                Fq d;
                aff_kind<FIRST>(entries, pts, sl, x1, d);
                run = fp_mul(run, d);
                pre[s] = run;
            }
        }
        tree[AF_T + threadIdx.x] = run;
    }
    __syncthreads();
    for (int w = AF_T / 2; w >= 1; w >>= 1) {
        if ((int)threadIdx.x < w) {
            const int i = w + threadIdx.x;
            tree[i] = fp_mul(tree[2 * i], tree[2 * i + 1]);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) tree[1] = fp_inv_euclid(tree[1]);
    __syncthreads();
    for (int w = 1; w <= AF_T / 2; w <<= 1) {
        if ((int)threadIdx.x < w) {
            const int i = w + threadIdx.x;
            const Fq l = tree[2 * i], r = tree[2 * i + 1], iv = tree[i];
            tree[2 * i] = fp_mul(iv, r);
            tree[2 * i + 1] = fp_mul(iv, l);
        }
        __syncthreads();
    }
    if (s1 <= s0) return;
    Fq inv_run = tree[AF_T + threadIdx.x];
    for (int j = AF_M - 1; j >= 0; --j) {
        const uint32_t s = s0 + j;
        if (s >= s1) continue;
        uint32_t b = bs[j];
        const AffSlot sl = aff_slot(s, b, wo, base, cnt, GB);
        const Fq x1 = aff_load_x<FIRST>(entries, pts, sl.i0);
        Fq d;
        const int kind = aff_kind<FIRST>(entries, pts, sl, x1, d);
        const Fq prev = j > 0 ? pre[s - 1] : Fq::one();
        const Fq inv_d = fp_mul(inv_run, prev);
        inv_run = fp_mul(inv_run, d);
        G1Affine r;
        if (kind == AK_COPY) {
            r.x = x1;
            r.y = aff_load_y<FIRST>(entries, pts, sl.i0);
        } else if (kind == AK_TAKE2) {
            r.x = aff_load_x<FIRST>(entries, pts, sl.i0 + 1);
            r.y = aff_load_y<FIRST>(entries, pts, sl.i0 + 1);
        } else if (kind == AK_INF) {
            r = G1Affine::inf();
        } else {
            const Fq y1 = aff_load_y<FIRST>(entries, pts, sl.i0);
            Fq lambda, x2;
            if (kind == AK_DBL) {
                const Fq xx = fp_sqr(x1);
                lambda = fp_mul(fp_add(fp_dbl(xx), xx), inv_d);   // 3 x^2 / 2y
                x2 = x1;
            } else {
                x2 = aff_load_x<FIRST>(entries, pts, sl.i0 + 1);
                const Fq y2 = aff_load_y<FIRST>(entries, pts, sl.i0 + 1);
                lambda = fp_mul(fp_sub(y2, y1), inv_d);
            }
            r.x = fp_sub(fp_sub(fp_sqr(lambda), x1), x2);
            r.y = fp_sub(fp_mul(lambda, fp_sub(x1, r.x)), y1);
        }
        if (sl.last) store_xyzz(buckets + sl.b, G1XYZZ::from_affine(r));
        else out[s] = r;
    }
}

// ---- what is left of buckets longer than 2^R entries -----------------------------------------------------------------------
// a few points: the thread that finds them sums them; many (a 65 536-entry bucket still has hundreds): one block each
__global__ void __launch_bounds__(128) k_aff_left_list(const uint32_t* __restrict__ cnt_prev, const uint32_t* __restrict__ wo_prev,
                                                       const G1Affine* __restrict__ pts, uint32_t GB, uint32_t* __restrict__ count,
                                                       uint32_t* __restrict__ list, G1XYZZ* __restrict__ buckets) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= GB) return;
    const uint32_t n = (cnt_prev[b] + 1) / 2;   // points the last round left for this bucket (1 = already final)
    if (n <= 1) return;
    if (n > (uint32_t)AF_LEFT_SERIAL) { list[atomicAdd(count, 1u)] = b; return; }
    G1XYZZ acc = G1XYZZ::inf();
    const uint32_t first = wo_prev[b];
    for (uint32_t i = 0; i < n; ++i) g1_madd(acc, load_affine(pts + first + i));
    store_xyzz(buckets + b, acc);
}

template <int THREADS>
SONIC_D G1XYZZ aff_block_sum(G1XYZZ v, G1XYZZ* smem) {
    smem[threadIdx.x] = v;
    __syncthreads();
    for (int s = THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            g1_add(v, smem[threadIdx.x + s]);
            smem[threadIdx.x] = v;
        }
        __syncthreads();
    }
    return v;
}

__global__ void __launch_bounds__(MSM_RED_THREADS)
k_aff_left_sum(const uint32_t* __restrict__ cnt_prev, const uint32_t* __restrict__ wo_prev, const G1Affine* __restrict__ pts,
               const uint32_t* __restrict__ count, const uint32_t* __restrict__ list, G1XYZZ* __restrict__ buckets) {
    __shared__ G1XYZZ smem[MSM_RED_THREADS];
    const uint32_t nl = *count;
    for (uint32_t h = blockIdx.x; h < nl; h += gridDim.x) {
        const uint32_t b = list[h];
        const uint32_t n = (cnt_prev[b] + 1) / 2, first = wo_prev[b];
        G1XYZZ acc = G1XYZZ::inf();
        for (uint32_t i = threadIdx.x; i < n; i += MSM_RED_THREADS) g1_madd(acc, load_affine(pts + first + i));
        acc = aff_block_sum<MSM_RED_THREADS>(acc, smem);
        if (threadIdx.x == 0) store_xyzz(buckets + b, acc);
        __syncthreads();
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------------
// entries sorted by bucket (offsets[GB+1], at most `entries_max` of them) -> buckets[GB]
void launch_accumulate_affine(Ctx& cx, uint64_t entries_max, const uint32_t* entries, const uint32_t* offsets, uint32_t GB,
                              const G1Affine* points, G1XYZZ* buckets, double max_mean_bucket) {
    Arena& ar = cx.arena;
    // rounds: enough for a bucket six standard deviations above the longest job's mean size
    int AF_ROUNDS = 4;
    while (AF_ROUNDS < AF_MAX_ROUNDS && (double)(1u << AF_ROUNDS) < max_mean_bucket + 6.0 * sqrt(max_mean_bucket) + 2.0) ++AF_ROUNDS;
    // the last `aff_tail` halvings are left to the serial tail (k_aff_left_list: a thread sums the <= 2^aff_tail points a bucket
    // still has with mixed XYZZ additions): the late rounds hold few additions each and cost their fixed ~0.2 ms anyway
    AF_ROUNDS = std::max(1, AF_ROUNDS - cx.opt_aff_tail);
    // upper bounds of the work slots per round: every bucket halves, rounding up
    uint64_t smax[AF_MAX_ROUNDS];
    uint64_t prev = entries_max;
    for (int r = 0; r < AF_ROUNDS; ++r) {
        smax[r] = (prev + GB) / 2 + 1;
        prev = smax[r];
    }
    uint32_t* base[2] = {ar.get<uint32_t>((size_t)GB + 1), ar.get<uint32_t>((size_t)GB + 1)};
    uint32_t* cnt[2] = {ar.get<uint32_t>((size_t)GB + 1), ar.get<uint32_t>((size_t)GB + 1)};
    uint32_t* wo[2] = {ar.get<uint32_t>((size_t)GB + 1), ar.get<uint32_t>((size_t)GB + 1)};
    G1Affine* buf[2] = {ar.get<G1Affine>(smax[0]), ar.get<G1Affine>(AF_ROUNDS > 1 ? smax[1] : 1)};
    Fq* pre = ar.get<Fq>(smax[0]);
    const uint64_t nb_max = (smax[0] + AF_T * 8 - 1) / (AF_T * 8);
    Fq* tot = ar.get<Fq>(nb_max);
    Fq* inv = ar.get<Fq>(nb_max);
    Fq* scratch = ar.get<Fq>(nb_max);
    SONIC_CUDA(cudaFuncSetAttribute(k_aff_fused<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_FSMEM));
    SONIC_CUDA(cudaFuncSetAttribute(k_aff_fused<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)AF_FSMEM));
    for (int r = 0; r < AF_ROUNDS; ++r) {
        const int cur = r & 1, prv = cur ^ 1;
        if (r == 0) SONIC_LAUNCH(k_aff_prepare_first, div_up((uint64_t)GB + 1, 256), 256, 0, offsets, GB, base[cur], cnt[cur], wo[cur], buckets);
        else SONIC_LAUNCH(k_aff_prepare_next, div_up((uint64_t)GB + 1, 256), 256, 0, cnt[prv], wo[prv], GB, base[cur], cnt[cur], wo[cur]);
        exclusive_scan_u32(ar, wo[cur], wo[cur], GB + 1);
        G1Affine* out = buf[cur];
        if (cx.opt_aff_fused == 1) {
            const unsigned fblocks = div_up(smax[r], AF_FSLOTS);
            if (r == 0) SONIC_LAUNCH(k_aff_fused<true>, fblocks, AF_T, AF_FSMEM, entries, points, wo[cur], base[cur], cnt[cur], GB, out, buckets);
            else SONIC_LAUNCH(k_aff_fused<false>, fblocks, AF_T, AF_FSMEM, (const uint32_t*)nullptr, (const G1Affine*)buf[prv], wo[cur], base[cur], cnt[cur], GB, out, buckets);
            continue;
        }
        // slots per thread by the size of the round: long per-thread runs amortise the block's product tree (32 slots: the
        // barrier stalls halve against 16), short ones keep the last wave of a small round short
        const int m = cx.opt_aff_m > 0 ? cx.opt_aff_m : smax[r] >= (16u << 20) ? 64 : smax[r] >= (8u << 20) ? 32 : smax[r] >= (2u << 20) ? 16 : 8;
        const unsigned blocks = div_up(smax[r], (uint64_t)AF_T * m);
        const uint32_t* e = r == 0 ? entries : nullptr;
        const G1Affine* in = r == 0 ? points : buf[prv];
        if (cx.opt_aff_fused == 2) {
#define AF_FROUND(FIRST, M) SONIC_LAUNCH((k_aff_round<FIRST, M>), blocks, AF_T, 0, e, in, wo[cur], base[cur], cnt[cur], GB, pre, out, buckets)
            if (m == 64) {
                const unsigned blocks64 = div_up(smax[r], (uint64_t)AF_T * 64);
                if (r == 0) SONIC_LAUNCH((k_aff_round<true, 64>), blocks64, AF_T, 0, e, in, wo[cur], base[cur], cnt[cur], GB, pre, out, buckets);
                else SONIC_LAUNCH((k_aff_round<false, 64>), blocks64, AF_T, 0, e, in, wo[cur], base[cur], cnt[cur], GB, pre, out, buckets);
            } else if (r == 0) { if (m == 32) AF_FROUND(true, 32); else if (m == 16) AF_FROUND(true, 16); else AF_FROUND(true, 8); }
            else { if (m == 32) AF_FROUND(false, 32); else if (m == 16) AF_FROUND(false, 16); else AF_FROUND(false, 8); }
#undef AF_FROUND
            continue;
        }
        const unsigned b_blocks = std::max(1u, std::min((unsigned)cx.sm_count, (unsigned)div_up((uint64_t)blocks, (uint64_t)AF_BT)));
#define AF_ROUND(FIRST, M)                                                                                                              \
        do {                                                                                                                            \
            SONIC_LAUNCH((k_aff_prefix<FIRST, M>), blocks, AF_T, 0, e, in, wo[cur], base[cur], cnt[cur], GB, pre, tot);                 \
            SONIC_LAUNCH(k_aff_inverses, b_blocks, AF_BT, 0, tot, wo[cur], GB, (uint32_t)(AF_T * M), scratch, inv);                       \
            SONIC_LAUNCH((k_aff_add<FIRST, M>), blocks, AF_T, 0, e, in, wo[cur], base[cur], cnt[cur], GB, pre, inv, out, buckets);      \
        } while (0)
        if (r == 0) { if (m == 64) AF_ROUND(true, 64); else if (m == 32) AF_ROUND(true, 32); else if (m == 16) AF_ROUND(true, 16); else AF_ROUND(true, 8); }
        else { if (m == 64) AF_ROUND(false, 64); else if (m == 32) AF_ROUND(false, 32); else if (m == 16) AF_ROUND(false, 16); else AF_ROUND(false, 8); }
#undef AF_ROUND
    }
    // buckets that still hold more than one point after the last round
    const int last = (AF_ROUNDS - 1) & 1;
    uint32_t* left_count = ar.get<uint32_t>(1);
    uint32_t* left_list = ar.get<uint32_t>((size_t)GB);
    SONIC_CUDA(cudaMemsetAsync(left_count, 0, 4, cx.stream));
    SONIC_LAUNCH(k_aff_left_list, div_up(GB, 128), 128, 0, cnt[last], wo[last], buf[last], GB, left_count, left_list, buckets);
    SONIC_LAUNCH(k_aff_left_sum, cx.sm_count * 2, MSM_RED_THREADS, 0, cnt[last], wo[last], buf[last], left_count, left_list, buckets);
}

}  // namespace sonic
