// Stages 5-7 of the batched MSM (see msm.cu for the pipeline): fold the bucket pieces that straddle
// chunk borders, reduce every bucket set to sum_b (b+1) * B_b, and finish each job (fold, one
// inversion, affine + compressed encoding).
#include "msm_acc.cuh"
#include "g1coop.cuh"

namespace sonic {

// ---- block-wide sum of XYZZ points through shared memory --------------------------------
template <int THREADS>
SONIC_D G1XYZZ block_sum_xyzz(G1XYZZ v, G1XYZZ* smem) {
    smem[threadIdx.x] = v;
    __syncthreads();
    for (int s = THREADS / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) {
            g1_add(v, smem[threadIdx.x + s]);
            smem[threadIdx.x] = v;
        }
        __syncthreads();
    }
    return v;  // valid in thread 0
}

// ---- stage 5: fold the pieces of buckets that straddle chunk borders ---------------------
__global__ void __launch_bounds__(128)
k_msm_fixup(const uint32_t* __restrict__ offsets, uint32_t GB, uint32_t L, G1XYZZ* __restrict__ buckets,
            const G1XYZZ* __restrict__ head, const G1XYZZ* __restrict__ tail,
            uint32_t* __restrict__ heavy_count, uint32_t* __restrict__ heavy_list) {
    const uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x;
    if (gb >= GB) return;
    const uint32_t a = offsets[gb], b = offsets[gb + 1];
    if (a == b) { store_xyzz(buckets + gb, G1XYZZ::inf()); return; }
    const uint32_t t0 = a / L, t1 = (b - 1) / L;
    if (t0 == t1) return;  // written whole by the accumulate kernel
    if (t1 - t0 + 1 > (uint32_t)MSM_HEAVY_PIECES) {
        heavy_list[atomicAdd(heavy_count, 1u)] = gb;
        return;
    }
    G1XYZZ acc = load_xyzz(tail + t0);
    for (uint32_t t = t0 + 1; t <= t1; ++t) g1_add(acc, load_xyzz(head + t));
    store_xyzz(buckets + gb, acc);
}

__global__ void __launch_bounds__(MSM_RED_THREADS)
k_msm_heavy(const uint32_t* __restrict__ offsets, uint32_t L, G1XYZZ* __restrict__ buckets,
            const G1XYZZ* __restrict__ head, const G1XYZZ* __restrict__ tail,
            const uint32_t* __restrict__ heavy_count, const uint32_t* __restrict__ heavy_list) {
    __shared__ G1XYZZ smem[MSM_RED_THREADS];
    const uint32_t nh = *heavy_count;
    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        const uint32_t gb = heavy_list[h];
        const uint32_t a = offsets[gb], b = offsets[gb + 1];
        const uint32_t t0 = a / L, t1 = (b - 1) / L;
        G1XYZZ acc = G1XYZZ::inf();
        for (uint32_t t = t0 + threadIdx.x; t <= t1; t += MSM_RED_THREADS)
            g1_add(acc, load_xyzz(t == t0 ? tail + t0 : head + t));
        acc = block_sum_xyzz<MSM_RED_THREADS>(acc, smem);
        if (threadIdx.x == 0) store_xyzz(buckets + gb, acc);
        __syncthreads();
    }
}

// Heavy buckets by quads, in two steps (option heavy_mode = 0; NOT the default): (heavy bucket, part) work items --
// HEAVY_SPLIT parts per bucket, each folded by one block of 64 quads (pieces strided over the quads, tree through
// shared memory) -- and one quad per bucket to fold the parts.  A 65 536-entry bucket (the all-ones weight rows of the
// synthetic circuits make dozens of them) has ~1 000 pieces.  Measured against one 128-thread block per bucket
// (k_msm_heavy): nothing at 8 GPUs (0.25 ms per rank either way: two rounds of items per SM), slower on one GPU
// (400 heavy buckets: 0.27 -> 0.47 ms, profiles/r02h_others_ncu.md): at 190 registers only 8 warps fit an SM and a
// quad spends four lanes per addition, so with many buckets the plain kernel's thread-level parallelism wins.
constexpr int HQ_THREADS = 256;
constexpr int HQ_QUADS = HQ_THREADS / 4;
constexpr int HEAVY_SPLIT = 4;

__global__ void __launch_bounds__(HQ_THREADS, 1)
k_msm_heavy_quad(const uint32_t* __restrict__ offsets, uint32_t L, const G1XYZZ* __restrict__ head, const G1XYZZ* __restrict__ tail,
                 const uint32_t* __restrict__ heavy_count, const uint32_t* __restrict__ heavy_list, G1XYZZ* __restrict__ parts) {
    __shared__ G1XYZZ park[HQ_QUADS];
    const Quad q;
    const uint32_t qi = threadIdx.x >> 2;
    const uint32_t items = *heavy_count * HEAVY_SPLIT;
    for (uint32_t item = blockIdx.x; item < items; item += gridDim.x) {
        const uint32_t gb = heavy_list[item / HEAVY_SPLIT], part = item % HEAVY_SPLIT;
        const uint32_t a = offsets[gb], b = offsets[gb + 1];
        const uint32_t t0 = a / L, t1 = (b - 1) / L;
        G1XYZZ acc = G1XYZZ::inf();
        for (uint32_t t = t0 + part * HQ_QUADS + qi; t <= t1; t += HQ_QUADS * HEAVY_SPLIT)
            g1_add_quad(q, acc, load_xyzz(t == t0 ? tail + t0 : head + t));
        if (q.lane == 0) park[qi] = acc;
        __syncthreads();
        for (uint32_t s = HQ_QUADS / 2; s > 0; s >>= 1) {
            if (qi < s) g1_add_quad(q, acc, park[qi + s]);
            __syncthreads();
            if (qi < s && q.lane == 0) park[qi] = acc;
            __syncthreads();
        }
        if (threadIdx.x == 0) store_xyzz(parts + item, acc);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(128)
k_msm_heavy_fin(const uint32_t* __restrict__ heavy_count, const uint32_t* __restrict__ heavy_list, const G1XYZZ* __restrict__ parts,
                G1XYZZ* __restrict__ buckets) {
    const Quad q;
    const uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    if (h >= *heavy_count) return;
    G1XYZZ acc = load_xyzz(parts + (size_t)h * HEAVY_SPLIT);
    for (int k = 1; k < HEAVY_SPLIT; ++k) g1_add_quad(q, acc, load_xyzz(parts + (size_t)h * HEAVY_SPLIT + k));
    if (q.lane == 0) store_xyzz(buckets + heavy_list[h], acc);
}

// ---- stage 6: bucket reduction  sum_b (b+1) * B_b  per bucket set, level by level ----------------
// Items come in sets of `count` consecutive elements; a thread folds K consecutive items of one
// set with the running-sum trick (2 additions per item) and emits
//   out_x = sum of its x items                      (still to be weighted by the levels above)
//   out_y = sum of its y items + 2^shift * sum_k (k + one_based) * x_k
// where y carries what is already fully weighted.  The next level sees count/K items whose index
// weight is worth K times more (shift += log2 K).  About 2.2 additions per bucket in total, no
// per-thread scalar multiplication, and every level is as wide as it has items.
__global__ void __launch_bounds__(128)
k_msm_bucket_level(const G1XYZZ* __restrict__ in_x, const G1XYZZ* __restrict__ in_y, uint32_t count, uint32_t K,
                   uint32_t shift, uint32_t one_based, G1XYZZ* __restrict__ out_x, G1XYZZ* __restrict__ out_y,
                   uint32_t total_threads) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_threads) return;
    const uint32_t per_set = count / K;
    const uint32_t set = t / per_set, u = t - set * per_set;
    const size_t base = (size_t)set * count + (size_t)u * K;
    G1XYZZ run = G1XYZZ::inf(), acc = G1XYZZ::inf();
    for (uint32_t k = K; k-- > 0;) {
        g1_add(run, load_xyzz(in_x + base + k));
        if (k > 0 || one_based) g1_add(acc, run);
    }
    for (uint32_t i = 0; i < shift; ++i) acc = g1_dbl(acc);
    if (in_y)
        for (uint32_t k = 0; k < K; ++k) g1_add(acc, load_xyzz(in_y + base + k));
    store_xyzz(out_x + t, run);
    store_xyzz(out_y + t, acc);
}

// ---- stage 6 (flat variant): per-window bucket reduction  sum_b (b+1) * B_b ------------------------------
// grid = (blocks_per_window, windows_total); each thread owns K consecutive buckets.
// Blocks of two warps, four resident per SM: a single warp per scheduler keeps the multiplier pipe
// about 40 % busy (dependent carry chains), two reach 75 %, and small blocks spread a few hundred
// thread-columns evenly over the 148 SMs.
constexpr int RED_T = 64;

// The weighted accumulator `acc` is parked in shared memory (word-major per-thread column, conflict
// free) while the running sum absorbs the next bucket, so that at most two points are live in
// registers at any time (three spill: 255 registers, 604 bytes of spill stores before).
SONIC_D G1XYZZ park_load(const uint32_t* __restrict__ col) {
    G1XYZZ r;
    Fq* f[4] = {&r.x, &r.y, &r.zz, &r.zzz};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int l = 0; l < 12; ++l) f[k]->l[l] = col[(k * 12 + l) * RED_T];
    return r;
}
SONIC_D void park_store(uint32_t* __restrict__ col, const G1XYZZ& v) {
    const Fq* f[4] = {&v.x, &v.y, &v.zz, &v.zzz};
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int l = 0; l < 12; ++l) col[(k * 12 + l) * RED_T] = f[k]->l[l];
}

template <int MINB>
__global__ void __launch_bounds__(RED_T, MINB)
k_msm_bucket_reduce(const G1XYZZ* __restrict__ buckets, uint32_t B, uint32_t K, G1XYZZ* __restrict__ partial) {
    __shared__ G1XYZZ smem[RED_T];
    static_assert(sizeof(G1XYZZ) == 48 * sizeof(uint32_t), "park layout");
    uint32_t* col = reinterpret_cast<uint32_t*>(smem) + threadIdx.x;
    const uint32_t window = blockIdx.y;
    const uint32_t first = (blockIdx.x * RED_T + threadIdx.x) * K;  // 0-based bucket index in window
    park_store(col, G1XYZZ::inf());
    if (first < B) {
        G1XYZZ run = G1XYZZ::inf();
        const G1XYZZ* bp = buckets + (size_t)window * B + first;
        const uint32_t kmax = (B - first < K) ? (B - first) : K;
        for (uint32_t k = kmax; k-- > 0;) {
            g1_add(run, load_xyzz(bp + k));
            G1XYZZ acc = park_load(col);
            g1_add(acc, run);
            park_store(col, acc);
        }
        // weights are (bucket index + 1): add first * (sum of the K buckets)
        if (first) {
            G1XYZZ m = G1XYZZ::inf();
            for (int bit = 31 - __clz(first); bit >= 0; --bit) {
                m = g1_dbl(m);
                if ((first >> bit) & 1) g1_add(m, run);
            }
            G1XYZZ acc = park_load(col);
            g1_add(acc, m);
            park_store(col, acc);
        }
    }
    G1XYZZ acc = park_load(col);
    __syncthreads();  // the tree below reuses the parking area
    acc = block_sum_xyzz<RED_T>(acc, smem);
    if (threadIdx.x == 0) store_xyzz(partial + (size_t)window * gridDim.x + blockIdx.x, acc);
}

// ---- stage 6, quad variant (the default): the same K-buckets-per-worker running sums, but a worker is a QUAD of
// lanes sharing every point operation (g1coop.cuh).  The stage is a chain of dependent full additions --
// 2K running-sum steps, the offset multiple first * (sum of the K buckets) (15 doublings + a few additions), the
// block tree -- whose length barely depends on K, and a lone thread spends ~20 us on each link (14 dependent
// multiplications at ~2 800 cycles: the carry chains of one multiplication cannot overlap).  A quad brings a
// link down to 4.5 multiplication latencies (add) / 3.5 (double).  Measured for one job of 2^15 buckets:
// 0.74 ms -> see profiles/r02*_others_ncu.md.
// The running sum and the multiple live in registers; the weighted accumulator is parked in shared memory (one
// slot per quad, read by all four lanes as a broadcast) so that two points are live at a time.
constexpr int RQ_THREADS = 128;
constexpr int RQ_QUADS = RQ_THREADS / 4;
constexpr int RQ_MINB = 2;

__global__ void __launch_bounds__(RQ_THREADS, RQ_MINB)
k_msm_bucket_reduce_quad(const G1XYZZ* __restrict__ buckets, uint32_t B, uint32_t K, G1XYZZ* __restrict__ partial) {
    __shared__ G1XYZZ park[RQ_QUADS];
    const Quad q;
    const uint32_t window = blockIdx.y;
    const uint32_t qi = threadIdx.x >> 2;
    const uint32_t first = (blockIdx.x * RQ_QUADS + qi) * K;  // 0-based bucket index in window
    if (q.lane == 0) park[qi] = G1XYZZ::inf();
    __syncwarp(q.mask);
    if (first < B) {
        G1XYZZ run = G1XYZZ::inf();
        const G1XYZZ* bp = buckets + (size_t)window * B + first;
        const uint32_t kmax = (B - first < K) ? (B - first) : K;
        for (uint32_t k = kmax; k-- > 0;) {
            g1_add_quad(q, run, load_xyzz(bp + k));
            G1XYZZ acc = park[qi];
            g1_add_quad(q, acc, run);
            __syncwarp(q.mask);
            if (q.lane == 0) park[qi] = acc;
            __syncwarp(q.mask);
        }
        // weights are (bucket index + 1): add first * (sum of the K buckets)
        if (first) {
            G1XYZZ m = G1XYZZ::inf();
            for (int bit = 31 - __clz(first); bit >= 0; --bit) {
                m = g1_dbl_quad(q, m);
                if ((first >> bit) & 1) g1_add_quad(q, m, run);
            }
            G1XYZZ acc = park[qi];
            g1_add_quad(q, acc, m);
            __syncwarp(q.mask);
            if (q.lane == 0) park[qi] = acc;
            __syncwarp(q.mask);
        }
    }
    // tree over the quads of the block
    __syncthreads();
    for (uint32_t s = RQ_QUADS / 2; s > 0; s >>= 1) {
        G1XYZZ acc;
        if (qi < s) {
            acc = park[qi];
            g1_add_quad(q, acc, park[qi + s]);
        }
        __syncthreads();
        if (qi < s && q.lane == 0) park[qi] = acc;
        __syncthreads();
    }
    if (threadIdx.x == 0) store_xyzz(partial + (size_t)window * gridDim.x + blockIdx.x, park[0]);
}

// ---- stage 7, quad variant, one bucket set per job (window tables): fold the S block partials, one
// inversion, affine + compressed
__global__ void __launch_bounds__(RQ_THREADS)
k_msm_finish_quad(const G1XYZZ* __restrict__ partial, uint32_t S, G1Affine* __restrict__ out_aff, uint8_t* __restrict__ out_comp) {
    __shared__ G1XYZZ park[RQ_QUADS];
    const Quad q;
    const uint32_t job = blockIdx.x, qi = threadIdx.x >> 2;
    const G1XYZZ* pp = partial + (size_t)job * S;
    G1XYZZ v = G1XYZZ::inf();
    for (uint32_t s = qi; s < S; s += RQ_QUADS) g1_add_quad(q, v, load_xyzz(pp + s));
    if (q.lane == 0) park[qi] = v;
    __syncthreads();
    for (uint32_t s = RQ_QUADS / 2; s > 0; s >>= 1) {
        if (qi < s) g1_add_quad(q, v, park[qi + s]);
        __syncthreads();
        if (qi < s && q.lane == 0) park[qi] = v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        G1Affine a = g1_to_affine_single(v);
        if (out_aff) out_aff[job] = a;
        if (out_comp) g1_compress(a, out_comp + (size_t)job * 48);
    }
}

// ---- stage 7: per job: fold window partials, Horner over windows, affine, compress ---------
__global__ void __launch_bounds__(64)
k_msm_finish(const G1XYZZ* __restrict__ partial, uint32_t S, int W, int c, G1Affine* __restrict__ out_aff,
             uint8_t* __restrict__ out_comp) {
    __shared__ G1XYZZ win[64];
    const uint32_t job = blockIdx.x;
    G1XYZZ acc;
    if (W == 1) {
        // one bucket set per job (precomputed tables): the S block partials fold as a tree
        const G1XYZZ* pp = partial + (size_t)job * S;
        G1XYZZ v = G1XYZZ::inf();
        for (uint32_t s = threadIdx.x; s < S; s += 64) g1_add(v, load_xyzz(pp + s));
        acc = block_sum_xyzz<64>(v, win);
    } else {
        if ((int)threadIdx.x < W) {
            const G1XYZZ* pp = partial + ((size_t)job * W + threadIdx.x) * S;
            G1XYZZ v = load_xyzz(pp);
            for (uint32_t s = 1; s < S; ++s) g1_add(v, load_xyzz(pp + s));
            win[threadIdx.x] = v;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        if (W > 1) {
            acc = win[W - 1];
            for (int w = W - 2; w >= 0; --w) {
                for (int i = 0; i < c; ++i) acc = g1_dbl(acc);
                g1_add(acc, win[w]);
            }
        }
        G1Affine a = g1_to_affine_single(acc);
        if (out_aff) out_aff[job] = a;
        if (out_comp) g1_compress(a, out_comp + (size_t)job * 48);
    }
}

// ---- host side ---------------------------------------------------------------------------
void msm_reduce_stage(Ctx& cx, const MsmPlan& p, int M, const uint32_t* offsets, uint32_t chunks, G1XYZZ* buckets,
                      const G1XYZZ* head, const G1XYZZ* tail, G1Affine* d_out_aff, uint8_t* d_out_comp, cudaEvent_t fixup_done, bool buckets_final) {
    Arena& ar = cx.arena;
    cudaStream_t st = cx.stream;
    if (!buckets_final) {
    uint32_t* heavy_count = ar.get<uint32_t>(1);
    const size_t heavy_max = (size_t)chunks / MSM_HEAVY_PIECES + 2;
    uint32_t* heavy_list = ar.get<uint32_t>(heavy_max);
    SONIC_CUDA(cudaMemsetAsync(heavy_count, 0, 4, st));
    SONIC_LAUNCH(k_msm_fixup, div_up(p.GB, 128), 128, 0, offsets, p.GB, p.L, buckets, head, tail, heavy_count, heavy_list);
    if (cx.opt_heavy_mode == 0) {
        G1XYZZ* parts = ar.get<G1XYZZ>(heavy_max * HEAVY_SPLIT);
        SONIC_LAUNCH(k_msm_heavy_quad, cx.sm_count, HQ_THREADS, 0, offsets, p.L, head, tail, heavy_count, heavy_list, parts);
        SONIC_LAUNCH(k_msm_heavy_fin, div_up(heavy_max * 4, 128), 128, 0, heavy_count, heavy_list, parts, buckets);
    } else {
        SONIC_LAUNCH(k_msm_heavy, cx.sm_count * 2, MSM_RED_THREADS, 0, offsets, p.L, buckets, head, tail, heavy_count, heavy_list);
    }
    }
    SONIC_CUDA(cudaEventRecord(fixup_done, st));

    // Automatic (reduce_mode 0): quads while the stage is latency-bound -- few bucket sets: a standalone MSM, a small
    // proof, one rank's share of a sharded proof (one job of 2^15 buckets: 0.97 -> 0.55 ms; prove at n = 2^13: 1.33 -> 0.95) --
    // and one thread per K buckets once there are enough buckets to fill the multiplier pipe without help (39 sets of
    // 2^15: 2.52 ms against 2.72 with quads; 17 / 23 sets, a rank of two: 1.86 / 1.93 against 1.48 / 1.74 -- the switch
    // sits at 2^20 buckets).
    const bool quads = cx.opt_reduce_mode == 3 || (cx.opt_reduce_mode == 0 && (uint64_t)M * p.sets * p.B <= (1ull << 20));
    if (quads) {
        // quads of lanes share every point operation.  K buckets per quad, as large as keeping ALL blocks
        // resident at once allows (RQ_MINB per SM): a second wave would double the time of this latency-bound stage.
        const uint32_t nsets = (uint32_t)M * p.sets;
        uint32_t K = (uint32_t)cx.opt_reduce_k;
        if (K == 0) {
            uint32_t S_max = (uint32_t)cx.sm_count * RQ_MINB / nsets;
            if (S_max < 1) S_max = 1;
            K = div_up(p.B, (uint64_t)RQ_QUADS * S_max);
            if (K < 2) K = 2;
        }
        if (K > p.B) K = p.B;
        const uint32_t S = div_up(p.B, (uint64_t)RQ_QUADS * K);
        G1XYZZ* partial = ar.get<G1XYZZ>((size_t)nsets * S);
        SONIC_LAUNCH(k_msm_bucket_reduce_quad, dim3(S, nsets), RQ_THREADS, 0, buckets, p.B, K, partial);
        if (p.sets == 1) SONIC_LAUNCH(k_msm_finish_quad, M, RQ_THREADS, 0, partial, S, d_out_aff, d_out_comp);
        else SONIC_LAUNCH(k_msm_finish, M, 64, 0, partial, S, p.sets, p.c, d_out_aff, d_out_comp);
    } else if (cx.opt_reduce_mode != 1) {
        // flat: each thread K buckets + its offset multiple, block tree, per-job fold of the block partials
        // buckets per thread: every thread pays ~29 extra point operations (its offset multiple and
        // the block tree) on top of 2 per bucket, so K is as large as filling the machine allows
        // Every thread pays ~37 extra point operations (its offset multiple and the block tree) on top
        // of 2 per bucket, so K is as large as keeping ALL blocks resident at once allows (4 per SM at
        // 244 registers): a second wave of blocks would double the time of this latency-bound stage.
        const uint32_t nsets = (uint32_t)M * p.sets;
        uint32_t K = (uint32_t)cx.opt_reduce_k;
        if (K == 0) {
            uint32_t S_max = (uint32_t)cx.sm_count * 4 / nsets;
            if (S_max < 1) S_max = 1;
            K = div_up(p.B, (uint64_t)RED_T * S_max);
            if (K < 4) K = 4;
        }
        if (K > p.B / RED_T) K = p.B / RED_T;
        if (K < 1) K = 1;
        const uint32_t S = div_up(p.B, (uint64_t)RED_T * K);
        G1XYZZ* partial = ar.get<G1XYZZ>((size_t)nsets * S);
        SONIC_LAUNCH(k_msm_bucket_reduce<4>, dim3(S, nsets), RED_T, 0, buckets, p.B, K, partial);
        SONIC_LAUNCH(k_msm_finish, M, 64, 0, partial, S, p.sets, p.c, d_out_aff, d_out_comp);
    } else {
    // level-by-level reduction of every bucket set to one point
        const uint32_t nsets = (uint32_t)M * p.sets;
        const G1XYZZ* lx = buckets;
        const G1XYZZ* ly = nullptr;
        uint32_t count = p.B, shift = 0, level = 0;
        while (count > 1 || level == 0) {
            uint32_t K = count >= 16 ? 16 : count;
            if (count > 16 && count / 16 < 8) K = count / 8 >= 2 ? count / 8 : count;  // keep the last level at 8 items
            const uint32_t threads = nsets * (count / K);
            G1XYZZ* ox = ar.get<G1XYZZ>(threads);
            G1XYZZ* oy = ar.get<G1XYZZ>(threads);
            SONIC_LAUNCH(k_msm_bucket_level, div_up(threads, 128), 128, 0, lx, ly, count, K, shift, level == 0 ? 1u : 0u, ox, oy, threads);
            lx = ox;
            ly = oy;
            uint32_t lg = 0;
            while ((1u << lg) < K) ++lg;
            shift += lg;
            count /= K;
            ++level;
        }
        SONIC_LAUNCH(k_msm_finish, M, 64, 0, ly, 1u, p.sets, p.c, d_out_aff, d_out_comp);
    }
}

}  // namespace sonic
