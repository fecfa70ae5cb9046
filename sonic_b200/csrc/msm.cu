// Batched signed-digit Pippenger MSM over resident SRS bases.
//
// Stands in for the reference's commitment fold, one full scalar multiplication plus one
// point addition per term (src/Sonic/CommitmentScheme.hs:26-29 and :45-48).  The value
// is the same group element; only the summation order differs, and the group is abelian.
//
// A launch handles a batch of M independent MSMs ("jobs"): all 4Q+7 commitments and
// openings of one proof go through the pipeline together so that every stage fills the
// 148 SMs even though a single job has only 2e5-5e5 terms.
//
// Stages (all on one stream):
//   1. digits+count : the terms of a job are cut into tiles, one block per tile: every term is
//                     recoded into W = ceil(255/c) signed digits and the block histograms its
//                     tile's bucket ids in SHARED memory (a job's whole bucket set, <= 128 KB),
//                     then stores the histogram (tile-major: coalesced).
//   2. scan         : per bucket, running sum over the job's tiles (each tile's slot inside the
//                     bucket) and the bucket total; exclusive prefix sum of the totals -> bucket
//                     offsets.
//   3. scatter      : same recode, the block's slots come from shared-memory cursors (bucket offset
//                     + tile slot) -> entries sorted by bucket (entry = point id | sign bit).
//                     No global atomics at all.
//                     (`sort_mode=0`, and bucket sets too large for shared memory: one thread per
//                     term with global atomics on the bucket counters, the first implementation.)
//   4. accumulate   : the hot kernel.  The sorted entry list is cut into equal chunks of L
//                     entries, one thread per chunk, so every lane of a warp performs the
//                     same number of mixed additions regardless of how skewed the digit
//                     histogram is.  Buckets that lie inside one chunk are written directly;
//                     pieces of buckets that straddle chunk borders go to head/tail slots.
//   5. fix-up       : one thread per straddling bucket folds its pieces; buckets with more
//                     than HEAVY pieces (skewed scalars: zeros, +-1) are folded by a whole
//                     block each.
//   6. bucket reduce: sum_b (b+1) * B_b per bucket set by segmented running sums + block tree
//                     (option reduce_mode=1: level by level, no per-thread scalar multiple).
//   7. finish       : per job, fold of the partials (Horner over the windows when every window
//                     has its own bucket set), one inversion, affine + compressed.
//
// With the precomputed window multiples of the SRS (MsmTables: level j of the point array holds
// 2^(c j) * P) all windows of a job share ONE bucket set and the window index selects the
// table level: W-fold fewer buckets to sort, reduce and fold, and no doubling tail.
#include <algorithm>

#include "msm_acc.cuh"
#include "scalar.cuh"

namespace sonic {

// ---- stage 1 / 3: digits -> histogram or scatter ----------------------------------------
template <bool SCATTER>
__global__ void __launch_bounds__(256) k_msm_digits(const uint32_t* __restrict__ scalars, MsmJobTable tab,
                                                    uint32_t n_tot, int c, int W, uint32_t B,
                                                    uint32_t level_stride,  // 0: one bucket set per window
                                                    uint32_t* __restrict__ counters,
                                                    uint32_t* __restrict__ entries) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tot) return;
    int j = 0;
    while (j + 1 < tab.M && tab.prefix[j + 1] <= t) ++j;
    const uint32_t local = t - tab.prefix[j];
    const uint4* sp = reinterpret_cast<const uint4*>(scalars + (size_t)(tab.job[j].scalar_off + local) * 8);
    uint4 s0 = __ldg(sp), s1 = __ldg(sp + 1);
    uint32_t s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) return;
    ScalarDigits sd(s, c);
    uint32_t pid = tab.job[j].point_base + local;
    // with precomputed multiples every window of a job feeds the same bucket set and the window
    // index selects the table level instead
    uint32_t gb0 = level_stride ? (uint32_t)j * B : (uint32_t)j * (uint32_t)W * B;
    const uint32_t gb_step = level_stride ? 0u : B;
    if (W <= 16) {
        // all digits first, then all atomics back to back (independent, so their latencies overlap)
        uint32_t gb[16], val[16];
#pragma unroll
        for (int w = 0; w < 16; ++w) {
            gb[w] = 0xffffffffu;
            if (w < W) {
                int32_t d = sd.next();
                if (d != 0) {
                    uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
                    gb[w] = gb0 + (uint32_t)w * gb_step + mag - 1;
                    val[w] = (pid + (uint32_t)w * level_stride) | (d < 0 ? 0x80000000u : 0u);
                }
            }
        }
        if (SCATTER) {
            uint32_t pos[16];
#pragma unroll
            for (int w = 0; w < 16; ++w) pos[w] = gb[w] != 0xffffffffu ? atomicAdd(&counters[gb[w]], 1u) : 0u;
#pragma unroll
            for (int w = 0; w < 16; ++w) if (gb[w] != 0xffffffffu) entries[pos[w]] = val[w];
        } else {
#pragma unroll
            for (int w = 0; w < 16; ++w) if (gb[w] != 0xffffffffu) atomicAdd(&counters[gb[w]], 1u);
        }
        return;
    }
    for (int w = 0; w < W; ++w, gb0 += gb_step, pid += level_stride) {
        int32_t d = sd.next();
        if (d == 0) continue;
        uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        uint32_t gb = gb0 + mag - 1;
        if (SCATTER) {
            uint32_t pos = atomicAdd(&counters[gb], 1u);
            entries[pos] = pid | (d < 0 ? 0x80000000u : 0u);
        } else {
            atomicAdd(&counters[gb], 1u);
        }
    }
}

// ---- stage 1 / 3, tiled: counting sort with per-tile histograms in shared memory ---------------
// grid = (tiles, bucket sets per job); a tile is a run of `tile_terms` consecutive terms of ONE job.
// hist_g is tile-major: [(tile * sets + set) * B + bucket].
// COUNT pass: hist_g receives the tile's bucket counts.  k_msm_tile_prefix then turns every count
// into the tile's first slot inside its bucket; the SCATTER pass loads offset + slot as cursors.
constexpr int SORT_THREADS = 1024;
struct MsmTileTable {
    uint32_t tile_prefix[MSM_MAX_JOBS + 1];
    uint32_t tile_terms;
};

template <bool SCATTER>
__global__ void __launch_bounds__(SORT_THREADS)
k_msm_sort_tiles(const uint32_t* __restrict__ scalars, MsmJobTable tab, MsmTileTable tt, int c, int W, uint32_t B,
                 uint32_t level_stride,  // 0: one bucket set per window (blockIdx.y = window)
                 uint32_t* __restrict__ hist_g, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ entries,
                 uint32_t* __restrict__ reserve) {   // non-null: bucket totals (COUNT) / bucket cursors (SCATTER) shared by all tiles
    extern __shared__ uint32_t hist[];
    const uint32_t tile = blockIdx.x, set = blockIdx.y, sets = gridDim.y;
    int j = 0;
    while (j + 1 < tab.M && tt.tile_prefix[j + 1] <= tile) ++j;
    const uint32_t first = (tile - tt.tile_prefix[j]) * tt.tile_terms;
    const uint32_t nj = tab.job[j].n;
    const uint32_t cnt = nj - first < tt.tile_terms ? nj - first : tt.tile_terms;
    uint32_t* g = hist_g + ((size_t)tile * sets + set) * B;
    const uint32_t* off = SCATTER && !reserve ? offsets + ((size_t)j * sets + set) * B : nullptr;
    uint32_t* res = reserve ? reserve + ((size_t)j * sets + set) * B : nullptr;
    if (SCATTER && res) {
        // the tile takes its slots of every bucket from the bucket's cursor: which tile comes first inside a bucket is
        // left to the order of arrival (a sum does not care)
        for (uint32_t b = threadIdx.x; b < B; b += SORT_THREADS) {
            const uint32_t h = g[b];
            hist[b] = h ? atomicAdd(&res[b], h) : 0u;
        }
    } else {
        for (uint32_t b = threadIdx.x; b < B; b += SORT_THREADS) hist[b] = SCATTER ? g[b] + off[b] : 0u;
    }
    __syncthreads();
    const uint4* sp = reinterpret_cast<const uint4*>(scalars + (size_t)(tab.job[j].scalar_off + first) * 8);
    const uint32_t pid0 = tab.job[j].point_base + first;
    for (uint32_t i = threadIdx.x; i < cnt; i += SORT_THREADS) {
        const uint4 s0 = __ldg(sp + 2 * (size_t)i), s1 = __ldg(sp + 2 * (size_t)i + 1);
        uint32_t s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
        if ((s[0] | s[1] | s[2] | s[3] | s[4] | s[5] | s[6] | s[7]) == 0) continue;
        ScalarDigits sd(s, c);
        const uint32_t pid = pid0 + i;
        if (level_stride) {
            // precomputed multiples: every window feeds the job's single bucket set, the window
            // index selects the table level
            for (int w = 0; w < W; ++w) {
                const int32_t d = sd.next();
                if (d == 0) continue;
                const uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
                const uint32_t pos = atomicAdd(&hist[mag - 1], 1u);
                if (SCATTER) entries[pos] = (pid + (uint32_t)w * level_stride) | (d < 0 ? 0x80000000u : 0u);
            }
        } else {
            int32_t d = 0;
            for (uint32_t w = 0; w <= set; ++w) d = sd.next();  // the carry chain runs through the lower windows
            if (d == 0) continue;
            const uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
            const uint32_t pos = atomicAdd(&hist[mag - 1], 1u);
            if (SCATTER) entries[pos] = pid | (d < 0 ? 0x80000000u : 0u);
        }
    }
    if (!SCATTER) {
        __syncthreads();
        for (uint32_t b = threadIdx.x; b < B; b += SORT_THREADS) {
            const uint32_t h = hist[b];
            g[b] = h;
            if (res && h) atomicAdd(&res[b], h);
        }
    }
}

// thread per global bucket: counts of the job's tiles -> each tile's first slot inside the bucket
// (in place), and the bucket's total.  Consecutive threads read consecutive words of every tile.
__global__ void __launch_bounds__(256)
k_msm_tile_prefix(uint32_t* __restrict__ hist_g, MsmTileTable tt, uint32_t per_job /* sets * B */, uint32_t GB,
                  uint32_t* __restrict__ counts) {
    const uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x;
    if (gb > GB) return;
    if (gb == GB) { counts[gb] = 0; return; }
    const uint32_t j = gb / per_job, within = gb - j * per_job;
    const uint32_t t0 = tt.tile_prefix[j], t1 = tt.tile_prefix[j + 1];
    uint32_t* h = hist_g + (size_t)t0 * per_job + within;
    uint32_t run = 0;
    uint32_t t = t0;
    // eight independent loads in flight per thread before the (in-place) stores
    for (; t + 8 <= t1; t += 8, h += 8 * (size_t)per_job) {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = h[i * (size_t)per_job];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            h[i * (size_t)per_job] = run;
            run += v[i];
        }
    }
    for (; t < t1; ++t, h += per_job) {
        const uint32_t v = *h;
        *h = run;
        run += v;
    }
    counts[gb] = run;
}

// ---- stage 2: exclusive scan of u32 (three-kernel, 4096 items per block) ------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tile(const uint32_t* __restrict__ in, uint32_t* __restrict__ out,
                                                            uint32_t n, uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
    uint32_t v[SCAN_ITEMS];
    uint32_t sum = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0u;
        sum += v[i];
    }
    // inclusive warp scan of the per-thread sums
    uint32_t inc = sum;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_sums[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        uint32_t ws = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0u;
#pragma unroll
        for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, ws, o);
            if (lane >= o) ws += y;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = ws;
    }
    __syncthreads();
    uint32_t excl = inc - sum + (wid ? warp_sums[wid - 1] : 0u);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = excl;
        excl += v[i];
    }
    if (threadIdx.x == SCAN_THREADS - 1 && tile_sums) tile_sums[blockIdx.x] = excl;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(uint32_t* __restrict__ data, uint32_t n,
                                                           const uint32_t* __restrict__ tile_offsets) {
    const uint32_t add = tile_offsets[blockIdx.x];
    const uint32_t base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) data[base + i] += add;
}

// out may alias in
void exclusive_scan_u32(Arena& ar, const uint32_t* in, uint32_t* out, uint32_t n) {
    const unsigned tiles = div_up(n, SCAN_TILE);
    if (tiles == 1) {
        SONIC_LAUNCH(k_scan_tile, 1, SCAN_THREADS, 0, in, out, n, (uint32_t*)nullptr);
        return;
    }
    uint32_t* sums = ar.get<uint32_t>(tiles);
    SONIC_LAUNCH(k_scan_tile, tiles, SCAN_THREADS, 0, in, out, n, sums);
    exclusive_scan_u32(ar, sums, sums, tiles);
    SONIC_LAUNCH(k_scan_add, tiles, SCAN_THREADS, 0, out, n, sums);
}

// ---- stage 4: chunked bucket accumulation (the hot kernel): msm_acc_compact.cu / msm_acc_regs.cu ----
// ---- stages 5-7 (fix-up, bucket reduction, finish): msm_reduce.cu ----

// ---- host side ---------------------------------------------------------------------------

static MsmPlan msm_plan(const Ctx& cx, uint32_t n_tot, int M, const MsmTables& tables) {
    MsmPlan p;
    int best_c = 8;
    double best = 1e300;
    for (int c = 4; c <= 20; ++c) {
        int W = msm_num_windows(c);
        if (W > 64) continue;
        double B = double(1u << (c - 1));
        double buckets = double(M) * W * B;
        if (buckets > 3.0e8) continue;
        // mixed add = 1.0, full add = 1.4; reduction ~ 2 adds + fix-up traffic per bucket
        double cost = double(n_tot) * W + buckets * (2 * 1.4 + 0.6);
        if (cost < best) { best = cost; best_c = c; }
    }
    p.c = cx.opt_window_bits > 0 ? cx.opt_window_bits : best_c;
    if (p.c < 4) p.c = 4;
    if (p.c > 20) p.c = 20;
    if (tables.c > 0) p.c = tables.c;  // the table levels fix the window
    p.W = msm_num_windows(p.c);
    p.B = 1u << (p.c - 1);
    p.sets = tables.c > 0 ? 1 : p.W;   // bucket sets per job
    p.GB = (uint32_t)M * p.sets * p.B;
    p.n_tot = n_tot;
    // chunk length: every thread performs exactly L additions, so the grid runs in lock-step waves of
    // `resident` threads (4 blocks x 128 threads per SM); L <= 64 is picked so that the chunks fill a
    // whole number of waves (a proof sharded 8 ways has ~4 waves: a half-empty fifth cost 10 %)
    uint64_t entries = (uint64_t)n_tot * p.W;
    const uint64_t resident = (uint64_t)cx.sm_count * 512;
    const uint64_t Lmax = (uint64_t)(cx.opt_chunk_max > 0 ? cx.opt_chunk_max : 64);
    const uint64_t waves = std::max<uint64_t>(1, (entries + resident * Lmax - 1) / (resident * Lmax));
    uint64_t L = (entries + resident * waves - 1) / (resident * waves);
    if (L < 8) L = 8;
    if (L > Lmax) L = Lmax;
    p.L = cx.opt_chunk > 0 ? (uint32_t)cx.opt_chunk : (uint32_t)L;
    return p;
}

// d_scalars: canonical little-endian scalars, 8 words each.  Results: one affine point
// (Montgomery form) and/or one 48-byte compressed encoding per job, in device memory.
void msm_run(Ctx& cx, const G1Affine* d_points, const MsmTables& tables, const uint32_t* d_scalars,
             const std::vector<MsmJob>& jobs, G1Affine* d_out_aff, uint8_t* d_out_comp, const MsmSync* sync) {
    const bool second = sync && sync->second;
    // stage events: 0 sort begins, 1 sort done, 2 fix-up done, 3 finish done, 4 / 5 around the accumulate kernel
    cudaEvent_t* E = second ? &cx.ev[16] : nullptr;
    auto mark = [&](int k) {
        static const int first_half[6] = {0, 1, 2, 3, 8, 9};
        SONIC_CUDA(cudaEventRecord(second ? E[k] : cx.ev[first_half[k]], cx.stream));
    };
    if (!second) cx.msm_second_half = false;
    const int M = (int)jobs.size();
    if (M == 0) return;
    if (M > MSM_MAX_JOBS) throw CudaError{cudaErrorInvalidValue, "too many MSM jobs", __LINE__};
    MsmJobTable tab;
    memset(&tab, 0, sizeof tab);
    tab.M = M;
    uint64_t n_tot64 = 0;
    for (int i = 0; i < M; ++i) {
        tab.job[i] = jobs[i];
        tab.prefix[i] = (uint32_t)n_tot64;
        n_tot64 += jobs[i].n;
    }
    tab.prefix[M] = (uint32_t)n_tot64;
    if (n_tot64 >= (1ull << 31)) throw CudaError{cudaErrorInvalidValue, "MSM batch too large", __LINE__};
    const uint32_t n_tot = (uint32_t)n_tot64;
    Arena& ar = cx.arena;
    MsmPlan p = msm_plan(cx, n_tot, M, tables);
    const uint32_t level_stride = tables.c > 0 ? tables.stride : 0u;
    if ((uint64_t)n_tot * p.W >= (1ull << 32)) throw CudaError{cudaErrorInvalidValue, "MSM batch too large", __LINE__};
    cudaStream_t st = cx.stream;

    mark(0);
    uint32_t* offsets = ar.get<uint32_t>((size_t)p.GB + 1);
    // zero digits are not stored, so n_tot*W is only an upper bound of the entry count; the
    // accumulate grid is sized for the bound and reads the true count from offsets[GB]
    const uint64_t total_max = (uint64_t)n_tot * p.W;
    uint32_t* entries = ar.get<uint32_t>(total_max ? total_max : 1);
    // Tiles never cross a job border.  Blocks start in index order, so the tiles in flight belong to
    // a few consecutive jobs; with many small tiles per job (about ten per SM in total) the slice of
    // the entry array being scattered into at any moment (those jobs' entries) stays inside L2 and
    // the 4-byte stores merge into whole sectors before they reach HBM.  With two tiles per SM the
    // scatter moved 5.7 GB through DRAM for 0.48 GB of entries (profiles/r01e).
    MsmTileTable tt;
    memset(&tt, 0, sizeof tt);
    const uint32_t target_tiles = (uint32_t)cx.sm_count * (cx.opt_sort_mode >= 2 ? (uint32_t)cx.opt_sort_mode : 10u);
    tt.tile_terms = div_up(n_tot ? n_tot : 1, target_tiles > (uint32_t)M ? target_tiles - (uint32_t)M : 1);
    if (tt.tile_terms < 2048) tt.tile_terms = 2048;
    for (int i = 0; i < M; ++i) tt.tile_prefix[i + 1] = tt.tile_prefix[i] + div_up(jobs[i].n, tt.tile_terms);
    const uint32_t tiles = tt.tile_prefix[M];
    const uint64_t hist_len = (uint64_t)tiles * p.sets * p.B;
    const size_t sort_smem = (size_t)p.B * sizeof(uint32_t);
    const bool tiled = cx.opt_sort_mode != 0 && n_tot > 0 && sort_smem <= (size_t(128) << 10) && hist_len <= (1ull << 26);  // <= 256 MB of histograms
    if (!second) cx.timing_ms["msm.sort_tiles"] = tiled ? tiles : 0;
    if (tiled) {
        if (sort_smem > (size_t(48) << 10)) {  // opt-in above 48 KB; a host-side attribute, set per launch so that it follows the device
            SONIC_CUDA(cudaFuncSetAttribute(k_msm_sort_tiles<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 << 10));
            SONIC_CUDA(cudaFuncSetAttribute(k_msm_sort_tiles<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 << 10));
        }
        uint32_t* hist = ar.get<uint32_t>(hist_len);
        const dim3 grid(tiles, (unsigned)p.sets);
        if (cx.opt_sort_reserve) {
            // bucket totals by atomic adds of the tiles' counts; after the scan every tile reserves its slots of a bucket
            // from the bucket's cursor (no pass over the tile histograms in between: k_msm_tile_prefix was 0.33 ms)
            uint32_t* cursors = ar.get<uint32_t>((size_t)p.GB + 1);
            SONIC_CUDA(cudaMemsetAsync(offsets, 0, ((size_t)p.GB + 1) * 4, st));
            SONIC_LAUNCH(k_msm_sort_tiles<false>, grid, SORT_THREADS, sort_smem, d_scalars, tab, tt, p.c, p.W, p.B, level_stride, hist, (const uint32_t*)nullptr, (uint32_t*)nullptr, offsets);
            exclusive_scan_u32(ar, offsets, offsets, p.GB + 1);
            SONIC_CUDA(cudaMemcpyAsync(cursors, offsets, ((size_t)p.GB + 1) * 4, cudaMemcpyDeviceToDevice, st));
            SONIC_LAUNCH(k_msm_sort_tiles<true>, grid, SORT_THREADS, sort_smem, d_scalars, tab, tt, p.c, p.W, p.B, level_stride, hist, (const uint32_t*)nullptr, entries, cursors);
        } else {
            SONIC_LAUNCH(k_msm_sort_tiles<false>, grid, SORT_THREADS, sort_smem, d_scalars, tab, tt, p.c, p.W, p.B, level_stride, hist, (const uint32_t*)nullptr, (uint32_t*)nullptr, (uint32_t*)nullptr);
            SONIC_LAUNCH(k_msm_tile_prefix, div_up((uint64_t)p.GB + 1, 256), 256, 0, hist, tt, (uint32_t)p.sets * p.B, p.GB, offsets);
            exclusive_scan_u32(ar, offsets, offsets, p.GB + 1);
            SONIC_LAUNCH(k_msm_sort_tiles<true>, grid, SORT_THREADS, sort_smem, d_scalars, tab, tt, p.c, p.W, p.B, level_stride, hist, offsets, entries, (uint32_t*)nullptr);
        }
    } else {
        uint32_t* cursors = ar.get<uint32_t>((size_t)p.GB + 1);
        SONIC_CUDA(cudaMemsetAsync(offsets, 0, ((size_t)p.GB + 1) * 4, st));
        if (n_tot) SONIC_LAUNCH(k_msm_digits<false>, div_up(n_tot, 256), 256, 0, d_scalars, tab, n_tot, p.c, p.W, p.B, level_stride, offsets, (uint32_t*)nullptr);
        exclusive_scan_u32(ar, offsets, offsets, p.GB + 1);
        SONIC_CUDA(cudaMemcpyAsync(cursors, offsets, ((size_t)p.GB + 1) * 4, cudaMemcpyDeviceToDevice, st));
        if (n_tot) SONIC_LAUNCH(k_msm_digits<true>, div_up(n_tot, 256), 256, 0, d_scalars, tab, n_tot, p.c, p.W, p.B, level_stride, cursors, entries);
    }
    mark(1);

    // acc_mode 3 (automatic, the default): the affine stage with batched inversions for batches of at least ~2^24.3 entries
    // (a whole n = 2^16 proof 43.4 against 49.0 ms, a rank of four 13.6 against 14.1 ms, a 2^24-point MSM 75.4 against
    // 81.4 ms), the XYZZ chunk kernel below that (the rounds cost ~0.1 ms each in small kernels and barrier waits: a rank
    // of eight 8.3 against 8.0 ms, a 2^20-point MSM 7.5 against 6.9 ms).  What counts are the entries that exist -- zero
    // digits make none, and half of the scalars of the skewed sweep rows are zero -- so when the upper bound passes the
    // threshold the actual count is read back (4 bytes; the host has to wait for the sort, ~20 us of idle GPU).
    constexpr uint64_t AFFINE_MIN_ENTRIES = 5ull << 22;
    uint64_t entries_bound = total_max;
    bool affine = cx.opt_acc_mode == 2;
    if (cx.opt_acc_mode == 3 && tables.c > 0 && total_max >= AFFINE_MIN_ENTRIES) {
        uint32_t actual = 0;
        SONIC_CUDA(cudaMemcpyAsync(&actual, offsets + p.GB, 4, cudaMemcpyDeviceToHost, st));
        SONIC_CUDA(cudaStreamSynchronize(st));
        affine = actual >= AFFINE_MIN_ENTRIES;
        if (affine) entries_bound = actual;
    }
    const uint32_t chunks = affine ? 0u : div_up(total_max, p.L);
    G1XYZZ* buckets = ar.get<G1XYZZ>(p.GB);
    G1XYZZ* head = ar.get<G1XYZZ>(chunks ? chunks : 1);
    G1XYZZ* tail = ar.get<G1XYZZ>(chunks ? chunks : 1);
    if (sync && sync->wait_before_acc) SONIC_CUDA(cudaStreamWaitEvent(st, sync->wait_before_acc, 0));
    mark(4);
    if (affine) {
        // mean entries per bucket of the longest job: its W digits land in `sets` bucket sets of B buckets
        double mean = 1.0;
        for (int i = 0; i < M; ++i) mean = std::max(mean, (double)jobs[i].n * p.W / ((double)p.sets * p.B));
        launch_accumulate_affine(cx, entries_bound, entries, offsets, p.GB, d_points, buckets, mean);
    } else if (chunks) {
        if (cx.opt_acc_mode != 0) launch_accumulate_compact(cx, chunks, entries, offsets, p.GB, p.L, d_points, buckets, head, tail);
        else launch_accumulate_regs(cx, chunks, entries, offsets, p.GB, p.L, d_points, buckets, head, tail);
    }
    mark(5);
    if (sync && sync->signal_after_acc) SONIC_CUDA(cudaEventRecord(sync->signal_after_acc, st));
    if (!second) {
        cx.timing_ms["msm.window_bits"] = p.c;
        cx.timing_ms["msm.windows"] = p.W;
        cx.timing_ms["msm.precomputed"] = tables.c > 0 ? 1 : 0;
        cx.timing_ms["msm.terms"] = n_tot;
        cx.timing_ms["msm.jobs"] = M;
        cx.timing_ms["msm.chunk"] = p.L;
        cx.timing_ms["msm.affine"] = affine ? 1 : 0;
        cx.timing_ms["msm.buckets"] = p.GB;
        cx.msm_offsets_total = offsets + p.GB;
    } else {
        cx.timing_ms["msm.terms"] += n_tot;
        cx.timing_ms["msm.jobs"] += M;
        cx.timing_ms["msm.buckets"] += p.GB;
        cx.msm_offsets_total2 = offsets + p.GB;
        cx.msm_second_half = true;
    }
    msm_reduce_stage(cx, p, M, offsets, chunks, buckets, head, tail, d_out_aff, d_out_comp, second ? E[2] : cx.ev[2], affine);
    mark(3);
}

// Stage times of the last batch.  When it ran as two overlapped halves (prove.cu) the stages are those of the pair:
// sort = the first half's (the second's is hidden), accumulate = first sort done -> second fix-up done, reduce = the
// second half's exposed tail, accumulate_kernel / entries / terms = both halves together.
void msm_collect_timing(Ctx& cx) {
    const bool two = cx.msm_second_half;
    cudaEvent_t* E = &cx.ev[16];
    float a = 0, b = 0, c = 0;
    if (cudaEventElapsedTime(&a, cx.ev[0], cx.ev[1]) == cudaSuccess &&
        cudaEventElapsedTime(&b, cx.ev[1], two ? E[2] : cx.ev[2]) == cudaSuccess &&
        cudaEventElapsedTime(&c, two ? E[2] : cx.ev[2], two ? E[3] : cx.ev[3]) == cudaSuccess) {
        cx.timing_ms["msm.sort"] = a;
        cx.timing_ms["msm.accumulate"] = b;
        cx.timing_ms["msm.reduce"] = c;
        cx.timing_ms["msm"] = a + b + c;
    }
    float k = 0, k2 = 0;
    if (cudaEventElapsedTime(&k, cx.ev[8], cx.ev[9]) == cudaSuccess) {
        if (two && cudaEventElapsedTime(&k2, E[4], E[5]) == cudaSuccess) k += k2;
        cx.timing_ms["msm.accumulate_kernel"] = k;
    }
    if (two) {
        float t1 = 0;
        if (cudaEventElapsedTime(&t1, cx.ev[2], cx.ev[3]) == cudaSuccess) cx.timing_ms["msm.reduce_hidden"] = t1;   // the first half's tail, under the second's accumulation
    }
    if (cx.msm_offsets_total) {
        uint32_t total = 0, total2 = 0;
        if (cudaMemcpy(&total, cx.msm_offsets_total, 4, cudaMemcpyDeviceToHost) == cudaSuccess) {
            if (two && cx.msm_offsets_total2) cudaMemcpy(&total2, cx.msm_offsets_total2, 4, cudaMemcpyDeviceToHost);
            cx.timing_ms["msm.entries"] = (double)total + (double)total2;
        }
    }
}


}  // namespace sonic
