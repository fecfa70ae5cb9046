// Device-side orchestration of the Sonic prover (one call per proof).
//
// Reference: `prove` (src/Sonic/Protocol.hs:47-109) and `hscProve`
// (src/Sonic/Signature.hs:32-72).  The reference builds bivariate sparse polynomials and
// evaluates one variable at a time; every polynomial it ever commits to or opens is
// univariate, so the device builds those univariate vectors directly:
//
//   r'(X,1)   Constraints.hs:23-31 + Protocol.hs:58-62      dense over X^{-2n-4..n}
//   s(X,y)    Constraints.hs:34-53 then Utils.hs:20-21      dense over X^{-n..2n}
//   s(u,Y)    Constraints.hs:34-53 then Utils.hs:17-18      dense over Y^{-n..n+Q}
//   t(X,y)    Constraints.hs:56-65 then Utils.hs:20-21      = r'(X,1) * (r'(X,y) + s(X,y)) - k(y),
//             one cyclic convolution of length 2^ceil(log2(7n+9)) by NTT, dense over X^{-4n-8..3n}
//
// All 2Q+8 random field elements arrive up front in the reference's draw order (they are
// sampled, not derived from the transcript), so the whole proof is one call: the Fr work
// first, then every commitment and opening of the proof as ONE batched MSM launch.
#include <algorithm>
#include <vector>

#include <nvtx3/nvToolsExt.h>

#include "internal.h"

namespace sonic {

// ---- vector builders -----------------------------------------------------------------------
// r'(X,1): slot k <-> exponent k - 2n - 4
__global__ void __launch_bounds__(256) k_build_r(const Fr* __restrict__ aL, const Fr* __restrict__ aR, const Fr* __restrict__ aO,
                                                 const Fr* __restrict__ cns, uint32_t n, Fr* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 3 * n + 5) return;
    const int64_t e = (int64_t)k - 2 * (int64_t)n - 4;
    Fr v = Fr::zero();
    if (e >= 1) v = aL[e - 1];                                   // a_i X^i
    else if (e <= -1 && e >= -(int64_t)n) v = aR[-e - 1];        // b_i X^-i
    else if (e < -(int64_t)n && e >= -2 * (int64_t)n) v = aO[-e - n - 1];  // c_i X^(-i-n)
    else if (e < -2 * (int64_t)n) v = cns[-e - 2 * (int64_t)n - 1];        // c_(n+i) X^(-2n-i)
    out[k] = v;
}

// s(X,y) for several y at once: grid.y selects the evaluation point.
// out[b][k], slot k <-> exponent k - n, k in [0, 3n]
__global__ void __launch_bounds__(128) k_build_sxy(const Fr* __restrict__ wL, const Fr* __restrict__ wR, const Fr* __restrict__ wO,
                                                   uint32_t n, uint32_t Q, const Fr* __restrict__ tabs, uint64_t tl,
                                                   const uint32_t* __restrict__ fwd_idx, const uint32_t* __restrict__ inv_idx,
                                                   const uint32_t* __restrict__ slot_idx, Fr* __restrict__ out) {
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= n) return;
    const uint32_t b = blockIdx.y;
    const Fr* yt = tabs + (size_t)fwd_idx[b] * tl;   // y^k
    const Fr* yi = tabs + (size_t)inv_idx[b] * tl;   // y^-k
    const uint32_t i = i0 + 1;
    Fr su = Fr::zero(), sv = Fr::zero(), sw = Fr::zero();
    for (uint32_t q = 1; q <= Q; ++q) {
        const Fr yp = yt[n + q];
        const size_t w = (size_t)(q - 1) * n + i0;
        su = fp_add(su, fp_mul(yp, wL[w]));
        sv = fp_add(sv, fp_mul(yp, wR[w]));
        sw = fp_add(sw, fp_mul(yp, wO[w]));
    }
    sw = fp_sub(fp_sub(sw, yt[i]), yi[i]);
    Fr* o = out + (size_t)slot_idx[b] * (3 * (size_t)n + 1);
    o[n - i] = su;        // u_i(y) X^-i
    o[n + i] = sv;        // v_i(y) X^i
    o[2 * n + i] = sw;    // w_i(y) X^(i+n)
    if (i0 == 0) o[n] = Fr::zero();
}

// r'(X,y) + s(X,y): slot k <-> exponent k - 2n - 4, k in [0, 4n+4]
__global__ void __launch_bounds__(256) k_build_rs(const Fr* __restrict__ rX1, const Fr* __restrict__ sXy, const Fr* __restrict__ yt,
                                                  const Fr* __restrict__ yi, uint32_t n, Fr* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= 4 * n + 5) return;
    const int64_t e = (int64_t)k - 2 * (int64_t)n - 4;
    Fr v = Fr::zero();
    if (e <= (int64_t)n) {
        const Fr p = e >= 0 ? yt[e] : yi[-e];
        v = fp_mul(rX1[k], p);  // r'(X,y) = r'(Xy,1)
    }
    if (e >= -(int64_t)n) v = fp_add(v, sXy[e + n]);
    out[k] = v;
}

// t(X,y) slot 4n+8 (X^0) -= k(y), k(y) = sum_q cs[q] y^(n+1+q)   (Constraints.hs:65,67-68)
__global__ void k_t_fix(Fr* __restrict__ t, const Fr* __restrict__ cs, const Fr* __restrict__ yt, uint32_t n, uint32_t Q) {
    if (threadIdx.x || blockIdx.x) return;
    Fr ky = Fr::zero();
    for (uint32_t q = 0; q < Q; ++q) ky = fp_add(ky, fp_mul(cs[q], yt[n + 1 + q]));
    t[4 * n + 8] = fp_sub(t[4 * n + 8], ky);
}

// s(u,Y): slot k <-> exponent k - n, k in [0, 2n+Q].  +-i slots: -u^(i+n)
__global__ void __launch_bounds__(256) k_build_suy_pm(const Fr* __restrict__ ut, uint32_t n, uint32_t Q, Fr* __restrict__ out) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k > 2 * n) return;
    if (k == n) { out[k] = Fr::zero(); return; }
    const uint32_t i = k > n ? k - n : n - k;
    out[k] = fp_neg(ut[i + n]);
    (void)Q;
}

// Y^(n+q) slots: sum_i u^-i wL[q][i] + u^i wR[q][i] + u^(i+n) wO[q][i].
// grid = (SUY_PARTS, Q): every block sums a slice of i into partial[q][part]; k_build_suy_fin folds.
constexpr int SUY_PARTS = 32;
__global__ void __launch_bounds__(256) k_build_suy_dot(const Fr* __restrict__ wL, const Fr* __restrict__ wR, const Fr* __restrict__ wO,
                                                       uint32_t n, const Fr* __restrict__ ut, const Fr* __restrict__ ui,
                                                       Fr* __restrict__ partial) {
    __shared__ Fr smem[8];
    const uint32_t q = blockIdx.y;  // 0-based
    Fr acc = Fr::zero();
    for (uint32_t i0 = blockIdx.x * 256 + threadIdx.x; i0 < n; i0 += 256 * SUY_PARTS) {
        const uint32_t i = i0 + 1;
        const size_t w = (size_t)q * n + i0;
        acc = fp_add(acc, fp_mul(ui[i], wL[w]));
        acc = fp_add(acc, fp_mul(ut[i], wR[w]));
        acc = fp_add(acc, fp_mul(ut[i + n], wO[w]));
    }
    // block reduction (8 warps)
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Fr y;
#pragma unroll
        for (int k = 0; k < 8; ++k) y.l[k] = __shfl_down_sync(0xffffffffu, acc.l[k], o);
        acc = fp_add(acc, y);
    }
    if (lane == 0) smem[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) acc = fp_add(acc, smem[w]);
        partial[q * SUY_PARTS + blockIdx.x] = acc;
    }
}

__global__ void k_build_suy_fin(const Fr* __restrict__ partial, uint32_t n, uint32_t Q, Fr* __restrict__ out) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    Fr acc = partial[q * SUY_PARTS];
    for (int k = 1; k < SUY_PARTS; ++k) acc = fp_add(acc, partial[q * SUY_PARTS + k]);
    out[2 * n + 1 + q] = acc;  // exponent n + (q+1)
}

// ---- sparse (CSR / CSC) variants of the two s-builders ------------------
struct SparseView {
    const uint32_t* ptr;  // col_ptr (CSC) or row_ptr (CSR)
    const uint32_t* idx;  // row (CSC) or col (CSR)
    const Fr* val;
};

// s(X,y) from column-major weights: thread per gate i, batched over the evaluation points
__global__ void __launch_bounds__(128) k_build_sxy_csc(SparseView L, SparseView R_, SparseView O, uint32_t n,
                                                       const Fr* __restrict__ tabs, uint64_t tl,
                                                       const uint32_t* __restrict__ fwd_idx, const uint32_t* __restrict__ inv_idx,
                                                       const uint32_t* __restrict__ slot_idx, Fr* __restrict__ out) {
    const uint32_t i0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (i0 >= n) return;
    const uint32_t b = blockIdx.y;
    const Fr* yt = tabs + (size_t)fwd_idx[b] * tl;
    const Fr* yi = tabs + (size_t)inv_idx[b] * tl;
    const uint32_t i = i0 + 1;
    Fr su = Fr::zero(), sv = Fr::zero(), sw = Fr::zero();
    for (uint32_t k = L.ptr[i0]; k < L.ptr[i0 + 1]; ++k) su = fp_add(su, fp_mul(yt[n + 1 + L.idx[k]], L.val[k]));
    for (uint32_t k = R_.ptr[i0]; k < R_.ptr[i0 + 1]; ++k) sv = fp_add(sv, fp_mul(yt[n + 1 + R_.idx[k]], R_.val[k]));
    for (uint32_t k = O.ptr[i0]; k < O.ptr[i0 + 1]; ++k) sw = fp_add(sw, fp_mul(yt[n + 1 + O.idx[k]], O.val[k]));
    sw = fp_sub(fp_sub(sw, yt[i]), yi[i]);
    Fr* o = out + (size_t)slot_idx[b] * (3 * (size_t)n + 1);
    o[n - i] = su;
    o[n + i] = sv;
    o[2 * n + i] = sw;
    if (i0 == 0) o[n] = Fr::zero();
}

// the Y^(n+q) coefficients of s(u,Y) from row-major weights; grid = (SUY_PARTS, Q)
__global__ void __launch_bounds__(256) k_build_suy_dot_csr(SparseView L, SparseView R_, SparseView O, uint32_t n,
                                                           const Fr* __restrict__ ut, const Fr* __restrict__ ui,
                                                           Fr* __restrict__ partial) {
    __shared__ Fr smem[8];
    const uint32_t q = blockIdx.y;
    const uint32_t stride = 256 * SUY_PARTS, first = blockIdx.x * 256 + threadIdx.x;
    Fr acc = Fr::zero();
    for (uint32_t k = L.ptr[q] + first; k < L.ptr[q + 1]; k += stride) acc = fp_add(acc, fp_mul(ui[L.idx[k] + 1], L.val[k]));
    for (uint32_t k = R_.ptr[q] + first; k < R_.ptr[q + 1]; k += stride) acc = fp_add(acc, fp_mul(ut[R_.idx[k] + 1], R_.val[k]));
    for (uint32_t k = O.ptr[q] + first; k < O.ptr[q + 1]; k += stride) acc = fp_add(acc, fp_mul(ut[O.idx[k] + 1 + n], O.val[k]));
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Fr y;
#pragma unroll
        for (int k = 0; k < 8; ++k) y.l[k] = __shfl_down_sync(0xffffffffu, acc.l[k], o);
        acc = fp_add(acc, y);
    }
    if (lane == 0) smem[wid] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) acc = fp_add(acc, smem[w]);
        partial[q * SUY_PARTS + blockIdx.x] = acc;
    }
}

// derived scalars: pts[] = the evaluation points (Montgomery), pts[np..2np) their inverses
// rnd_m: 2M+8 draws (Montgomery), M = number of (y_j, z_j) pairs.  Layout of pts: y, z, yz, y_1..y_M, z_1..z_M, u, v
__global__ void k_prove_points(const Fr* __restrict__ rnd_m, uint32_t M, int has_main, const uint32_t* __restrict__ sel, Fr* __restrict__ pts) {
    const uint32_t np = 2 * M + 5;
    const uint32_t t = sel[blockIdx.x];   // one single-thread block per point: the Euclid inverse branches on the data
    if (t >= np || threadIdx.x != 0) return;
    Fr v;
    if (t == 0) v = has_main ? rnd_m[4] : Fr::one();
    else if (t == 1) v = has_main ? rnd_m[5] : Fr::one();
    else if (t == 2) v = has_main ? fp_mul(rnd_m[4], rnd_m[5]) : Fr::one();
    else v = rnd_m[6 + (t - 3)];  // ys, zs, u, v are contiguous in the draw order
    pts[t] = v;
    pts[np + t] = fp_inv_euclid(v);
}

// first index in [a, b) (relative to s) whose scalar is non-zero -> atomicMin into *out
__global__ void __launch_bounds__(256) k_first_nonzero_job(const Fr* __restrict__ s, uint32_t a, uint32_t b, uint32_t* __restrict__ out) {
    const uint32_t i = a + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    if (!s[i].is_zero()) atomicMin(out, i);
}

}  // namespace sonic

namespace sonic {
// resident circuit (weights in Montgomery form), one replica per device
// One sparse weight matrix on the device, both ways: by row (CSR: the Y^(n+q) dot products of
// s(u,Y) walk rows) and by column (CSC: every X-coefficient of s(X,y) sums one column).
struct SparseMat {
    uint32_t* row_ptr = nullptr;  // Q+1
    uint32_t* col = nullptr;      // nnz, gate index (0-based)
    sonic::Fr* val = nullptr;     // nnz, Montgomery, CSR order
    uint32_t* col_ptr = nullptr;  // n+1
    uint32_t* row = nullptr;      // nnz, constraint index (0-based)
    sonic::Fr* cval = nullptr;    // nnz, Montgomery, CSC order
    uint64_t nnz = 0;
};

struct CircuitRep {
    uint64_t n = 0, Q = 0;
    bool sparse = false;
    sonic::Fr* w = nullptr;   // dense: wL | wR | wO, each Q*n row-major, then cs (Q); sparse: cs only
    SparseMat sp[3];          // sparse: wL, wR, wO
    char* blob = nullptr;     // one allocation behind the sparse arrays
    sonic::Fr* wL() const { return w; }
    sonic::Fr* wR() const { return w + Q * n; }
    sonic::Fr* wO() const { return w + 2 * Q * n; }
    sonic::Fr* cs() const { return sparse ? w : w + 3 * Q * n; }
};

int circuit_load(Ctx& cx, uint64_t n, uint64_t Q, const uint8_t* wL, const uint8_t* wR, const uint8_t* wO,
                 const uint8_t* cs, CircuitRep** out) {
    const uint64_t m = Q * n;
    CircuitRep* c = new CircuitRep;
    c->n = n;
    c->Q = Q;
    cudaError_t e = cudaMalloc((void**)&c->w, (3 * m + Q) * sizeof(Fr));
    if (e != cudaSuccess) { delete c; throw CudaError{e, "cudaMalloc(circuit)", __LINE__}; }
    try {
        Fr* stage = cx.arena.get<Fr>(3 * m + Q);
        SONIC_CUDA(cudaMemcpyAsync(stage, wL, m * 32, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(stage + m, wR, m * 32, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(stage + 2 * m, wO, m * 32, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(stage + 3 * m, cs, Q * 32, cudaMemcpyHostToDevice, cx.stream));
        uint32_t* bad = cx.arena.get<uint32_t>(1);
        SONIC_CUDA(cudaMemsetAsync(bad, 0, 4, cx.stream));
        fr_to_mont(cx, stage, c->w, 3 * m + Q, bad);
        uint32_t h = 0;
        SONIC_CUDA(cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        if (h) {
            cudaFree(c->w);
            delete c;
            return fail(SONIC_ERR_NONCANONICAL, "a circuit weight is not a canonical residue (>= r)");
        }
    } catch (...) {
        cudaFree(c->w);
        delete c;
        throw;
    }
    *out = c;
    return SONIC_OK;
}

// Sparse load: host pointers to three CSR matrices (row_ptr as u64[Q+1], col as u32[nnz], val as
// nnz x 32 canonical bytes).  The column-major copy is built here on the host (index shuffling
// only; every field operation stays on the device).
int circuit_load_csr(Ctx& cx, uint64_t n, uint64_t Q, const uint64_t* const row_ptr[3], const uint32_t* const col[3],
                     const uint8_t* const val[3], const uint8_t* cs, CircuitRep** out) {
    uint64_t nnz[3], total = 0;
    for (int m = 0; m < 3; ++m) {
        if (row_ptr[m][0] != 0) return fail(SONIC_ERR_INVALID_ARG, "CSR row_ptr must start at 0");
        for (uint64_t q = 0; q < Q; ++q)
            if (row_ptr[m][q + 1] < row_ptr[m][q]) return fail(SONIC_ERR_INVALID_ARG, "CSR row_ptr must be non-decreasing");
        nnz[m] = row_ptr[m][Q];
        if (nnz[m] >= (1ull << 31)) return fail(SONIC_ERR_INVALID_ARG, "too many non-zeros");
        for (uint64_t k = 0; k < nnz[m]; ++k)
            if (col[m][k] >= n) return fail(SONIC_ERR_INVALID_ARG, "CSR column index %u out of range (n = %llu)", col[m][k], (unsigned long long)n);
        total += nnz[m];
    }
    // host staging: [per matrix: row_ptr u32 | col u32 | col_ptr u32 | row u32] then values (CSR order, CSC order) and cs
    const size_t idx_words = 3 * (Q + 1) + 3 * (n + 1) + 2 * total;
    std::vector<uint32_t> idx(idx_words);
    std::vector<uint8_t> vals((2 * total + Q) * 32);
    size_t iw = 0, vw = 0;
    size_t off_rowptr[3], off_col[3], off_colptr[3], off_row[3], off_val[3], off_cval[3];
    for (int m = 0; m < 3; ++m) {
        off_rowptr[m] = iw;
        for (uint64_t q = 0; q <= Q; ++q) idx[iw++] = (uint32_t)row_ptr[m][q];
        off_col[m] = iw;
        for (uint64_t k = 0; k < nnz[m]; ++k) idx[iw++] = col[m][k];
        // counting sort by column -> CSC
        off_colptr[m] = iw;
        uint32_t* cp = &idx[iw];
        iw += n + 1;
        for (uint64_t k = 0; k < nnz[m]; ++k) cp[col[m][k] + 1]++;
        for (uint64_t i = 0; i < n; ++i) cp[i + 1] += cp[i];
        off_row[m] = iw;
        uint32_t* rw = &idx[iw];
        iw += nnz[m];
        off_val[m] = vw;
        memcpy(&vals[vw * 32], val[m], nnz[m] * 32);
        vw += nnz[m];
        off_cval[m] = vw;
        std::vector<uint32_t> cursor(cp, cp + n);
        for (uint64_t q = 0; q < Q; ++q)
            for (uint64_t k = row_ptr[m][q]; k < row_ptr[m][q + 1]; ++k) {
                const uint32_t dst = cursor[col[m][k]]++;
                rw[dst] = (uint32_t)q;
                memcpy(&vals[(vw + dst) * 32], val[m] + k * 32, 32);
            }
        vw += nnz[m];
    }
    const size_t off_cs = vw;
    memcpy(&vals[vw * 32], cs, Q * 32);
    vw += Q;

    CircuitRep* c = new CircuitRep;
    c->n = n;
    c->Q = Q;
    c->sparse = true;
    const size_t idx_bytes = (idx_words * 4 + 255) & ~size_t(255);
    const size_t val_bytes = vw * sizeof(Fr);
    cudaError_t e = cudaMalloc((void**)&c->blob, idx_bytes + val_bytes);
    if (e != cudaSuccess) { delete c; throw CudaError{e, "cudaMalloc(circuit)", __LINE__}; }
    try {
        uint32_t* d_idx = (uint32_t*)c->blob;
        Fr* d_val = (Fr*)(c->blob + idx_bytes);
        Fr* stage = cx.arena.get<Fr>(vw ? vw : 1);
        SONIC_CUDA(cudaMemcpyAsync(d_idx, idx.data(), idx_words * 4, cudaMemcpyHostToDevice, cx.stream));
        SONIC_CUDA(cudaMemcpyAsync(stage, vals.data(), vw * 32, cudaMemcpyHostToDevice, cx.stream));
        uint32_t* bad = cx.arena.get<uint32_t>(1);
        SONIC_CUDA(cudaMemsetAsync(bad, 0, 4, cx.stream));
        fr_to_mont(cx, stage, d_val, vw, bad);
        uint32_t h = 0;
        SONIC_CUDA(cudaMemcpyAsync(&h, bad, 4, cudaMemcpyDeviceToHost, cx.stream));
        SONIC_CUDA(cudaStreamSynchronize(cx.stream));
        if (h) {
            cudaFree(c->blob);
            delete c;
            return fail(SONIC_ERR_NONCANONICAL, "a circuit weight is not a canonical residue (>= r)");
        }
        for (int m = 0; m < 3; ++m) {
            c->sp[m].row_ptr = d_idx + off_rowptr[m];
            c->sp[m].col = d_idx + off_col[m];
            c->sp[m].col_ptr = d_idx + off_colptr[m];
            c->sp[m].row = d_idx + off_row[m];
            c->sp[m].val = d_val + off_val[m];
            c->sp[m].cval = d_val + off_cval[m];
            c->sp[m].nnz = nnz[m];
        }
        c->w = d_val + off_cs;
    } catch (...) {
        cudaFree(c->blob);
        delete c;
        throw;
    }
    *out = c;
    return SONIC_OK;
}

void circuit_free(CircuitRep* c) {
    if (!c) return;
    if (c->sparse) { if (c->blob) cudaFree(c->blob); }
    else if (c->w) cudaFree(c->w);
    delete c;
}

namespace {

struct Window {  // a dense vector and the exponent of its slot 0
    Fr* mont = nullptr;
    Fr* canon = nullptr;
    int64_t lo = 0;
    uint32_t len = 0;
};

// One MSM of the proof, in the order the proof record lists its G1 fields.
struct ProofMsm {
    int family;
    const Fr* scal;   // canonical scalars, slot 0 <-> exponent lo (null: not built on this rank)
    int64_t lo;       // exponent of scalar 0 (already shifted for commits)
    uint32_t len;
    bool commit_text;
};

}  // namespace

// Proof bytes from the G1 encodings (record order, 48 B each) and the Fr values (record order).
static void assemble_proof(bool has_main, uint32_t M, const uint8_t* g48, const uint8_t* f32, uint8_t* out) {
    uint8_t* o = out;
    auto putG = [&]() { memcpy(o, g48, 48); o += 48; g48 += 48; };
    auto putF = [&]() { memcpy(o, f32, 32); o += 32; f32 += 32; };
    if (has_main) { putG(); putG(); putF(); putG(); putF(); putG(); putG(); putF(); }  // prR prT prA prWa prB prWb prWt prS
    for (uint32_t j = 0; j < M; ++j) { putG(); putF(); putG(); }                        // (S_j, (s_j, W_j))
    for (uint32_t j = 0; j < M; ++j) { putF(); putG(); putG(); }                        // (s'_j, W'_j, Q_j)
    putG(); putG(); putF(); putF();                                                     // hscQv hscC hscU hscV
}

// affine Montgomery -> raw 96 bytes (canonical little-endian x || y; infinity = zeros)
__global__ void k_points_to_raw(const G1Affine* __restrict__ pts, uint32_t n, Fq* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[2 * i] = fp_from_mont(pts[i].x);
    out[2 * i + 1] = fp_from_mont(pts[i].y);
}

// The fold of a sharded proof: `recs` holds one exchange record per rank (rec_bytes apart).
//   blocks 0..nm-1   out48[m] = compress(sum_r raw[r][m]).  When at most one rank contributes (an MSM
//                    owned whole by one rank: everybody else holds the identity) the point is already
//                    affine and is only re-encoded: no addition, no second inversion.  At most world-1
//                    MSMs are split between ranks and take the addition path.
//   block nm         the field values (exactly one rank computes each; the others hold zeros, so the
//                    bytes are OR-ed) and the status words (first violating index per MSM range = min
//                    over the ranks that scanned it; encoding flag = OR).
// Every rank that folds the same records reaches the same verdict.
__global__ void k_fold_records(const uint8_t* __restrict__ recs, size_t rec_bytes, uint32_t nm, uint32_t nv, uint32_t world,
                               size_t rec_vals, size_t rec_status, uint8_t* __restrict__ out, size_t out_vals, size_t out_status) {
    const uint32_t m = blockIdx.x;
    if (m == nm) {
        const uint32_t nw = nv * 8;
        for (uint32_t i = threadIdx.x; i < nw; i += blockDim.x) {
            uint32_t v = 0;
            for (uint32_t r = 0; r < world; ++r) v |= reinterpret_cast<const uint32_t*>(recs + r * rec_bytes + rec_vals)[i];
            reinterpret_cast<uint32_t*>(out + out_vals)[i] = v;
        }
        for (uint32_t i = threadIdx.x; i < 3 * nm + 3; i += blockDim.x) {   // flags, then srsD (the same on every rank)
            uint32_t v = i < 3 * nm ? 0xffffffffu : 0u;
            for (uint32_t r = 0; r < world; ++r) {
                const uint32_t w = reinterpret_cast<const uint32_t*>(recs + r * rec_bytes + rec_status)[i];
                v = i < 3 * nm ? (w < v ? w : v) : (v | w);
            }
            reinterpret_cast<uint32_t*>(out + out_status)[i] = v;
        }
        return;
    }
    if (threadIdx.x != 0) return;   // one thread per MSM: the Euclid inverse branches on the data
    uint32_t contributors = 0;
    G1Affine only = G1Affine::inf();
    for (uint32_t r = 0; r < world; ++r) {
        const Fq* p = reinterpret_cast<const Fq*>(recs + r * rec_bytes) + 2 * (size_t)m;
        if (!(p[0].is_zero() && p[1].is_zero())) {
            ++contributors;
            only.x = p[0];
            only.y = p[1];
        }
    }
    if (contributors <= 1) {
        if (contributors) { only.x = fp_to_mont(only.x); only.y = fp_to_mont(only.y); }
        g1_compress(only, out + (size_t)m * 48);
        return;
    }
    G1XYZZ acc = G1XYZZ::inf();
    for (uint32_t r = 0; r < world; ++r) {
        const Fq* p = reinterpret_cast<const Fq*>(recs + r * rec_bytes) + 2 * (size_t)m;
        G1Affine a;
        a.x = fp_to_mont(p[0]);
        a.y = fp_to_mont(p[1]);
        g1_madd(acc, a);  // (0,0) marks infinity and is skipped
    }
    g1_compress(g1_to_affine_single(acc), out + (size_t)m * 48);
}

void prove_fold_enqueue(Ctx& cx, const ProveLayout& lay, uint32_t world, const uint8_t* d_records, uint8_t* d_out) {
    (void)cx;
    SONIC_LAUNCH(k_fold_records, lay.nm + 1, 64, 0, d_records, lay.rec_bytes(), lay.nm, lay.nv, world, lay.rec_vals(), lay.rec_status(),
                 d_out, lay.out_vals(), lay.out_status());
}

// The status words decide, identically for every rank that holds the folded buffer, whether the call
// succeeds: an encoding >= r, or the reference's `index` panic for the first MSM in record order with
// a non-zero coefficient outside what the SRS holds (CommitmentScheme.hs:70-73; ascending exponents).
int prove_finish(const ProveLayout& lay, const uint8_t* h_out, const uint8_t* rnd_host, uint8_t* proof, uint64_t cap,
                 uint64_t* written) {
    const uint64_t need = lay.proof_bytes();
    if (written) *written = need;
    if (cap < need) return fail(SONIC_ERR_BUFFER_TOO_SMALL, "proof needs %llu bytes", (unsigned long long)need);
    const uint32_t* st = reinterpret_cast<const uint32_t*>(h_out + lay.out_status());
    const uint64_t d = (uint64_t)st[3 * (size_t)lay.nm + 1] | ((uint64_t)st[3 * (size_t)lay.nm + 2] << 32);
    if (st[3 * (size_t)lay.nm]) return fail(SONIC_ERR_NONCANONICAL, "an Fr encoding is not a canonical residue (>= r)");
    for (uint32_t i = 0; i < lay.nm; ++i)
        for (int k = 0; k < 3; ++k) {
            const uint32_t v = st[3 * (size_t)i + k];
            if (v == 0xffffffffu) continue;
            // the flag holds the offending exponent biased by 2^30; which MSMs are commitments follows from the record order
            const int64_t e = (int64_t)v - (int64_t)(1u << 30);
            bool commit;
            if (lay.has_main && i < 5) commit = i < 2;
            else {
                const uint32_t h = i - (lay.has_main ? 5u : 0u);
                commit = h < 2 * lay.M ? (h % 2 == 0) : (h == 4 * lay.M + 1);
            }
            if (commit) {
                if (e > 0) return fail(SONIC_ERR_SRS_TOO_SHORT, "commitPoly: gPositiveAlphaX is not long enough: %lld >= %llu", (long long)(e - 1), (unsigned long long)d);
                return fail(SONIC_ERR_SRS_TOO_SHORT, "commitPoly: gNegativeAlphaX is not long enough: %lld >= %llu", (long long)((e < 0 ? -e : e) - 1), (unsigned long long)d);
            }
            if (e >= 0) return fail(SONIC_ERR_SRS_TOO_SHORT, "openPoly: gPositiveX is not long enough: %lld >= %llu", (long long)e, (unsigned long long)(d + 1));
            return fail(SONIC_ERR_SRS_TOO_SHORT, "openPoly: gNegativeX is not long enough: %lld >= %llu", (long long)(-e - 1), (unsigned long long)d);
        }
    // field values in record order (Protocol.hs:28-38, Signature.hs:22-29): device values, then hscU, hscV from the draws
    std::vector<uint8_t> f32((size_t)lay.nF * 32);
    memcpy(f32.data(), h_out + lay.out_vals(), (size_t)lay.nv * 32);
    memcpy(f32.data() + (size_t)lay.nv * 32, rnd_host + 32 * (6 + 2 * (size_t)lay.M), 64);
    assemble_proof(lay.has_main, lay.M, h_out, f32.data(), proof);
    return SONIC_OK;
}

// first non-zero scalar of s[a, b) -> its exponent (lo + index) biased by 2^30, atomicMin into *out
__global__ void __launch_bounds__(256) k_first_violation(const Fr* __restrict__ s, uint32_t a, uint32_t b, int64_t lo, uint32_t* __restrict__ out) {
    const uint32_t i = a + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= b) return;
    if (!s[i].is_zero()) atomicMin(out, (uint32_t)(lo + (int64_t)i + (int64_t)(1u << 30)));
}

// The G1 half of a proof: range scans and the batched MSM over this rank's part [piece_lo, piece_hi) of every
// window.  world == 1 (`sharded` false): compressed results straight into the result buffer; else raw
// partial sums into the exchange record (the identity for MSMs this rank holds no part of).
static int enqueue_msms(Ctx& cx, SrsRep& srs, const std::vector<ProofMsm>& pm, const std::vector<int64_t>& piece_lo,
                        const std::vector<int64_t>& piece_hi, uint32_t* viol, bool sharded, uint8_t* d_result) {
    Arena& ar = cx.arena;
    cudaStream_t st = cx.stream;
    const int64_t d = (int64_t)srs.d;
    const uint32_t nm = (uint32_t)pm.size();
    // range / hole checks: first non-zero scalar outside what the SRS holds, per MSM, by the ranks that
    // built the vector (the fold takes the minimum, so every rank reaches the same verdict)
    std::vector<MsmJob> jobs(nm);
    std::vector<int64_t> slice_lo(nm, 0);
    for (uint32_t i = 0; i < nm; ++i) {
        const ProofMsm& m = pm[i];
        const int64_t lo = m.lo, hi = m.lo + (int64_t)m.len;
        struct Rng { int64_t a, b; } r[3] = {{0, 0}, {0, 0}, {0, 0}};
        if (lo < -d) r[0] = {lo, std::min(hi, -d)};
        if (m.family == SONIC_FAMILY_ALPHA && lo <= 0 && 0 < hi) r[1] = {0, 1};
        if (hi > d + 1) r[2] = {std::max(lo, d + 1), hi};
        for (int k = 0; k < 3; ++k) {
            if (r[k].b > r[k].a && m.scal) {
                const uint32_t a = (uint32_t)(r[k].a - lo), b = (uint32_t)(r[k].b - lo);
                SONIC_LAUNCH(k_first_violation, div_up(b - a, 256), 256, 0, m.scal, a, b, lo, viol + 3 * (size_t)i + k);
            }
        }
        // the part of this window inside the rank's run of terms (the whole clipped window when not sharded)
        const int64_t clo = piece_lo[i], chi = piece_hi[i];
        slice_lo[i] = clo;
        jobs[i].point_base = 0;
        jobs[i].n = (uint32_t)(chi - clo);
        jobs[i].scalar_off = 0;
        jobs[i].pad = 0;
    }
    // ---- bases: the full-range window tables, else tables restricted to this circuit size, else none ----
    const G1Affine* d_points = srs.points;
    MsmTables tables = srs.tables;
    bool restricted = false;
    if (srs.tables.c == 0 && srs.rt.points) {
        restricted = true;
        for (uint32_t i = 0; i < nm && restricted; ++i)
            if (jobs[i].n && srs.rt.find(pm[i].family, slice_lo[i], jobs[i].n) < 0) restricted = false;
        if (restricted) { d_points = srs.rt.points; tables = srs.rt.tables; }
    }
    for (uint32_t i = 0; i < nm; ++i) {
        if (!jobs[i].n) continue;
        jobs[i].point_base = restricted ? (uint32_t)srs.rt.find(pm[i].family, slice_lo[i], jobs[i].n) : (uint32_t)srs.index(pm[i].family, slice_lo[i]);
    }
    // all scalar vectors live in the arena; express them as offsets from the lowest address
    const Fr* sbase = nullptr;
    for (uint32_t i = 0; i < nm; ++i) {
        if (jobs[i].n == 0) continue;
        if (!pm[i].scal) return fail(SONIC_ERR_INVALID_ARG, "internal: MSM %u has terms on this rank but no scalars", i);
        if (!sbase || pm[i].scal < sbase) sbase = pm[i].scal;
    }
    for (uint32_t i = 0; i < nm; ++i)
        if (jobs[i].n) jobs[i].scalar_off = (uint32_t)((pm[i].scal - sbase) + (slice_lo[i] - pm[i].lo));
    nvtxRangePushA("sonic.prove.msm");
    // The jobs with terms on this rank, in record order.  Option "overlap" (off: measured slower) cuts the batch in two
    // halves of about equal terms that run on two streams: the second half's sort runs under the first half's
    // accumulation, and its accumulation -- which waits for the first half's -- runs over the first half's
    // latency-bound tail (fix-up, bucket reduction, finish).  On a B200 the tail's blocks and the accumulate kernel's
    // (216 KB of shared memory and all registers of an SM at 4 blocks) evict each other instead of sharing the SM:
    // prove() at n = 2^16 went 49.0 -> 52.2 ms, a rank of eight 7.8 -> 8.9 ms.
    std::vector<uint32_t> own;
    for (uint32_t i = 0; i < nm; ++i) if (jobs[i].n > 0) own.push_back(i);
    G1Affine* d_aff = nullptr;
    G1Affine* d_mine = nullptr;
    if (sharded) {
        // only this rank's MSMs (whole or partial) enter the pipeline; the others contribute the identity (zeros)
        d_aff = ar.get<G1Affine>(nm);
        SONIC_CUDA(cudaMemsetAsync(d_aff, 0, (size_t)nm * sizeof(G1Affine), st));
        d_mine = ar.get<G1Affine>(own.size() ? own.size() : 1);
    } else if (own.size() != nm) {
        // (jobs without terms -- a window clipped away entirely -- still encode as the identity)
        own.clear();
        for (uint32_t i = 0; i < nm; ++i) own.push_back(i);
    }
    auto run_range = [&](size_t first, size_t cnt, const MsmSync* sync) {
        std::vector<MsmJob> part;
        for (size_t k = 0; k < cnt; ++k) part.push_back(jobs[own[first + k]]);
        if (sharded) msm_run(cx, d_points, tables, (const uint32_t*)sbase, part, d_mine + first, nullptr, sync);
        else msm_run(cx, d_points, tables, (const uint32_t*)sbase, part, nullptr, d_result + (size_t)first * 48, sync);
    };
    uint64_t own_terms = 0;
    for (uint32_t i : own) own_terms += jobs[i].n;
    size_t cut = 0;
    if (cx.opt_overlap && cx.stream2 && own.size() >= 2 && own.size() <= (size_t)MSM_MAX_JOBS && own_terms >= (1u << 16)) {
        uint64_t acc_terms = 0;
        while (cut + 1 < own.size() && 2 * (acc_terms + jobs[own[cut]].n) <= own_terms + jobs[own[cut]].n) acc_terms += jobs[own[cut++]].n;
        if (cut == 0) cut = 1;
    }
    if (cut > 0) {
        SONIC_CUDA(cudaEventRecord(cx.ovl[0], st));                      // everything the MSMs read has been enqueued before this point
        MsmSync a_sync, b_sync;
        a_sync.signal_after_acc = cx.ovl[1];
        b_sync.wait_before_acc = cx.ovl[1];
        b_sync.second = true;
        SONIC_CUDA(cudaStreamWaitEvent(cx.stream2, cx.ovl[0], 0));
        run_range(0, cut, &a_sync);
        std::swap(cx.stream, cx.stream2);                                  // the launchers use the context's current stream
        try {
            run_range(cut, own.size() - cut, &b_sync);
            SONIC_CUDA(cudaEventRecord(cx.ovl[2], cx.stream));
        } catch (...) {
            std::swap(cx.stream, cx.stream2);
            throw;
        }
        std::swap(cx.stream, cx.stream2);
        SONIC_CUDA(cudaStreamWaitEvent(st, cx.ovl[2], 0));
        cx.timing_ms["msm.overlap"] = 1;
    } else {
        for (size_t first = 0; first < own.size(); first += MSM_MAX_JOBS) run_range(first, std::min<size_t>(MSM_MAX_JOBS, own.size() - first), nullptr);
    }
    if (sharded) {
        // runs of consecutive record indices move in one copy
        for (size_t k = 0; k < own.size();) {
            size_t e = k + 1;
            while (e < own.size() && own[e] == own[e - 1] + 1) ++e;
            SONIC_CUDA(cudaMemcpyAsync(d_aff + own[k], d_mine + k, (e - k) * sizeof(G1Affine), cudaMemcpyDeviceToDevice, st));
            k = e;
        }
        SONIC_LAUNCH(k_points_to_raw, div_up(nm, 64), 64, 0, d_aff, nm, reinterpret_cast<Fq*>(d_result));
    }
    nvtxRangePop();
    return SONIC_OK;
}

int prove_enqueue(Ctx& cx, SrsRep& srs, const CircuitRep& circ_, const Fr* d_in, const Fr* d_rnd, uint32_t M, bool has_main,
                  uint32_t rank, uint32_t world, uint8_t* d_result) {
    const CircuitRep* circ = &circ_;
    const uint32_t n = (uint32_t)circ->n, Q = (uint32_t)circ->Q;
    const int64_t d = (int64_t)srs.d;
    const ProveLayout lay(M, has_main);
    const bool sharded = world > 1;
    Arena& ar = cx.arena;
    cudaStream_t st = cx.stream;
    nvtxRangePushA("sonic.prove.poly");
    SONIC_CUDA(cudaEventRecord(cx.ev[4], st));

    // ---- the result buffer (or this rank's exchange record): values zero, range flags "none", encoding flag clear
    const uint32_t nm = lay.nm;
    Fr* d_vals = reinterpret_cast<Fr*>(d_result + (sharded ? lay.rec_vals() : lay.out_vals()));
    uint32_t* viol = reinterpret_cast<uint32_t*>(d_result + (sharded ? lay.rec_status() : lay.out_status()));
    uint32_t* bad = viol + 3 * (size_t)nm;
    SONIC_CUDA(cudaMemsetAsync(d_result, 0, sharded ? lay.rec_bytes() : lay.out_bytes(), st));
    SONIC_CUDA(cudaMemsetAsync(viol, 0xff, 12 * (size_t)nm, st));
    {
        const uint32_t dw[2] = {(uint32_t)srs.d, (uint32_t)(srs.d >> 32)};   // 8 pageable bytes: staged before the call returns
        SONIC_CUDA(cudaMemcpyAsync(bad + 1, dw, 8, cudaMemcpyHostToDevice, st));
    }

    // ---- which MSMs does this rank sum? -------------------------------------------------------------
    // The shapes of the proof's MSMs (exponent of scalar 0 and length, record order) depend only on
    // n, Q, M and d, so the dealing is known before any polynomial exists, and a rank of a sharded
    // proof builds only what feeds its own MSMs (below).
    // Sharding: the (clipped) exponent windows of all MSMs, concatenated in
    // record order, are cut into `world` runs of terms.  A rank so owns a few whole MSMs plus
    // at most two partial ones -- sorting, bucket reduction and the tail shrink with the rank count,
    // the loads differ by at most one term, and at most world-1 MSMs are split (their partial sums
    // meet in the fold).  A deterministic function of the sizes: all ranks agree.  Mirrored by
    // sonic_b200/dist.py:deal_terms.
    struct Shape { int64_t lo; uint32_t len; };
    std::vector<Shape> shape;
    const uint32_t slen = 3 * n + 1;
    {
        const int64_t nn = (int64_t)n;
        const uint32_t rlen = 3 * n + 5, tlen = 7 * n + 9, ulen = 2 * n + Q + 1;
        const int64_t rlo = -2 * nn - 4, tlo = -4 * nn - 8;
        if (has_main) {
            shape.push_back({rlo + (d - nn), rlen});   // prR
            shape.push_back({tlo, tlen});              // prT
            shape.push_back({rlo, rlen - 1});          // prWa
            shape.push_back({rlo, rlen - 1});          // prWb
            shape.push_back({tlo, tlen - 1});          // prWt
        }
        for (uint32_t j = 0; j < M; ++j) { shape.push_back({-nn, slen}); shape.push_back({-nn, slen - 1}); }      // S_j, W_j
        for (uint32_t j = 0; j < M; ++j) { shape.push_back({-nn, slen - 1}); shape.push_back({-nn, ulen - 1}); }  // W'_j, Q_j
        shape.push_back({-nn, ulen - 1});              // Q_v
        shape.push_back({-nn, ulen});                  // C
    }
    const uint32_t nshape = (uint32_t)shape.size();
    if (nshape != nm) return fail(SONIC_ERR_INVALID_ARG, "internal: MSM list does not match the proof layout");
    std::vector<int64_t> piece_lo(nshape), piece_hi(nshape);  // this rank's part [lo, hi) of every window, as exponents
    std::vector<uint32_t> first_owner(nshape, 0);             // lowest rank holding terms of the MSM (0 for an empty window)
    {
        // clipped windows and their positions in the concatenation
        std::vector<int64_t> clo(nshape);
        std::vector<uint64_t> pos(nshape + 1, 0);
        for (uint32_t i = 0; i < nshape; ++i) {
            const int64_t lo = shape[i].lo, hi = shape[i].lo + (int64_t)shape[i].len;
            clo[i] = std::max(lo, -d);
            const int64_t chi = std::max(clo[i], std::min(hi, d + 1));
            pos[i + 1] = pos[i] + (uint64_t)(chi - clo[i]);
        }
        const uint64_t total_len = pos[nshape];
        const uint32_t W_ = sharded ? world : 1;
        // No slivers: a boundary that would leave only a few terms of an MSM on one side moves to that MSM's
        // border.  A sliver costs its rank a whole bucket set (sort, reduction, finish) and the polynomial
        // behind it for nothing -- 27 terms of one W'_j cost a rank 1.3 ms of 7.7 before this rule.
        auto snap = [&](std::vector<uint64_t>& bd) {
            const uint64_t floor_ = std::max<uint64_t>(4096, total_len / ((uint64_t)W_ * 64));
            for (uint32_t r = 1; r < W_; ++r) {
                const uint64_t b = bd[r];
                for (uint32_t i = 0; i < nshape; ++i) {
                    if (!(pos[i] < b && b < pos[i + 1])) continue;
                    const uint64_t minp = std::min<uint64_t>((pos[i + 1] - pos[i]) / 2, floor_);
                    if (b - pos[i] < minp) bd[r] = pos[i];
                    else if (pos[i + 1] - b < minp) bd[r] = pos[i + 1];
                    break;
                }
            }
        };
        // run boundaries: equal runs first
        std::vector<uint64_t> bound(W_ + 1);
        for (uint32_t r = 0; r <= W_; ++r) bound[r] = total_len * r / W_;
        snap(bound);
        // A rank that owns part of prT / prWt also builds t(X,y) (three NTTs and a 7n-long opening).  With the
        // Fr side built by ownership everywhere that is worth about n/2 terms of MSM work at n = 2^16 (measured
        // per device at 8 GPUs: the t-owners hold fewer, longer MSMs and so also spend less in the bucket
        // reduction): those ranks are dealt that much less, if the shorter runs leave the same ranks in charge
        // of t (otherwise the equal runs stay).
        auto t_owners = [&](const std::vector<uint64_t>& bd) {
            std::vector<char> o(W_, 0);
            for (uint32_t r = 0; r < W_; ++r)
                for (uint32_t i : {1u, 4u}) {
                    const uint64_t a = std::max(bd[r], pos[i]), b = std::min(bd[r + 1], pos[i + 1]);
                    if (b > a) o[r] = 1;
                }
            return o;
        };
        if (sharded && has_main) {
            // (n/2 at eight ranks, where the t-owners hold 4 long MSMs against 6-7 short ones elsewhere; nothing at two, where
            // the t-owner's extra polynomial work and the other rank's extra bucket sets cancel: measured per rank)
            const uint64_t V = (uint64_t)n / 2 * std::min<uint64_t>(W_ > 2 ? W_ - 2 : 0, 6) / 6;
            const std::vector<char> o1 = t_owners(bound);
            uint64_t k = 0;
            for (char c : o1) k += c;
            const uint64_t padded = total_len + k * V;
            std::vector<uint64_t> b2(W_ + 1, 0);
            bool ok = k > 0 && k < W_;
            for (uint32_t r = 0; r < W_ && ok; ++r) {
                const uint64_t cap = padded * (r + 1) / W_ - padded * r / W_;
                if (o1[r] && cap < V) { ok = false; break; }
                b2[r + 1] = b2[r] + cap - (o1[r] ? V : 0);
            }
            if (ok && b2[W_] == total_len) {
                snap(b2);
                if (t_owners(b2) == o1) bound = b2;
            }
        }
        const uint64_t run_lo = sharded ? bound[rank] : 0, run_hi = sharded ? bound[rank + 1] : total_len;
        for (uint32_t i = 0; i < nshape; ++i) {
            const uint64_t span = pos[i + 1] - pos[i];
            const uint64_t a = run_lo > pos[i] ? std::min(run_lo - pos[i], span) : 0;
            const uint64_t b = run_hi > pos[i] ? std::min(run_hi - pos[i], span) : 0;
            piece_lo[i] = clo[i] + (int64_t)a;
            piece_hi[i] = clo[i] + (int64_t)b;
            for (uint32_t r = 0; r < W_; ++r)
                if (std::min(bound[r + 1], pos[i + 1]) > std::max(bound[r], pos[i])) { first_owner[i] = r; break; }
        }
    }
    // `mine(i)`: this rank builds the scalar vector of MSM i (it sums part of it, or nobody does and rank 0
    // stands in so that the range scan of CommitmentScheme.hs:70-73 still happens); `lead(i)`: this rank
    // also contributes the field value that opening i yields.
    auto owns = [&](uint32_t i) { return piece_hi[i] > piece_lo[i]; };
    auto lead = [&](uint32_t i) { return first_owner[i] == rank; };
    auto mine = [&](uint32_t i) { return owns(i) || lead(i); };
    const uint32_t mbase = has_main ? 5u : 0u;  // record index of S_1
    auto iS = [&](uint32_t j) { return mbase + 2 * j; };
    auto iW = [&](uint32_t j) { return mbase + 2 * j + 1; };
    auto iWp = [&](uint32_t j) { return mbase + 2 * M + 2 * j; };
    auto iQ = [&](uint32_t j) { return mbase + 2 * M + 2 * j + 1; };
    const uint32_t iQv = mbase + 4 * M, iC = mbase + 4 * M + 1;

    // ---- what has to exist on this rank -----------------------------------------------------------------
    // Everything below is demand-driven: a vector, a power table or an opening is built only if one of
    // this rank's MSMs (or a field value it leads) needs it.  With one rank everything is needed.
    const bool need_t = has_main && (mine(1) || mine(4));
    const bool lead_s = has_main && lead(1);                       // prS = s(z, y) comes from the rank that leads prT
    const bool need_rx1 = has_main && (mine(0) || mine(2) || mine(3) || need_t);
    const bool need_s0 = need_t || lead_s;
    std::vector<char> need_sj(M, 0);
    bool need_suy = mine(iQv) || mine(iC);
    for (uint32_t j = 0; j < M; ++j) {
        need_sj[j] = mine(iS(j)) || mine(iW(j)) || mine(iWp(j));
        need_suy = need_suy || mine(iQ(j));
    }

    // ---- inputs to Montgomery form (flags non-canonical encodings; every rank checks) ---------------------
    const uint32_t nr = 2 * M + 8;
    Fr* rnd_m = ar.get<Fr>(nr);
    fr_to_mont(cx, d_rnd, rnd_m, nr, bad);
    Fr* in_m = nullptr;
    if (has_main) {
        in_m = ar.get<Fr>(3 * (size_t)n);
        fr_to_mont(cx, d_in, in_m, 3 * (size_t)n, bad);
    }

    // ---- evaluation points and their power tables ---------------------------------------------
    const uint32_t np = 2 * M + 5;
    enum { PT_Y = 0, PT_Z = 1, PT_YZ = 2, PT_YJ = 3 };
    const uint32_t PT_ZJ = 3 + M, PT_U = 3 + 2 * M, PT_V = 4 + 2 * M;
    std::vector<char> need_pt(np, 0);
    bool need_zlong = false;   // z also opens t(X,y), 7n+9 long
    if (has_main) {
        if (need_s0 || need_t) need_pt[PT_Y] = 1;
        if (mine(4)) need_zlong = true;
        if (mine(2) || lead_s || need_zlong) need_pt[PT_Z] = 1;
        if (mine(3)) need_pt[PT_YZ] = 1;
    }
    for (uint32_t j = 0; j < M; ++j) {
        if (need_sj[j] || mine(iQ(j))) need_pt[PT_YJ + j] = 1;
        if (mine(iW(j))) need_pt[PT_ZJ + j] = 1;
        if (mine(iWp(j))) need_pt[PT_U] = 1;
    }
    if (need_suy) need_pt[PT_U] = 1;
    if (mine(iQv)) need_pt[PT_V] = 1;
    Fr* pts = ar.get<Fr>(2 * np);
    const uint64_t tl = std::max<uint64_t>(3 * (uint64_t)n + 8, 2 * (uint64_t)n + Q + 4);
    Fr* tabs = ar.get<Fr>(2 * (size_t)np * tl);
    {
        std::vector<uint32_t> sel_p, sel_t;
        for (uint32_t p = 0; p < np; ++p)
            if (need_pt[p]) { sel_p.push_back(p); }
        for (uint32_t p : sel_p) sel_t.push_back(p);
        for (uint32_t p : sel_p) sel_t.push_back(np + p);
        if (!sel_p.empty()) {
            uint32_t* d_sel = ar.get<uint32_t>(3 * sel_p.size());
            std::vector<uint32_t> h(sel_p);
            h.insert(h.end(), sel_t.begin(), sel_t.end());
            SONIC_CUDA(cudaMemcpyAsync(d_sel, h.data(), 4 * h.size(), cudaMemcpyHostToDevice, st));
            SONIC_LAUNCH(k_prove_points, (unsigned)sel_p.size(), 32, 0, rnd_m, M, has_main ? 1 : 0, d_sel, pts);
            pow_tables(cx, pts, (int)sel_t.size(), tabs, tl, tl, d_sel + sel_p.size());
        }
    }
    auto fwd = [&](uint32_t p) { return tabs + (size_t)p * tl; };
    auto inv = [&](uint32_t p) { return tabs + (size_t)(np + p) * tl; };
    const uint64_t tlz = 7 * (uint64_t)n + 12;
    const Fr* zf = fwd(PT_Z);
    const Fr* zi = inv(PT_Z);
    if (need_zlong) {
        Fr* ztabs = ar.get<Fr>(2 * tlz);
        Fr* zb = ar.get<Fr>(2);
        SONIC_CUDA(cudaMemcpyAsync(zb, pts + PT_Z, sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        SONIC_CUDA(cudaMemcpyAsync(zb + 1, pts + np + PT_Z, sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        pow_tables(cx, zb, 2, ztabs, tlz, tlz);
        zf = ztabs;
        zi = ztabs + tlz;
    }

    // ---- s(X,y) for y (main) and the y_j this rank needs ----------------------------------------------
    const uint32_t nb = M + 1;  // batch slots: 0 = y, 1..M = y_j
    Fr* sxy_m = ar.get<Fr>((size_t)nb * slen);
    Fr* sxy_c = ar.get<Fr>((size_t)nb * slen);
    {
        std::vector<uint32_t> slots;
        if (need_s0) slots.push_back(0);
        for (uint32_t j = 0; j < M; ++j) if (need_sj[j]) slots.push_back(1 + j);
        const uint32_t ns = (uint32_t)slots.size();
        if (ns) {
            std::vector<uint32_t> h_idx(3 * (size_t)ns);
            for (uint32_t k = 0; k < ns; ++k) {
                const uint32_t b = slots[k];
                const uint32_t p = b == 0 ? (uint32_t)PT_Y : (uint32_t)PT_YJ + (b - 1);
                h_idx[k] = p;
                h_idx[ns + k] = np + p;
                h_idx[2 * ns + k] = b;
            }
            uint32_t* d_idx = ar.get<uint32_t>(3 * (size_t)ns);
            SONIC_CUDA(cudaMemcpyAsync(d_idx, h_idx.data(), 12 * (size_t)ns, cudaMemcpyHostToDevice, st));
            if (circ->sparse) {
                SparseView cl{circ->sp[0].col_ptr, circ->sp[0].row, circ->sp[0].cval}, cr{circ->sp[1].col_ptr, circ->sp[1].row, circ->sp[1].cval},
                    co{circ->sp[2].col_ptr, circ->sp[2].row, circ->sp[2].cval};
                SONIC_LAUNCH(k_build_sxy_csc, dim3(div_up(n, 128), ns), 128, 0, cl, cr, co, n, tabs, tl, d_idx, d_idx + ns, d_idx + 2 * ns, sxy_m);
            } else {
                SONIC_LAUNCH(k_build_sxy, dim3(div_up(n, 128), ns), 128, 0, circ->wL(), circ->wR(), circ->wO(), n, Q, tabs, tl, d_idx, d_idx + ns, d_idx + 2 * ns, sxy_m);
            }
            // canonical copies only where the vector itself is committed to (S_j); runs of adjacent slots in one launch
            for (uint32_t k = 0; k < ns;) {
                if (slots[k] == 0 || !mine(iS(slots[k] - 1))) { ++k; continue; }
                uint32_t e = k + 1;
                while (e < ns && slots[e] == slots[e - 1] + 1 && mine(iS(slots[e] - 1))) ++e;
                fr_from_mont(cx, sxy_m + (size_t)slots[k] * slen, sxy_c + (size_t)slots[k] * slen, (size_t)(e - k) * slen);
                k = e;
            }
        }
    }
    auto sxy = [&](uint32_t b) { Window w; w.mont = sxy_m + (size_t)b * slen; w.canon = sxy_c + (size_t)b * slen; w.lo = -(int64_t)n; w.len = slen; return w; };

    // ---- s(u,Y) -------------------------------------------------------------------------------
    Window suy;
    suy.len = 2 * n + Q + 1;
    suy.lo = -(int64_t)n;
    if (need_suy) {
        suy.mont = ar.get<Fr>(suy.len);
        suy.canon = mine(iC) ? ar.get<Fr>(suy.len) : nullptr;   // committed to (and range-scanned) only where C is summed
        SONIC_LAUNCH(k_build_suy_pm, div_up(2 * n + 1, 256), 256, 0, fwd(PT_U), n, Q, suy.mont);
        Fr* part = ar.get<Fr>((size_t)Q * SUY_PARTS);
        if (circ->sparse) {
            SparseView rl{circ->sp[0].row_ptr, circ->sp[0].col, circ->sp[0].val}, rr{circ->sp[1].row_ptr, circ->sp[1].col, circ->sp[1].val},
                ro{circ->sp[2].row_ptr, circ->sp[2].col, circ->sp[2].val};
            SONIC_LAUNCH(k_build_suy_dot_csr, dim3(SUY_PARTS, Q), 256, 0, rl, rr, ro, n, fwd(PT_U), inv(PT_U), part);
        } else {
            SONIC_LAUNCH(k_build_suy_dot, dim3(SUY_PARTS, Q), 256, 0, circ->wL(), circ->wR(), circ->wO(), n, fwd(PT_U), inv(PT_U), part);
        }
        SONIC_LAUNCH(k_build_suy_fin, div_up(Q, 64), 64, 0, part, n, Q, suy.mont);
        if (mine(iC)) fr_from_mont(cx, suy.mont, suy.canon, suy.len);
    }

    // ---- r'(X,1), t(X,y) ------------------------------------------------------------------------
    Window rx1, txy;
    rx1.len = 3 * n + 5;
    rx1.lo = -2 * (int64_t)n - 4;
    txy.len = 7 * n + 9;
    txy.lo = -4 * (int64_t)n - 8;
    if (need_rx1) {
        uint32_t logL = 1;
        while ((1ull << logL) < txy.len) ++logL;
        const uint64_t L = need_t ? 1ull << logL : rx1.len;
        Fr* A = ar.get<Fr>(L);
        if (need_t) SONIC_CUDA(cudaMemsetAsync(A, 0, L * sizeof(Fr), st));
        SONIC_LAUNCH(k_build_r, div_up(rx1.len, 256), 256, 0, in_m, in_m + n, in_m + 2 * (size_t)n, rnd_m, n, A);
        rx1.mont = A;
        if (need_t) {
            rx1.mont = ar.get<Fr>(rx1.len);
            SONIC_CUDA(cudaMemcpyAsync(rx1.mont, A, rx1.len * sizeof(Fr), cudaMemcpyDeviceToDevice, st));
        }
        if (mine(0)) {
            rx1.canon = ar.get<Fr>(rx1.len);
            fr_from_mont(cx, rx1.mont, rx1.canon, rx1.len);
        }
        if (need_t) {
            Fr* B = ar.get<Fr>(L);
            SONIC_CUDA(cudaMemsetAsync(B, 0, L * sizeof(Fr), st));
            SONIC_LAUNCH(k_build_rs, div_up(4 * n + 5, 256), 256, 0, rx1.mont, sxy_m, fwd(PT_Y), inv(PT_Y), n, B);
            NttPlan plan = ntt_prepare(cx, logL);
            ntt_forward(plan, A);
            ntt_forward(plan, B);
            fr_mul_pointwise(cx, A, B, (uint32_t)L);
            ntt_inverse(plan, A);
            SONIC_LAUNCH(k_t_fix, 1, 32, 0, A, circ->cs(), fwd(PT_Y), n, Q);
            txy.mont = A;
            if (mine(1)) {
                txy.canon = ar.get<Fr>(txy.len);
                fr_from_mont(cx, txy.mont, txy.canon, txy.len);
            }
        }
    }

    // ---- openings: values and quotient vectors --------------------------------------------------
    // a value goes straight into its slot of the result buffer when this rank leads the opening
    // (record order: prA prB prS | s_j | s'_j), else into scratch
    std::vector<OpenJob> ojobs;
    Fr* scratch = ar.get<Fr>(3 * (size_t)M + 8);
    uint32_t nscratch = 0;
    struct Quot { Fr* q; int64_t lo; uint32_t len; };
    auto add_open = [&](const Window& f, const Fr* pz, const Fr* pzi, bool want_q, Fr* value_slot) -> Quot {
        Quot qt{nullptr, f.lo, f.len - 1};
        if (!want_q && !value_slot) return qt;
        OpenJob jb;
        jb.f = f.mont;
        jb.pz = pz;
        jb.pzi = pzi;
        jb.q_canon = want_q ? ar.get<Fr>(f.len) : nullptr;
        jb.value_canon = value_slot ? value_slot : scratch + nscratch++;
        jb.len = f.len;
        jb.lo = (int32_t)f.lo;
        jb.z_is_zero = 0;
        jb.pad = 0;
        ojobs.push_back(jb);
        qt.q = jb.q_canon;
        return qt;
    };
    const uint32_t vbase = has_main ? 3u : 0u;
    Quot q_a{nullptr, rx1.lo, rx1.len - 1}, q_b{nullptr, rx1.lo, rx1.len - 1}, q_t{nullptr, txy.lo, txy.len - 1}, q_v{};
    std::vector<Quot> q_wj(M), q_wpj(M), q_qj(M);
    if (has_main) {
        q_a = add_open(rx1, zf, zi, mine(2), lead(2) ? d_vals + 0 : nullptr);                    // Protocol.hs:79
        q_b = add_open(rx1, fwd(PT_YZ), inv(PT_YZ), mine(3), lead(3) ? d_vals + 1 : nullptr);    // Protocol.hs:80
        q_t = add_open(txy, zf, zi, mine(4), nullptr);      // Protocol.hs:81 (t(z,y) itself is not a proof field)
        if (lead_s) add_open(sxy(0), zf, zi, false, d_vals + 2);                                 // Protocol.hs:83
    }
    for (uint32_t j = 0; j < M; ++j) {
        q_wj[j] = add_open(sxy(1 + j), fwd(PT_ZJ + j), inv(PT_ZJ + j), mine(iW(j)), lead(iW(j)) ? d_vals + vbase + j : nullptr);     // Signature.hs:43
        q_wpj[j] = add_open(sxy(1 + j), fwd(PT_U), inv(PT_U), mine(iWp(j)), nullptr);                                                // Signature.hs:54
        q_qj[j] = add_open(suy, fwd(PT_YJ + j), inv(PT_YJ + j), mine(iQ(j)), lead(iQ(j)) ? d_vals + vbase + M + j : nullptr);        // Signature.hs:55
    }
    q_v = add_open(suy, fwd(PT_V), inv(PT_V), mine(iQv), nullptr);                                                                   // Signature.hs:63
    open_batch(cx, ojobs);
    SONIC_CUDA(cudaEventRecord(cx.ev[5], st));
    nvtxRangePop();

    // ---- the proof's MSMs, in record order -----------------------------------------------------
    std::vector<ProofMsm> pm;
    auto commit = [&](const Window& f, int64_t maxm) { pm.push_back(ProofMsm{SONIC_FAMILY_ALPHA, f.canon, f.lo + (d - maxm), f.len, true}); };
    auto opening = [&](const Quot& q) { pm.push_back(ProofMsm{SONIC_FAMILY_PLAIN, q.q, q.lo, q.len, false}); };
    if (has_main) {
        commit(rx1, (int64_t)n);  // prR   Protocol.hs:63
        commit(txy, d);           // prT   Protocol.hs:73
        opening(q_a);             // prWa
        opening(q_b);             // prWb
        opening(q_t);             // prWt
    }
    for (uint32_t j = 0; j < M; ++j) {   // hscS   Signature.hs:42-43
        Window w = sxy(1 + j);
        if (!mine(iS(j))) w.canon = nullptr;
        commit(w, d);
        opening(q_wj[j]);
    }
    for (uint32_t j = 0; j < M; ++j) { opening(q_wpj[j]); opening(q_qj[j]); }          // hscW   Signature.hs:54-55
    opening(q_v);                                                                       // hscQv  Signature.hs:63
    commit(suy, d);                                                                     // hscC   Signature.hs:52
    if ((uint32_t)pm.size() != nshape) return fail(SONIC_ERR_INVALID_ARG, "internal: MSM list does not match its shape table");
    for (uint32_t i = 0; i < nm; ++i)
        if (pm[i].lo != shape[i].lo || pm[i].len != shape[i].len) return fail(SONIC_ERR_INVALID_ARG, "internal: MSM %u does not match its shape", i);

    return enqueue_msms(cx, srs, pm, piece_lo, piece_hi, viol, sharded, d_result);
}

// ---- hscProve on a general sparse s(X,Y) (src/Sonic/Signature.hs:32-72) ----------------------------------------
// The reference's hscProve takes any `BiVLaurent Fr`; its test feeds `sPoly weights`
// (test/Test/Signature.hs:30-36).  Here s(X,Y) = sum_t c_t X^(a_t) Y^(b_t) arrives as a term list and the
// two univariate families are evaluated from it directly:
//   s(X, y_j)[a] = sum over the terms with a_t = a of c_t y_j^(b_t)        (evalY, Utils.hs:20-21)
//   s(u, Y)[b]   = sum over the terms with b_t = b of c_t u^(a_t)          (evalX, Utils.hs:17-18)
// (terms sorted by the surviving exponent on the host; one thread per coefficient of the result).
struct TermRows {
    const uint32_t* ptr;   // rows + 1
    const int32_t* exp;    // exponent of the variable being evaluated, per term
    const Fr* val;         // coefficient, Montgomery
};

__global__ void __launch_bounds__(128) k_eval_term_rows(TermRows rows, uint32_t nrows, const Fr* __restrict__ tabs, uint64_t tl,
                                                        const uint32_t* __restrict__ fwd_idx, const uint32_t* __restrict__ inv_idx,
                                                        Fr* __restrict__ out, uint32_t out_stride) {
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= nrows) return;
    const uint32_t b = blockIdx.y;
    const Fr* pf = tabs + (size_t)fwd_idx[b] * tl;
    const Fr* pi = tabs + (size_t)inv_idx[b] * tl;
    Fr acc = Fr::zero();
    for (uint32_t k = rows.ptr[row]; k < rows.ptr[row + 1]; ++k) {
        const int32_t e = rows.exp[k];
        acc = fp_add(acc, fp_mul(rows.val[k], e >= 0 ? pf[e] : pi[-e]));
    }
    out[(size_t)b * out_stride + row] = acc;
}

int hsc_terms_enqueue(Ctx& cx, SrsRep& srs, uint64_t nterms, const int64_t* eX, const int64_t* eY, const uint8_t* coeff32,
                      uint32_t M, const Fr* d_rnd, uint8_t* d_result) {
    const ProveLayout lay(M, false);
    const int64_t d = (int64_t)srs.d;
    Arena& ar = cx.arena;
    cudaStream_t st = cx.stream;
    const uint32_t nm = lay.nm;
    Fr* d_vals = reinterpret_cast<Fr*>(d_result + lay.out_vals());
    uint32_t* viol = reinterpret_cast<uint32_t*>(d_result + lay.out_status());
    uint32_t* bad = viol + 3 * (size_t)nm;
    SONIC_CUDA(cudaMemsetAsync(d_result, 0, lay.out_bytes(), st));
    SONIC_CUDA(cudaMemsetAsync(viol, 0xff, 12 * (size_t)nm, st));
    {
        const uint32_t dw[2] = {(uint32_t)srs.d, (uint32_t)(srs.d >> 32)};
        SONIC_CUDA(cudaMemcpyAsync(bad + 1, dw, 8, cudaMemcpyHostToDevice, st));
    }
    // ---- the term list: zero coefficients are not terms (the sparse normal form drops them) ----------------
    std::vector<uint32_t> live;
    int64_t xlo = 0, xhi = 0, ylo = 0, yhi = 0;   // both windows contain exponent 0 (openPoly subtracts f(z) there)
    const int64_t lim = int64_t(1) << 26;
    for (uint64_t t = 0; t < nterms; ++t) {
        const uint64_t* w = reinterpret_cast<const uint64_t*>(coeff32 + 32 * t);
        uint64_t c[4];
        memcpy(c, w, 32);
        if ((c[0] | c[1] | c[2] | c[3]) == 0) continue;
        if (eX[t] < -lim || eX[t] > lim || eY[t] < -lim || eY[t] > lim) return fail(SONIC_ERR_INVALID_ARG, "term %llu: exponent out of range (|e| <= 2^26)", (unsigned long long)t);
        live.push_back((uint32_t)t);
        xlo = std::min(xlo, eX[t]); xhi = std::max(xhi, eX[t]);
        ylo = std::min(ylo, eY[t]); yhi = std::max(yhi, eY[t]);
    }
    const uint32_t nl = (uint32_t)live.size();
    const uint32_t xlen = (uint32_t)(xhi - xlo + 1), ylen = (uint32_t)(yhi - ylo + 1);
    // rows by X (the exponent that survives evalY) and by Y (survives evalX); duplicates of a monomial add up
    std::vector<uint32_t> ptrX(xlen + 1, 0), ptrY(ylen + 1, 0), ordX(nl), ordY(nl);
    for (uint32_t k = 0; k < nl; ++k) { ptrX[eX[live[k]] - xlo + 1]++; ptrY[eY[live[k]] - ylo + 1]++; }
    for (uint32_t i = 0; i < xlen; ++i) ptrX[i + 1] += ptrX[i];
    for (uint32_t i = 0; i < ylen; ++i) ptrY[i + 1] += ptrY[i];
    {
        std::vector<uint32_t> cx_(ptrX.begin(), ptrX.end() - 1), cy_(ptrY.begin(), ptrY.end() - 1);
        for (uint32_t k = 0; k < nl; ++k) { ordX[cx_[eX[live[k]] - xlo]++] = live[k]; ordY[cy_[eY[live[k]] - ylo]++] = live[k]; }
    }
    std::vector<int32_t> expX(nl ? nl : 1), expY(nl ? nl : 1);   // X-rows carry Y exponents and vice versa
    std::vector<uint8_t> coef(2 * (size_t)(nl ? nl : 1) * 32);
    for (uint32_t k = 0; k < nl; ++k) {
        expX[k] = (int32_t)eY[ordX[k]];
        expY[k] = (int32_t)eX[ordY[k]];
        memcpy(&coef[32 * (size_t)k], coeff32 + 32 * (size_t)ordX[k], 32);
        memcpy(&coef[32 * ((size_t)nl + k)], coeff32 + 32 * (size_t)ordY[k], 32);
    }
    uint32_t* d_ptrX = ar.get<uint32_t>(xlen + 1);
    uint32_t* d_ptrY = ar.get<uint32_t>(ylen + 1);
    int32_t* d_expX = ar.get<int32_t>(nl ? nl : 1);
    int32_t* d_expY = ar.get<int32_t>(nl ? nl : 1);
    Fr* d_coef_c = ar.get<Fr>(2 * (size_t)(nl ? nl : 1));
    Fr* d_coef = ar.get<Fr>(2 * (size_t)(nl ? nl : 1));
    SONIC_CUDA(cudaMemcpyAsync(d_ptrX, ptrX.data(), 4 * (size_t)(xlen + 1), cudaMemcpyHostToDevice, st));
    SONIC_CUDA(cudaMemcpyAsync(d_ptrY, ptrY.data(), 4 * (size_t)(ylen + 1), cudaMemcpyHostToDevice, st));
    if (nl) {
        SONIC_CUDA(cudaMemcpyAsync(d_expX, expX.data(), 4 * (size_t)nl, cudaMemcpyHostToDevice, st));
        SONIC_CUDA(cudaMemcpyAsync(d_expY, expY.data(), 4 * (size_t)nl, cudaMemcpyHostToDevice, st));
        SONIC_CUDA(cudaMemcpyAsync(d_coef_c, coef.data(), 64 * (size_t)nl, cudaMemcpyHostToDevice, st));
        fr_to_mont(cx, d_coef_c, d_coef, 2 * (size_t)nl, bad);
    }
    SONIC_CUDA(cudaStreamSynchronize(st));   // the host vectors above are large and pageable: do not let them die in flight

    // ---- evaluation points y_j, z_j, u, v and their power tables ------------------------------------------------
    const uint32_t nr = 2 * M + 8, np = 2 * M + 5;
    const uint32_t PT_YJ = 3, PT_ZJ = 3 + M, PT_U = 3 + 2 * M, PT_V = 4 + 2 * M;
    Fr* rnd_m = ar.get<Fr>(nr);
    fr_to_mont(cx, d_rnd, rnd_m, nr, bad);
    Fr* pts = ar.get<Fr>(2 * np);
    const uint64_t tl = (uint64_t)std::max(xlen, ylen) + 2;   // the openings index z^k for every slot k of a window
    Fr* tabs = ar.get<Fr>(2 * (size_t)np * tl);
    {
        std::vector<uint32_t> h;
        for (uint32_t p = 3; p < np; ++p) h.push_back(p);
        const uint32_t ns = (uint32_t)h.size();
        for (uint32_t k = 0; k < ns; ++k) h.push_back(h[k]);
        for (uint32_t k = 0; k < ns; ++k) h.push_back(np + h[k]);
        uint32_t* d_sel = ar.get<uint32_t>(h.size());
        SONIC_CUDA(cudaMemcpyAsync(d_sel, h.data(), 4 * h.size(), cudaMemcpyHostToDevice, st));
        SONIC_LAUNCH(k_prove_points, ns, 32, 0, rnd_m, M, 0, d_sel, pts);
        pow_tables(cx, pts, (int)(2 * ns), tabs, tl, tl, d_sel + ns);
    }
    auto fwd = [&](uint32_t p) { return tabs + (size_t)p * tl; };
    auto inv = [&](uint32_t p) { return tabs + (size_t)(np + p) * tl; };

    // ---- s(X, y_j) and s(u, Y) ------------------------------------------------------------------------------------
    Fr* sxy_m = ar.get<Fr>((size_t)(M ? M : 1) * xlen);
    Fr* sxy_c = ar.get<Fr>((size_t)(M ? M : 1) * xlen);
    if (M) {
        std::vector<uint32_t> h(2 * (size_t)M);
        for (uint32_t j = 0; j < M; ++j) { h[j] = PT_YJ + j; h[M + j] = np + PT_YJ + j; }
        uint32_t* d_idx = ar.get<uint32_t>(2 * (size_t)M);
        SONIC_CUDA(cudaMemcpyAsync(d_idx, h.data(), 8 * (size_t)M, cudaMemcpyHostToDevice, st));
        SONIC_LAUNCH(k_eval_term_rows, dim3(div_up(xlen, 128), M), 128, 0, TermRows{d_ptrX, d_expX, d_coef}, xlen, tabs, tl, d_idx, d_idx + M, sxy_m, xlen);
        fr_from_mont(cx, sxy_m, sxy_c, (size_t)M * xlen);
    }
    Window suy;
    suy.lo = ylo;
    suy.len = ylen;
    suy.mont = ar.get<Fr>(ylen);
    suy.canon = ar.get<Fr>(ylen);
    {
        const uint32_t h[2] = {PT_U, np + PT_U};
        uint32_t* d_idx = ar.get<uint32_t>(2);
        SONIC_CUDA(cudaMemcpyAsync(d_idx, h, 8, cudaMemcpyHostToDevice, st));
        SONIC_LAUNCH(k_eval_term_rows, dim3(div_up(ylen, 128), 1), 128, 0, TermRows{d_ptrY, d_expY, d_coef + nl}, ylen, tabs, tl, d_idx, d_idx + 1, suy.mont, ylen);
        fr_from_mont(cx, suy.mont, suy.canon, ylen);
    }
    auto sxy = [&](uint32_t j) { Window w; w.mont = sxy_m + (size_t)j * xlen; w.canon = sxy_c + (size_t)j * xlen; w.lo = xlo; w.len = xlen; return w; };

    // ---- openings (Signature.hs:43,54-55,63) and the MSM list in record order -----------------------------------------
    std::vector<OpenJob> ojobs;
    Fr* scratch = ar.get<Fr>((size_t)M + 2);
    uint32_t nscratch = 0;
    struct Quot { Fr* q; int64_t lo; uint32_t len; };
    auto add_open = [&](const Window& f, const Fr* pz, const Fr* pzi, Fr* value_slot) -> Quot {
        OpenJob jb;
        jb.f = f.mont; jb.pz = pz; jb.pzi = pzi;
        jb.q_canon = ar.get<Fr>(f.len);
        jb.value_canon = value_slot ? value_slot : scratch + nscratch++;
        jb.len = f.len; jb.lo = (int32_t)f.lo; jb.z_is_zero = 0; jb.pad = 0;
        ojobs.push_back(jb);
        return Quot{jb.q_canon, f.lo, f.len - 1};
    };
    std::vector<Quot> q_wj(M), q_wpj(M), q_qj(M);
    for (uint32_t j = 0; j < M; ++j) {
        q_wj[j] = add_open(sxy(j), fwd(PT_ZJ + j), inv(PT_ZJ + j), d_vals + j);
        q_wpj[j] = add_open(sxy(j), fwd(PT_U), inv(PT_U), nullptr);
        q_qj[j] = add_open(suy, fwd(PT_YJ + j), inv(PT_YJ + j), d_vals + M + j);
    }
    const Quot q_v = add_open(suy, fwd(PT_V), inv(PT_V), nullptr);
    open_batch(cx, ojobs);
    std::vector<ProofMsm> pm;
    auto commit = [&](const Window& f) { pm.push_back(ProofMsm{SONIC_FAMILY_ALPHA, f.canon, f.lo, f.len, true}); };   // max = d: no shift
    auto opening = [&](const Quot& q) { pm.push_back(ProofMsm{SONIC_FAMILY_PLAIN, q.q, q.lo, q.len, false}); };
    for (uint32_t j = 0; j < M; ++j) { commit(sxy(j)); opening(q_wj[j]); }
    for (uint32_t j = 0; j < M; ++j) { opening(q_wpj[j]); opening(q_qj[j]); }
    opening(q_v);
    commit(suy);
    std::vector<int64_t> piece_lo(pm.size()), piece_hi(pm.size());
    for (size_t i = 0; i < pm.size(); ++i) {
        piece_lo[i] = std::max(pm[i].lo, -d);
        piece_hi[i] = std::max(piece_lo[i], std::min(pm[i].lo + (int64_t)pm[i].len, d + 1));
    }
    return enqueue_msms(cx, srs, pm, piece_lo, piece_hi, viol, false, d_result);
}

void prove_collect_timing(Ctx& cx) {
    msm_collect_timing(cx);
    float ms = 0;
    if (cudaEventElapsedTime(&ms, cx.ev[4], cx.ev[5]) == cudaSuccess) cx.timing_ms["poly"] = ms;
}

}  // namespace sonic
