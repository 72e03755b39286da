// sls_gather.cu — kernel 1: the bit-exact gather of every column's (d,T)-localized index sets.
//
// Replaces, per disturbance column j (reference src/reduction.jl:11-27 and src/synthesis.jl:57-60):
//   s_x = unique(findnz((𝓢ₓ[end]*(A.≠0))[:,j])[1])      -> sorted union of 𝓢ₓ[T][:,k] over k with A[k,j] ≠ 0
//   s_u = unique(findnz((𝓢ᵤ[end]*(A.≠0))[:,j])[1])
//   ii_x = indexin(s_x, [j]) .≠ nothing                  -> jloc
//   masks 𝓢ₓ[t][s_x,j] == 1, 𝓢ᵤ[t][s_u,j] == 1           -> bit words per tap
// and emits per-tap stored-entry counts (-> Julia colptr by an in-place scan) plus the sizes kernel 2
// needs for its shared-memory carve-up.  The reference recomputes the global sparse product
// 𝓢[T]*(A.≠0) for every column; here a column only touches the few support columns it needs.
//
// Integer-only, HBM-bound: one warp per column, coalesced reads of contiguous CSC segments (adjacent
// columns of a block read adjacent memory), sorted-unique by a bitonic sort in shared memory.
#include "sls_common.h"
#include "sls_kernels.h"
#include <climits>

#define GATHER_WPB 4

// Sorted-unique union of the rows of S[T-1][:,k] over the stored nonzeros k of A[:,j].
// Result in buf[0..ret); ret = -1 if the candidate list overflows `cap` (a power of two).
__device__ int sls_warp_index_set(const SlsCsc& A, const SlsSupports& S, int T, int64_t Nx, int64_t j,
                                  int32_t* buf, int cap) {
    const int lane = sls_lane();
    const int64_t a0 = A.colptr[j] - 1, a1 = A.colptr[j + 1] - 1;
    const int64_t* scp = S.colptr + (int64_t)(T - 1) * (Nx + 1);
    const int64_t base = S.tap_off[T - 1];
    int C = 0;
    for (int64_t e = a0; e < a1; e++) {
        if (A.nzval[e] == 0.0) continue;                      // `A .≠ 0` drops stored zeros
        const int64_t k = A.rowval[e] - 1;
        const int64_t s0 = base + scp[k] - 1, s1 = base + scp[k + 1] - 1;
        const int len = (int)(s1 - s0);
        if (C + len > cap) return -1;
        for (int i = lane; i < len; i += 32) buf[C + i] = (int32_t)(S.rowval[s0 + i] - 1);
        C += len;
    }
    int P = 32;
    while (P < C) P <<= 1;
    if (P > cap) P = cap;                                     // cap >= 32 and a power of two
    for (int i = C + lane; i < P; i += 32) buf[i] = INT_MAX;
    __syncwarp();
    for (int k = 2; k <= P; k <<= 1) {
        for (int jj = k >> 1; jj > 0; jj >>= 1) {
            for (int i = lane; i < P; i += 32) {
                const int ixj = i ^ jj;
                if (ixj > i) {
                    const int a = buf[i], b = buf[ixj];
                    const bool asc = (i & k) == 0;
                    if ((a > b) == asc) { buf[i] = b; buf[ixj] = a; }
                }
            }
            __syncwarp();
        }
    }
    int n = 0, prev_last = -1;
    for (int b0 = 0; b0 < P; b0 += 32) {
        const int v = buf[b0 + lane];
        int pv = __shfl_up_sync(SLS_FULL, v, 1);
        if (lane == 0) pv = prev_last;
        const bool keep = (v != INT_MAX) && (v != pv);
        const unsigned bal = __ballot_sync(SLS_FULL, keep);
        __syncwarp();
        if (keep) buf[n + __popc(bal & ((1u << lane) - 1u))] = v;
        n += __popc(bal);
        prev_last = __shfl_sync(SLS_FULL, v, 31);
    }
    __syncwarp();
    return n;
}

// kernel 1a: sizes only (n, m) so that the host can size the strided metadata exactly.
__global__ void __launch_bounds__(GATHER_WPB * 32)
sls_gather_count_kernel(SlsCsc A, SlsSupports Sx, SlsSupports Su, int T, int64_t Nx,
                        const int32_t* __restrict__ cols, int ncols, int32_t* n_out, int32_t* m_out, int cap) {
    SLS_DYN_SMEM(smem);
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int c = blockIdx.x * GATHER_WPB + warp;
    if (c >= ncols) return;
    int32_t* buf = reinterpret_cast<int32_t*>(smem) + (size_t)warp * cap;
    const int64_t j = cols[c];
    const int n = sls_warp_index_set(A, Sx, T, Nx, j, buf, cap);
    __syncwarp();
    const int m = sls_warp_index_set(A, Su, T, Nx, j, buf, cap);
    if (lane == 0) { n_out[c] = n; m_out[c] = m; }
}

// Set the mask bits of one tap for one column: rows of S[t][:,j] (with nzval true) found in s.
__device__ __forceinline__ void sls_warp_mask(const SlsSupports& S, int t, int64_t Nx, int64_t j,
                                              const int32_t* s, int n, uint32_t* words) {
    const int lane = sls_lane();
    const int64_t* cp = S.colptr + (int64_t)t * (Nx + 1);
    const int64_t base = S.tap_off[t];
    const int64_t s0 = base + cp[j] - 1, s1 = base + cp[j + 1] - 1;
    for (int64_t e = s0 + lane; e < s1; e += 32) {
        const bool on = S.nzval ? (S.nzval[e] != 0) : true;   // `𝓢[t][s,cⱼ] .≠ 1` fixes the entry (synthesis.jl:58-59)
        const int idx = sls_find(s, n, (int32_t)(S.rowval[e] - 1));
        if (on && idx >= 0) atomicOr(&words[idx >> 5], 1u << (idx & 31));
    }
}

// kernel 1b: index sets, ii_x, masks, per-tap counts, packed-result offsets, local nnz.
// shared memory per warp: cap + nstride + mstride int32, wx + wu uint32.
__global__ void __launch_bounds__(GATHER_WPB * 32)
sls_gather_fill_kernel(SlsCsc A, SlsCsc B2, SlsSupports Sx, SlsSupports Su, SlsPlanView P, int cap) {
    SLS_DYN_SMEM(smem);
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int c = blockIdx.x * GATHER_WPB + warp;
    if (c >= P.ncols) return;
    const int per_warp = cap + P.nstride + P.mstride + P.wx + P.wu;
    int32_t* buf = reinterpret_cast<int32_t*>(smem) + (size_t)warp * per_warp;
    int32_t* sx = buf + cap;
    int32_t* su = sx + P.nstride;
    uint32_t* mw = reinterpret_cast<uint32_t*>(su + P.mstride);
    const int T = P.T;
    const int64_t Nx = P.Nx;
    const int64_t j = P.cols[c];

    const int n = sls_warp_index_set(A, Sx, T, Nx, j, buf, cap);
    for (int i = lane; i < n; i += 32) { sx[i] = buf[i]; P.sx[(size_t)c * P.nstride + i] = buf[i]; }
    __syncwarp();
    const int m = sls_warp_index_set(A, Su, T, Nx, j, buf, cap);
    for (int i = lane; i < m; i += 32) { su[i] = buf[i]; P.su[(size_t)c * P.mstride + i] = buf[i]; }
    __syncwarp();
    if (lane == 0) { P.n[c] = n; P.m[c] = m; P.jloc[c] = sls_find(sx, n, (int32_t)j); }

    uint32_t run = 0;
    for (int t = 0; t < T; t++) {
        for (int w = lane; w < P.wx + P.wu; w += 32) mw[w] = 0u;
        __syncwarp();
        sls_warp_mask(Sx, t, Nx, j, sx, n, mw);
        sls_warp_mask(Su, t, Nx, j, su, m, mw + P.wx);
        __syncwarp();
        uint32_t cx = 0, cu = 0;
        for (int w = 0; w < P.wx; w++) cx += __popc(mw[w]);
        for (int w = 0; w < P.wu; w++) cu += __popc(mw[P.wx + w]);
        for (int w = lane; w < P.wx; w += 32) P.maskx[((size_t)c * T + t) * P.wx + w] = mw[w];
        for (int w = lane; w < P.wu; w += 32) P.masku[((size_t)c * T + t) * P.wu + w] = mw[P.wx + w];
        if (lane == 0) {
            P.colptr[(int64_t)t * (Nx + 1) + j] = cx;
            P.colptr[(int64_t)(T + t) * (Nx + 1) + j] = cu;
            P.pre[(size_t)c * 2 * T + 2 * t] = run;
            P.pre[(size_t)c * 2 * T + 2 * t + 1] = run + cx;
        }
        run += cx + cu;
        __syncwarp();
    }
    // local nnz: A[sx,sx] by columns, B2[sx,su] by columns, and the full column lengths of B2[:,su]
    int cntA = 0, cntB = 0, cntBfull = 0;
    for (int l = lane; l < n; l += 32) {
        const int64_t e0 = A.colptr[sx[l]] - 1, e1 = A.colptr[sx[l] + 1] - 1;
        for (int64_t e = e0; e < e1; e++) cntA += (sls_find(sx, n, (int32_t)(A.rowval[e] - 1)) >= 0);
    }
    for (int l = lane; l < m; l += 32) {
        const int64_t e0 = B2.colptr[su[l]] - 1, e1 = B2.colptr[su[l] + 1] - 1;
        cntBfull += (int)(e1 - e0);
        for (int64_t e = e0; e < e1; e++) cntB += (sls_find(sx, n, (int32_t)(B2.rowval[e] - 1)) >= 0);
    }
    cntA = __reduce_add_sync(SLS_FULL, cntA);
    cntB = __reduce_add_sync(SLS_FULL, cntB);
    cntBfull = __reduce_add_sync(SLS_FULL, cntBfull);
    if (lane == 0) {
        P.tot[c] = run;
        P.lnnz[c] = cntA + cntB;
        P.lnnzA[c] = cntA;
        P.lnnzB[c] = cntBfull;
    }
}

// In-place exclusive scans: blocks 0..narr-1 turn per-column counts into 1-based Julia colptr arrays
// (length len = Nx+1); block narr scans tot (length totlen = ncols+1) into 0-based packed offsets.
#define SCAN_THREADS 1024
#define SCAN_ITEMS 8
__global__ void __launch_bounds__(SCAN_THREADS)
sls_scan_kernel(int64_t* colptr, int64_t len, int narr, int64_t* tot, int64_t totlen) {
    __shared__ int64_t wsum[SCAN_THREADS / 32];
    __shared__ int64_t carry_s, tile_total_s;
    const int b = blockIdx.x;
    int64_t* arr = (b < narr) ? colptr + (int64_t)b * len : tot;
    const int64_t L = (b < narr) ? len : totlen;
    const int64_t add = (b < narr) ? 1 : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int64_t tile = 0; tile < L; tile += (int64_t)SCAN_THREADS * SCAN_ITEMS) {
        const int64_t i0 = tile + (int64_t)tid * SCAN_ITEMS;
        int64_t v[SCAN_ITEMS];
        int64_t sum = 0;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { v[k] = (i0 + k < L) ? arr[i0 + k] : 0; sum += v[k]; }
        int64_t incl = sum;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int64_t o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            const int64_t w = wsum[lane];
            int64_t wi = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int64_t o = __shfl_up_sync(SLS_FULL, wi, d); if (lane >= d) wi += o; }
            wsum[lane] = wi - w;                               // exclusive warp offsets
            if (lane == 31) tile_total_s = wi;
        }
        __syncthreads();
        int64_t run = carry_s + wsum[warp] + (incl - sum) + add;
#pragma unroll
        for (int k = 0; k < SCAN_ITEMS; k++) { if (i0 + k < L) arr[i0 + k] = run; run += v[k]; }
        __syncthreads();
        if (tid == 0) carry_s += tile_total_s;
        __syncthreads();
    }
}

// Dense local blocks of one planned column, for sls_local_problem (test / inspection of kernel 1 and
// of the sub-plant gather used by kernel 2): At (n x n, column-major), B2t (n x m, column-major),
// Mx (T x n), Mu (T x m) row-major Bool.
__global__ void sls_local_dense_kernel(SlsCsc A, SlsCsc B2, SlsPlanView P, int c,
                                       double* At, double* B2t, uint8_t* Mx, uint8_t* Mu) {
    const int n = P.n[c], m = P.m[c];
    const int32_t* sx = P.sx + (size_t)c * P.nstride;
    const int32_t* su = P.su + (size_t)c * P.mstride;
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int i = tid; i < n * n; i += nt) At[i] = 0.0;
    for (int i = tid; i < n * m; i += nt) B2t[i] = 0.0;
    __syncthreads();
    for (int l = tid; l < n; l += nt) {
        for (int64_t e = A.colptr[sx[l]] - 1; e < A.colptr[sx[l] + 1] - 1; e++) {
            const int r = sls_find(sx, n, (int32_t)(A.rowval[e] - 1));
            if (r >= 0) At[(size_t)l * n + r] = A.nzval[e];
        }
    }
    for (int l = tid; l < m; l += nt) {
        for (int64_t e = B2.colptr[su[l]] - 1; e < B2.colptr[su[l] + 1] - 1; e++) {
            const int r = sls_find(sx, n, (int32_t)(B2.rowval[e] - 1));
            if (r >= 0) B2t[(size_t)l * n + r] = B2.nzval[e];
        }
    }
    for (int i = tid; i < P.T * n; i += nt) {
        const int t = i / n, r = i % n;
        Mx[i] = (P.maskx[((size_t)c * P.T + t) * P.wx + (r >> 5)] >> (r & 31)) & 1u;
    }
    for (int i = tid; i < P.T * m; i += nt) {
        const int t = i / m, r = i % m;
        Mu[i] = (P.masku[((size_t)c * P.T + t) * P.wu + (r >> 5)] >> (r & 31)) & 1u;
    }
}

// ---- host launchers -------------------------------------------------------------------------
int sls_launch_gather_count(const SlsCsc& A, const SlsSupports& Sx, const SlsSupports& Su, int T, int64_t Nx,
                            const int32_t* cols, int ncols, int32_t* n_out, int32_t* m_out, int cap,
                            sls_stream_t stream) {
    if (ncols == 0) return 0;
    const int grid = (ncols + GATHER_WPB - 1) / GATHER_WPB;
    const size_t smem = (size_t)GATHER_WPB * cap * sizeof(int32_t);
    auto kp = &sls_gather_count_kernel;
    if (smem > 48 * 1024) { int e = rt::set_max_smem(kp, (int)smem); if (e) return e; }
    SLS_LAUNCH(kp, grid, GATHER_WPB * 32, smem, stream, A, Sx, Su, T, Nx, cols, ncols, n_out, m_out, cap);
    return rt::last_error();
}

int sls_launch_gather_fill(const SlsCsc& A, const SlsCsc& B2, const SlsSupports& Sx, const SlsSupports& Su,
                           const SlsPlanView& P, int cap, sls_stream_t stream) {
    if (P.ncols == 0) return 0;
    const int grid = (P.ncols + GATHER_WPB - 1) / GATHER_WPB;
    const size_t smem = (size_t)GATHER_WPB * (cap + P.nstride + P.mstride + P.wx + P.wu) * sizeof(int32_t);
    auto kp = &sls_gather_fill_kernel;
    if (smem > 48 * 1024) { int e = rt::set_max_smem(kp, (int)smem); if (e) return e; }
    SLS_LAUNCH(kp, grid, GATHER_WPB * 32, smem, stream, A, B2, Sx, Su, P, cap);
    return rt::last_error();
}

int sls_launch_scan(int64_t* colptr, int64_t len, int narr, int64_t* tot, int64_t totlen, sls_stream_t stream) {
    auto kp = &sls_scan_kernel;
    SLS_LAUNCH(kp, narr + 1, SCAN_THREADS, 0, stream, colptr, len, narr, tot, totlen);
    return rt::last_error();
}

int sls_launch_local_dense(const SlsCsc& A, const SlsCsc& B2, const SlsPlanView& P, int c,
                           double* At, double* B2t, uint8_t* Mx, uint8_t* Mu, sls_stream_t stream) {
    auto kp = &sls_local_dense_kernel;
    SLS_LAUNCH(kp, 1, 128, 0, stream, A, B2, P, c, At, B2t, Mx, Mu);
    return rt::last_error();
}
