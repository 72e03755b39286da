// sls_common.h — device-side views of the plant, the supports and the per-plan metadata.
// All global index arrays keep the caller's Julia layout (int64, 1-based) so that sls_set_* is a
// plain H2D copy; kernels subtract 1 on the fly.  Per-column metadata produced by kernel 1 is int32.
#pragma once
#include "sls_rt.h"

#define SLS_WARP 32
#define SLS_FULL 0xffffffffu
// per-column status bits (mirror include/slsb200.h)
#define SLS_ST_RESIDUAL 1
#define SLS_ST_RANK_SKIPPED 2
#define SLS_ST_NO_IMPULSE 4
#define SLS_ST_COST_SINGULAR 8
#define SLS_ST_GROUP_PATTERN 16

struct SlsCsc {                 // Julia SparseMatrixCSC{Float64,Int64} on the device
    const int64_t* colptr;      // ncol+1, 1-based
    const int64_t* rowval;      // nnz, 1-based, ascending per column
    const double*  nzval;
};

struct SlsSupports {            // T Bool CSC matrices, concatenated
    const int64_t* colptr;      // T x (Nx+1), 1-based within the tap
    const int64_t* rowval;      // all taps back to back
    const uint8_t* nzval;       // same indexing as rowval, or nullptr (= all true)
    const int64_t* tap_off;     // T+1 offsets of each tap inside rowval
    int64_t row_off;            // stored row - row_off = row of the handle's plant (0 unless the plant is a sub-plant window
                                // and the supports keep their global row numbers: sls_synthesize_window)
};

struct SlsPlanView {            // per planned column c (0 <= c < ncols), written by kernel 1
    int64_t Nx, Nu;
    int32_t T, ncols;
    int32_t nstride, mstride;   // row strides of sx / su (max n, max m over the plan)
    int32_t wx, wu;             // mask words per tap (ceil(nstride/32), ceil(mstride/32))
    const int32_t* cols;        // 0-based disturbance column j of planned column c
    int32_t* n;  int32_t* m;    // |s_x|, |s_u|
    int32_t* sx; int32_t* su;   // 0-based global rows, ascending
    uint32_t* maskx; uint32_t* masku;   // [c][t][w]
    int32_t* jloc;              // local index of j inside s_x, or -1 (ii_x, reduction.jl:22)
    int32_t* lnnz;              // nnz(A[sx,sx]) + nnz(B2[sx,su])  (local CSR/CSC capacity)
    int32_t* lnnzA;             // nnz(A[sx,sx])
    int32_t* lnnzB;             // nnz(B2[:,su]) for the SURVEY §8d byte count
    uint32_t* pre;              // [c][2T]: offsets of (x_t, u_t) segments inside the column's packed result
    int64_t* tot;               // [ncols+1]: packed entries per column -> exclusive scan = result offsets
    int64_t* colptr;            // [2T][Nx+1]: counts at index j -> scanned in place to Julia colptr (1-based)
};

__device__ __forceinline__ int sls_find(const int32_t* s, int n, int key) {
    int lo = 0, hi = n;
    while (lo < hi) { int mid = (lo + hi) >> 1; if (s[mid] < key) lo = mid + 1; else hi = mid; }
    return (lo < n && s[lo] == key) ? lo : -1;
}

__device__ __forceinline__ int sls_lane() { return threadIdx.x & 31; }

// In-place ascending sort + unique of buf[0..cnt) by one warp (bitonic network in shared memory, padded with INT_MAX to
// the next power of two <= cap; cap >= 32 is a power of two and cnt <= cap).  Returns the number of distinct values
// (INT_MAX entries are padding and are dropped).  Shared by kernel 1 (index sets), the support generator and the
// simulation windows.
__device__ __forceinline__ int sls_warp_sort_unique(int32_t* buf, int cnt, int cap) {
    const int lane = sls_lane();
    int P = 32;
#pragma unroll 1
    while (P < cnt) P <<= 1;
    if (P > cap) P = cap;
#pragma unroll 1
    for (int i = cnt + lane; i < P; i += 32) buf[i] = 0x7fffffff;
    __syncwarp();
#pragma unroll 1
    for (int k = 2; k <= P; k <<= 1)
#pragma unroll 1
        for (int j = k >> 1; j > 0; j >>= 1) {
#pragma unroll 1
            for (int i = lane; i < P; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const int x = buf[i], y = buf[ixj];
                    const bool asc = (i & k) == 0;
                    if ((x > y) == asc) { buf[i] = y; buf[ixj] = x; }
                }
            }
            __syncwarp();
        }
    int n = 0, prev_last = -1;
#pragma unroll 1
    for (int b0 = 0; b0 < P; b0 += 32) {
        const int v = buf[b0 + lane];
        int pv = __shfl_up_sync(SLS_FULL, v, 1);
        if (lane == 0) pv = prev_last;
        const bool keep = (v != 0x7fffffff) && (v != pv);
        const unsigned bal = __ballot_sync(SLS_FULL, keep);
        __syncwarp();
        if (keep) buf[n + __popc(bal & ((1u << lane) - 1u))] = v;
        n += __popc(bal);
        prev_last = __shfl_sync(SLS_FULL, v, 31);
        __syncwarp();
    }
    return n;
}

// out-of-line copy for cold code (per-column setup of kernel 2): keeps the hot loops' code footprint small
static __device__ __noinline__ int sls_find_cold(const int32_t* s, int n, int key) { return sls_find(s, n, key); }
