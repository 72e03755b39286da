// sls_scatter.cu — kernel 3: the index half of the scatter into the per-tap CSC patterns.
//
// Replaces reference src/synthesis.jl:65-67: `sparse(rows=s_x, cols=j, vals=value.(Φ̃ₓ[t]) .* mask)` followed
// by the Θ(Nx) `Φ̃ += …` re-allocation per column.  The values (nzval) are written straight into the 2T tap-major
// CSC arrays by kernel 2's back substitution; this kernel writes the matching rowval (Int64, 1-based, ascending),
// derived from the mask bits and s_x / s_u.  It does not depend on kernel 2 and runs beside it on a second
// stream.  One warp handles 32 consecutive planned columns of one tap and flattens their (column, rank) pairs
// over the lanes so that rowval is written as full coalesced 256-byte stores.  HBM-bound.
#include "sls_common.h"
#include "sls_kernels.h"

#define SCATTER_WPB 8

__device__ __forceinline__ int sls_select_bit(const uint32_t* words, int nwords, int r) {
    // index of the r-th (0-based) set bit across `nwords` mask words
    int basebit = 0;
    for (int w = 0; w < nwords; w++) {
        const uint32_t v = words[w];
        const int pc = __popc(v);
        if (r < pc) return basebit + (int)__fns(v, 0, r + 1);
        r -= pc; basebit += 32;
    }
    return -1;
}

__global__ void __launch_bounds__(SCATTER_WPB * 32)
sls_scatter_kernel(SlsScatterArgs a) {
    __shared__ int32_t pref_s[SCATTER_WPB][33];
    const SlsPlanView& P = a.P;
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int tt = blockIdx.y;                                // 0..T-1: Φx taps, T..2T-1: Φu taps
    const int T = P.T;
    const bool isu = tt >= T;
    const int t = isu ? tt - T : tt;
    const int c0 = a.c_begin + (blockIdx.x * SCATTER_WPB + warp) * 32;
    if (c0 >= a.c_end) return;
    int32_t* pref = pref_s[warp];
    const int c = c0 + lane;
    const bool valid = c < a.c_end;
    const int nw = isu ? P.wu : P.wx;
    const uint32_t* mbase = isu ? P.masku : P.maskx;
    int cnt = 0;
    if (valid) {
        const uint32_t* mw = mbase + ((size_t)c * T + t) * nw;
        for (int w = 0; w < nw; w++) cnt += __popc(mw[w]);
    }
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
    pref[lane + 1] = incl;
    if (lane == 0) pref[0] = 0;
    __syncwarp();
    const int E = pref[32];
    const int64_t* colptr = P.colptr + (int64_t)tt * (P.Nx + 1);
    int64_t* rowval = a.rowval[tt];
    for (int e = lane; e < E; e += 32) {
        int lo = 0, hi = 32;                                  // largest cc with pref[cc] <= e
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pref[mid] <= e) lo = mid; else hi = mid; }
        const int cc = c0 + lo;
        const int r = e - pref[lo];
        const int64_t j = P.cols[cc];
        const int64_t dst = colptr[j] - 1 + r;
        const uint32_t* mw = mbase + ((size_t)cc * T + t) * nw;
        const int li = sls_select_bit(mw, nw, r);
        const int32_t grow = isu ? P.su[(size_t)cc * P.mstride + li] : P.sx[(size_t)cc * P.nstride + li];
        rowval[dst] = (int64_t)grow + 1;
    }
}

int sls_launch_scatter(const SlsScatterArgs& a, sls_stream_t stream) {
    if (a.c_end <= a.c_begin) return 0;
    const int per_block = SCATTER_WPB * 32;
    dim3 grid((a.c_end - a.c_begin + per_block - 1) / per_block, 2 * a.P.T, 1);
    auto kp = &sls_scatter_kernel;
    SLS_LAUNCH(kp, grid, SCATTER_WPB * 32, 0, stream, a);
    return rt::last_error();
}
