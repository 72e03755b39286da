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

#define SCATTER_MAXW 8

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

// One warp handles 32 consecutive planned columns of one tap.  Their index-set rows (s_x or s_u: 32 x stride ints,
// contiguous in the plan) are copied to shared memory with coalesced loads; every lane then expands ITS column's mask bits
// into row indices (find-first-set loop, ~6 instructions per entry against ~40 for a per-entry binary search +
// select-nth-bit) into a per-warp staging area at the offset a warp scan gives it, and the warp streams the staged rows
// out as full 256-byte coalesced stores.  When the planned columns ascend (a.sorted: every plan but an unsorted explicit
// column list) the 32 columns' segments are contiguous in rowval and the store address is base + e.
// Shared memory per warp: 33 + 2 * 32 * stride ints (stride = nstride for the Φx taps, mstride for the Φu taps); the
// launcher picks the warps per block that fit.
// RUNS: every index set of the plan is a run of consecutive rows (banded plants; kernel 1a's statistics): a row is the set's
// first row plus the mask bit, the index sets are not staged and a warp needs half the shared memory (twice the resident warps).
template <bool RUNS>
__global__ void __launch_bounds__(SCATTER_MAXW * 32)
sls_scatter_kernel(SlsScatterArgs a, int wpb, int per_warp) {
    SLS_DYN_SMEM(smem);
    const SlsPlanView& P = a.P;
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int tt = blockIdx.y;                                // 0..T-1: Φx taps, T..2T-1: Φu taps
    const int T = P.T;
    const bool isu = tt >= T;
    const int t = isu ? tt - T : tt;
    const int c0 = a.c_begin + (blockIdx.x * wpb + warp) * 32;
    if (c0 >= a.c_end) return;
    const int stride = isu ? P.mstride : P.nstride;
    int32_t* pref = reinterpret_cast<int32_t*>(smem) + (size_t)warp * per_warp;
    int32_t* rows = pref + 33;                                // [32][stride] index sets of the 32 columns (not with RUNS)
    int32_t* stage = RUNS ? rows : rows + 32 * stride;
    const int c = c0 + lane;
    const bool valid = c < a.c_end;
    const int ncol = min(32, a.c_end - c0);
    const int nw = isu ? P.wu : P.wx;
    const uint32_t* mbase = isu ? P.masku : P.maskx;
    const uint32_t* mw = mbase + ((size_t)(valid ? c : c0) * T + t) * nw;
    int32_t row0 = 0;
    if (RUNS) { if (valid) row0 = ((isu ? P.su : P.sx))[(size_t)c * stride]; }
    else {
        const int32_t* src = (isu ? P.su : P.sx) + (size_t)c0 * stride;
#pragma unroll 4
        for (int i = lane; i < ncol * stride; i += 32) rows[i] = src[i];
    }
    int cnt = 0;
    if (valid) for (int w = 0; w < nw; w++) cnt += __popc(mw[w]);
    int incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
    pref[lane + 1] = incl;
    if (lane == 0) pref[0] = 0;
    __syncwarp();
    const int E = pref[32];
    const int64_t* colptr = P.colptr + (int64_t)tt * (P.Nx + 1);
    int64_t* rowval = a.rowval[tt];
    const int64_t roff = 1 + (isu ? a.row_off_u : a.row_off_x);
    if (valid) {
        const int32_t* myrows = rows + lane * stride;
        int o = incl - cnt;
#pragma unroll 1
        for (int w = 0; w < nw; w++) {
            uint32_t v = mw[w];
#pragma unroll 1
            while (v) { const int b = __ffs(v) - 1; stage[o++] = RUNS ? row0 + 32 * w + b : myrows[32 * w + b]; v &= v - 1; }
        }
    }
    __syncwarp();
    if (a.sorted) {
        int64_t* dst = rowval + (colptr[P.cols[c0]] - 1);
#pragma unroll 4
        for (int e = lane; e < E; e += 32) dst[e] = (int64_t)stage[e] + roff;
    } else {
#pragma unroll 1
        for (int e = lane; e < E; e += 32) {
            int lo = 0, hi = 32;                              // largest cc with pref[cc] <= e
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (pref[mid] <= e) lo = mid; else hi = mid; }
            rowval[colptr[P.cols[c0 + lo]] - 1 + (e - pref[lo])] = (int64_t)stage[e] + roff;
        }
    }
}

int sls_launch_scatter(const SlsScatterArgs& a, sls_stream_t stream) {
    if (a.c_end <= a.c_begin) return 0;
    const int stride = a.P.nstride > a.P.mstride ? a.P.nstride : a.P.mstride;
    const bool runs = a.runs != 0;
    const int per_warp = (33 + (runs ? 1 : 2) * 32 * stride + 3) & ~3;     // ints
    int wpb = (int)((96 * 1024) / ((size_t)per_warp * sizeof(int32_t)));
    if (wpb > SCATTER_MAXW) wpb = SCATTER_MAXW;
    if (wpb < 1) wpb = 1;
    const size_t smem = (size_t)wpb * per_warp * sizeof(int32_t);
    const int per_block = wpb * 32;
    dim3 grid((a.c_end - a.c_begin + per_block - 1) / per_block, 2 * a.P.T, 1);
    if (runs) {
        auto kp = &sls_scatter_kernel<true>;
        if (smem > 48 * 1024) { int e = rt::set_max_smem(kp, (int)smem); if (e) return e; }
        SLS_LAUNCH(kp, grid, wpb * 32, smem, stream, a, wpb, per_warp);
    } else {
        auto kp = &sls_scatter_kernel<false>;
        if (smem > 48 * 1024) { int e = rt::set_max_smem(kp, (int)smem); if (e) return e; }
        SLS_LAUNCH(kp, grid, wpb * 32, smem, stream, a, wpb, per_warp);
    }
    return rt::last_error();
}

// colptr of one column chunk of the pipelined one-shot path (sls_synthesize): the chunk's plan holds, per tap, a
// colptr scanned over its own columns only; the caller's colptr is that plus the entries of the earlier chunks.
//     dst[tt][j] = src[tt][j] + base[tt]      for j in [j0, j1]   (0-based positions inside the Nx+1 array)
__global__ void sls_colptr_rebase_kernel(int64_t* dst, const int64_t* src, const int64_t* base, int64_t len, int64_t j0, int64_t j1) {
    const int tt = blockIdx.y;
    const int64_t j = j0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j <= j1) dst[(int64_t)tt * len + j] = src[(int64_t)tt * len + j] + base[tt];
}

int sls_launch_colptr_rebase(int64_t* dst, const int64_t* src, const int64_t* base, int64_t len, int narr, int64_t j0, int64_t j1,
                             sls_stream_t stream) {
    if (j1 < j0 || narr <= 0) return 0;
    dim3 grid((unsigned)((j1 - j0 + 1 + 255) / 256), narr, 1);
    auto kp = &sls_colptr_rebase_kernel;
    SLS_LAUNCH(kp, grid, 256, 0, stream, dst, src, base, len, j0, j1);
    return rt::last_error();
}

// Small metadata transfers of the plan (column lists, per-column sizes) as kernels over mapped pinned host memory
// instead of copy-engine commands: in the pipelined path the copy engines' queues hold megabytes of support /
// output slices, and a 50 KB size read-back queued behind them would stall kernel 1 of the next chunk for
// milliseconds.  Either pointer may be device memory or pinned host memory (UVA); 4-byte granularity.
__global__ void sls_kcopy_kernel(uint32_t* dst, const uint32_t* src, size_t nwords) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nwords; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
// 16-byte variant: SM-initiated PCIe reads are latency-bound per request, so wide requests matter when src is host memory
__global__ void sls_kcopy16_kernel(uint4* dst, const uint4* src, size_t nvec) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
int sls_kcopy(void* dst, const void* src, size_t bytes, sls_stream_t stream) {
    if (!bytes) return 0;
    const size_t nvec = ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) ? bytes / 16 : 0;
    if (nvec) {
        const int grid = (int)((nvec + 255) / 256 < 296 ? (nvec + 255) / 256 : 296);
        auto kp = &sls_kcopy16_kernel;
        SLS_LAUNCH(kp, grid, 256, 0, stream, reinterpret_cast<uint4*>(dst), reinterpret_cast<const uint4*>(src), nvec);
    }
    const size_t done = nvec * 16;
    if (done < bytes) {
        const size_t nwords = (bytes - done + 3) / 4;
        const int grid = (int)((nwords + 255) / 256 < 296 ? (nwords + 255) / 256 : 296);
        auto kp = &sls_kcopy_kernel;
        SLS_LAUNCH(kp, grid, 256, 0, stream, reinterpret_cast<uint32_t*>((char*)dst + done), reinterpret_cast<const uint32_t*>((const char*)src + done), nwords);
    }
    return rt::last_error();
}
// dst[i] = src[i * pitch], i < count  (the last colptr entry of every tap)
__global__ void sls_kgather64_kernel(int64_t* dst, const int64_t* src, size_t pitch, int count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) dst[i] = src[(size_t)i * pitch];
}
int sls_kgather64(int64_t* dst, const int64_t* src, size_t pitch, int count, sls_stream_t stream) {
    if (count <= 0) return 0;
    auto kp = &sls_kgather64_kernel;
    SLS_LAUNCH(kp, (count + 127) / 128, 128, 0, stream, dst, src, pitch, count);
    return rt::last_error();
}

// Diagonal cost weights (sls_set_weights): the weighted problem  min sum (wx.x)^2 + (wu.u)^2  is the unweighted one for
// the similarity-scaled plant  A^ = Wx A Wx^-1,  B2^ = Wx B2 Wu^-1  (uploaded by sls_set_plant) with
//     Phi_x[t][i,j] = (wx_j / wx_i) Phi^_x[t][i,j],     Phi_u[t][l,j] = (wx_j / wu_l) Phi^_u[t][l,j].
// This pass applies the back-scaling to the stored values of the planned columns [c_begin, c_end): one warp per column.
__global__ void __launch_bounds__(256)
sls_unscale_kernel(SlsPlanView P, int c_begin, int c_end, const int64_t* const* rowval, double* const* nzval,
                   const double* wx, const double* wu, int64_t row_off_x, int64_t row_off_u) {
    const int c = c_begin + blockIdx.x * 8 + (threadIdx.x >> 5);
    if (c >= c_end) return;
    const int lane = sls_lane();
    const int64_t j = P.cols[c];
    const double wj = wx[j];
    for (int tt = 0; tt < 2 * P.T; tt++) {
        const int64_t* cp = P.colptr + (int64_t)tt * (P.Nx + 1);
        const int64_t e0 = cp[j] - 1, e1 = cp[j + 1] - 1;
        const int64_t* rv = rowval[tt]; double* nz = nzval[tt];
        const double* w = (tt < P.T) ? wx : wu;
        const int64_t ro = (tt < P.T) ? row_off_x : row_off_u;
        for (int64_t e = e0 + lane; e < e1; e += 32) nz[e] *= wj / w[rv[e] - 1 - ro];
    }
}
int sls_launch_unscale(const SlsPlanView& P, int c_begin, int c_end, const int64_t* const* rowval, double* const* nzval,
                       const double* wx, const double* wu, int64_t row_off_x, int64_t row_off_u, sls_stream_t stream) {
    if (c_end <= c_begin) return 0;
    auto kp = &sls_unscale_kernel;
    SLS_LAUNCH(kp, (c_end - c_begin + 7) / 8, 256, 0, stream, P, c_begin, c_end, rowval, nzval, wx, wu, row_off_x, row_off_u);
    return rt::last_error();
}

// dst[i] = first + i  (the column list of a plan over all columns / a contiguous range: nothing to copy from the host)
__global__ void sls_iota_kernel(int32_t* dst, int32_t first, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = first + i;
}
int sls_launch_iota(int32_t* dst, int32_t first, int n, sls_stream_t stream) {
    if (n <= 0) return 0;
    auto kp = &sls_iota_kernel;
    SLS_LAUNCH(kp, (n + 255) / 256, 256, 0, stream, dst, first, n);
    return rt::last_error();
}
