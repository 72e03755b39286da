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
#include <algorithm>

#define GATHER_WPB 4

// Sorted-unique union of the rows of S[T-1][:,k] over the stored nonzeros k of A[:,j].
// Result in buf[0..ret); ret = -1 if the candidate list overflows `cap` (a power of two).
//
// Fast path (banded plants: every segment is a run of consecutive rows).  Support rows are strictly ascending (validated on
// the device behind their upload, sls_validate_supports_kernel), so a segment whose last row minus its first equals its length
// minus one IS that run and only its two end rows are read.  The union of the runs is built as a bitmap over [min first, max
// last] (one 32-bit word per lane) and expanded in order: no candidate list, no sort.  Anything else — a segment with a gap, more
// than 32 segments, a span above 1024 rows — takes the general path below with identical results.
__device__ int sls_warp_index_set(const SlsCsc& A, const SlsSupports& S, int T, int64_t Nx, int64_t j,
                                  int32_t* buf, int cap) {
    const int lane = sls_lane();
    const int64_t a0 = A.colptr[j] - 1, a1 = A.colptr[j + 1] - 1;
    const int64_t* scp = S.colptr + (int64_t)(T - 1) * (Nx + 1);
    const int64_t base = S.tap_off[T - 1];
    if (a1 - a0 <= 32) {
        int len = 0, first = 0, last = -1;
        if (lane < a1 - a0 && A.nzval[a0 + lane] != 0.0) {    // `A .≠ 0` drops stored zeros
            const int64_t k = A.rowval[a0 + lane] - 1;
            const int64_t s0 = base + scp[k] - 1, s1 = base + scp[k + 1] - 1;
            len = (int)(s1 - s0);
            if (len > 0) {
                first = (int32_t)(S.rowval[s0] - 1 - S.row_off);
                last = (int32_t)(S.rowval[s1 - 1] - 1 - S.row_off);
            }
        }
        const bool isrun = (len == 0) || (last - first == len - 1);
        const int C = __reduce_add_sync(SLS_FULL, len);
        if (C > cap) return -1;                               // same overflow rule as the general path (candidates, not distinct rows)
        if (C == 0) return 0;
        if (__all_sync(SLS_FULL, isrun)) {
            const int lo = -__reduce_max_sync(SLS_FULL, len > 0 ? -first : INT_MIN);
            const int hi = __reduce_max_sync(SLS_FULL, len > 0 ? last : INT_MIN);
            if (hi - lo < 1024) {
                const int nseg = (int)(a1 - a0);
                uint32_t word = 0u;                           // rows lo + 32*lane .. lo + 32*lane + 31
#pragma unroll 1
                for (int e = 0; e < nseg; e++) {
                    const int le = __shfl_sync(SLS_FULL, len, e);
                    const int f = __shfl_sync(SLS_FULL, first, e) - lo - 32 * lane, l = __shfl_sync(SLS_FULL, last, e) - lo - 32 * lane;
                    if (le > 0 && l >= 0 && f <= 31) {
                        const uint32_t upto = (l >= 31) ? 0xffffffffu : ((2u << l) - 1u);
                        const uint32_t from = (f <= 0) ? 0xffffffffu : ~((1u << f) - 1u);
                        word |= upto & from;
                    }
                }
                const int cw = __popc(word);
                int incl = cw;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
                const int excl = incl - cw;
                const int nw = ((hi - lo) >> 5) + 1;
                __syncwarp();
#pragma unroll 1
                for (int w = 0; w < nw; w++) {
                    const uint32_t ww = __shfl_sync(SLS_FULL, word, w);
                    const int off = __shfl_sync(SLS_FULL, excl, w);
                    if ((ww >> lane) & 1u) buf[off + __popc(ww & ((1u << lane) - 1u))] = lo + 32 * w + lane;
                }
                __syncwarp();
                return __shfl_sync(SLS_FULL, incl, 31);
            }
        }
    }
    int C = 0;
#pragma unroll 1
    for (int64_t e = a0; e < a1; e++) {
        if (A.nzval[e] == 0.0) continue;                      // `A .≠ 0` drops stored zeros
        const int64_t k = A.rowval[e] - 1;
        const int64_t s0 = base + scp[k] - 1, s1 = base + scp[k + 1] - 1;
        const int len = (int)(s1 - s0);
        if (C + len > cap) return -1;
#pragma unroll 1
        for (int i = lane; i < len; i += 32) buf[C + i] = (int32_t)(S.rowval[s0 + i] - 1 - S.row_off);
        C += len;
    }
    __syncwarp();
    return sls_warp_sort_unique(buf, C, cap);
}

// kernel 1a: sizes only (n, m) so that the host can size the strided metadata exactly.
// Plan statistics are reduced inside the block (shared-memory atomics) and flushed once per block: with one global
// atomic per column and statistic, 100 000 columns queued ~500 000 atomics on five addresses of one L2 sector, and the
// kernel ran at the L2's same-address atomic rate (0.28 ms) whatever else it did.
#define COUNT_MAX_WPB 16
__global__ void __launch_bounds__(COUNT_MAX_WPB * 32)
sls_gather_count_kernel(SlsCsc A, SlsSupports Sx, SlsSupports Su, int T, int64_t Nx,
                        const int32_t* __restrict__ cols, int ncols, int32_t* n_out, int32_t* m_out, int cap,
                        int32_t* keep_x, int32_t* keep_u, const int32_t* __restrict__ bad_supports, int32_t* stats) {
    SLS_DYN_SMEM(smem);
    __shared__ int s_stats[36];
    const int wpb = blockDim.x >> 5;
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int c = blockIdx.x * wpb + warp;
    if (threadIdx.x < 36) s_stats[threadIdx.x] = 0;
    __syncthreads();
    // malformed support rows (sls_validate_supports_kernel): index nothing
    if (c < ncols && !(bad_supports && *bad_supports)) {
        int32_t* buf = reinterpret_cast<int32_t*>(smem) + (size_t)warp * cap;
        const int64_t j = cols[c];
        // the index sets are kept (row stride cap) so that kernel 1b does not gather and sort them a second time
        const int n = sls_warp_index_set(A, Sx, T, Nx, j, buf, cap);
        // [3]: some index set of the plan is NOT a run of consecutive rows (kernel 3 then stages the index sets; for banded
        // plants it derives a row from the set's first row and the mask bit)
        bool notrun = n > 1 && (buf[n - 1] - buf[0] != n - 1);
        if (keep_x)
#pragma unroll 1
            for (int i = lane; i < n; i += 32) keep_x[(size_t)c * cap + i] = buf[i];
        __syncwarp();
        const int m = sls_warp_index_set(A, Su, T, Nx, j, buf, cap);
        notrun = notrun || (m > 1 && (buf[m - 1] - buf[0] != m - 1));
        if (keep_u)
#pragma unroll 1
            for (int i = lane; i < m; i += 32) keep_u[(size_t)c * cap + i] = buf[i];
        if (lane == 0) {
            n_out[c] = n; m_out[c] = m;
            // plan statistics on the device (the host reads 36 integers instead of 2 x ncols sizes): [0] max n, [1] max m,
            // [2] overflow, [3] some index set is not a run of consecutive rows, [4 + k] columns of size class k (n in 8k+1..8k+8; class 15 = everything larger),
            // [20 + k] largest m of class k
            if (stats) {
                if (n < 0 || m < 0) atomicOr(s_stats + 2, 1);
                else {
                    if (notrun) atomicOr(s_stats + 3, 1);
                    atomicMax(s_stats + 0, n); atomicMax(s_stats + 1, m);
                    int k = n > 0 ? (n - 1) >> 3 : 0; if (k > 15) k = 15;
                    atomicAdd(s_stats + 4 + k, 1); atomicMax(s_stats + 20 + k, m);
                }
            }
        }
    }
    __syncthreads();
    if (stats && threadIdx.x < 36) {
        const int i = threadIdx.x, v = s_stats[i];
        if (v != 0) {                                         // every statistic starts at 0 and only grows
            if (i == 2 || i == 3) atomicOr(stats + i, v);
            else if (i >= 4 && i < 20) atomicAdd(stats + i, v);
            else atomicMax(stats + i, v);
        }
    }
}

// Kernel-2 work lists: block k writes the planned-column indices of size class k, ascending, behind those of the smaller
// classes (stable stream compaction: 4 columns per thread per round, warp-shuffle scans, one pass over n[] per non-empty
// class).
__global__ void __launch_bounds__(1024)
sls_class_lists_kernel(const int32_t* __restrict__ n, int ncols, const int32_t* __restrict__ stats, int nclass, int32_t* lists) {
    __shared__ int wsum[32];
    __shared__ int carry_s;
    const int k = blockIdx.x;
    if (k >= nclass || stats[4 + k] == 0) return;
    int off = 0;
    for (int kk = 0; kk < k; kk++) off += stats[4 + kk];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int i0 = 0; i0 < ncols; i0 += 4096) {
        const int i = i0 + 4 * tid;
        bool mine[4];
        int cnt = 0;
#pragma unroll
        for (int e = 0; e < 4; e++) {
            mine[e] = false;
            if (i + e < ncols) { const int nn = n[i + e]; int kc = nn > 0 ? (nn - 1) >> 3 : 0; if (kc > 15) kc = 15; mine[e] = (kc == k); }
            cnt += mine[e];
        }
        int incl = cnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
        if (lane == 31) wsum[warp] = incl;
        __syncthreads();
        int wv = wsum[lane];                                   // 32 warps: one more warp-level scan, done redundantly by every warp
        int winc = wv;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const int o = __shfl_up_sync(SLS_FULL, winc, d); if (lane >= d) winc += o; }
        const int before = __shfl_sync(SLS_FULL, winc - wv, warp);
        const int total = __shfl_sync(SLS_FULL, winc, 31);
        int pos = off + carry_s + before + incl - cnt;
#pragma unroll
        for (int e = 0; e < 4; e++) if (mine[e]) lists[pos++] = i + e;
        __syncthreads();
        if (tid == 0) carry_s += total;
        __syncthreads();
    }
}

int sls_launch_class_lists(const int32_t* n, int ncols, const int32_t* stats, int nclass, int32_t* lists, sls_stream_t stream) {
    if (ncols == 0 || nclass <= 0) return 0;
    auto kp = &sls_class_lists_kernel;
    SLS_LAUNCH(kp, nclass, 1024, 0, stream, n, ncols, stats, nclass, lists);
    return rt::last_error();
}

// 32-bit words of shared memory one warp of kernel 1b needs (kept even so that the int64 section stays aligned)
__host__ __device__ int sls_gather_fill_words(int cap, int nstride, int mstride, int wx, int wu, int T) {
    const int w = 4 * T + cap + nstride + mstride + T * (wx + wu) + (2 * T + 1) + 2 * T;
    return (w + 1) & ~1;
}

// every one of the len bytes at p is non-zero: aligned 4-byte words in the middle (a zero byte shows as a set high bit of
// (v - 0x01010101) & ~v), single bytes up to the first and after the last aligned word — nothing outside [p, p + len) is read
__device__ __forceinline__ bool sls_all_true(const uint8_t* p, int len) {
    bool ok = true;
    int i = 0;
    const int head = (int)((4 - (reinterpret_cast<uintptr_t>(p) & 3)) & 3);
#pragma unroll 1
    for (; i < head && i < len; i++) ok = ok && (p[i] != 0);
#pragma unroll 1
    for (; i + 4 <= len; i += 4) {
        const uint32_t v = *reinterpret_cast<const uint32_t*>(p + i);
        ok = ok && (((v - 0x01010101u) & ~v & 0x80808080u) == 0u);
    }
#pragma unroll 1
    for (; i < len; i++) ok = ok && (p[i] != 0);
    return ok;
}

// kernel 1b: index sets, ii_x, masks, per-tap counts, packed-result offsets, local nnz.
// The 2T support columns 𝓢ₓ[t][:,j], 𝓢ᵤ[t][:,j] of a disturbance column are read as ONE flattened list
// (segment table + prefix sums in shared memory), so all their loads are independent and in flight together
// instead of 2T dependent colptr -> rowval round trips.
__global__ void __launch_bounds__(GATHER_WPB * 32)
sls_gather_fill_kernel(SlsCsc A, SlsCsc B2, SlsSupports Sx, SlsSupports Su, SlsPlanView P, int cap,
                       const int32_t* __restrict__ keep_x, const int32_t* __restrict__ keep_u) {
    SLS_DYN_SMEM(smem);
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int c = blockIdx.x * GATHER_WPB + warp;
    if (c >= P.ncols) return;
    const int T = P.T;
    const int nmw = P.wx + P.wu;
    const int per_warp = sls_gather_fill_words(cap, P.nstride, P.mstride, P.wx, P.wu, T);
    int32_t* wbase = reinterpret_cast<int32_t*>(smem) + (size_t)warp * per_warp;
    int64_t* seg0 = reinterpret_cast<int64_t*>(wbase);          // [2T] first entry of each support column
    int32_t* buf = wbase + 4 * T;
    int32_t* sx = buf + cap;
    int32_t* su = sx + P.nstride;
    uint32_t* mw = reinterpret_cast<uint32_t*>(su + P.mstride); // [T][wx+wu] mask words of every tap
    int32_t* pfx = reinterpret_cast<int32_t*>(mw + T * nmw);    // [2T+1] prefix sums of the segment lengths
    int32_t* cnt = pfx + 2 * T + 1;                             // [2T] stored entries per tap (x taps, then u taps)
    const int64_t Nx = P.Nx;
    const int64_t j = P.cols[c];

    // segment table first: these loads overlap with the index-set work below
#pragma unroll 1
    for (int tt = lane; tt < 2 * T; tt += 32) {
        const SlsSupports& S = (tt < T) ? Sx : Su;
        const int t = (tt < T) ? tt : tt - T;
        const int64_t* cp = S.colptr + (int64_t)t * (Nx + 1);
        const int64_t s0 = S.tap_off[t] + cp[j] - 1, s1 = S.tap_off[t] + cp[j + 1] - 1;
        seg0[tt] = s0; cnt[tt] = (int)(s1 - s0);
    }
#pragma unroll 1
    for (int w = lane; w < T * nmw; w += 32) mw[w] = 0u;

    int n, m;
    if (keep_x && keep_u) {                                   // index sets left by kernel 1a (P.n / P.m hold their sizes)
        n = P.n[c]; m = P.m[c];
#pragma unroll 1
        for (int i = lane; i < n; i += 32) { const int32_t r = keep_x[(size_t)c * cap + i]; sx[i] = r; P.sx[(size_t)c * P.nstride + i] = r; }
#pragma unroll 1
        for (int i = lane; i < m; i += 32) { const int32_t r = keep_u[(size_t)c * cap + i]; su[i] = r; P.su[(size_t)c * P.mstride + i] = r; }
        __syncwarp();
    } else {
        n = sls_warp_index_set(A, Sx, T, Nx, j, buf, cap);
#pragma unroll 1
        for (int i = lane; i < n; i += 32) { sx[i] = buf[i]; P.sx[(size_t)c * P.nstride + i] = buf[i]; }
        __syncwarp();
        m = sls_warp_index_set(A, Su, T, Nx, j, buf, cap);
#pragma unroll 1
        for (int i = lane; i < m; i += 32) { su[i] = buf[i]; P.su[(size_t)c * P.mstride + i] = buf[i]; }
        __syncwarp();
    }
    if (lane == 0) { P.n[c] = n; P.m[c] = m; P.jloc[c] = sls_find(sx, n, (int32_t)j); }
    // an index set that is a run of consecutive rows (banded plants) maps a row to its local index by subtraction
    const bool runx = n > 0 && sx[n - 1] - sx[0] == n - 1, runu = m > 0 && su[m - 1] - su[0] == m - 1;

    // Fast path, one lane per support column 𝓢[t][:,j]: when the index set is a run and the column's stored rows are a run too
    // (strictly ascending rows: last - first == len - 1 proves it from the two end rows) with every stored value true, the tap's
    // mask is the bit range [first - s[0], last - s[0]] clipped to the index set: no per-entry work.  Such columns leave the
    // flattened list below (length 0); everything else (gaps, stored-false entries, scattered index sets) stays on it.
#pragma unroll 1
    for (int tt = lane; tt < 2 * T; tt += 32) {
        const int len = cnt[tt];
        const bool isu = tt >= T;
        bool fast = (len == 0);
        if (len > 0 && (isu ? runu : runx)) {
            const SlsSupports& S = isu ? Su : Sx;
            const int64_t s0 = seg0[tt];
            const int32_t first = (int32_t)(S.rowval[s0] - 1 - S.row_off), last = (int32_t)(S.rowval[s0 + len - 1] - 1 - S.row_off);
            if (last - first == len - 1) {
                const bool allon = !S.nzval || sls_all_true(S.nzval + s0, len);
                if (allon) {
                    const int b0 = isu ? su[0] : sx[0], nn = isu ? m : n;
                    const int lo = max(first - b0, 0), hi = min(last - b0, nn - 1);
                    uint32_t* w = mw + (isu ? tt - T : tt) * nmw + (isu ? P.wx : 0);
#pragma unroll 1
                    for (int q = lo >> 5; q <= (hi >> 5); q++) {           // empty when lo > hi
                        const int l = max(lo - 32 * q, 0), h = min(hi - 32 * q, 31);
                        if (l <= h) w[q] = ((h == 31) ? 0xffffffffu : ((2u << h) - 1u)) & ~((1u << l) - 1u);
                    }
                    fast = true;
                }
            }
        }
        if (fast) cnt[tt] = 0;
    }
    __syncwarp();

    // prefix sums of the segment lengths still on the list
    {
        int carry = 0;
#pragma unroll 1
        for (int t0 = 0; t0 < 2 * T; t0 += 32) {
            const int tt = t0 + lane;
            const int v = (tt < 2 * T) ? cnt[tt] : 0;
            int incl = v;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
            if (tt < 2 * T) pfx[tt] = carry + incl - v;
            carry += __shfl_sync(SLS_FULL, incl, 31);
        }
        if (lane == 0) pfx[2 * T] = carry;
    }
    __syncwarp();
    const int total = pfx[2 * T];
    int tt = 0;                                               // a lane's entries f = lane, lane+32, ... move forward through the taps
#pragma unroll 1
    for (int f = lane; f < total; f += 32) {
#pragma unroll 1
        while (pfx[tt + 1] <= f) tt++;                        // largest tt with pfx[tt] <= f  (pfx[2T] = total > f)
        const bool isu = tt >= T;
        const int t = isu ? tt - T : tt;
        const SlsSupports& S = isu ? Su : Sx;
        const int64_t e = seg0[tt] + (f - pfx[tt]);
        const bool on = S.nzval ? (S.nzval[e] != 0) : true;   // `𝓢[t][s,cⱼ] .≠ 1` fixes the entry (synthesis.jl:58-59)
        const int32_t row = (int32_t)(S.rowval[e] - 1 - S.row_off);
        int idx;
        if (isu) { if (runu) { idx = row - su[0]; if (idx < 0 || idx >= m) idx = -1; } else idx = sls_find(su, m, row); }
        else { if (runx) { idx = row - sx[0]; if (idx < 0 || idx >= n) idx = -1; } else idx = sls_find(sx, n, row); }
        if (on && idx >= 0) atomicOr(&mw[t * nmw + (isu ? P.wx : 0) + (idx >> 5)], 1u << (idx & 31));
    }
    __syncwarp();
    // per-tap counts, masks to global, colptr counts
#pragma unroll 1
    for (int tt = lane; tt < 2 * T; tt += 32) {
        const bool isu = tt >= T;
        const int t = isu ? tt - T : tt;
        const uint32_t* w = mw + t * nmw + (isu ? P.wx : 0);
        const int nwd = isu ? P.wu : P.wx;
        int k = 0;
#pragma unroll 1
        for (int q = 0; q < nwd; q++) k += __popc(w[q]);
        cnt[tt] = k;
        P.colptr[(int64_t)tt * (Nx + 1) + j] = k;
        uint32_t* dst = isu ? P.masku + ((size_t)c * T + t) * P.wu : P.maskx + ((size_t)c * T + t) * P.wx;
#pragma unroll 1
        for (int q = 0; q < nwd; q++) dst[q] = w[q];
    }
    __syncwarp();
    // offsets of the (x_t, u_t) segments inside the column's packed result: order x_0 u_0 x_1 u_1 ...
    uint32_t run = 0;
#pragma unroll 1
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        const int cx = (t < T) ? cnt[t] : 0, cu = (t < T) ? cnt[T + t] : 0;
        int incl = cx + cu;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
        if (t < T) {
            const uint32_t off = run + (uint32_t)(incl - cx - cu);
            P.pre[(size_t)c * 2 * T + 2 * t] = off;
            P.pre[(size_t)c * 2 * T + 2 * t + 1] = off + (uint32_t)cx;
        }
        run += (uint32_t)__shfl_sync(SLS_FULL, incl, 31);
    }
    // local nnz: A[sx,sx] by columns, B2[sx,su] by columns, and the full column lengths of B2[:,su]
    int cntA = 0, cntB = 0, cntBfull = 0;
#pragma unroll 1
    for (int l = lane; l < n; l += 32) {
        const int64_t e0 = A.colptr[sx[l]] - 1, e1 = A.colptr[sx[l] + 1] - 1;
#pragma unroll 1
        for (int64_t e = e0; e < e1; e++) {
            const int32_t r = (int32_t)(A.rowval[e] - 1);
            cntA += runx ? (r >= sx[0] && r <= sx[n - 1]) : (sls_find(sx, n, r) >= 0);
        }
    }
#pragma unroll 1
    for (int l = lane; l < m; l += 32) {
        const int64_t e0 = B2.colptr[su[l]] - 1, e1 = B2.colptr[su[l] + 1] - 1;
        cntBfull += (int)(e1 - e0);
#pragma unroll 1
        for (int64_t e = e0; e < e1; e++) {
            const int32_t r = (int32_t)(B2.rowval[e] - 1);
            cntB += runx ? (r >= sx[0] && r <= sx[n - 1]) : (sls_find(sx, n, r) >= 0);
        }
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
// Only the window [lo, hi) of a colptr array can hold non-zero counts (the planned columns): it is scanned, the
// entries before it are set to 1 and the entries from hi on to 1 + total, so the serial part of the scan does not
// grow with Nx when one GPU plans a shard of a larger plant (weak scaling).  Nothing outside [lo, hi) is read.
// With a scratch array (`partial`, (narr+1) x nchunk sums written by sls_scan_sum_kernel) every array is cut into nchunk
// pieces of `chunk` entries scanned by separate blocks (grid (narr+1) x nchunk): the 2T + 1 serial block-loops of the
// one-block-per-array form (13 tiles of four barriers each at Nx = 100 000) become one tile per block.
// entries [T0, L) of array b (a colptr array: the window [lo, hi); the tot array: [0, totlen)) that chunk ch covers
__device__ __forceinline__ void sls_scan_range(int b, int narr, int64_t totlen, int64_t lo, int64_t hi, int64_t chunk,
                                               int ch, int64_t& T0, int64_t& L) {
    const int64_t a0 = (b < narr) ? lo : 0, a1 = (b < narr) ? hi : totlen;
    T0 = a0 + (int64_t)ch * chunk; if (T0 > a1) T0 = a1;
    L = T0 + chunk; if (L > a1) L = a1;
}

__global__ void __launch_bounds__(SCAN_THREADS)
sls_scan_sum_kernel(const int64_t* colptr, int64_t len, int narr, const int64_t* tot, int64_t totlen, int64_t lo, int64_t hi,
                    int64_t chunk, int64_t* partial) {
    __shared__ int64_t wsum[SCAN_THREADS / 32];
    const int b = blockIdx.x, ch = blockIdx.y;
    const int64_t* arr = (b < narr) ? colptr + (int64_t)b * len : tot;
    int64_t T0, L;
    sls_scan_range(b, narr, totlen, lo, hi, chunk, ch, T0, L);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int64_t sum = 0;
#pragma unroll 1
    for (int64_t i = T0 + tid; i < L; i += SCAN_THREADS) sum += arr[i];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) sum += __shfl_xor_sync(SLS_FULL, sum, d);
    if (lane == 0) wsum[warp] = sum;
    __syncthreads();
    if (warp == 0) {
        int64_t w = wsum[lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(SLS_FULL, w, d);
        if (lane == 0) partial[(int64_t)b * gridDim.y + ch] = w;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS)
sls_scan_kernel(int64_t* colptr, int64_t len, int narr, int64_t* tot, int64_t totlen, int64_t lo, int64_t hi,
                int64_t chunk, const int64_t* partial) {
    __shared__ int64_t wsum[SCAN_THREADS / 32];
    __shared__ int64_t carry_s, tile_total_s;
    const int b = blockIdx.x, ch = blockIdx.y;
    int64_t* arr = (b < narr) ? colptr + (int64_t)b * len : tot;
    int64_t T0, L;
    sls_scan_range(b, narr, totlen, lo, hi, chunk, ch, T0, L);
    const int64_t add = (b < narr) ? 1 : 0;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {   // carry into this chunk: the sums of the chunks before it
        int64_t pre = 0;
        if (partial)
#pragma unroll 1
            for (int c = tid; c < ch; c += SCAN_THREADS) pre += partial[(int64_t)b * gridDim.y + c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) pre += __shfl_xor_sync(SLS_FULL, pre, d);
        if (lane == 0) wsum[warp] = pre;
        __syncthreads();
        if (warp == 0) {
            int64_t w = wsum[lane];
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) w += __shfl_xor_sync(SLS_FULL, w, d);
            if (lane == 0) carry_s = w;
        }
        __syncthreads();
    }
    for (int64_t tile = T0; tile < L; tile += (int64_t)SCAN_THREADS * SCAN_ITEMS) {
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
    // the value of every entry from hi on (sls_colptr_fill_kernel): written by the block of the last chunk
    if (b < narr && ch == (int)gridDim.y - 1 && tid == 0 && hi < len) arr[hi] = carry_s + 1;
}

// entries outside the scanned window: 1 before it, arr[hi] (= 1 + total, written by the scan) after it
__global__ void sls_colptr_fill_kernel(int64_t* colptr, int64_t len, int64_t lo, int64_t hi) {
    int64_t* arr = colptr + (int64_t)blockIdx.y * len;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < lo) arr[i] = 1;
    else if (i > hi && i < len) arr[i] = arr[hi];
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
#pragma unroll 1
    for (int i = tid; i < n * n; i += nt) At[i] = 0.0;
#pragma unroll 1
    for (int i = tid; i < n * m; i += nt) B2t[i] = 0.0;
    __syncthreads();
#pragma unroll 1
    for (int l = tid; l < n; l += nt) {
#pragma unroll 1
        for (int64_t e = A.colptr[sx[l]] - 1; e < A.colptr[sx[l] + 1] - 1; e++) {
            const int r = sls_find(sx, n, (int32_t)(A.rowval[e] - 1));
            if (r >= 0) At[(size_t)l * n + r] = A.nzval[e];
        }
    }
#pragma unroll 1
    for (int l = tid; l < m; l += nt) {
#pragma unroll 1
        for (int64_t e = B2.colptr[su[l]] - 1; e < B2.colptr[su[l] + 1] - 1; e++) {
            const int r = sls_find(sx, n, (int32_t)(B2.rowval[e] - 1));
            if (r >= 0) B2t[(size_t)l * n + r] = B2.nzval[e];
        }
    }
#pragma unroll 1
    for (int i = tid; i < P.T * n; i += nt) {
        const int t = i / n, r = i % n;
        Mx[i] = (P.maskx[((size_t)c * P.T + t) * P.wx + (r >> 5)] >> (r & 31)) & 1u;
    }
#pragma unroll 1
    for (int i = tid; i < P.T * m; i += nt) {
        const int t = i / m, r = i % m;
        Mu[i] = (P.masku[((size_t)c * P.T + t) * P.wu + (r >> 5)] >> (r & 31)) & 1u;
    }
}

// Layout check of the uploaded support columns [c0, c1) of every tap: row indices within 1..nrow and strictly ascending
// inside a column (the binary searches and the run-of-consecutive-rows shortcut of kernel 1 rely on it; out-of-range
// rows would index A / B2 out of bounds).  One thread per (tap, column); sets *flag on the first violation.
__global__ void __launch_bounds__(256)
sls_validate_supports_kernel(SlsSupports S, int64_t Nx, int64_t nrow, int64_t c0, int64_t c1, int32_t* flag) {
    const int t = blockIdx.y;
    const int64_t c = c0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    const int64_t* cp = S.colptr + (int64_t)t * (Nx + 1);
    const int64_t e0 = S.tap_off[t] + cp[c] - 1, e1 = S.tap_off[t] + cp[c + 1] - 1;
    int64_t prev = 0;
    bool bad = false;
#pragma unroll 1
    for (int64_t e = e0; e < e1; e++) {
        const int64_t r = S.rowval[e] - S.row_off;
        bad |= (r <= prev) | (r > nrow);
        prev = r;
    }
    if (bad) *flag = 1;
}

int sls_launch_validate_supports(const SlsSupports& S, int T, int64_t Nx, int64_t nrow, int64_t c0, int64_t c1, int32_t* flag,
                                 sls_stream_t stream) {
    if (c1 <= c0 || T <= 0) return 0;
    dim3 grid((unsigned)((c1 - c0 + 255) / 256), T, 1);
    auto kp = &sls_validate_supports_kernel;
    SLS_LAUNCH(kp, grid, 256, 0, stream, S, Nx, nrow, c0, c1, flag);
    return rt::last_error();
}

// Output-feedback branch of sparsity_dim_reduction (reference src/reduction.jl:16-18): for one support family
//     s = unique(findnz(𝓢ⱼ[end][:,cⱼ])[1])
// i.e. the stored rows of the listed columns of the LAST tap, first occurrences in CSC order (cⱼ order, rows ascending
// inside a column) — no product with the pattern of A on this branch.  `rows` holds the q column segments back to back
// (1-based, seg[s]..seg[s+1]).  One block: the segments are taken in order; rows are distinct inside a segment, so one
// test-and-set per row on a bitmap decides "seen before", and a block-wide rank keeps the order of first occurrence.
// Malformed input (row outside 1..nrow, not strictly ascending inside a column) sets *flag.
#define UNION_THREADS 256
__global__ void __launch_bounds__(UNION_THREADS)
sls_union_first_kernel(const int64_t* __restrict__ rows, const int64_t* __restrict__ seg, int q, int64_t nrow,
                       uint32_t* seen, int64_t* out, int64_t* n_out, int32_t* flag) {
    __shared__ int wsum[UNION_THREADS / 32];
    __shared__ int64_t base_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) base_s = 0;
    __syncthreads();
#pragma unroll 1
    for (int s = 0; s < q; s++) {
        const int64_t e0 = seg[s], e1 = seg[s + 1];
#pragma unroll 1
        for (int64_t i0 = e0; i0 < e1; i0 += UNION_THREADS) {
            const int64_t i = i0 + tid;
            bool keep = false;
            int64_t r = 0;
            if (i < e1) {
                r = rows[i];
                const int64_t prev = (i > e0) ? rows[i - 1] : 0;
                if (r < 1 || r > nrow || r <= prev) *flag = 1;
                else {
                    const uint32_t bit = 1u << ((r - 1) & 31);
                    keep = !(atomicOr(&seen[(r - 1) >> 5], bit) & bit);
                }
            }
            const uint32_t b = __ballot_sync(SLS_FULL, keep);
            if (lane == 0) wsum[warp] = __popc(b);
            __syncthreads();
            int before = 0, total = 0;
#pragma unroll
            for (int w = 0; w < UNION_THREADS / 32; w++) { if (w < warp) before += wsum[w]; total += wsum[w]; }
            if (keep) out[base_s + before + __popc(b & ((1u << lane) - 1u))] = r;
            __syncthreads();
            if (tid == 0) base_s += total;
            __syncthreads();
        }
    }
    if (tid == 0) *n_out = base_s;
}

int sls_launch_union_first(const int64_t* rows, const int64_t* seg, int q, int64_t nrow, uint32_t* seen, int64_t* out,
                           int64_t* n_out, int32_t* flag, sls_stream_t stream) {
    auto kp = &sls_union_first_kernel;
    SLS_LAUNCH(kp, 1, UNION_THREADS, 0, stream, rows, seg, q, nrow, seen, out, n_out, flag);
    return rt::last_error();
}

// ---- host launchers -------------------------------------------------------------------------
int sls_launch_gather_count(const SlsCsc& A, const SlsSupports& Sx, const SlsSupports& Su, int T, int64_t Nx,
                            const int32_t* cols, int ncols, int32_t* n_out, int32_t* m_out, int cap,
                            int32_t* keep_x, int32_t* keep_u, const int32_t* bad_supports, int32_t* stats, sls_stream_t stream) {
    if (ncols == 0) return 0;
    int wpb = COUNT_MAX_WPB;                                  // warps per block: as many as 32 KB of candidate buffers allow
    while (wpb > 1 && (size_t)wpb * cap * sizeof(int32_t) > 32 * 1024) wpb >>= 1;
    const int grid = (ncols + wpb - 1) / wpb;
    const size_t smem = (size_t)wpb * cap * sizeof(int32_t);
    auto kp = &sls_gather_count_kernel;
    if (smem > 48 * 1024) { int e = rt::set_max_smem(kp, (int)smem); if (e) return e; }
    SLS_LAUNCH(kp, grid, wpb * 32, smem, stream, A, Sx, Su, T, Nx, cols, ncols, n_out, m_out, cap, keep_x, keep_u, bad_supports, stats);
    return rt::last_error();
}

int sls_launch_gather_fill(const SlsCsc& A, const SlsCsc& B2, const SlsSupports& Sx, const SlsSupports& Su,
                           const SlsPlanView& P, int cap, const int32_t* keep_x, const int32_t* keep_u, sls_stream_t stream) {
    if (P.ncols == 0) return 0;
    const int grid = (P.ncols + GATHER_WPB - 1) / GATHER_WPB;
    const size_t smem = (size_t)GATHER_WPB * sls_gather_fill_words(cap, P.nstride, P.mstride, P.wx, P.wu, P.T) * sizeof(int32_t);
    auto kp = &sls_gather_fill_kernel;
    if (smem > 48 * 1024) { int e = rt::set_max_smem(kp, (int)smem); if (e) return e; }
    SLS_LAUNCH(kp, grid, GATHER_WPB * 32, smem, stream, A, B2, Sx, Su, P, cap, keep_x, keep_u);
    return rt::last_error();
}

size_t sls_scan_scratch_bytes(int narr, int64_t span) {
    const int64_t tile = (int64_t)SCAN_THREADS * SCAN_ITEMS;
    return (size_t)(narr + 1) * (size_t)((span + tile - 1) / tile + 1) * sizeof(int64_t);
}

// `scratch` (sls_scan_scratch_bytes(narr, max(hi - lo, totlen)) bytes) may be null: one block per array then.
int sls_launch_scan(int64_t* colptr, int64_t len, int narr, int64_t* tot, int64_t totlen, int64_t lo, int64_t hi, sls_stream_t stream,
                    int64_t* scratch) {
    auto kp = &sls_scan_kernel;
    if (lo < 0) lo = 0;
    if (hi > len || hi < lo) hi = len;
    const int64_t tile = (int64_t)SCAN_THREADS * SCAN_ITEMS;
    const int64_t span = std::max<int64_t>(narr > 0 ? hi - lo : 0, totlen);
    int64_t nchunk = scratch ? (span + tile - 1) / tile : 1;
    if (nchunk < 1) nchunk = 1;
    if (nchunk > 65535) nchunk = 65535;
    const int64_t chunk = nchunk == 1 ? span : ((span + nchunk - 1) / nchunk + tile - 1) / tile * tile;
    dim3 grid((unsigned)(narr + 1), (unsigned)nchunk, 1);
    if (nchunk > 1) {
        auto ks = &sls_scan_sum_kernel;
        SLS_LAUNCH(ks, grid, SCAN_THREADS, 0, stream, (const int64_t*)colptr, len, narr, (const int64_t*)tot, totlen, lo, hi, chunk, scratch);
    }
    SLS_LAUNCH(kp, grid, SCAN_THREADS, 0, stream, colptr, len, narr, tot, totlen, lo, hi, chunk, (const int64_t*)(nchunk > 1 ? scratch : nullptr));
    if (narr > 0 && (lo > 0 || hi + 1 < len)) {
        auto kf = &sls_colptr_fill_kernel;
        dim3 gridf((unsigned)((len + 255) / 256), narr, 1);
        SLS_LAUNCH(kf, gridf, 256, 0, stream, colptr, len, lo, hi);
    }
    return rt::last_error();
}

int sls_launch_local_dense(const SlsCsc& A, const SlsCsc& B2, const SlsPlanView& P, int c,
                           double* At, double* B2t, uint8_t* Mx, uint8_t* Mu, sls_stream_t stream) {
    auto kp = &sls_local_dense_kernel;
    SLS_LAUNCH(kp, 1, 128, 0, stream, A, B2, P, c, At, B2t, Mx, Mu);
    return rt::last_error();
}
