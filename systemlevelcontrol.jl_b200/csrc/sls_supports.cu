// sls_supports.cu — (d,T)-localized support generation on the device.
//
// The reference leaves the construction of the supports to user code (README.md:52-54):
//     𝓢ₓ[t] = (A .≠ 0)^min(d,   floor(α(t-1))) .≠ 0
//     𝓢ᵤ[t] = (B₂' .≠ 0) * (A .≠ 0)^min(d+1, floor(α(t-1))) .≠ 0          t = 1..T
// i.e. Boolean sparse matrix powers, built on the host and (for the drop-in path) shipped over PCIe: 0.6 GB at
// Nx = 100 000.  Here column j of (A .≠ 0)^s is obtained as the s-step reachability set R_s(j) of the pattern graph
// (R_0 = {j}, R_{s+1} = union of the stored-nonzero rows of A[:,c], c in R_s) — exactly the Boolean power, no
// assumption on the diagonal — and the rows of 𝓢ᵤ as the actuators with a stored nonzero in a row of R_s.
// One warp per column; sets are kept sorted/unique in shared memory (append + bitonic sort + unique).
// Taps that share an exponent share one pattern in HBM (7 + 8 distinct patterns instead of 2 x 29 at config #3).
#include "sls_common.h"
#include "sls_kernels.h"
#include <climits>

#define SUPP_WPB 4

// dst = sorted unique union over c in src[0..ns) of the stored-nonzero rows of M[:,c]; returns size or -1 (overflow).
// One lane per source column: the colptr loads of 32 columns are in flight together, then every lane copies its own
// rows to the offset a warp scan gives it.
__device__ int sls_supp_expand(const SlsCsc& M, const int32_t* src, int ns, int32_t* dst, int cap) {
    const int lane = sls_lane();
    int cnt = 0;
#pragma unroll 1
    for (int k0 = 0; k0 < ns; k0 += 32) {
        const int k = k0 + lane;
        int64_t e0 = 0;
        int len = 0;
        if (k < ns) { const int64_t c = src[k]; e0 = M.colptr[c] - 1; len = (int)(M.colptr[c + 1] - 1 - e0); }
        int incl = len;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
        const int total = __shfl_sync(SLS_FULL, incl, 31);
        if (cnt + total > cap) return -1;
        const int base = cnt + incl - len;
#pragma unroll 1
        for (int i = 0; i < len; i++)
            dst[base + i] = (M.nzval[e0 + i] != 0.0) ? (int32_t)(M.rowval[e0 + i] - 1) : INT_MAX;   // `.≠ 0` drops stored zeros
        cnt += total;
    }
    __syncwarp();
    return sls_warp_sort_unique(dst, cnt, cap);
}

template <bool FILL>
__global__ void __launch_bounds__(SUPP_WPB * 32)
sls_supp_kernel(SlsSuppArgs a) {
    SLS_DYN_SMEM(smem);
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int64_t j = (int64_t)blockIdx.x * SUPP_WPB + warp;
    if (j >= a.Nx) return;
    int32_t* cur = reinterpret_cast<int32_t*>(smem) + (size_t)warp * 3 * a.cap;
    int32_t* nxt = cur + a.cap;
    int32_t* ub = nxt + a.cap;
    if (lane == 0) cur[0] = (int32_t)j;
    __syncwarp();
    int n = 1;
#pragma unroll 1
    for (int s = 0; s <= a.nlev; s++) {
        const int xs = a.ex_slot[s], us = a.eu_slot[s];
        if (xs >= 0) {
            int64_t* cp = a.xcolptr + (int64_t)xs * (a.Nx + 1);
            if (!FILL) { if (lane == 0) { cp[j] = n; if (xs == a.xlast_slot) atomicMax(a.maxlen, n); } }
            else {
                int64_t* dst = a.xrow + a.xoff[xs] + cp[j] - 1;
#pragma unroll 1
                for (int i = lane; i < n; i += 32) dst[i] = (int64_t)cur[i] + 1;
            }
        }
        if (us >= 0) {
            const int m = sls_supp_expand(a.B2r, cur, n, ub, a.cap);
            if (m < 0) { if (lane == 0) atomicOr(a.overflow, 1); return; }
            int64_t* cp = a.ucolptr + (int64_t)us * (a.Nx + 1);
            if (!FILL) { if (lane == 0) { cp[j] = m; if (us == a.ulast_slot) atomicMax(a.maxlen + 1, m); } }
            else {
                int64_t* dst = a.urow + a.uoff[us] + cp[j] - 1;
#pragma unroll 1
                for (int i = lane; i < m; i += 32) dst[i] = (int64_t)ub[i] + 1;
            }
            __syncwarp();
        }
        if (s < a.nlev) {
            const int nn = sls_supp_expand(a.A, cur, n, nxt, a.cap);
            if (nn < 0) { if (lane == 0) atomicOr(a.overflow, 1); return; }
            int32_t* t = cur; cur = nxt; nxt = t;
            n = nn;
        }
    }
}

int sls_launch_supports(const SlsSuppArgs& a, bool fill, sls_stream_t stream) {
    if (a.Nx == 0) return 0;
    const int grid = (int)((a.Nx + SUPP_WPB - 1) / SUPP_WPB);
    const size_t smem = (size_t)SUPP_WPB * 3 * a.cap * sizeof(int32_t);
    if (fill) {
        auto kp = &sls_supp_kernel<true>;
        if (smem > 48 * 1024) { int e = rt::set_max_smem(kp, (int)smem); if (e) return e; }
        SLS_LAUNCH(kp, grid, SUPP_WPB * 32, smem, stream, a);
    } else {
        auto kp = &sls_supp_kernel<false>;
        if (smem > 48 * 1024) { int e = rt::set_max_smem(kp, (int)smem); if (e) return e; }
        SLS_LAUNCH(kp, grid, SUPP_WPB * 32, smem, stream, a);
    }
    return rt::last_error();
}
