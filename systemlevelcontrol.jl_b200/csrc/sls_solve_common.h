// sls_solve_common.h — pieces shared by the two kernel-2 variants: sls_solve.cu (one warp per column, n <= 32) and
// sls_solve_wide.cu (one CTA per column, 32 < n <= 64).
#pragma once
#include "sls_common.h"
#include "sls_kernels.h"

#include <cmath>
#include <cstdlib>
#define SLS_INF ((double)INFINITY)

__device__ __forceinline__ void sls_dmma(double& c0, double& c1, double a, double b) {
#ifdef SLS_EMU
    emu::dmma(c0, c1, a, b);
#else
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
#endif
}
// 1/x for x > 0 to 1 ulp: hardware seed (rcp.approx.ftz.f64) + one cubic correction
__device__ __forceinline__ double sls_rcp(double x) {
#ifdef SLS_EMU
    return 1.0 / x;
#else
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    // the seed is good to 2^-19.9 (measured): with e = 1 - x*y,  1/x = y (1 + e + e^2 + O(e^3)),  e^3 < 2^-59.
    // Three dependent FMAs instead of two Newton steps (four); 1 ulp worst case (tools/rcp_test.cu).
    const double e = fma(-x, y, 1.0);
    return fma(y, fma(e, e, e), y);
#endif
}
// max that keeps a NaN from either side (fmax would drop it: a NaN residual must end as SLS_STATUS_RESIDUAL)
__device__ __forceinline__ double sls_nanmax(double a, double b) { return (b > a || b != b) ? b : a; }
// acc += X * W^T for two 8x8 blocks in accumulator layout
__device__ __forceinline__ void sls_mma_xwt(double (&acc)[2], const double (&x)[2], const double (&w)[2]) {
    sls_dmma(acc[0], acc[1], x[0], w[0]);
    sls_dmma(acc[0], acc[1], x[1], w[1]);
}

struct SlsSolveLayout {
    int N, NB, mcap, nnzcap, T, wu;
    int oAblk, oBBt, oY0, oY1, oG, oDref, oLam, oTx, oRhs, oXa, oXb, oUc, oValR, oValC;
    int oSx, oSu, oMx, oMu, oPre, oRptr, oRsplit, oColR, oCptr, oRowC, total;
    __host__ __device__ SlsSolveLayout(int N_, int mcap_, int nnzcap_, int T_, int wu_) {
        N = N_; NB = N_ / 8; mcap = mcap_ < 1 ? 1 : mcap_; nnzcap = nnzcap_ < 1 ? 1 : nnzcap_; T = T_; wu = wu_ < 1 ? 1 : wu_;
        int o = 0;
#define SLS_SEC(name, bytes) name = o; o += (bytes); o = (o + 15) & ~15
        // sections whose size depends on N only come first: their offsets are compile-time constants in the kernel
        SLS_SEC(oAblk, 8 * N * N);
        SLS_SEC(oBBt, 8 * N * N);
        SLS_SEC(oY0, 8 * N);
        SLS_SEC(oY1, 8 * N);
        SLS_SEC(oG, 8 * N);
        SLS_SEC(oDref, 8 * N);
        SLS_SEC(oLam, 8 * N);
        SLS_SEC(oRhs, 8 * N);
        SLS_SEC(oXa, 8 * N);
        SLS_SEC(oXb, 8 * N);
        SLS_SEC(oSx, 4 * N);
        SLS_SEC(oRptr, 2 * (N + 2));
        SLS_SEC(oRsplit, 2 * (N + 2));
        SLS_SEC(oTx, 8 * (N + mcap));
        SLS_SEC(oUc, 8 * (N + mcap));                       // [x_k | u_k]: one vector for the residual's gathers
        SLS_SEC(oValR, 8 * nnzcap);
        SLS_SEC(oValC, 8 * nnzcap);
        SLS_SEC(oSu, 4 * mcap);
        SLS_SEC(oMx, 8 * (T + 1));
        SLS_SEC(oMu, 4 * T * wu);
        SLS_SEC(oPre, 8 * 2 * T);
        SLS_SEC(oColR, 2 * nnzcap);
        SLS_SEC(oCptr, 2 * (N + mcap + 2));
        SLS_SEC(oRowC, 2 * nnzcap);
#undef SLS_SEC
        total = (o + 15) & ~15;
    }
};

__device__ __forceinline__ bool sls_bit(const uint32_t* w, int i) { return (w[i >> 5] >> (i & 31)) & 1u; }
// index of element (r,c) inside a blocked (accumulator-layout friendly) N x N shared-memory matrix
__device__ __forceinline__ int sls_blk(int NB, int r, int c) { return (((r >> 3) * NB + (c >> 3)) << 6) + ((r & 7) << 3) + (c & 7); }

// free-variable mask of one tap over the local state index: 32 bits for n <= 32, 64 bits for n <= 64
template <int NB> struct SlsMask { typedef uint32_t type; };
template <> struct SlsMask<5> { typedef uint64_t type; };
template <> struct SlsMask<6> { typedef uint64_t type; };
template <> struct SlsMask<7> { typedef uint64_t type; };
template <> struct SlsMask<8> { typedef uint64_t type; };
__device__ __forceinline__ bool sls_mbit(uint32_t m, int r) { return (m >> r) & 1u; }
__device__ __forceinline__ bool sls_mbit(uint64_t m, int r) { return (m >> r) & 1ull; }
__device__ __forceinline__ int sls_mrank(uint32_t m, int r) { return __popc(m & ((1u << r) - 1u)); }
__device__ __forceinline__ int sls_mrank(uint64_t m, int r) { return __popcll(m & ((1ull << r) - 1ull)); }

