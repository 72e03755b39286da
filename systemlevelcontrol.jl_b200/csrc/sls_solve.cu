// sls_solve.cu — kernel 2 for local state dimensions n <= 32: batched FP64 equality-constrained least squares, one warp
// per column, dense O(n^3) work on the FP64 tensor cores (DMMA m8n8k4), everything register-resident.
// (32 < n <= 64: one CTA per column, sls_solve_wide.cu.)
//
// Replaces the JuMP/Ipopt solve of reference src/synthesis.jl:46-62 for plants built by
// Plant(A,B1=I,B2) ([C1 D12] = I, D11 = 0; GeneralizedPlant.jl:105-110).  Per column the reference's QP is
//     min  sum_t ||x_t||^2 + ||u_t||^2      over the in-mask entries (others fixed to 0, synthesis.jl:57-60)
//     s.t. x_1 = e_j,  x_{t+1} = A~ x_t + B~ u_t (t<T),  0 = A~ x_T + B~ u_T     (synthesis.jl:53-55)
// i.e. the minimum-norm solution v = G^T lambda, (G G^T) lambda = h, of a rank-deficient but consistent
// system.  G G^T is block tridiagonal in time with n x n blocks
//     D_0 = diag(mx_1),   D_k = diag(mx_{k+1}) + A~ diag(mx_k) A~^T + B~ diag(mu_k) B~^T,   E_k = -A~ diag(mx_k)
// Forward block elimination (k = 0..T):   S_k = D_k - E_k Z_{k-1} E_k^T,   Z_k = S_k^+ :
//     S_k = diag(mx_{k+1}) + B~ diag(mu_k) B~^T + A~ Pi_k A~^T,      Pi_k = mx_k o (I - Z_{k-1}) o mx_k
// S_k^+ is the inverse on the rows whose pivot exceeds rank_tol * diag(D_k); redundant constraint rows are
// eliminated (what Ipopt's inertia correction does implicitly).
//
// Data layout: an n x n matrix (n <= N = 8*NB <= 32) lives in the registers of one warp as NB x NB blocks
// of 8 x 8 in the DMMA accumulator layout: lane 4g+q holds X[8bi+g][8bj+2q], X[8bi+g][8bj+2q+1].
// With the k index of the MMA mapped to even / odd columns, a product  X * W^T  of two blocks that are both
// in accumulator layout needs NO data movement: A-fragment = (x0 | x1), B-fragment = (w0 | w1).  All
// products of the recursion have that form because Pi, Z and the pivot-block inverses are symmetric:
//     T = A~ Pi :  T_ij += A~_il * (Pi_jl)^T         Y = T A~^T :  Y_ij += T_il * (A~_jl)^T
//     block symmetric sweep (Gauss-Jordan) with pivot block p, P = S_pp^+ :
//         S_ij -= (S_ip P) * (S_jp)^T,   S_ip <- S_ip P = S_ip * P^T,   S_pi <- P * (S_ip)^T,   S_pp <- -P
// The 8 x 8 pivot blocks are inverted by 8 scalar symmetric sweeps with rank skipping through warp shuffles.
// -Z_k goes to a per-warp history slot (lower-triangle blocks in register order, L2/HBM) for the back substitution
//     lambda_{k-1} = Z_{k-1}(yh_{k-1} + mx_k o A~^T lambda_k),
//     x_k = mx_k o (lambda_{k-1} - A~^T lambda_k),   u_k = -mu_k o (B~^T lambda_k).
// Sparse pieces (A~ x, B~ diag(mu) B~^T, diagonals) use a local CSR/CSC of C~ = [A~ B~] gathered from the
// plant into shared memory.  The achievability residual of the computed (x,u) is checked per column.
#include "sls_common.h"
#include "sls_kernels.h"


#include "sls_solve_common.h"

#ifndef SLS_K2_PIVOT_SMEM
#define SLS_K2_PIVOT_SMEM 1
#endif

// warps per block / minimum resident blocks per SM of each size class (registers: 128 per thread up to N = 24)
template <int NB> struct SlsCfg { static const int WPB = 4; static const int MINB = (NB <= 3) ? 4 : 2; };

template <int NB>
__global__ void __launch_bounds__(SlsCfg<NB>::WPB * 32, SlsCfg<NB>::MINB)
sls_solve_kernel(SlsSolveArgs a) {
    constexpr int N = 8 * NB;
    constexpr int WPB = SlsCfg<NB>::WPB;
    constexpr int RPL = (N + 31) / 32;       // rows per lane in the sparse (lane-per-row) phases
    typedef typename SlsMask<NB>::type mask_t;
    SLS_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int g = lane >> 2, q = lane & 3;   // accumulator-layout coordinates
    const SlsPlanView& P = a.P;
    const int T = P.T;
    const SlsSolveLayout L(N, a.mcap, a.nnzcap, T, P.wu);
    unsigned char* base = smem_raw + (size_t)warp * L.total;
    double* Am = reinterpret_cast<double*>(base + L.oAblk);     // A~ diag(mx_k) as dense 8x8 blocks (accumulator order)
    double* Dp = reinterpret_cast<double*>(base + L.oBBt);      // diag(mx_{k+1}) + B~ diag(mu_k) B~^T, same layout
    double* ybuf0 = reinterpret_cast<double*>(base + L.oY0);
    double* ybuf1 = reinterpret_cast<double*>(base + L.oY1);
    double* gbuf = reinterpret_cast<double*>(base + L.oG);
    double* thrs = reinterpret_cast<double*>(base + L.oDref);   // pivot thresholds rank_tol*diag(D_k) (+inf: empty row)
    double* lam_s = reinterpret_cast<double*>(base + L.oLam);
    double* tx = reinterpret_cast<double*>(base + L.oTx);
    double* rhs_s = reinterpret_cast<double*>(base + L.oRhs);
    double* xa = reinterpret_cast<double*>(base + L.oXa);
    double* xb = reinterpret_cast<double*>(base + L.oXb);
    double* ucur = reinterpret_cast<double*>(base + L.oUc);
    double* valr = reinterpret_cast<double*>(base + L.oValR);
    double* valc = reinterpret_cast<double*>(base + L.oValC);
    int32_t* sx = reinterpret_cast<int32_t*>(base + L.oSx);
    int32_t* su = reinterpret_cast<int32_t*>(base + L.oSu);
    mask_t* mxs = reinterpret_cast<mask_t*>(base + L.oMx);
    uint32_t* mus = reinterpret_cast<uint32_t*>(base + L.oMu);
    int64_t* obase = reinterpret_cast<int64_t*>(base + L.oPre);   // [2T] first stored entry of column j in every output tap
    uint16_t* rptr = reinterpret_cast<uint16_t*>(base + L.oRptr);
    uint16_t* rsplit = reinterpret_cast<uint16_t*>(base + L.oRsplit);
    uint16_t* colr = reinterpret_cast<uint16_t*>(base + L.oColR);
    uint16_t* cptr = reinterpret_cast<uint16_t*>(base + L.oCptr);
    uint16_t* rowc = reinterpret_cast<uint16_t*>(base + L.oRowC);
    const int wu = P.wu;

    const int slot = blockIdx.x * WPB + warp;
    const int nslots = gridDim.x * WPB;
    double* hist = a.hist + (size_t)slot * a.hist_stride;      // [T+1][N] yh history, then [T][tri(N)] packed -Z_k
    double* yhist = hist;
    double* zhist = hist + (size_t)(T + 1) * N;
    constexpr int HB = NB * (NB + 1) / 2 * 64;             // doubles of one stored -Z_k: lower-triangle blocks in register order

    for (int idx = slot; idx < a.count; idx += nslots) {
        const int c = a.list[idx];
        const int n = P.n[c], m = P.m[c], jl = P.jloc[c];
        __syncwarp();
        // ------------------------------------------------------------------ metadata -> shared
#pragma unroll 1
        for (int i = lane; i < n; i += 32) sx[i] = P.sx[(size_t)c * P.nstride + i];
#pragma unroll 1
        for (int l = lane; l < m; l += 32) su[l] = P.su[(size_t)c * P.mstride + l];
#pragma unroll 1
        for (int t = lane; t < T; t += 32) {
            const uint32_t* w = P.maskx + ((size_t)c * T + t) * P.wx;
            mask_t v = (mask_t)w[0];
            if (sizeof(mask_t) == 8 && P.wx > 1) v |= (mask_t)((uint64_t)w[1] << 32);
            mxs[t] = v;
        }
        if (lane == 0) mxs[T] = (mask_t)0;
#pragma unroll 1
        for (int t = lane; t < T * wu; t += 32) mus[t] = P.masku[(size_t)c * T * wu + t];
#pragma unroll 1
        for (int t = lane; t < 2 * T; t += 32) obase[t] = P.colptr[(int64_t)t * (P.Nx + 1) + P.cols[c]] - 1;
#pragma unroll 1
        for (int e = 2 * lane; e < N * N; e += 64) *reinterpret_cast<double2*>(Am + e) = make_double2(0.0, 0.0);
        __syncwarp();
        // ------------------------------------------------------------------ local CSR of C~ = [A~ B~]
        uint64_t nzA = 0ull;                                  // bit bi*NB+bj: block (bi,bj) of A~ has a stored entry
        {
            int carry = 0;
#pragma unroll 1
            for (int rr = 0; rr < RPL; rr++) {
                const int i = lane + 32 * rr;
                const bool act = i < n;
                int cntA = 0, cntB = 0;
                int64_t ea0 = 0, ea1 = 0, eb0 = 0, eb1 = 0;
                // the local index of a row's first 8 entries found by the counting pass is kept (one byte each,
                // 0xff = outside the index set) so that the filling pass does not repeat the binary searches
                uint64_t fa = ~0ull, fb = ~0ull;
                if (act) {
                    const int64_t gr = sx[i];
                    ea0 = a.Ar.colptr[gr] - 1; ea1 = a.Ar.colptr[gr + 1] - 1;
                    eb0 = a.B2r.colptr[gr] - 1; eb1 = a.B2r.colptr[gr + 1] - 1;
#pragma unroll 1
                    for (int64_t e = ea0; e < ea1; e++) {
                        const int l = sls_find_cold(sx, n, (int32_t)(a.Ar.rowval[e] - 1));
                        cntA += (l >= 0);
                        if (e - ea0 < 8) fa = (fa & ~(0xffull << (8 * (e - ea0)))) | ((uint64_t)(l & 0xff) << (8 * (e - ea0)));
                    }
#pragma unroll 1
                    for (int64_t e = eb0; e < eb1; e++) {
                        const int l = sls_find_cold(su, m, (int32_t)(a.B2r.rowval[e] - 1));
                        cntB += (l >= 0);
                        if (e - eb0 < 8) fb = (fb & ~(0xffull << (8 * (e - eb0)))) | ((uint64_t)(l & 0xff) << (8 * (e - eb0)));
                    }
                }
                int incl = cntA + cntB;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
                const int excl = carry + incl - (cntA + cntB);
                if (act) {
                    rptr[i] = (uint16_t)excl; rsplit[i] = (uint16_t)(excl + cntA);
                    int o = excl;
#pragma unroll 1
                    for (int64_t e = ea0; e < ea1; e++) {
                        int l;
                        if (e - ea0 < 8) { l = (int)((fa >> (8 * (e - ea0))) & 0xff); if (l == 0xff) l = -1; }
                        else l = sls_find_cold(sx, n, (int32_t)(a.Ar.rowval[e] - 1));
                        if (l >= 0) {
                            colr[o] = (uint16_t)l; valr[o] = a.Ar.nzval[e]; o++;
                            nzA |= 1ull << ((i >> 3) * NB + (l >> 3));
                        }
                    }
#pragma unroll 1
                    for (int64_t e = eb0; e < eb1; e++) {
                        int l;
                        if (e - eb0 < 8) { l = (int)((fb >> (8 * (e - eb0))) & 0xff); if (l == 0xff) l = -1; }
                        else l = sls_find_cold(su, m, (int32_t)(a.B2r.rowval[e] - 1));
                        if (l >= 0) { colr[o] = (uint16_t)(n + l); valr[o] = a.B2r.nzval[e]; o++; }
                    }
                }
                carry += __shfl_sync(SLS_FULL, incl, 31);
            }
            if (lane == 0) rptr[n] = (uint16_t)carry;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) nzA |= __shfl_xor_sync(SLS_FULL, nzA, d);
        // ------------------------------------------------------------------ local CSC of C~ (columns 0..n+m)
        {
            int carry = 0;
#pragma unroll 1
            for (int p0 = 0; p0 < n + m; p0 += 32) {
                const int p = p0 + lane;
                int cnt = 0;
                int64_t e0 = 0, e1 = 0;
                uint64_t fc = ~0ull;
                const SlsCsc& M = (p < n) ? a.A : a.B2;
                if (p < n + m) {
                    const int64_t gc = (p < n) ? sx[p] : su[p - n];
                    e0 = M.colptr[gc] - 1; e1 = M.colptr[gc + 1] - 1;
#pragma unroll 1
                    for (int64_t e = e0; e < e1; e++) {
                        const int r = sls_find_cold(sx, n, (int32_t)(M.rowval[e] - 1));
                        cnt += (r >= 0);
                        if (e - e0 < 8) fc = (fc & ~(0xffull << (8 * (e - e0)))) | ((uint64_t)(r & 0xff) << (8 * (e - e0)));
                    }
                }
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
                if (p < n + m) {
                    int o = carry + incl - cnt;
                    cptr[p] = (uint16_t)o;
#pragma unroll 1
                    for (int64_t e = e0; e < e1; e++) {
                        int r;
                        if (e - e0 < 8) { r = (int)((fc >> (8 * (e - e0))) & 0xff); if (r == 0xff) r = -1; }
                        else r = sls_find_cold(sx, n, (int32_t)(M.rowval[e] - 1));
                        if (r >= 0) { rowc[o] = (uint16_t)r; valc[o] = M.nzval[e]; o++; }
                    }
                }
                carry += __shfl_sync(SLS_FULL, incl, 31);
            }
            if (lane == 0) cptr[n + m] = (uint16_t)carry;
        }
        __syncwarp();

        // ------------------------------------------------------------------ tap 0:  S_0 = diag(mx_1), yh_0 = e
        double S[NB][NB][2];                                  // -Z_k ("swept" S_k) in accumulator layout
        {
            const mask_t m1 = mxs[0];
#pragma unroll
            for (int bi = 0; bi < NB; bi++) {
                const int r = 8 * bi + g;
#pragma unroll
                for (int bj = 0; bj < NB; bj++) {
                    const int c0 = 8 * bj + 2 * q;
                    S[bi][bj][0] = (r == c0 && sls_mbit(m1, r)) ? -1.0 : 0.0;
                    S[bi][bj][1] = (r == c0 + 1 && sls_mbit(m1, r)) ? -1.0 : 0.0;
                }
            }
        }
        const double dg0 = (g == 2 * q) ? 1.0 : 0.0, dg1 = (g == 2 * q + 1) ? 1.0 : 0.0;   // identity inside a diagonal block
        double* yprev = ybuf0;
        double* ycur = ybuf1;
#pragma unroll
        for (int rr = 0; rr < RPL; rr++) {
            const int i = lane + 32 * rr;
            if (i < N) { const double y0 = (i == jl) ? 1.0 : 0.0; yprev[i] = y0; yhist[i] = y0; }
        }
        int status = (jl < 0) ? SLS_ST_NO_IMPULSE : 0;
        uint32_t mu_prev[4] = {0u, 0u, 0u, 0u};               // wu <= 4 is guaranteed by the launcher (m <= 128)
        mask_t mk_prev = 0, mk1_prev = 0;
        bool cache_valid = false;
        uint64_t nonempty = 0ull;
        __syncwarp();

        // ------------------------------------------------------------------ forward elimination, taps 1..T
        for (int k = 1; k <= T; k++) {
            const mask_t mk = mxs[k - 1];
            const mask_t mk1 = mxs[k];                        // mxs[T] = 0
            const uint32_t* muk = mus + (size_t)(k - 1) * wu;
            // (1) g = mx_k o (Z_{k-1} yh_{k-1}) = -(mx_k o (S yprev))
#pragma unroll
            for (int bi = 0; bi < NB; bi++) {
                double acc = 0.0;
#pragma unroll
                for (int bj = 0; bj < NB; bj++) {
                    const double2 yv = *reinterpret_cast<const double2*>(yprev + 8 * bj + 2 * q);
                    acc = fma(S[bi][bj][0], yv.x, acc);
                    acc = fma(S[bi][bj][1], yv.y, acc);
                }
                acc += __shfl_xor_sync(SLS_FULL, acc, 1);
                acc += __shfl_xor_sync(SLS_FULL, acc, 2);
                const int r = 8 * bi + g;
                if (q == 0) gbuf[r] = sls_mbit(mk, r) ? -acc : 0.0;
            }
            // (2) mask-dependent operands, rebuilt only when a mask changed (constant once the band is full):
            //     Am = A~ diag(mx_k),  Dp = diag(mx_{k+1}) + B~ diag(mu_k) B~^T,  thr = rank_tol * diag(D_k)
            bool same = cache_valid && mk == mk_prev && mk1 == mk1_prev;
#pragma unroll
            for (int w = 0; w < 4; w++) if (w < wu) same = same && (muk[w] == mu_prev[w]);
            if (!same) {
#pragma unroll 1
                for (int e = 2 * lane; e < N * N; e += 64) *reinterpret_cast<double2*>(Dp + e) = make_double2(0.0, 0.0);
                __syncwarp();
                nonempty = 0ull;
#pragma unroll
                for (int rr = 0; rr < RPL; rr++) {
                    const int i = lane + 32 * rr;
                    double dref = 0.0;
                    if (i < n) {
                        const int r0 = rptr[i], r1 = rsplit[i], r2 = rptr[i + 1];
                        if (sls_mbit(mk1, i)) { dref = 1.0; Dp[sls_blk(NB, i, i)] = 1.0; }
#pragma unroll 1
                        for (int e = r0; e < r1; e++) {
                            const int l = colr[e]; const double v = valr[e];
                            const bool on = sls_mbit(mk, l);
                            Am[sls_blk(NB, i, l)] = on ? v : 0.0;
                            if (on) dref = fma(v, v, dref);
                        }
#pragma unroll 1
                        for (int e = r1; e < r2; e++) {
                            const int l = colr[e] - n;
                            if (!sls_bit(muk, l)) continue;
                            const double v = valr[e];
                            dref = fma(v, v, dref);
#pragma unroll 1
                            for (int e2 = cptr[n + l]; e2 < cptr[n + l + 1]; e2++) Dp[sls_blk(NB, i, rowc[e2])] += v * valc[e2];
                        }
                    }
                    if (i < N) thrs[i] = (dref != 0.0) ? a.rank_tol * dref : SLS_INF;
                    nonempty |= (uint64_t)__ballot_sync(SLS_FULL, dref != 0.0) << (32 * rr);
                }
#pragma unroll
                for (int w = 0; w < 4; w++) if (w < wu) mu_prev[w] = muk[w];
                mk_prev = mk; mk1_prev = mk1; cache_valid = true;
            }
            __syncwarp();
            // (3) yh_k = A~ g
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) {
                const int i = lane + 32 * rr;
                double acc = 0.0;
                if (i < n) {
                    const int r0 = rptr[i], r1 = rsplit[i];
#pragma unroll 1
                    for (int e = r0; e < r1; e++) acc = fma(valr[e], gbuf[colr[e]], acc);
                }
                if (i < N) { ycur[i] = acc; yhist[(size_t)k * N + i] = acc; }
            }
            // (4) Pi' = I - Z_{k-1}  (the masks live in Am)
#pragma unroll
            for (int bi = 0; bi < NB; bi++) { S[bi][bi][0] += dg0; S[bi][bi][1] += dg1; }
            // (5) Tm = Am Pi'
            double Tm[NB][NB][2];
#pragma unroll
            for (int bi = 0; bi < NB; bi++) {
#pragma unroll
                for (int bj = 0; bj < NB; bj++) { Tm[bi][bj][0] = 0.0; Tm[bi][bj][1] = 0.0; }
#pragma unroll
                for (int bl = 0; bl < NB; bl++) {
                    if (!((nzA >> (bi * NB + bl)) & 1ull)) continue;
                    const double2 av = *reinterpret_cast<const double2*>(Am + ((bi * NB + bl) << 6) + (g << 3) + 2 * q);
                    // even-k halves of all NB accumulators first, then the odd-k halves: back-to-back DMMAs are independent
#pragma unroll
                    for (int bj = 0; bj < NB; bj++) sls_dmma(Tm[bi][bj][0], Tm[bi][bj][1], av.x, S[bj][bl][0]);
#pragma unroll
                    for (int bj = 0; bj < NB; bj++) sls_dmma(Tm[bi][bj][0], Tm[bi][bj][1], av.y, S[bj][bl][1]);
                }
            }
            // (6) S_k = Dp + Tm Am^T
#pragma unroll
            for (int bi = 0; bi < NB; bi++)
#pragma unroll
                for (int bj = 0; bj < NB; bj++) {
                    const double2 bv = *reinterpret_cast<const double2*>(Dp + ((bi * NB + bj) << 6) + (g << 3) + 2 * q);
                    S[bi][bj][0] = bv.x; S[bi][bj][1] = bv.y;
                }
#pragma unroll
            for (int bj = 0; bj < NB; bj++)
#pragma unroll
                for (int bl = 0; bl < NB; bl++) {
                    if (!((nzA >> (bj * NB + bl)) & 1ull)) continue;
                    const double2 av = *reinterpret_cast<const double2*>(Am + ((bj * NB + bl) << 6) + (g << 3) + 2 * q);
#pragma unroll
                    for (int bi = 0; bi < NB; bi++) sls_dmma(S[bi][bj][0], S[bi][bj][1], Tm[bi][bl][0], av.x);
#pragma unroll
                    for (int bi = 0; bi < NB; bi++) sls_dmma(S[bi][bj][0], S[bi][bj][1], Tm[bi][bl][1], av.y);
                }
            // (7) block symmetric sweep with rank skipping:  S <- -(S^+).  A skipped pivot (<= threshold) gets
            //     pinv = 0, which zeroes its row and column at once (they never feed a kept row).
            int nkept = 0;                                        // pivots kept in this tap (a kept pivot leaves a non-zero diagonal of -P)
#pragma unroll
            for (int p = 0; p < NB; p++) {
                double d0 = S[p][p][0], d1 = S[p][p][1];
                const uint32_t actp = (uint32_t)(nonempty >> (8 * p)) & 0xffu;   // non-empty rows of D_k in this block
                // rows beyond n are never touched; an empty row inside the block (threshold +inf) takes the skip path of
                // the arithmetic below (pinv = 0 on an all-zero row and column), so no per-pivot branch is needed
                const int npair = (actp == 0u) ? 0 : ((min(8, n - 8 * p) + 1) >> 1);
#pragma unroll 1
                for (int rp = 0; rp < npair; rp++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int r = 2 * rp + e;
#if SLS_K2_PIVOT_SMEM
                        // pivot row through shared memory (two alternating 8-double buffers: lam_s and rhs_s, which only the back
                        // substitution uses; N >= 8 doubles each in every size class): the
                        // four lanes of row r post it, everybody reads its two row entries, the pivot and — by symmetry,
                        // S[g][r] = S[r][g] — its column entry from it.  1 STS + 3 LDS instead of 8 SHFL and the ~10 register
                        // moves that re-pair the shuffled 32-bit halves.
                        double* rb = e ? rhs_s : lam_s;
                        if (g == r) *reinterpret_cast<double2*>(rb + 2 * q) = make_double2(d0, d1);
                        __syncwarp();
                        const double2 rrv = *reinterpret_cast<const double2*>(rb + 2 * q);
                        const double rr0 = rrv.x, rr1 = rrv.y;
                        const double cg = rb[g];
                        const double dv = rb[r];
#else
                        const double src = e ? d1 : d0;
                        const double dv = __shfl_sync(SLS_FULL, src, 4 * r + rp);
                        const double cg = __shfl_sync(SLS_FULL, src, (lane & ~3) | rp);
                        const double rr0 = __shfl_sync(SLS_FULL, d0, 4 * r + q);
                        const double rr1 = __shfl_sync(SLS_FULL, d1, 4 * r + q);
#endif
                        const bool ok = dv > thrs[8 * p + r];
                        const double pinv = ok ? sls_rcp(dv) : 0.0;
                        const bool isrow = (g == r);
                        const double cp = cg * pinv;
                        const double f = isrow ? (1.0 - pinv) : cp;
                        double n0 = fma(-f, rr0, d0), n1 = fma(-f, rr1, d1);
                        const double cv = isrow ? -pinv : cp;
                        if (q == rp) { if (e) n1 = cv; else n0 = cv; }
                        d0 = n0; d1 = n1;
                    }
                }
                // a kept pivot ends with a strictly negative diagonal entry of -P, a skipped one (and every row the sweep never
                // touched: all-zero) with an exact zero, so the kept pivots are counted from the diagonal in one ballot
                nkept += __popc(__ballot_sync(SLS_FULL, (q == (g >> 1)) && (((g & 1) ? d1 : d0) != 0.0)));
                const double Pn[2] = {d0, d1};                // -P, symmetric
                double Tn[NB][2], Tt[NB][2];                  // Tn_i = -S_ip P,  Tt_i = -P S_pi = -(S_ip P)^T
#pragma unroll
                for (int bi = 0; bi < NB; bi++) { Tn[bi][0] = 0.0; Tn[bi][1] = 0.0; Tt[bi][0] = 0.0; Tt[bi][1] = 0.0; }
#pragma unroll
                for (int e = 0; e < 2; e++) {
#pragma unroll
                    for (int bi = 0; bi < NB; bi++) {
                        if (bi == p) continue;
                        sls_dmma(Tn[bi][0], Tn[bi][1], S[bi][p][e], Pn[e]);
                        sls_dmma(Tt[bi][0], Tt[bi][1], Pn[e], S[bi][p][e]);
                    }
                }
#pragma unroll
                for (int e = 0; e < 2; e++) {
#pragma unroll
                    for (int bi = 0; bi < NB; bi++) {
                        if (bi == p) continue;
#pragma unroll
                        for (int bj = 0; bj < NB; bj++) {
                            if (bj == p) continue;
                            sls_dmma(S[bi][bj][0], S[bi][bj][1], Tn[bi][e], S[bj][p][e]);   // S_ij -= (S_ip P) S_jp^T
                        }
                    }
                }
#pragma unroll
                for (int bi = 0; bi < NB; bi++) {
                    if (bi == p) continue;
                    S[bi][p][0] = -Tn[bi][0]; S[bi][p][1] = -Tn[bi][1];
                    S[p][bi][0] = -Tt[bi][0]; S[p][bi][1] = -Tt[bi][1];
                }
                S[p][p][0] = d0; S[p][p][1] = d1;
            }
            if (nkept != __popcll(nonempty)) status |= SLS_ST_RANK_SKIPPED;   // kept pivots are a subset of the non-empty rows
            // (8) history: the lower-triangle blocks of -Z_k in REGISTER order (block b at hk + 64 b, a lane's two entries at
            //     2*lane: one coalesced 16-byte store per block, read back the same way — no index arithmetic, no predicates;
            //     rows / columns beyond n are exact zeros), k = 1..T-1 (Z_0 is diag(mx_1)); -Z_T, which the back substitution
            //     uses straight from the registers, goes to slot 0 for the refinement sweeps
            {
                double* hk = zhist + (size_t)(k < T ? k : 0) * HB + 2 * lane;
                int b = 0;
#pragma unroll
                for (int bi = 0; bi < NB; bi++)
#pragma unroll
                    for (int bj = 0; bj <= bi; bj++) {
                        *reinterpret_cast<double2*>(hk + 64 * b) = make_double2(S[bi][bj][0], S[bi][bj][1]);
                        b++;
                    }
            }
            __syncwarp();
            double* tsw = yprev; yprev = ycur; ycur = tsw;
        }

        // ------------------------------------------------------------------ lambda_T = Z_T yh_T
#pragma unroll
        for (int bi = 0; bi < NB; bi++) {
            double acc = 0.0;
#pragma unroll
            for (int bj = 0; bj < NB; bj++) {
                const double2 yv = *reinterpret_cast<const double2*>(yprev + 8 * bj + 2 * q);
                acc = fma(S[bi][bj][0], yv.x, acc);
                acc = fma(S[bi][bj][1], yv.y, acc);
            }
            acc += __shfl_xor_sync(SLS_FULL, acc, 1);
            acc += __shfl_xor_sync(SLS_FULL, acc, 2);
            const int r = 8 * bi + g;
            if (q == 0) lam_s[r] = (r < n) ? -acc : 0.0;
        }
        __syncwarp();

        // ------------------------------------------------------------------ back substitution, k = T..1
        // lam_s holds lambda_k (zero beyond n) at the top of every iteration
        double* const xcur = ucur; // [x_k | u_k] (shared: the residual gathers both through the local CSR, whose B~ columns are n + l)
        double xnext_r[RPL];       // x_{k+1}[i] of the lane's own rows stays in registers
        double resmax = 0.0, resprev = SLS_INF;
        // Refinement (cold path): the recursion works on the normal equations G G^T, so lambda — and with it v = G^T lambda —
        // carries an error ~eps cond(G)^2.  The residual rho = h - G v of the COMPUTED v is accurate to eps ||G|| ||v||, and
        // v += G^T dlambda with (G G^T) dlambda = rho (stored factors: one forward sweep over the -Z_k history, one more back
        // substitution) brings the error down to ~eps cond(G) (Bjorck's corrected normal equations).  Taken only by columns
        // whose residual exceeds refine_tol (none on the BASELINE chain / grid configs).
#pragma unroll 1
        for (int pass = 0;; pass++) {
#pragma unroll
        for (int rr = 0; rr < RPL; rr++) xnext_r[rr] = 0.0;
        resmax = 0.0;
        // History double buffer: -Z_{k-1} and yh_{k-1} of the NEXT tap are loaded into registers one full iteration ahead
        // (the matrix registers of the forward pass are dead here), so the L2/HBM latency of the history hides behind a
        // whole tap instead of heading its serial chain.
        constexpr int NZB = NB * (NB + 1) / 2;
        double zn[NZB][2], yn[RPL];
        auto load_hist = [&](int kk, double (&z)[NZB][2], double (&yv)[RPL]) {
            const double* hk = zhist + (size_t)kk * HB + 2 * lane;
#pragma unroll
            for (int i = 0; i < NZB; i++) {
                double2 v = make_double2(0.0, 0.0);
                if (kk >= 1) v = *reinterpret_cast<const double2*>(hk + 64 * i);
                z[i][0] = v.x; z[i][1] = v.y;
            }
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) { const int i = lane + 32 * rr; yv[rr] = (i < n) ? yhist[(size_t)kk * N + i] : 0.0; }
        };
        load_hist(T - 1, zn, yn);
        for (int k = T; k >= 1; k--) {
            const mask_t mk = mxs[k - 1];
            const uint32_t* muk = mus + (size_t)(k - 1) * wu;
            double zc[NZB][2], yc[RPL];
#pragma unroll
            for (int i = 0; i < NZB; i++) { zc[i][0] = zn[i][0]; zc[i][1] = zn[i][1]; }
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) yc[rr] = yn[rr];
            if (k >= 2) load_hist(k - 2, zn, yn);
            __syncwarp();
            // t = C~^T lambda_k
#pragma unroll 1
            for (int p = lane; p < n + m; p += 32) {
                double acc = 0.0;
                const int e0 = cptr[p], e1 = cptr[p + 1];
#pragma unroll 1
                for (int e = e0; e < e1; e++) acc = fma(valc[e], lam_s[rowc[e]], acc);
                tx[p] = acc;
            }
            __syncwarp();
            // u_k = -mu_k o t_u  -> packed result
            {
                int rank0 = 0;
#pragma unroll 1
                for (int l0 = 0; l0 < m; l0 += 32) {
                    const int l = l0 + lane;
                    const uint32_t word = muk[l0 >> 5];
                    const bool on = (l < m) && ((word >> lane) & 1u);
                    double u = on ? -tx[n + l] : 0.0;
                    if (on) {
                        double* dst = a.out_nzval[T + k - 1] + (obase[T + k - 1] + rank0 + __popc(word & ((1u << lane) - 1u)));
                        if (pass) u += *dst;                    // refinement: u_k += du_k
                        *dst = u;
                    }
                    if (l < m) ucur[n + l] = u;
                    rank0 += __popc(word);
                }
            }
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) {
                const int i = lane + 32 * rr;
                const double yprev_k = yc[rr];
                if (i < N) rhs_s[i] = (i < n) ? (yprev_k + (sls_mbit(mk, i) ? tx[i] : 0.0)) : 0.0;
            }
            __syncwarp();
            // lambda_{k-1} = Z_{k-1} rhs  (history read back in accumulator layout; symmetric mat-vec) -> gbuf
            if (k - 1 == 0) {
#pragma unroll
                for (int rr = 0; rr < RPL; rr++) {
                    const int i = lane + 32 * rr;
                    if (i < N) gbuf[i] = (i < n && sls_mbit(mxs[0], i)) ? -rhs_s[i] : 0.0;
                }
            } else {
                double racc[NB], cacc[NB][2];
#pragma unroll
                for (int b = 0; b < NB; b++) { racc[b] = 0.0; cacc[b][0] = 0.0; cacc[b][1] = 0.0; }
                int zi = 0;
#pragma unroll
                for (int bi = 0; bi < NB; bi++) {
                    const int r = 8 * bi + g;
                    const double vr = rhs_s[r];
#pragma unroll
                    for (int bj = 0; bj <= bi; bj++) {
                        const int c0 = 8 * bj + 2 * q;
                        const double2 vc = *reinterpret_cast<const double2*>(rhs_s + c0);
                        const double z0 = zc[zi][0], z1 = zc[zi][1];
                        zi++;
                        if (bj < bi) {
                            cacc[bj][0] = fma(z0, vr, cacc[bj][0]);
                            cacc[bj][1] = fma(z1, vr, cacc[bj][1]);
                        }
                        racc[bi] = fma(z0, vc.x, racc[bi]);
                        racc[bi] = fma(z1, vc.y, racc[bi]);
                    }
                }
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    racc[b] += __shfl_xor_sync(SLS_FULL, racc[b], 1);
                    racc[b] += __shfl_xor_sync(SLS_FULL, racc[b], 2);
                    if (b < NB - 1) {
#pragma unroll
                        for (int d = 4; d < 32; d <<= 1) {
                            cacc[b][0] += __shfl_xor_sync(SLS_FULL, cacc[b][0], d);
                            cacc[b][1] += __shfl_xor_sync(SLS_FULL, cacc[b][1], d);
                        }
                    }
                }
                // row part of row 8b+g sits in every lane of quad g; column part of rows 8b+2q, 8b+2q+1 in every g
                if (q == 0) {
#pragma unroll
                    for (int b = 0; b < NB; b++) gbuf[8 * b + g] = racc[b];
                }
                __syncwarp();
                if (g == 0) {
#pragma unroll
                    for (int b = 0; b < NB - 1; b++) { gbuf[8 * b + 2 * q] += cacc[b][0]; gbuf[8 * b + 2 * q + 1] += cacc[b][1]; }
                }
            }
            __syncwarp();
            // gbuf = -lambda_{k-1} (history stores -Z);  x_k = mx_k o (lambda_{k-1} - A~^T lambda_k)
            double xk_r[RPL];
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) {
                const int i = lane + 32 * rr;
                const bool on = (i < n) && sls_mbit(mk, i);
                const double lam_prev = (i < n) ? -gbuf[i] : 0.0;
                double x = on ? (lam_prev - tx[i]) : 0.0;
                if (on) {
                    double* dst = a.out_nzval[k - 1] + (obase[k - 1] + sls_mrank(mk, i));
                    if (pass) x += *dst;                        // refinement: x_k += dx_k
                    *dst = x;
                }
                xk_r[rr] = x;
                if (i < n) xcur[i] = x;
                if (i < N) lam_s[i] = lam_prev;
            }
            __syncwarp();
            // residual of constraint block k: A~ x_k + B~ u_k - x_{k+1}
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) {
                const int i = lane + 32 * rr;
                if (i < n) {
                    double r = -xnext_r[rr];
                    const int r0 = rptr[i], r2 = rptr[i + 1];
#pragma unroll 1
                    for (int e = r0; e < r2; e++) r = fma(valr[e], xcur[colr[e]], r);
                    const double ar = fabs(r);
                    resmax = sls_nanmax(resmax, ar);
                }
            }
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) xnext_r[rr] = xk_r[rr];
        }
        __syncwarp();
        // block 0: x_1 = e
#pragma unroll
        for (int rr = 0; rr < RPL; rr++) {
            const int i = lane + 32 * rr;
            if (i < n) { const double ar = fabs(((i == jl) ? 1.0 : 0.0) - xnext_r[rr]); resmax = sls_nanmax(resmax, ar); }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) resmax = sls_nanmax(resmax, __shfl_xor_sync(SLS_FULL, resmax, d));
        // stop: accurate enough, out of sweeps, no longer contracting (infeasible column, or cond(G)^2 eps >~ 1), or NaN
        // (a residual above 1e-2 of the unit impulse is an infeasible localisation, not a conditioning problem: no sweep)
        if (!(resmax > a.refine_tol) || pass >= a.max_refine || !(resmax < 0.5 * resprev) || !(resmax <= 1e-2)) break;
        resprev = resmax;
        if (pass == 0 && lane == 0 && a.refined) atomicAdd(a.refined, 1);
        __syncwarp();
        // ---- (R1) rho_k = h_k - (G v)_k of the stored solution -> yhist[k]   (rho_0 = e - x_1, rho_k = A~ x_k + B~ u_k - x_{k+1})
#pragma unroll
        for (int rr = 0; rr < RPL; rr++) xnext_r[rr] = 0.0;
#pragma unroll 1
        for (int k = T; k >= 1; k--) {
            const mask_t mk = mxs[k - 1];
            const uint32_t* muk = mus + (size_t)(k - 1) * wu;
            int rank0 = 0;
#pragma unroll 1
            for (int l0 = 0; l0 < m; l0 += 32) {
                const int l = l0 + lane;
                const uint32_t word = muk[l0 >> 5];
                const bool on = (l < m) && ((word >> lane) & 1u);
                if (l < m) ucur[n + l] = on ? a.out_nzval[T + k - 1][obase[T + k - 1] + rank0 + __popc(word & ((1u << lane) - 1u))] : 0.0;
                rank0 += __popc(word);
            }
            double xk_r[RPL];
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) {
                const int i = lane + 32 * rr;
                const bool on = (i < n) && sls_mbit(mk, i);
                const double x = on ? a.out_nzval[k - 1][obase[k - 1] + sls_mrank(mk, i)] : 0.0;
                xk_r[rr] = x;
                if (i < n) xcur[i] = x;
            }
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) {
                const int i = lane + 32 * rr;
                if (i < N) {
                    double r = 0.0;
                    if (i < n) {
                        r = -xnext_r[rr];
                        const int r0 = rptr[i], r2 = rptr[i + 1];
#pragma unroll 1
                        for (int e = r0; e < r2; e++) r = fma(valr[e], xcur[colr[e]], r);
                    }
                    yhist[(size_t)k * N + i] = r;
                }
                xnext_r[rr] = xk_r[rr];
            }
            __syncwarp();
        }
#pragma unroll
        for (int rr = 0; rr < RPL; rr++) {
            const int i = lane + 32 * rr;
            if (i < N) { const double r = (i < n) ? (((i == jl) ? 1.0 : 0.0) - xnext_r[rr]) : 0.0; yhist[i] = r; rhs_s[i] = r; }
        }
        __syncwarp();
        // ---- (R2) forward sweep with the stored factors: yh_0 = rho_0, yh_k = rho_k + A~ (mx_k o Z_{k-1} yh_{k-1}) -> yhist[k]
        //      rhs_s holds yh_{k-1} at the top of every iteration
#pragma unroll 1
        for (int k = 1; k <= T + 1; k++) {
            // gbuf = (-Z_{k-1}) rhs_s  (k-1 = 0: Z_0 = diag(mx_1); k-1 = T: slot 0)
            if (k == 1) {
#pragma unroll
                for (int rr = 0; rr < RPL; rr++) {
                    const int i = lane + 32 * rr;
                    if (i < N) gbuf[i] = (i < n && sls_mbit(mxs[0], i)) ? -rhs_s[i] : 0.0;
                }
            } else {
                const double* hk = zhist + (size_t)(k - 1 < T ? k - 1 : 0) * HB + 2 * lane;
                int zi = 0;
                double racc[NB], cacc[NB][2];
#pragma unroll
                for (int b = 0; b < NB; b++) { racc[b] = 0.0; cacc[b][0] = 0.0; cacc[b][1] = 0.0; }
#pragma unroll
                for (int bi = 0; bi < NB; bi++) {
                    const int r = 8 * bi + g;
                    const double vr = rhs_s[r];
#pragma unroll
                    for (int bj = 0; bj <= bi; bj++) {
                        const int c0 = 8 * bj + 2 * q;
                        const double2 vc = *reinterpret_cast<const double2*>(rhs_s + c0);
                        const double2 zv = *reinterpret_cast<const double2*>(hk + 64 * zi);
                        const double z0 = zv.x, z1 = zv.y;
                        zi++;
                        if (bj < bi) { cacc[bj][0] = fma(z0, vr, cacc[bj][0]); cacc[bj][1] = fma(z1, vr, cacc[bj][1]); }
                        racc[bi] = fma(z0, vc.x, racc[bi]);
                        racc[bi] = fma(z1, vc.y, racc[bi]);
                    }
                }
#pragma unroll
                for (int b = 0; b < NB; b++) {
                    racc[b] += __shfl_xor_sync(SLS_FULL, racc[b], 1);
                    racc[b] += __shfl_xor_sync(SLS_FULL, racc[b], 2);
                    if (b < NB - 1) {
#pragma unroll
                        for (int d = 4; d < 32; d <<= 1) {
                            cacc[b][0] += __shfl_xor_sync(SLS_FULL, cacc[b][0], d);
                            cacc[b][1] += __shfl_xor_sync(SLS_FULL, cacc[b][1], d);
                        }
                    }
                }
                __syncwarp();
                if (q == 0) {
#pragma unroll
                    for (int b = 0; b < NB; b++) gbuf[8 * b + g] = racc[b];
                }
                __syncwarp();
                if (g == 0) {
#pragma unroll
                    for (int b = 0; b < NB - 1; b++) { gbuf[8 * b + 2 * q] += cacc[b][0]; gbuf[8 * b + 2 * q + 1] += cacc[b][1]; }
                }
            }
            __syncwarp();
            if (k == T + 1) {                                   // lambda_T = Z_T yh_T
#pragma unroll
                for (int rr = 0; rr < RPL; rr++) {
                    const int i = lane + 32 * rr;
                    if (i < N) lam_s[i] = (i < n) ? -gbuf[i] : 0.0;
                }
                break;
            }
            // g = mx_k o (Z_{k-1} yh_{k-1}) = -(mx_k o gbuf);  yh_k = rho_k + A~ g
            const mask_t mk = mxs[k - 1];
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) {
                const int i = lane + 32 * rr;
                if (i < N) xa[i] = (i < n && sls_mbit(mk, i)) ? -gbuf[i] : 0.0;
            }
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < RPL; rr++) {
                const int i = lane + 32 * rr;
                if (i < N) {
                    double acc = 0.0;
                    if (i < n) {
                        acc = yhist[(size_t)k * N + i];
                        const int r0 = rptr[i], r1 = rsplit[i];
#pragma unroll 1
                        for (int e = r0; e < r1; e++) acc = fma(valr[e], xa[colr[e]], acc);
                    }
                    yhist[(size_t)k * N + i] = acc;
                    rhs_s[i] = acc;
                }
            }
            __syncwarp();
        }
        __syncwarp();
        }   // refinement loop
        if (!(resmax <= a.feas_tol)) status |= SLS_ST_RESIDUAL;
        if (lane == 0) { a.status[c] = status; a.resid[c] = resmax; }
    }
}

// ---- host side ------------------------------------------------------------------------------
static const int kClasses[] = {8, 16, 24, 32, 40, 48, 56, 64, 72, 80, 88, 96, 104, 112, 120, 128};
int sls_solve_class_count() { return (int)(sizeof(kClasses) / sizeof(int)); }
int sls_solve_class_N(int k) { return kClasses[k]; }
int sls_solve_class_of(int n) {
    for (int k = 0; k < sls_solve_class_count(); k++) if (kClasses[k] >= n) return kClasses[k];
    return -1;
}
size_t sls_solve_smem_per_warp(int N, int mcap, int nnzcap, int T, int wu) {
    return (size_t)SlsSolveLayout(N, mcap, nnzcap, T, wu).total;
}
// size classes above 32 run one CTA per column (sls_solve_wide.cu).  Measured both ways on B200: the one-warp kernel
// needs 278 ms for the 300 x 300 grid (matrix spilled to local memory) against 131 ms here; the CTA kernel needs 19.5 ms for
// the chain's N = 24 class against 12.5 ms for the one-warp kernel (profiles/r02_ab_logs.txt).
static bool use_wide(int N) { return N > 32 && N <= 64; }
// 64 < n <= 128: one CTA per column on matrices in global scratch (sls_solve_huge.cu) — a completeness path, not a DMMA kernel
// (also every class of a plan whose input masks need more than four words, m > 128: the DMMA kernels keep mu_k in four registers)
static bool use_huge(int N, int wu) { return N > 64 || wu > 4; }
int64_t sls_solve_hist_doubles(int N, int T, int wu) {
    if (use_huge(N, wu)) return sls_solve_huge_hist_doubles(N, T);
    if (use_wide(N)) return sls_solve_wide_hist_doubles(N, T);
    return (int64_t)(T + 1) * N + (int64_t)T * ((N / 8) * (N / 8 + 1) / 2 * 64);
}

template <int NB>
static int resident_warps(int mcap, int nnzcap, int T, int wu, int sm_count, int max_smem) {
    constexpr int WPB = SlsCfg<NB>::WPB;
    if (wu > 4) return -1;
    const size_t per_block = sls_solve_smem_per_warp(8 * NB, mcap, nnzcap, T, wu) * WPB;
    if (per_block > (size_t)max_smem) return -1;
    auto kp = &sls_solve_kernel<NB>;
    if (per_block > 48 * 1024) { if (rt::set_max_smem(kp, (int)per_block)) return -1; }
    int bps = rt::max_blocks_per_sm(kp, WPB * 32, per_block);
    if (bps < 1) return -1;
    if (const char* cap = getenv("SLS_MAX_BPS")) { const int c = atoi(cap); if (c >= 1 && c < bps) bps = c; }   // tuning knob
    return bps * sm_count * WPB;
}

template <int NB>
static int launch_class(const SlsSolveArgs& a, int sm_count, int max_smem, int* warps_launched, sls_stream_t stream) {
    constexpr int WPB = SlsCfg<NB>::WPB;
    const size_t per_block = sls_solve_smem_per_warp(8 * NB, a.mcap, a.nnzcap, a.P.T, a.P.wu) * WPB;
    const int maxw = resident_warps<NB>(a.mcap, a.nnzcap, a.P.T, a.P.wu, sm_count, max_smem);
    if (maxw <= 0) return -5;
    auto kp = &sls_solve_kernel<NB>;
    int blocks = (a.count + WPB - 1) / WPB;
    if (blocks > maxw / WPB) blocks = maxw / WPB;
    if (blocks < 1) blocks = 1;
    *warps_launched = blocks * WPB;
    SLS_LAUNCH(kp, blocks, WPB * 32, per_block, stream, a);
    return rt::last_error();
}

#define SLS_FOR_CLASSES(X) X(1) X(2) X(3) X(4)

int sls_solve_max_resident_warps(int N, int mcap, int nnzcap, int T, int wu, int sm_count, int max_smem) {
    if (use_huge(N, wu)) return sls_solve_huge_max_slots(N, mcap, nnzcap, T, 4, wu, sm_count, max_smem);
    if (use_wide(N)) return sls_solve_wide_max_slots(N, mcap, nnzcap, T, wu, sm_count, max_smem);
    switch (N / 8) {
#define X(nb) case nb: return resident_warps<nb>(mcap, nnzcap, T, wu, sm_count, max_smem);
        SLS_FOR_CLASSES(X)
#undef X
        default: return -1;
    }
}

int sls_launch_solve(int N, const SlsSolveArgs& a, int sm_count, int max_smem, int* warps_launched,
                     sls_stream_t stream) {
    *warps_launched = 0;
    if (a.count == 0) return 0;
    if (use_huge(N, a.P.wu)) return sls_launch_solve_huge(N, a, sm_count, max_smem, warps_launched, stream);
    if (use_wide(N)) return sls_launch_solve_wide(N, a, sm_count, max_smem, warps_launched, stream);
    switch (N / 8) {
#define X(nb) case nb: return launch_class<nb>(a, sm_count, max_smem, warps_launched, stream);
        SLS_FOR_CLASSES(X)
#undef X
        default: return -6;
    }
}
