// sls_solve_huge.cu — kernel 2 for local state dimensions 64 < n <= 128: one CTA per column, matrices in the CTA's global
// scratch slot (L2-resident), plain FP64 SIMT.  A completeness path: the reference (src/synthesis.jl:34-72) accepts any (d, plant),
// e.g. a 5-point grid with d = 5 has n = 85; the DMMA kernels stop at n = 64 (sls_solve.cu one warp per column, sls_solve_wide.cu one
// CTA per column with the matrix in registers).  Same mathematics as those kernels (see the header of sls_solve.cu):
//     S_k = diag(mx_{k+1}) + B~ diag(mu_k) B~^T + (A~ M_k)(I - Z_{k-1})(A~ M_k)^T,   Z_k = S_k^+ (rank skipping),
//     lambda_{k-1} = Z_{k-1}(yh_{k-1} + M_k A~^T lambda_k),  x_k = M_k(lambda_{k-1} - A~^T lambda_k),  u_k = -M^u_k B~^T lambda_k,
// with the products by the sparse A~ done as row gathers (O(nnz n) instead of O(n^3)) and S_k^+ by an in-place symmetric
// Gauss-Jordan sweep, one barrier pair per pivot.  -Z_k is swept in place in its history slot (no copy).  No refinement sweeps here:
// a column whose achievability residual exceeds feas_tol is flagged SLS_STATUS_RESIDUAL.
#include "sls_common.h"
#include "sls_kernels.h"
#include "sls_solve_common.h"

#define HUGE_THREADS 256

struct SlsHugeLayout {
    int N, mcap, nnzcap, T, wx, wu;
    int oY0, oY1, oG, oThr, oLam, oRhs, oTx, oXc, oRow, oValR, oValC, oPre, oSx, oSu, oMx, oMu, oRptr, oRsplit, oCptr, oColR, oRowC, oMisc, total;
    __host__ __device__ SlsHugeLayout(int N_, int mcap_, int nnzcap_, int T_, int wx_, int wu_) {
        N = N_; mcap = mcap_ < 1 ? 1 : mcap_; nnzcap = nnzcap_ < 1 ? 1 : nnzcap_; T = T_; wx = wx_ < 1 ? 1 : wx_; wu = wu_ < 1 ? 1 : wu_;
        int o = 0;
#define SLS_SEC(name, bytes) name = o; o += (bytes); o = (o + 15) & ~15
        SLS_SEC(oY0, 8 * N); SLS_SEC(oY1, 8 * N); SLS_SEC(oG, 8 * N); SLS_SEC(oThr, 8 * N); SLS_SEC(oLam, 8 * N); SLS_SEC(oRhs, 8 * N);
        SLS_SEC(oTx, 8 * (N + mcap)); SLS_SEC(oXc, 8 * (N + mcap)); SLS_SEC(oRow, 8 * N);
        SLS_SEC(oValR, 8 * nnzcap); SLS_SEC(oValC, 8 * nnzcap); SLS_SEC(oPre, 8 * 2 * T);
        SLS_SEC(oSx, 4 * N); SLS_SEC(oSu, 4 * mcap); SLS_SEC(oMx, 4 * (T + 1) * wx); SLS_SEC(oMu, 4 * T * wu);
        SLS_SEC(oRptr, 4 * (N + 2)); SLS_SEC(oRsplit, 4 * (N + 2)); SLS_SEC(oCptr, 4 * (N + mcap + 2));
        SLS_SEC(oColR, 2 * nnzcap); SLS_SEC(oRowC, 2 * nnzcap); SLS_SEC(oMisc, 64);
#undef SLS_SEC
        total = (o + 15) & ~15;
    }
};

__device__ __forceinline__ bool hbit(const uint32_t* w, int i) { return (w[i >> 5] >> (i & 31)) & 1u; }
__device__ __forceinline__ int hrank(const uint32_t* w, int i) {           // set bits below position i
    int r = 0;
    for (int k = 0; k < (i >> 5); k++) r += __popc(w[k]);
    return r + __popc(w[i >> 5] & ((1u << (i & 31)) - 1u));
}

__global__ void __launch_bounds__(HUGE_THREADS)
sls_solve_huge_kernel(SlsSolveArgs a, int N) {
    SLS_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x;
    const SlsPlanView& P = a.P;
    const int T = P.T, wx = P.wx, wu = P.wu;
    const SlsHugeLayout L(N, a.mcap, a.nnzcap, T, wx, wu);
    unsigned char* base = smem_raw;
    double* ybuf0 = reinterpret_cast<double*>(base + L.oY0);
    double* ybuf1 = reinterpret_cast<double*>(base + L.oY1);
    double* gbuf = reinterpret_cast<double*>(base + L.oG);
    double* thrs = reinterpret_cast<double*>(base + L.oThr);
    double* lam_s = reinterpret_cast<double*>(base + L.oLam);
    double* rhs_s = reinterpret_cast<double*>(base + L.oRhs);
    double* tx = reinterpret_cast<double*>(base + L.oTx);
    double* xc = reinterpret_cast<double*>(base + L.oXc);      // [x_k | u_k]
    double* prow = reinterpret_cast<double*>(base + L.oRow);    // pivot row of the sweep
    double* valr = reinterpret_cast<double*>(base + L.oValR);
    double* valc = reinterpret_cast<double*>(base + L.oValC);
    int64_t* obase = reinterpret_cast<int64_t*>(base + L.oPre);
    int32_t* sx = reinterpret_cast<int32_t*>(base + L.oSx);
    int32_t* su = reinterpret_cast<int32_t*>(base + L.oSu);
    uint32_t* mxs = reinterpret_cast<uint32_t*>(base + L.oMx);  // [T+1][wx], the last one all zero
    uint32_t* mus = reinterpret_cast<uint32_t*>(base + L.oMu);  // [T][wu]
    int32_t* rptr = reinterpret_cast<int32_t*>(base + L.oRptr);
    int32_t* rsplit = reinterpret_cast<int32_t*>(base + L.oRsplit);
    int32_t* cptr = reinterpret_cast<int32_t*>(base + L.oCptr);
    uint16_t* colr = reinterpret_cast<uint16_t*>(base + L.oColR);
    uint16_t* rowc = reinterpret_cast<uint16_t*>(base + L.oRowC);
    int32_t* misc = reinterpret_cast<int32_t*>(base + L.oMisc);  // [0] rank flag, [1] masks changed

    const int64_t NN = (int64_t)N * N;
    double* slot = a.hist + (size_t)blockIdx.x * a.hist_stride;
    double* yhist = slot;                                       // [T+1][N]
    double* zhist = slot + (size_t)(T + 1) * N;                 // [T+1][N*N]: -Z_k, k = 0..T, swept in place
    double* Dp = zhist + (size_t)(T + 1) * NN;                  // diag(mx_{k+1}) + B~ diag(mu_k) B~^T
    double* Tm = Dp + NN;                                       // (A~ M_k)(I - Z_{k-1})

    for (int idx = blockIdx.x; idx < a.count; idx += gridDim.x) {
        const int c = a.list[idx];
        const int n = P.n[c], m = P.m[c], jl = P.jloc[c];
        __syncthreads();
        for (int i = tid; i < n; i += HUGE_THREADS) sx[i] = P.sx[(size_t)c * P.nstride + i];
        for (int l = tid; l < m; l += HUGE_THREADS) su[l] = P.su[(size_t)c * P.mstride + l];
        for (int t = tid; t < T * wx; t += HUGE_THREADS) mxs[t] = P.maskx[(size_t)c * T * wx + t];
        for (int t = tid; t < wx; t += HUGE_THREADS) mxs[T * wx + t] = 0u;
        for (int t = tid; t < T * wu; t += HUGE_THREADS) mus[t] = P.masku[(size_t)c * T * wu + t];
        for (int t = tid; t < 2 * T; t += HUGE_THREADS) obase[t] = P.colptr[(int64_t)t * (P.Nx + 1) + P.cols[c]] - 1;
        if (tid == 0) { misc[0] = 0; }
        __syncthreads();
        // ---- local CSR of C~ = [A~ B~] (row i: A~ entries, then B~ entries with column n + l) and local CSC (columns 0..n+m)
        for (int i = tid; i < n; i += HUGE_THREADS) {
            const int64_t gr = sx[i];
            int ca = 0, cb = 0;
            for (int64_t e = a.Ar.colptr[gr] - 1; e < a.Ar.colptr[gr + 1] - 1; e++) ca += (sls_find(sx, n, (int32_t)(a.Ar.rowval[e] - 1)) >= 0);
            for (int64_t e = a.B2r.colptr[gr] - 1; e < a.B2r.colptr[gr + 1] - 1; e++) cb += (sls_find(su, m, (int32_t)(a.B2r.rowval[e] - 1)) >= 0);
            rptr[i + 1] = ca + cb; rsplit[i] = ca;
        }
        for (int p = tid; p < n + m; p += HUGE_THREADS) {
            const SlsCsc& M = (p < n) ? a.A : a.B2;
            const int64_t gc = (p < n) ? sx[p] : su[p - n];
            int cnt = 0;
            for (int64_t e = M.colptr[gc] - 1; e < M.colptr[gc + 1] - 1; e++) cnt += (sls_find(sx, n, (int32_t)(M.rowval[e] - 1)) >= 0);
            cptr[p + 1] = cnt;
        }
        __syncthreads();
        if (tid == 0) { rptr[0] = 0; for (int i = 0; i < n; i++) { rsplit[i] += rptr[i]; rptr[i + 1] += rptr[i]; } }
        if (tid == 32) { cptr[0] = 0; for (int p = 0; p < n + m; p++) cptr[p + 1] += cptr[p]; }
        __syncthreads();
        for (int i = tid; i < n; i += HUGE_THREADS) {
            const int64_t gr = sx[i];
            int o = rptr[i];
            for (int64_t e = a.Ar.colptr[gr] - 1; e < a.Ar.colptr[gr + 1] - 1; e++) {
                const int l = sls_find(sx, n, (int32_t)(a.Ar.rowval[e] - 1));
                if (l >= 0) { colr[o] = (uint16_t)l; valr[o] = a.Ar.nzval[e]; o++; }
            }
            for (int64_t e = a.B2r.colptr[gr] - 1; e < a.B2r.colptr[gr + 1] - 1; e++) {
                const int l = sls_find(su, m, (int32_t)(a.B2r.rowval[e] - 1));
                if (l >= 0) { colr[o] = (uint16_t)(n + l); valr[o] = a.B2r.nzval[e]; o++; }
            }
        }
        for (int p = tid; p < n + m; p += HUGE_THREADS) {
            const SlsCsc& M = (p < n) ? a.A : a.B2;
            const int64_t gc = (p < n) ? sx[p] : su[p - n];
            int o = cptr[p];
            for (int64_t e = M.colptr[gc] - 1; e < M.colptr[gc + 1] - 1; e++) {
                const int r = sls_find(sx, n, (int32_t)(M.rowval[e] - 1));
                if (r >= 0) { rowc[o] = (uint16_t)r; valc[o] = M.nzval[e]; o++; }
            }
        }
        // ---- tap 0: -Z_0 = -diag(mx_1), yh_0 = e
        double* S = zhist;
        for (int64_t e = tid; e < NN; e += HUGE_THREADS) {
            const int i = (int)(e / N), j = (int)(e % N);
            S[e] = (i == j && i < n && hbit(mxs, i)) ? -1.0 : 0.0;
        }
        double* yprev = ybuf0;
        double* ycur = ybuf1;
        for (int i = tid; i < N; i += HUGE_THREADS) { const double y0 = (i == jl) ? 1.0 : 0.0; yprev[i] = y0; yhist[i] = y0; }
        int status = (jl < 0) ? SLS_ST_NO_IMPULSE : 0;
        __syncthreads();

        // ---- forward elimination
        for (int k = 1; k <= T; k++) {
            const uint32_t* mk = mxs + (size_t)(k - 1) * wx;
            const uint32_t* mk1 = mxs + (size_t)k * wx;            // all zero for k = T
            const uint32_t* muk = mus + (size_t)(k - 1) * wu;
            double* Sn = zhist + (size_t)k * NN;
            // (1) g = -(mx_k o (S yprev))   (S symmetric: column reads are coalesced)
            for (int r = tid; r < N; r += HUGE_THREADS) {
                double acc = 0.0;
                if (r < n && hbit(mk, r)) for (int cc = 0; cc < n; cc++) acc = fma(S[(int64_t)cc * N + r], yprev[cc], acc);
                gbuf[r] = -acc;
            }
            // (2) mask-dependent operands, rebuilt only when a mask changed
            if (tid == 0) {
                // operands of tap k: (mx_k, mx_{k+1}, mu_k); those of tap k-1: (mx_{k-1}, mx_k, mu_{k-1})
                bool same = (k >= 2);
                if (same) {
                    for (int w = 0; w < wx; w++) same = same && (mk[w] == mxs[(size_t)(k - 2) * wx + w]) && (mk1[w] == mk[w]);
                    for (int w = 0; w < wu; w++) same = same && (muk[w] == mus[(size_t)(k - 2) * wu + w]);
                }
                misc[1] = same ? 0 : 1;
            }
            __syncthreads();
            if (misc[1]) {
                for (int64_t e = tid; e < NN; e += HUGE_THREADS) Dp[e] = 0.0;
                __syncthreads();
                for (int i = tid; i < N; i += HUGE_THREADS) {
                    double dref = 0.0;
                    if (i < n) {
                        if (hbit(mk1, i)) { dref = 1.0; Dp[(int64_t)i * N + i] = 1.0; }
                        for (int e = rptr[i]; e < rsplit[i]; e++) if (hbit(mk, colr[e])) dref = fma(valr[e], valr[e], dref);
                        for (int e = rsplit[i]; e < rptr[i + 1]; e++) {
                            const int l = colr[e] - n;
                            if (!hbit(muk, l)) continue;
                            const double v = valr[e];
                            dref = fma(v, v, dref);
                            for (int e2 = cptr[n + l]; e2 < cptr[n + l + 1]; e2++) Dp[(int64_t)i * N + rowc[e2]] += v * valc[e2];   // row i is this thread's alone
                        }
                    }
                    thrs[i] = (dref != 0.0) ? a.rank_tol * dref : SLS_INF;
                }
            }
            __syncthreads();
            // (3) yh_k = A~ g
            for (int i = tid; i < N; i += HUGE_THREADS) {
                double acc = 0.0;
                if (i < n) for (int e = rptr[i]; e < rsplit[i]; e++) acc = fma(valr[e], gbuf[colr[e]], acc);
                ycur[i] = acc; yhist[(size_t)k * N + i] = acc;
            }
            // (5) Tm = (A~ M_k)(I + S):  row i = sum_e a_e (e_{c_e} + S[c_e][:])
            for (int64_t e0 = tid; e0 < NN; e0 += HUGE_THREADS) {
                const int i = (int)(e0 / N), cc = (int)(e0 % N);
                double acc = 0.0;
                if (i < n) for (int e = rptr[i]; e < rsplit[i]; e++) {
                    const int l = colr[e];
                    if (!hbit(mk, l)) continue;
                    acc = fma(valr[e], S[(int64_t)l * N + cc] + ((l == cc) ? 1.0 : 0.0), acc);
                }
                Tm[e0] = acc;
            }
            __syncthreads();
            // (6) S_k = Dp + Tm (A~ M_k)^T:  entry (i, j) = Dp + sum_e a_{j,e} Tm[i][c_e]
            for (int64_t e0 = tid; e0 < NN; e0 += HUGE_THREADS) {
                const int i = (int)(e0 / N), j = (int)(e0 % N);
                double acc = Dp[e0];
                if (j < n) for (int e = rptr[j]; e < rsplit[j]; e++) {
                    const int l = colr[e];
                    if (hbit(mk, l)) acc = fma(valr[e], Tm[(int64_t)i * N + l], acc);
                }
                Sn[e0] = acc;
            }
            __syncthreads();
            // (7) in-place symmetric sweep with rank skipping: Sn <- -(Sn^+).  Empty rows (threshold +inf) are all zero: skipped.
            for (int r = 0; r < n; r++) {
                const double thr = thrs[r];
                if (thr == SLS_INF) continue;                      // block-uniform
                for (int cc = tid; cc < N; cc += HUGE_THREADS) prow[cc] = Sn[(int64_t)r * N + cc];
                __syncthreads();
                const double dv = prow[r];
                const bool ok = dv > thr;
                const double pinv = ok ? 1.0 / dv : 0.0;
                if (!ok && tid == 0) misc[0] = 1;
                for (int64_t e0 = tid; e0 < NN; e0 += HUGE_THREADS) {
                    const int i = (int)(e0 / N), cc = (int)(e0 % N);
                    double v;
                    if (i == r) v = (cc == r) ? -pinv : prow[cc] * pinv;
                    else {
                        const double cp = prow[i] * pinv;           // S[i][r] = S[r][i]
                        v = (cc == r) ? cp : fma(-cp, prow[cc], Sn[e0]);
                    }
                    Sn[e0] = v;
                }
                __syncthreads();
            }
            S = Sn;
            double* tsw = yprev; yprev = ycur; ycur = tsw;
        }
        // ---- lambda_T = Z_T yh_T
        for (int r = tid; r < N; r += HUGE_THREADS) {
            double acc = 0.0;
            if (r < n) for (int cc = 0; cc < n; cc++) acc = fma(S[(int64_t)cc * N + r], yprev[cc], acc);
            lam_s[r] = -acc;
        }
        __syncthreads();
        // ---- back substitution, k = T..1
        double resmax = 0.0;
        double xnext = 0.0;                                         // x_{k+1}[tid] (N <= HUGE_THREADS... guarded below)
        for (int k = T; k >= 1; k--) {
            const uint32_t* mk = mxs + (size_t)(k - 1) * wx;
            const uint32_t* muk = mus + (size_t)(k - 1) * wu;
            const double* Zp = zhist + (size_t)(k - 1) * NN;        // -Z_{k-1}
            for (int p = tid; p < n + m; p += HUGE_THREADS) {       // t = C~^T lambda_k
                double acc = 0.0;
                for (int e = cptr[p]; e < cptr[p + 1]; e++) acc = fma(valc[e], lam_s[rowc[e]], acc);
                tx[p] = acc;
            }
            __syncthreads();
            for (int l = tid; l < m; l += HUGE_THREADS) {           // u_k = -mu_k o t_u
                const bool on = hbit(muk, l);
                const double u = on ? -tx[n + l] : 0.0;
                if (on) a.out_nzval[T + k - 1][obase[T + k - 1] + hrank(muk, l)] = u;
                xc[n + l] = u;
            }
            for (int i = tid; i < N; i += HUGE_THREADS)
                rhs_s[i] = (i < n) ? (yhist[(size_t)(k - 1) * N + i] + (hbit(mk, i) ? tx[i] : 0.0)) : 0.0;
            __syncthreads();
            for (int r = tid; r < N; r += HUGE_THREADS) {           // gbuf = -lambda_{k-1} = (-Z_{k-1}) rhs
                double acc = 0.0;
                if (r < n) for (int cc = 0; cc < n; cc++) acc = fma(Zp[(int64_t)cc * N + r], rhs_s[cc], acc);
                gbuf[r] = acc;
            }
            __syncthreads();
            double xk = 0.0;
            if (tid < N) {                                          // x_k = mx_k o (lambda_{k-1} - A~^T lambda_k); N <= 128 <= HUGE_THREADS
                const int i = tid;
                const bool on = (i < n) && hbit(mk, i);
                const double lam_prev = (i < n) ? -gbuf[i] : 0.0;
                xk = on ? (lam_prev - tx[i]) : 0.0;
                if (on) a.out_nzval[k - 1][obase[k - 1] + hrank(mk, i)] = xk;
                if (i < n) xc[i] = xk;
                lam_s[i] = lam_prev;
            }
            __syncthreads();
            if (tid < n) {                                          // residual of constraint block k: A~ x_k + B~ u_k - x_{k+1}
                double r = -xnext;
                for (int e = rptr[tid]; e < rptr[tid + 1]; e++) r = fma(valr[e], xc[colr[e]], r);
                resmax = sls_nanmax(resmax, fabs(r));
            }
            xnext = xk;
            __syncthreads();
        }
        if (tid < n) resmax = sls_nanmax(resmax, fabs(((tid == jl) ? 1.0 : 0.0) - xnext));   // block 0: x_1 = e
        // block-wide NaN-keeping maximum (through shared memory: gbuf is free)
        if (tid < N) gbuf[tid] = (tid < n) ? resmax : 0.0;
        __syncthreads();
        if (tid == 0) {
            double rm = 0.0;
            for (int i = 0; i < n; i++) rm = sls_nanmax(rm, gbuf[i]);
            if (!(rm <= a.feas_tol)) status |= SLS_ST_RESIDUAL;
            if (misc[0]) status |= SLS_ST_RANK_SKIPPED;
            a.status[c] = status; a.resid[c] = rm;
        }
        __syncthreads();
    }
}

// ---- host side ------------------------------------------------------------------------------
size_t sls_solve_huge_smem(int N, int mcap, int nnzcap, int T, int wx, int wu) { return (size_t)SlsHugeLayout(N, mcap, nnzcap, T, wx, wu).total; }
int64_t sls_solve_huge_hist_doubles(int N, int T) { return (int64_t)(T + 1) * N + (int64_t)(T + 3) * N * N; }

int sls_solve_huge_max_slots(int N, int mcap, int nnzcap, int T, int wx, int wu, int sm_count, int max_smem) {
    const size_t smem = sls_solve_huge_smem(N, mcap, nnzcap, T, wx, wu);
    if (smem > (size_t)max_smem) return -1;
    auto kp = &sls_solve_huge_kernel;
    if (smem > 48 * 1024) { if (rt::set_max_smem(kp, (int)smem)) return -1; }
    int bps = rt::max_blocks_per_sm(kp, HUGE_THREADS, smem);
    if (bps < 1) return -1;
    if (bps > 4) bps = 4;                                       // history slots are 100s of KB each: enough CTAs to fill the SMs
    return bps * sm_count;
}

int sls_launch_solve_huge(int N, const SlsSolveArgs& a, int sm_count, int max_smem, int* slots_launched, sls_stream_t stream) {
    *slots_launched = 0;
    if (a.count == 0) return 0;
    const size_t smem = sls_solve_huge_smem(N, a.mcap, a.nnzcap, a.P.T, a.P.wx, a.P.wu);
    const int maxb = sls_solve_huge_max_slots(N, a.mcap, a.nnzcap, a.P.T, a.P.wx, a.P.wu, sm_count, max_smem);
    if (maxb <= 0) return -5;
    auto kp = &sls_solve_huge_kernel;
    int blocks = a.count < maxb ? a.count : maxb;
    *slots_launched = blocks;
    SLS_LAUNCH(kp, blocks, HUGE_THREADS, smem, stream, a, N);
    return rt::last_error();
}
