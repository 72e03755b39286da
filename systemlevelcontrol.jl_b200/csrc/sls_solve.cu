// sls_solve.cu — kernel 2: batched FP64 equality-constrained least squares, one warp per column.
//
// Replaces the JuMP/Ipopt solve of reference src/synthesis.jl:46-62 for plants built by
// Plant(A,B1=I,B2) ([C1 D12] = I, D11 = 0; GeneralizedPlant.jl:105-110).  Per column the reference's QP is
//     min  sum_t ||x_t||^2 + ||u_t||^2      over the in-mask entries (others fixed to 0, synthesis.jl:57-60)
//     s.t. x_1 = e_j,  x_{t+1} = A~ x_t + B~ u_t (t<T),  0 = A~ x_T + B~ u_T     (synthesis.jl:53-55)
// i.e. the minimum-norm solution v = G^T lambda, (G G^T) lambda = h, of a rank-deficient but consistent
// system.  G G^T is block tridiagonal in time with n x n blocks:
//     D_0 = diag(mx_1),   D_k = diag(mx_{k+1}) + A~ diag(mx_k) A~^T + B~ diag(mu_k) B~^T,   E_k = -A~ diag(mx_k)
// Forward block elimination (k = 0..T):   S_k = D_k - E_k Z_{k-1} E_k^T,  Z_k = S_k^+,
//     = diag(mx_{k+1}) + C~ blkdiag(mx_k o (I - Z_{k-1}) o mx_k, diag(mu_k)) C~^T,      C~ = [A~ B~]
// where S_k^+ is the inverse on the rows whose pivot exceeds rank_tol*diag(D_k) (redundant constraint
// rows are eliminated — what Ipopt's inertia correction does implicitly), computed by N symmetric
// Gauss-Jordan sweeps on a register-resident matrix: lane i owns row i (n <= N <= 32), the pivot
// column is broadcast through shared memory, every lane does N DFMAs per sweep.  The sparse triple
// product uses a local CSR of C~ gathered from the plant (this is where the local A~/B~ blocks and the
// FIR constraint block are built in shared memory).  -Z_k goes to a per-warp history slot (packed
// lower triangle, L2-resident) for the back substitution  lambda_{k-1} = Z_{k-1}(yh_{k-1} + mx_k o A~^T lambda_k),
// then  x_k = mx_k o (lambda_{k-1} - A~^T lambda_k),  u_k = -mu_k o (B~^T lambda_k).
// The achievability residual of the computed (x,u) is checked per column and reported in status/resid.
#include "sls_common.h"
#include "sls_kernels.h"

#define SOLVE_WPB 4

struct SlsSolveLayout {
    int N, LDW, LDP, NP, mcap, nnzcap, T, wu;
    int oWs, oPs, oYh, oCb, oDref, oLam, oTx, oRhs, oXa, oXb, oUc, oG, oValR, oValC;
    int oSx, oSu, oMx, oMu, oRptr, oRsplit, oColR, oCptr, oRowC, total;
    __host__ __device__ SlsSolveLayout(int N_, int mcap_, int nnzcap_, int T_, int wu_) {
        N = N_; mcap = mcap_ < 1 ? 1 : mcap_; nnzcap = nnzcap_ < 1 ? 1 : nnzcap_; T = T_; wu = wu_ < 1 ? 1 : wu_;
        LDW = N | 1; LDP = (N + 1) & ~1; NP = (N + 1) & ~1;
        int o = 0;
#define SLS_SEC(name, bytes) name = o; o += (bytes); o = (o + 15) & ~15
        SLS_SEC(oWs, 8 * N * LDW);
        SLS_SEC(oPs, 8 * (N + mcap) * LDP);
        SLS_SEC(oYh, 8 * (T + 1) * N);
        SLS_SEC(oCb, 8 * 2 * NP);
        SLS_SEC(oDref, 8 * NP);
        SLS_SEC(oLam, 8 * NP);
        SLS_SEC(oTx, 8 * (N + mcap));
        SLS_SEC(oRhs, 8 * NP);
        SLS_SEC(oXa, 8 * NP);
        SLS_SEC(oXb, 8 * NP);
        SLS_SEC(oUc, 8 * mcap);
        SLS_SEC(oG, 8 * NP);
        SLS_SEC(oValR, 8 * nnzcap);
        SLS_SEC(oValC, 8 * nnzcap);
        SLS_SEC(oSx, 4 * N);
        SLS_SEC(oSu, 4 * mcap);
        SLS_SEC(oMx, 4 * T);
        SLS_SEC(oMu, 4 * T * wu);
        SLS_SEC(oRptr, 2 * (N + 2));
        SLS_SEC(oRsplit, 2 * (N + 2));
        SLS_SEC(oColR, 2 * nnzcap);
        SLS_SEC(oCptr, 2 * (N + mcap + 2));
        SLS_SEC(oRowC, 2 * nnzcap);
#undef SLS_SEC
        total = (o + 15) & ~15;
    }
};

__device__ __forceinline__ bool sls_bit(const uint32_t* w, int i) { return (w[i >> 5] >> (i & 31)) & 1u; }

template <int N>
__global__ void __launch_bounds__(SOLVE_WPB * 32)
sls_solve_kernel(SlsSolveArgs a) {
    SLS_DYN_SMEM(smem_raw);
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int i = lane;
    const SlsPlanView& P = a.P;
    const int T = P.T;
    const SlsSolveLayout L(N, a.mcap, a.nnzcap, T, P.wu);
    unsigned char* base = smem_raw + (size_t)warp * L.total;
    double* Ws = reinterpret_cast<double*>(base + L.oWs);
    double* Ps = reinterpret_cast<double*>(base + L.oPs);
    double* yh = reinterpret_cast<double*>(base + L.oYh);
    double* cb = reinterpret_cast<double*>(base + L.oCb);
    double* drefs = reinterpret_cast<double*>(base + L.oDref);
    double* lam_s = reinterpret_cast<double*>(base + L.oLam);
    double* tx = reinterpret_cast<double*>(base + L.oTx);
    double* rhs_s = reinterpret_cast<double*>(base + L.oRhs);
    double* xa = reinterpret_cast<double*>(base + L.oXa);
    double* xb = reinterpret_cast<double*>(base + L.oXb);
    double* ucur = reinterpret_cast<double*>(base + L.oUc);
    double* gbuf = reinterpret_cast<double*>(base + L.oG);
    double* valr = reinterpret_cast<double*>(base + L.oValR);
    double* valc = reinterpret_cast<double*>(base + L.oValC);
    int32_t* sx = reinterpret_cast<int32_t*>(base + L.oSx);
    int32_t* su = reinterpret_cast<int32_t*>(base + L.oSu);
    uint32_t* mxs = reinterpret_cast<uint32_t*>(base + L.oMx);
    uint32_t* mus = reinterpret_cast<uint32_t*>(base + L.oMu);
    uint16_t* rptr = reinterpret_cast<uint16_t*>(base + L.oRptr);
    uint16_t* rsplit = reinterpret_cast<uint16_t*>(base + L.oRsplit);
    uint16_t* colr = reinterpret_cast<uint16_t*>(base + L.oColR);
    uint16_t* cptr = reinterpret_cast<uint16_t*>(base + L.oCptr);
    uint16_t* rowc = reinterpret_cast<uint16_t*>(base + L.oRowC);
    constexpr int LDW = (N | 1), LDP = ((N + 1) & ~1);
    const int wu = P.wu;

    const int slot = blockIdx.x * SOLVE_WPB + warp;
    const int nslots = gridDim.x * SOLVE_WPB;
    double* hist = a.hist + (size_t)slot * a.hist_stride;

    for (int idx = slot; idx < a.count; idx += nslots) {
        const int c = a.list[idx];
        const int n = P.n[c], m = P.m[c], jl = P.jloc[c];
        const int tri = n * (n + 1) / 2;
        const bool act = i < n;
        __syncwarp();
        // ------------------------------------------------------------------ metadata -> shared
        if (act) sx[i] = P.sx[(size_t)c * P.nstride + i];
        for (int l = lane; l < m; l += 32) su[l] = P.su[(size_t)c * P.mstride + l];
        for (int t = lane; t < T; t += 32) mxs[t] = P.maskx[((size_t)c * T + t) * P.wx];
        for (int t = lane; t < T * wu; t += 32) mus[t] = P.masku[(size_t)c * T * wu + t];
        __syncwarp();
        // ------------------------------------------------------------------ local CSR of C~ = [A~ B~]
        {
            int cntA = 0, cntB = 0;
            int64_t ea0 = 0, ea1 = 0, eb0 = 0, eb1 = 0;
            if (act) {
                const int64_t g = sx[i];
                ea0 = a.Ar.colptr[g] - 1; ea1 = a.Ar.colptr[g + 1] - 1;
                eb0 = a.B2r.colptr[g] - 1; eb1 = a.B2r.colptr[g + 1] - 1;
                for (int64_t e = ea0; e < ea1; e++) cntA += (sls_find(sx, n, (int32_t)(a.Ar.rowval[e] - 1)) >= 0);
                for (int64_t e = eb0; e < eb1; e++) cntB += (sls_find(su, m, (int32_t)(a.B2r.rowval[e] - 1)) >= 0);
            }
            int incl = cntA + cntB;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
            const int excl = incl - (cntA + cntB);
            if (act) { rptr[i] = (uint16_t)excl; rsplit[i] = (uint16_t)(excl + cntA); }
            if (i == n - 1 || (n == 0 && i == 0)) rptr[n] = (uint16_t)(n ? incl : 0);
            if (act) {
                int o = excl;
                for (int64_t e = ea0; e < ea1; e++) {
                    const int l = sls_find(sx, n, (int32_t)(a.Ar.rowval[e] - 1));
                    if (l >= 0) { colr[o] = (uint16_t)l; valr[o] = a.Ar.nzval[e]; o++; }
                }
                for (int64_t e = eb0; e < eb1; e++) {
                    const int l = sls_find(su, m, (int32_t)(a.B2r.rowval[e] - 1));
                    if (l >= 0) { colr[o] = (uint16_t)(n + l); valr[o] = a.B2r.nzval[e]; o++; }
                }
            }
        }
        // ------------------------------------------------------------------ local CSC of C~ (columns 0..n+m)
        {
            int carry = 0;
            for (int p0 = 0; p0 < n + m; p0 += 32) {
                const int p = p0 + lane;
                int cnt = 0;
                int64_t e0 = 0, e1 = 0;
                const SlsCsc& M = (p < n) ? a.A : a.B2;
                if (p < n + m) {
                    const int64_t g = (p < n) ? sx[p] : su[p - n];
                    e0 = M.colptr[g] - 1; e1 = M.colptr[g + 1] - 1;
                    for (int64_t e = e0; e < e1; e++) cnt += (sls_find(sx, n, (int32_t)(M.rowval[e] - 1)) >= 0);
                }
                int incl = cnt;
#pragma unroll
                for (int d = 1; d < 32; d <<= 1) { int o = __shfl_up_sync(SLS_FULL, incl, d); if (lane >= d) incl += o; }
                if (p < n + m) {
                    int o = carry + incl - cnt;
                    cptr[p] = (uint16_t)o;
                    for (int64_t e = e0; e < e1; e++) {
                        const int r = sls_find(sx, n, (int32_t)(M.rowval[e] - 1));
                        if (r >= 0) { rowc[o] = (uint16_t)r; valc[o] = M.nzval[e]; o++; }
                    }
                }
                carry += __shfl_sync(SLS_FULL, incl, 31);
            }
            if (lane == 0) cptr[n + m] = (uint16_t)carry;
        }
        __syncwarp();

        // ------------------------------------------------------------------ tap 0:  S_0 = diag(mx_1), yh_0 = e
        double s[N];                       // row i of -Z_k (the sweep works on -Z; signs folded below)
        const uint32_t m1 = mxs[0];
#pragma unroll
        for (int p = 0; p < N; p++) s[p] = (p == i && ((m1 >> p) & 1u)) ? -1.0 : 0.0;
        if (i < N) yh[i] = (i == jl) ? 1.0 : 0.0;
        {
            int hp = 0;
#pragma unroll
            for (int p = 0; p < N; p++) {
                if (p < n) { if (act && i >= p) hist[hp + i - p] = s[p]; hp += n - p; }
            }
        }
        int status = (jl < 0) ? SLS_ST_NO_IMPULSE : 0;
        __syncwarp();

        // ------------------------------------------------------------------ forward elimination, taps 1..T
        for (int k = 1; k <= T; k++) {
            const uint32_t mk = mxs[k - 1];
            const uint32_t mk1 = (k < T) ? mxs[k] : 0u;
            const uint32_t* muk = mus + (size_t)(k - 1) * wu;
            const bool bi = act && ((mk >> i) & 1u);
            const double* yprev = yh + (size_t)(k - 1) * N;
            // g = Z_{k-1} yh_{k-1} (masked by mx_k), W row = mx_k o (I - Z_{k-1}) o mx_k
            double g = 0.0;
#pragma unroll
            for (int p = 0; p < N; p++) g = fma(s[p], yprev[p], g);
            if (i < N) gbuf[i] = bi ? -g : 0.0;
            if (act) {
#pragma unroll
                for (int p = 0; p < N; p++) {
                    const double w = (bi && ((mk >> p) & 1u)) ? ((p == i ? 1.0 : 0.0) + s[p]) : 0.0;
                    Ws[i * LDW + p] = w;
                }
            }
            // zero the B rows of P
            for (int q = lane; q < m * LDP; q += 32) Ps[n * LDP + q] = 0.0;
            __syncwarp();
            // yh_k = A~ (mx_k o g)  (h_k = 0 for k >= 1);  dref = diag(D_k)
            double dref = 0.0;
            if (act) {
                double acc = 0.0;
                dref = ((mk1 >> i) & 1u) ? 1.0 : 0.0;
                const int r0 = rptr[i], r1 = rsplit[i], r2 = rptr[i + 1];
                for (int e = r0; e < r1; e++) {
                    const int l = colr[e]; const double v = valr[e];
                    acc = fma(v, gbuf[l], acc);
                    if ((mk >> l) & 1u) dref = fma(v, v, dref);
                }
                for (int e = r1; e < r2; e++) {
                    const int l = colr[e] - n; const double v = valr[e];
                    if (sls_bit(muk, l)) { dref = fma(v, v, dref); Ps[(n + l) * LDP + i] = v; }
                }
                yh[(size_t)k * N + i] = acc;
            } else if (i < N) {
                yh[(size_t)k * N + i] = 0.0;
            }
            if (i < L.NP) drefs[i] = dref;
            // stage 1: P = W A~^T  (row i of P; row jj of A~ is read uniformly by all lanes)
            if (act) {
#pragma unroll
                for (int jj = 0; jj < N; jj++) {
                    double acc = 0.0;
                    if (jj < n) {
                        const int r0 = rptr[jj], r1 = rsplit[jj];
                        for (int e = r0; e < r1; e++) acc = fma(Ws[i * LDW + colr[e]], valr[e], acc);
                    }
                    Ps[i * LDP + jj] = acc;
                }
            }
            __syncwarp();
            // stage 2: S_k row i = diag(mx_{k+1}) + sum_{e in row i of C~} val * P[col][:]
#pragma unroll
            for (int p = 0; p < N; p++) s[p] = (p == i && act && ((mk1 >> p) & 1u)) ? 1.0 : 0.0;
            if (act) {
                const int r0 = rptr[i], r2 = rptr[i + 1];
                for (int e = r0; e < r2; e++) {
                    const double v = valr[e];
                    const double2* prow = reinterpret_cast<const double2*>(Ps + (size_t)colr[e] * LDP);
#pragma unroll
                    for (int p2 = 0; p2 < N / 2; p2++) {
                        const double2 pv = prow[p2];
                        s[2 * p2] = fma(v, pv.x, s[2 * p2]);
                        s[2 * p2 + 1] = fma(v, pv.y, s[2 * p2 + 1]);
                    }
                    if (N & 1) s[N - 1] = fma(v, (Ps + (size_t)colr[e] * LDP)[N - 1], s[N - 1]);
                }
            }
            // N symmetric sweeps with rank skipping
            uint32_t keep = 0u;
#pragma unroll
            for (int kk = 0; kk < N; kk++) {
                if (kk < n) {
                    double* cbk = cb + (kk & 1) * L.NP;
                    const double cval = s[kk];
                    if (i < L.NP) cbk[i] = cval;
                    __syncwarp();
                    const double d = cbk[kk];
                    const double dr = drefs[kk];
                    if (dr != 0.0 && d > a.rank_tol * dr) {
                        const double pinv = 1.0 / d;
                        const double f = (i == kk) ? (1.0 - pinv) : cval * pinv;
                        const double2* cb2 = reinterpret_cast<const double2*>(cbk);
                        const double nf = -f;
#pragma unroll
                        for (int p2 = 0; p2 < N / 2; p2++) {
                            const double2 cv = cb2[p2];
                            s[2 * p2] = fma(nf, cv.x, s[2 * p2]);
                            s[2 * p2 + 1] = fma(nf, cv.y, s[2 * p2 + 1]);
                        }
                        if (N & 1) s[N - 1] = fma(nf, cbk[N - 1], s[N - 1]);
                        s[kk] = (i == kk) ? -pinv : cval * pinv;
                        keep |= 1u << kk;
                    }
                }
            }
            {
                const bool ki = (keep >> i) & 1u;
#pragma unroll
                for (int p = 0; p < N; p++) s[p] = (ki && ((keep >> p) & 1u)) ? s[p] : 0.0;
                const unsigned nonempty = __ballot_sync(SLS_FULL, act && dref != 0.0);
                if (keep != nonempty) status |= SLS_ST_RANK_SKIPPED;
            }
            // history: packed lower triangle of -Z_k, column p holds rows p..n-1
            {
                double* hk = hist + (size_t)k * tri;
                int hp = 0;
#pragma unroll
                for (int p = 0; p < N; p++) {
                    if (p < n) { if (act && i >= p) hk[hp + i - p] = s[p]; hp += n - p; }
                }
            }
            __syncwarp();
        }

        // ------------------------------------------------------------------ back substitution, k = T..1
        double lam;
        {
            const double* yT = yh + (size_t)T * N;
            double acc = 0.0;
#pragma unroll
            for (int p = 0; p < N; p++) acc = fma(s[p], yT[p], acc);
            lam = -acc;
        }
        double* xnext = xa;   // x_{k+1}
        double* xcur = xb;    // x_k
        if (i < L.NP) { xnext[i] = 0.0; }
        double resmax = 0.0;
        const int64_t xoff = P.tot[c];
        const uint32_t* pre = P.pre + (size_t)c * 2 * T;
        for (int k = T; k >= 1; k--) {
            const uint32_t mk = mxs[k - 1];
            const uint32_t* muk = mus + (size_t)(k - 1) * wu;
            if (i < L.NP) lam_s[i] = act ? lam : 0.0;
            __syncwarp();
            // t = C~^T lambda_k
            for (int p = lane; p < n + m; p += 32) {
                double acc = 0.0;
                const int e0 = cptr[p], e1 = cptr[p + 1];
                for (int e = e0; e < e1; e++) acc = fma(valc[e], lam_s[rowc[e]], acc);
                tx[p] = acc;
            }
            __syncwarp();
            // u_k = -mu_k o t_u  -> packed result
            {
                int rank0 = 0;
                for (int l0 = 0; l0 < m; l0 += 32) {
                    const int l = l0 + lane;
                    const uint32_t word = muk[l0 >> 5];
                    const bool on = (l < m) && ((word >> lane) & 1u);
                    const double u = on ? -tx[n + l] : 0.0;
                    if (l < m) ucur[l] = u;
                    if (on) a.xw[xoff + pre[2 * (k - 1) + 1] + rank0 + __popc(word & ((1u << lane) - 1u))] = u;
                    rank0 += __popc(word);
                }
            }
            const bool bi = act && ((mk >> i) & 1u);
            if (i < L.NP) rhs_s[i] = act ? (yh[(size_t)(k - 1) * N + i] + (bi ? tx[i] : 0.0)) : 0.0;
            __syncwarp();
            // lambda_{k-1} = Z_{k-1} rhs = -(hist_{k-1} rhs), symmetric packed read
            double acc = 0.0;
            if (act) {
                const double* hk = hist + (size_t)(k - 1) * tri;
                const int hpi = i * n - (i * (i - 1)) / 2;       // start of packed column i
                int hp = 0;
#pragma unroll
                for (int p = 0; p < N; p++) {
                    if (p < n) {
                        const double z = (p <= i) ? hk[hp + i - p] : hk[hpi + p - i];
                        acc = fma(z, rhs_s[p], acc);
                        hp += n - p;
                    }
                }
            }
            const double lam_prev = -acc;
            const double x = bi ? (lam_prev - tx[i]) : 0.0;
            if (i < L.NP) xcur[i] = x;
            if (bi) a.xw[xoff + pre[2 * (k - 1)] + __popc(mk & ((1u << i) - 1u))] = x;
            __syncwarp();
            // residual of constraint block k: A~ x_k + B~ u_k - x_{k+1}
            if (act) {
                double r = -xnext[i];
                const int r0 = rptr[i], r1 = rsplit[i], r2 = rptr[i + 1];
                for (int e = r0; e < r1; e++) r = fma(valr[e], xcur[colr[e]], r);
                for (int e = r1; e < r2; e++) r = fma(valr[e], ucur[colr[e] - n], r);
                resmax = fmax(resmax, fabs(r));
            }
            __syncwarp();
            double* tswap = xnext; xnext = xcur; xcur = tswap;
            lam = lam_prev;
        }
        // block 0: x_1 = e
        if (act) resmax = fmax(resmax, fabs(((i == jl) ? 1.0 : 0.0) - xnext[i]));
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) resmax = fmax(resmax, __shfl_xor_sync(SLS_FULL, resmax, d));
        if (!(resmax <= a.feas_tol)) status |= SLS_ST_RESIDUAL;
        if (lane == 0) { a.status[c] = status; a.resid[c] = resmax; }
    }
}

// ---- host side ------------------------------------------------------------------------------
static const int kClasses[] = {4, 8, 12, 16, 20, 21, 22, 24, 28, 32};
int sls_solve_class_count() { return (int)(sizeof(kClasses) / sizeof(int)); }
int sls_solve_class_N(int k) { return kClasses[k]; }
int sls_solve_class_of(int n) {
    for (int k = 0; k < sls_solve_class_count(); k++) if (kClasses[k] >= n) return kClasses[k];
    return -1;
}
size_t sls_solve_smem_per_warp(int N, int mcap, int nnzcap, int T, int wu) {
    return (size_t)SlsSolveLayout(N, mcap, nnzcap, T, wu).total;
}
int64_t sls_solve_hist_doubles(int N, int T) { return (int64_t)(T + 1) * (N * (N + 1) / 2); }

int sls_solve_max_resident_warps(int N, int mcap, int nnzcap, int T, int wu, int sm_count, int max_smem) {
    const size_t per_block = sls_solve_smem_per_warp(N, mcap, nnzcap, T, wu) * SOLVE_WPB;
    if (per_block > (size_t)max_smem) return -1;
    int bps = (int)((size_t)(max_smem) / (per_block + 1024));
    if (bps < 1) bps = 1;
    if (bps > 4) bps = 4;                       // register budget: 4 blocks x 128 threads
    return bps * sm_count * SOLVE_WPB;
}

template <int N>
static int launch_class(const SlsSolveArgs& a, int sm_count, int max_smem, int* warps_launched, sls_stream_t stream) {
    const size_t per_block = sls_solve_smem_per_warp(N, a.mcap, a.nnzcap, a.P.T, a.P.wu) * SOLVE_WPB;
    if (per_block > (size_t)max_smem) return -5;
    auto kp = &sls_solve_kernel<N>;
    if (per_block > 48 * 1024) { int e = rt::set_max_smem(kp, (int)per_block); if (e) return e; }
    const int maxw = sls_solve_max_resident_warps(N, a.mcap, a.nnzcap, a.P.T, a.P.wu, sm_count, max_smem);
    int blocks = (a.count + SOLVE_WPB - 1) / SOLVE_WPB;
    if (blocks > maxw / SOLVE_WPB) blocks = maxw / SOLVE_WPB;
    if (blocks < 1) blocks = 1;
    *warps_launched = blocks * SOLVE_WPB;
    SLS_LAUNCH(kp, blocks, SOLVE_WPB * 32, per_block, stream, a);
    return rt::last_error();
}

int sls_launch_solve(int N, const SlsSolveArgs& a, int sm_count, int max_smem, int* warps_launched,
                     sls_stream_t stream) {
    *warps_launched = 0;
    if (a.count == 0) return 0;
    switch (N) {
        case 4:  return launch_class<4>(a, sm_count, max_smem, warps_launched, stream);
        case 8:  return launch_class<8>(a, sm_count, max_smem, warps_launched, stream);
        case 12: return launch_class<12>(a, sm_count, max_smem, warps_launched, stream);
        case 16: return launch_class<16>(a, sm_count, max_smem, warps_launched, stream);
        case 20: return launch_class<20>(a, sm_count, max_smem, warps_launched, stream);
        case 21: return launch_class<21>(a, sm_count, max_smem, warps_launched, stream);
        case 22: return launch_class<22>(a, sm_count, max_smem, warps_launched, stream);
        case 24: return launch_class<24>(a, sm_count, max_smem, warps_launched, stream);
        case 28: return launch_class<28>(a, sm_count, max_smem, warps_launched, stream);
        case 32: return launch_class<32>(a, sm_count, max_smem, warps_launched, stream);
        default: return -6;
    }
}
