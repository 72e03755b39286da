// sls_solve_general.cu — kernel 2g: the reference's per-group problem in full generality: general
// cost matrices [C1 D12], D11, a non-diagonal B1 and column groups c_j with q >= 1 columns that couple through the cost
// or through the union index set.  One CTA per group, FP64 SIMT arithmetic on dense shared-memory matrices.
//
// Replaces, per group c_j (reference src/synthesis.jl:40-62, 76-77; src/reduction.jl:14-23):
//     min  sum_t || L [X_t; U_t] R + D ||_F^2          X_t (n x q), U_t (m x q), entries outside 𝓢[t][s,c_j] fixed to 0
//     s.t. X_1 = I~,   X_{t+1} = A~ X_t + B2~ U_t,   0 = A~ X_T + B2~ U_T
// with s_x / s_u the union of the member columns' index sets (reduction.jl:14), L = [C1~ D12~] = rows [s_x; Nx+s_u] of
// [C1[:,s_x] D12[:,s_u]] (reduction.jl:15, GeneralizedPlant.jl:266-271), R = B1[c_j, c_j] (synthesis.jl:42; groups are
// ascending), D = D11[[s_x; Nx+s_u], c_j].
//
// With z_t = vec([X_t; U_t]) (length pb = q(n+m)), Mb = (R R^T) (x) (L^T L), gb = vec(L^T D R^T) the problem is
//     min  sum_t 1/2 z_t^T Mb z_t + gb^T z_t     s.t.  E z_1 = e,  E z_{k+1} - Fb z_k = 0 (k < T),  -Fb z_T = 0
// (Fb = I_q (x) [A~ B2~], E picks the state entries).  W_t = inverse of Mb on the free entries of tap t; the multipliers
// solve the block-tridiagonal system S mu = h + G W g,
//     S_kk = Fb W_k Fb^T + E W_{k+1} E^T,      S_{k,k-1} = C_k = -Fb W_k E^T,
// by block elimination (Z_k = pseudo-inverse of the Schur complement by symmetric sweeps that skip the redundant rows:
// pivot <= rank_tol * diag), and z_t = W_t (E^T mu_{t-1} - Fb^T mu_t - gb).  Z_k, C_k and W_t stay in a per-CTA scratch
// slot (L2) so that the refinement sweeps on the residual h - G z reuse them.  The identity-cost single-column case of
// this recursion is what sls_solve.cu does on DMMA; this kernel is the general, slower path and is only launched when the
// caller installed a cost (sls_set_cost) or column groups (sls_set_groups).
#include "sls_common.h"
#include "sls_kernels.h"
#include <cmath>

#ifndef GEN_NT
#define GEN_NT 256
#endif
#define SLS_INF_D ((double)INFINITY)

struct GenLayout {
    int pc, Nc, nb, pb, qc, T, cap;      // caps: pc = max q(n+m), Nc = max q n, nb = max n, pb = max n+m, qc = max q
    size_t oF, oW, oFW, oS, oC, oZ, oT, oVec, oD, oR, oInt, oFree, total;
    __host__ __device__ GenLayout(int pc_, int Nc_, int nb_, int pb_, int qc_, int T_, int cap_) {
        pc = pc_ < 1 ? 1 : pc_; Nc = Nc_ < 1 ? 1 : Nc_; nb = nb_ < 1 ? 1 : nb_; pb = pb_ < 1 ? 1 : pb_; qc = qc_ < 1 ? 1 : qc_; T = T_; cap = cap_;
        size_t o = 0;
#define GEN_SEC(name, bytes) name = o; o += (size_t)(bytes); o = (o + 15) & ~(size_t)15
        GEN_SEC(oF, 8 * nb * pb);                 // [A~ B2~]
        size_t wbytes = (size_t)8 * pc * pc;       // W_t; during the set-up: L (pb x pb) and the sort buffer
        if (wbytes < (size_t)8 * pb * pb) wbytes = (size_t)8 * pb * pb;
        if (wbytes < (size_t)4 * cap) wbytes = (size_t)4 * cap;
        GEN_SEC(oW, wbytes);
        GEN_SEC(oFW, 8 * Nc * pc);
        GEN_SEC(oS, 8 * Nc * Nc);
        GEN_SEC(oC, 8 * Nc * Nc);
        GEN_SEC(oZ, 8 * Nc * Nc);
        GEN_SEC(oT, 8 * Nc * Nc);
        GEN_SEC(oVec, 8 * (6 * Nc + 5 * pc + 16));     // thr, colbuf(max(Nc,pc)) ... see the kernel
        GEN_SEC(oD, 8 * pb * qc);
        GEN_SEC(oR, 8 * 2 * qc * qc);
        GEN_SEC(oInt, 4 * (nb + pb + 4 * qc + 16));
        GEN_SEC(oFree, (size_t)(T + 1) * pc);
#undef GEN_SEC
        total = (o + 15) & ~(size_t)15;
    }
};

// doubles of one CTA's scratch slot
__host__ __device__ inline int64_t gen_scratch_doubles(int pc, int Nc, int T) {
    return (int64_t)pc * pc                       // Mb
         + (int64_t)(T + 1) * Nc * Nc             // -Z_k
         + (int64_t)(T + 1) * Nc * Nc             // C_k (k = 1..T)
         + (int64_t)(T + 1) * pc * pc             // W_t (t = 1..T)
         + (int64_t)3 * (T + 1) * Nc              // r, y, mu
         + (int64_t)2 * (T + 1) * pc;             // z, dz
}

// sorted union of the members' index sets (kernel 1's per-column lists); one warp.  Returns the size, -1 on overflow.
__device__ int gen_union(const int32_t* lists, int stride, const int32_t* sizes, int k0, int k1, int32_t* buf, int cap) {
    const int lane = sls_lane();
    int C = 0;
#pragma unroll 1
    for (int k = k0; k < k1; k++) {
        const int len = sizes[k];
        if (len < 0 || C + len > cap) return -1;
#pragma unroll 1
        for (int i = lane; i < len; i += 32) buf[C + i] = lists[(size_t)k * stride + i];
        C += len;
    }
    __syncwarp();
    return sls_warp_sort_unique(buf, C, cap);
}

// group sizes: one warp per group; stats[0..4] = max q, max n, max n+m, max q n, max q(n+m); stats[5] = overflow
__global__ void __launch_bounds__(128)
sls_general_size_kernel(SlsPlanView P, const int32_t* __restrict__ grp_ptr, int ngroups, int cap, int32_t* stats) {
    SLS_DYN_SMEM(smem);
    const int warp = threadIdx.x >> 5, lane = sls_lane();
    const int g = blockIdx.x * 4 + warp;
    if (g >= ngroups) return;
    int32_t* buf = reinterpret_cast<int32_t*>(smem) + (size_t)warp * cap;
    const int k0 = grp_ptr ? grp_ptr[g] : g, k1 = grp_ptr ? grp_ptr[g + 1] : g + 1;
    const int n = gen_union(P.sx, P.nstride, P.n, k0, k1, buf, cap);
    __syncwarp();
    const int m = gen_union(P.su, P.mstride, P.m, k0, k1, buf, cap);
    if (lane == 0) {
        if (n < 0 || m < 0) { atomicOr(stats + 5, 1); return; }
        const int q = k1 - k0;
        atomicMax(stats + 0, q); atomicMax(stats + 1, n); atomicMax(stats + 2, n + m);
        atomicMax(stats + 3, q * n); atomicMax(stats + 4, q * (n + m));
    }
}

// In-place symmetric sweep of the d x d matrix S (leading dimension ld): on return S = -(S_KK)^-1 embedded, K = the
// pivots that exceeded thr[p]; the other rows / columns are zero.  Returns (in *nskip) the number of skipped pivots with
// want[p] != 0 (the caller's "this one should have been regular").  colb: d doubles.  All threads of the CTA.
__device__ void gen_sweep(double* S, int d, int ld, const double* thr, const unsigned char* want, double* colb, int* nskip) {
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int p = 0; p < d; p++) {
        __syncthreads();
        const double piv = S[(size_t)p * ld + p];
        const bool skip = !(piv > thr[p]);
#pragma unroll 1
        for (int i = tid; i < d; i += GEN_NT) colb[i] = S[(size_t)i * ld + p];
        __syncthreads();
        if (skip) {
#pragma unroll 1
            for (int i = tid; i < d; i += GEN_NT) { S[(size_t)i * ld + p] = 0.0; S[(size_t)p * ld + i] = 0.0; }
            if (tid == 0 && want && want[p]) (*nskip)++;
            continue;
        }
        const double dinv = 1.0 / piv;
#pragma unroll 1
        for (int e = tid; e < d * d; e += GEN_NT) {
            const int i = e / d, j = e - i * d;
            double v;
            if (i == p && j == p) v = -dinv;
            else if (i == p) v = colb[j] * dinv;
            else if (j == p) v = colb[i] * dinv;
            else v = S[(size_t)i * ld + j] - colb[i] * colb[j] * dinv;
            S[(size_t)i * ld + j] = v;
        }
    }
    __syncthreads();
}

struct GenCtx {
    int n, m, p, q, Nb, pbar, T;
    const double* F;            // n x p
    const double* gb;           // pbar
    double* scratch;
    int64_t oZ, oC, oW, oR, oY, oMu, oZv, oDz;
    int Nc, pc;
};

// y[r] = alpha * sum_j M[r*ld + j] x[j] (+ add[r]) for r < rows: one warp per row, lanes across the row (coalesced), the
// partial sums combined by shuffles in a fixed order.  M may live in the scratch slot (L2) or in shared memory.
__device__ __forceinline__ void gen_matvec(const double* M, int ld, int rows, int cols, const double* x, double* y,
                                           double alpha, const double* add) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll 1
    for (int r = warp; r < rows; r += GEN_NT / 32) {
        double s = 0.0;
#pragma unroll 1
        for (int j = lane; j < cols; j += 32) s += M[(size_t)r * ld + j] * x[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(SLS_FULL, s, o);
        if (lane == 0) y[r] = alpha * s + (add ? add[r] : 0.0);
    }
}

// mu <- solution of S mu = r (r in scratch[oR], T+1 blocks of Nb): forward sweep with the stored C_k, -Z_k, then back
// substitution.  y, tmp: Nb doubles of shared memory each.
__device__ void gen_solve_mu(const GenCtx& c, double* y, double* tmp) {
    const int tid = threadIdx.x, Nb = c.Nb, T = c.T;
    double* r = c.scratch + c.oR; double* mu = c.scratch + c.oMu; double* rt = c.scratch + c.oY;
    // forward: rt_0 = r_0, rt_k = r_k + C_k (Zb_{k-1} rt_{k-1})   (Zb = -Z)
#pragma unroll 1
    for (int i = tid; i < Nb; i += GEN_NT) rt[i] = r[i];
    __syncthreads();
#pragma unroll 1
    for (int k = 1; k <= T; k++) {
        const double* Zb = c.scratch + c.oZ + (int64_t)(k - 1) * c.Nc * c.Nc;
        const double* Ck = c.scratch + c.oC + (int64_t)k * c.Nc * c.Nc;
        gen_matvec(Zb, c.Nc, Nb, Nb, rt + (size_t)(k - 1) * c.Nc, y, 1.0, nullptr);
        __syncthreads();
        gen_matvec(Ck, c.Nc, Nb, Nb, y, rt + (size_t)k * c.Nc, 1.0, r + (size_t)k * c.Nc);
        __syncthreads();
    }
    // back: mu_T = -Zb_T rt_T;  mu_k = -Zb_k (rt_k - C_{k+1}^T mu_{k+1})
#pragma unroll 1
    for (int k = T; k >= 0; k--) {
        const double* Zb = c.scratch + c.oZ + (int64_t)k * c.Nc * c.Nc;
#pragma unroll 1
        for (int i = tid; i < Nb; i += GEN_NT) {
            double s = rt[(size_t)k * c.Nc + i];
            if (k < T) {
                const double* Cn = c.scratch + c.oC + (int64_t)(k + 1) * c.Nc * c.Nc;
                const double* mn = mu + (size_t)(k + 1) * c.Nc;
#pragma unroll 4
                for (int j = 0; j < Nb; j++) s -= Cn[(size_t)j * c.Nc + i] * mn[j];     // column i of C: coalesced across threads
            }
            tmp[i] = s;
        }
        __syncthreads();
        gen_matvec(Zb, c.Nc, Nb, Nb, tmp, mu + (size_t)k * c.Nc, -1.0, nullptr);
        __syncthreads();
    }
}

// dst_t = W_t (E^T mu_{t-1} - Fb^T mu_t - (with_g ? gb : 0)) for t = 1..T; avec: pbar doubles of shared memory
__device__ void gen_primal(const GenCtx& c, double* dst, bool with_g, double* avec) {
    const int tid = threadIdx.x, n = c.n, p = c.p, pbar = c.pbar, T = c.T;
    const double* mu = c.scratch + c.oMu;
#pragma unroll 1
    for (int t = 1; t <= T; t++) {
        const double* m0 = mu + (size_t)(t - 1) * c.Nc; const double* m1 = mu + (size_t)t * c.Nc;
#pragma unroll 1
        for (int v = tid; v < pbar; v += GEN_NT) {
            const int cc = v / p, a = v - cc * p;
            double s = (a < n) ? m0[cc * n + a] : 0.0;
#pragma unroll 4
            for (int i = 0; i < n; i++) s -= c.F[(size_t)i * p + a] * m1[cc * n + i];
            if (with_g) s -= c.gb[v];
            avec[v] = s;
        }
        __syncthreads();
        gen_matvec(c.scratch + c.oW + (int64_t)t * c.pc * c.pc, c.pc, pbar, pbar, avec, dst + (size_t)t * c.pc, 1.0, nullptr);
        __syncthreads();
    }
}

// r_k = h_k - (G z)_k into scratch[oR]; returns max |r| (the same value in every thread).  red: GEN_NT/32 doubles of smem
__device__ double gen_residual(const GenCtx& c, const double* z, const double* e0, double* red) {
    const int tid = threadIdx.x, n = c.n, p = c.p, Nb = c.Nb, T = c.T;
    double* r = c.scratch + c.oR;
    double worst = 0.0;
#pragma unroll 1
    for (int k = 0; k <= T; k++) {
#pragma unroll 1
        for (int v = tid; v < Nb; v += GEN_NT) {
            const int cc = v / n, i = v - cc * n;
            double s = (k == 0) ? e0[v] : 0.0;
            if (k < T) s -= z[(size_t)(k + 1) * c.pc + cc * p + i];
            if (k >= 1) {
                const double* zk = z + (size_t)k * c.pc + cc * p;
                for (int a = 0; a < p; a++) s += c.F[(size_t)i * p + a] * zk[a];
            }
            r[(size_t)k * c.Nc + v] = s;
            const double as = fabs(s);
            if (as > worst || as != as) worst = as;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const double w = __shfl_xor_sync(SLS_FULL, worst, o); if (w > worst || w != w) worst = w; }
    __syncthreads();
    if ((tid & 31) == 0) red[tid >> 5] = worst;
    __syncthreads();
    double all = 0.0;
    for (int w = 0; w < GEN_NT / 32; w++) { const double x = red[w]; if (x > all || x != x) all = x; }
    __syncthreads();
    return all;
}

#ifndef GEN_MINB
#define GEN_MINB 4
#endif
__global__ void __launch_bounds__(GEN_NT, GEN_MINB)
sls_solve_general_kernel(SlsGeneralArgs a) {
    SLS_DYN_SMEM(smem);
    const int tid = threadIdx.x, warp = tid >> 5;
    const SlsPlanView& P = a.P;
    const int T = P.T;
    const GenLayout L(a.pc, a.Nc, a.nb, a.pb, a.qc, T, a.cap);
    double* F = reinterpret_cast<double*>(smem + L.oF);
    double* Wb = reinterpret_cast<double*>(smem + L.oW);
    double* FW = reinterpret_cast<double*>(smem + L.oFW);
    double* Sb = reinterpret_cast<double*>(smem + L.oS);
    double* Cb = reinterpret_cast<double*>(smem + L.oC);
    double* Zb = reinterpret_cast<double*>(smem + L.oZ);
    double* Tb = reinterpret_cast<double*>(smem + L.oT);
    double* vec = reinterpret_cast<double*>(smem + L.oVec);
    const int Nc = L.Nc, pc = L.pc;
    const int vmax = Nc > pc ? Nc : pc;
    double* thr = vec;                      // vmax
    double* colb = thr + vmax;              // vmax
    double* gb = colb + vmax;               // pc
    double* avec = gb + pc;                 // pc
    double* e0 = avec + pc;                 // Nc
    double* yv = e0 + Nc;                   // Nc
    double* tv = yv + Nc;                   // Nc
    double* red = tv + Nc;                  // 8
    double* Dq = reinterpret_cast<double*>(smem + L.oD);
    double* Rm = reinterpret_cast<double*>(smem + L.oR);       // R (q x q), then R R^T
    int32_t* ib = reinterpret_cast<int32_t*>(smem + L.oInt);
    int32_t* sx = ib;                       // nb
    int32_t* su = sx + L.nb;                // pb (>= m)
    int32_t* jcol = su + L.pb;              // qc: global column of member c
    int32_t* jloc = jcol + L.qc;            // qc
    int32_t* kcol = jloc + L.qc;            // qc: planned-column index
    int32_t* misc = kcol + L.qc + L.qc;     // [0] n, [1] m, [2] skipped-but-free count of W, [3] skipped pivots of S, [4] W changed, [5] pattern mismatch
    unsigned char* freeb = smem + L.oFree;  // [T+1][pc]: row t = free entries of tap t (1-based), row 0 unused
    double* scratch = a.scratch + (size_t)blockIdx.x * a.scratch_stride;
    const int64_t oMb = 0;
    const int64_t oZ = oMb + (int64_t)pc * pc;
    const int64_t oC = oZ + (int64_t)(T + 1) * Nc * Nc;
    const int64_t oWh = oC + (int64_t)(T + 1) * Nc * Nc;
    const int64_t oR = oWh + (int64_t)(T + 1) * pc * pc;
    const int64_t oY = oR + (int64_t)(T + 1) * Nc;
    const int64_t oMu = oY + (int64_t)(T + 1) * Nc;
    const int64_t oZv = oMu + (int64_t)(T + 1) * Nc;
    const int64_t oDz = oZv + (int64_t)(T + 1) * pc;

#pragma unroll 1
    for (int g = blockIdx.x; g < a.ngroups; g += gridDim.x) {
        __syncthreads();
        const int k0 = a.grp_ptr ? a.grp_ptr[g] : g, k1 = a.grp_ptr ? a.grp_ptr[g + 1] : g + 1;
        const int q = k1 - k0;
        // ------------------------------------------------------------------ union index sets (reduction.jl:14)
        if (warp == 0) {
            int32_t* buf = reinterpret_cast<int32_t*>(Wb);
            const int n_ = gen_union(P.sx, P.nstride, P.n, k0, k1, buf, a.cap);
            for (int i = sls_lane(); i < n_; i += 32) sx[i] = buf[i];
            __syncwarp();
            const int m_ = gen_union(P.su, P.mstride, P.m, k0, k1, buf, a.cap);
            for (int i = sls_lane(); i < m_; i += 32) su[i] = buf[i];
            if (sls_lane() == 0) { misc[0] = n_; misc[1] = m_; misc[2] = 0; misc[3] = 0; misc[5] = 0; }
        }
        __syncthreads();
        const int n = misc[0], m = misc[1];
        const int p = n + m, Nb = q * n, pbar = q * p;
        if (tid < q) {
            const int kc = k0 + tid;
            kcol[tid] = kc; jcol[tid] = P.cols[kc]; jloc[tid] = sls_find(sx, n, P.cols[kc]);
        }
        // ------------------------------------------------------------------ F = [A~ B2~] dense (n x p)
#pragma unroll 1
        for (int e = tid; e < n * p; e += GEN_NT) F[e] = 0.0;
#pragma unroll 1
        for (int e = tid; e < (T + 1) * pc; e += GEN_NT) freeb[e] = 0;
        __syncthreads();
#pragma unroll 1
        for (int col = tid; col < p; col += GEN_NT) {
            const bool isA = col < n;
            const SlsCsc& Mx = isA ? a.A : a.B2;
            const int64_t gc = isA ? sx[col] : su[col - n];
            const int64_t e0_ = Mx.colptr[gc] - 1, e1_ = Mx.colptr[gc + 1] - 1;
            for (int64_t e = e0_; e < e1_; e++) {
                const int i = sls_find(sx, n, (int)(Mx.rowval[e] - 1));
                if (i >= 0) F[(size_t)i * p + col] = Mx.nzval[e];
            }
        }
        // ------------------------------------------------------------------ free-entry masks 𝓢[t][s,c_j] (synthesis.jl:57-60)
#pragma unroll 1
        for (int w = tid; w < 2 * T * q; w += GEN_NT) {
            const int cc = w / (2 * T), tt = w - cc * 2 * T;
            const bool isx = tt < T; const int t = isx ? tt : tt - T;
            const SlsSupports& S = isx ? a.Sx : a.Su;
            const int64_t* scp = S.colptr + (int64_t)t * (P.Nx + 1);
            const int64_t base = S.tap_off[t];
            const int64_t j = jcol[cc];
            const int64_t s0 = base + scp[j] - 1, s1 = base + scp[j + 1] - 1;
            int cnt = 0;
            for (int64_t e = s0; e < s1; e++) {
                if (S.nzval && !S.nzval[e]) continue;
                const int r = (int)(S.rowval[e] - 1 - S.row_off);
                const int i = isx ? sls_find(sx, n, r) : sls_find(su, m, r);
                if (i >= 0) { freeb[(size_t)(t + 1) * pc + cc * p + (isx ? i : n + i)] = 1; cnt++; }
            }
            // the planned column's own pattern (kernel 1b / kernel 3) must describe the same entries
            const int64_t* ocp = P.colptr + (int64_t)tt * (P.Nx + 1);
            if ((int64_t)cnt != ocp[j + 1] - ocp[j]) misc[5] = 1;
        }
        // ------------------------------------------------------------------ cost blocks: L (p x p) into Wb, R, D
#pragma unroll 1
        for (int e = tid; e < p * p; e += GEN_NT) Wb[e] = 0.0;
#pragma unroll 1
        for (int e = tid; e < p * q; e += GEN_NT) Dq[e] = 0.0;
#pragma unroll 1
        for (int e = tid; e < 2 * q * q; e += GEN_NT) Rm[e] = 0.0;
        __syncthreads();
        if (a.has_cost) {
#pragma unroll 1
            for (int col = tid; col < p; col += GEN_NT) {
                const bool isx = col < n;
                const SlsCsc& Mx = isx ? a.C1 : a.D12;
                const int64_t gc = isx ? sx[col] : su[col - n];
                const int64_t e0_ = Mx.colptr[gc] - 1, e1_ = Mx.colptr[gc + 1] - 1;
                for (int64_t e = e0_; e < e1_; e++) {
                    const int64_t r = Mx.rowval[e] - 1;
                    int row = -1;
                    if (r < P.Nx) row = sls_find(sx, n, (int)r);
                    else { const int l = sls_find(su, m, (int)(r - P.Nx)); if (l >= 0) row = n + l; }
                    if (row >= 0) Wb[(size_t)row * p + col] = Mx.nzval[e];
                }
            }
        } else {
#pragma unroll 1
            for (int col = tid; col < p; col += GEN_NT) Wb[(size_t)col * p + col] = 1.0;
        }
        if (a.has_d11) {
#pragma unroll 1
            for (int cc = tid; cc < q; cc += GEN_NT) {
                const int64_t gc = jcol[cc];
                const int64_t e0_ = a.D11.colptr[gc] - 1, e1_ = a.D11.colptr[gc + 1] - 1;
                for (int64_t e = e0_; e < e1_; e++) {
                    const int64_t r = a.D11.rowval[e] - 1;
                    int row = -1;
                    if (r < P.Nx) row = sls_find(sx, n, (int)r);
                    else { const int l = sls_find(su, m, (int)(r - P.Nx)); if (l >= 0) row = n + l; }
                    if (row >= 0) Dq[(size_t)row * q + cc] = a.D11.nzval[e];
                }
            }
        }
        if (a.has_b1) {                                        // R[c][c'] = B1[j_c, j_c']
#pragma unroll 1
            for (int cc = tid; cc < q; cc += GEN_NT) {
                const int64_t gc = jcol[cc];
                const int64_t e0_ = a.B1.colptr[gc] - 1, e1_ = a.B1.colptr[gc + 1] - 1;
                for (int64_t e = e0_; e < e1_; e++) {
                    const int r = (int)(a.B1.rowval[e] - 1);
                    for (int c2 = 0; c2 < q; c2++) if (jcol[c2] == r) Rm[c2 * q + cc] = a.B1.nzval[e];
                }
            }
        } else if (tid < q) Rm[tid * q + tid] = 1.0;
        __syncthreads();
        // gb = vec(L^T D R^T):  G1[a][k] = sum_r L[r][a] D[r][k],  gb[(c,a)] = sum_k G1[a][k] R[c][k]
#pragma unroll 1
        for (int v = tid; v < pbar; v += GEN_NT) {
            const int cc = v / p, aa = v - cc * p;
            double s = 0.0;
            if (a.has_d11)
                for (int k = 0; k < q; k++) {
                    const double rk = Rm[cc * q + k];
                    if (rk == 0.0) continue;
                    double g1 = 0.0;
                    for (int r = 0; r < p; r++) g1 += Wb[(size_t)r * p + aa] * Dq[(size_t)r * q + k];
                    s += g1 * rk;
                }
            gb[v] = s;
        }
        // R R^T behind R
#pragma unroll 1
        for (int e = tid; e < q * q; e += GEN_NT) {
            const int c1 = e / q, c2 = e - c1 * q;
            double s = 0.0;
            for (int k = 0; k < q; k++) s += Rm[c1 * q + k] * Rm[c2 * q + k];
            Rm[q * q + e] = s;
        }
        __syncthreads();
        // Mb = (R R^T) (x) (L^T L) -> scratch
#pragma unroll 1
        for (int e = tid; e < p * p; e += GEN_NT) {
            const int aa = e / p, bb = e - aa * p;
            double s = 0.0;
            for (int r = 0; r < p; r++) s += Wb[(size_t)r * p + aa] * Wb[(size_t)r * p + bb];
            for (int c1 = 0; c1 < q; c1++)
                for (int c2 = 0; c2 < q; c2++)
                    scratch[oMb + (int64_t)(c1 * p + aa) * pc + (c2 * p + bb)] = Rm[q * q + c1 * q + c2] * s;
        }
        // impulses e (block row 0): column c starts at its own state (reduction.jl:22-23)
#pragma unroll 1
        for (int v = tid; v < Nb; v += GEN_NT) { const int cc = v / n, i = v - cc * n; e0[v] = (jloc[cc] == i) ? 1.0 : 0.0; }
        __syncthreads();
        int st = 0;
        for (int cc = 0; cc < q; cc++) if (jloc[cc] < 0) st |= SLS_ST_NO_IMPULSE;
        bool has_g = false;
        if (a.has_d11) for (int v = 0; v < pbar; v++) if (gb[v] != 0.0) { has_g = true; break; }

        GenCtx cx;
        cx.n = n; cx.m = m; cx.p = p; cx.q = q; cx.Nb = Nb; cx.pbar = pbar; cx.T = T; cx.F = F; cx.gb = gb; cx.scratch = scratch;
        cx.oZ = oZ; cx.oC = oC; cx.oW = oWh; cx.oR = oR; cx.oY = oY; cx.oMu = oMu; cx.oZv = oZv; cx.oDz = oDz; cx.Nc = Nc; cx.pc = pc;

        // ------------------------------------------------------------------ matrix pass: W_t, S, C, Z
        // A_prev (in Sb) = Fb W_{t-1} Fb^T (zero for t = 1);  Cb = C_{t-1};  Zb = -Z_{t-2}
#pragma unroll 1
        for (int e = tid; e < Nb * Nb; e += GEN_NT) { Sb[(size_t)(e / Nb) * Nc + (e % Nb)] = 0.0; }
        __syncthreads();
#pragma unroll 1
        for (int t = 1; t <= T + 1; t++) {
            if (t <= T) {
                // W_t: reuse W_{t-1} when the free set did not change
                if (tid == 0) misc[4] = (t == 1);
                __syncthreads();
                if (t > 1) {
                    bool diff = false;
#pragma unroll 1
                    for (int v = tid; v < pbar; v += GEN_NT) diff = diff || (freeb[(size_t)t * pc + v] != freeb[(size_t)(t - 1) * pc + v]);
                    if (diff) misc[4] = 1;
                    __syncthreads();
                }
                if (misc[4]) {
                    const unsigned char* fr = freeb + (size_t)t * pc;
#pragma unroll 1
                    for (int e = tid; e < pbar * pbar; e += GEN_NT) {
                        const int v = e / pbar, w = e - v * pbar;
                        Wb[(size_t)v * pc + w] = (fr[v] && fr[w]) ? scratch[oMb + (int64_t)v * pc + w] : 0.0;
                    }
                    __syncthreads();
#pragma unroll 1
                    for (int v = tid; v < pbar; v += GEN_NT) thr[v] = fr[v] ? 1e-13 * fabs(Wb[(size_t)v * pc + v]) : SLS_INF_D;
                    __syncthreads();
                    gen_sweep(Wb, pbar, pc, thr, fr, colb, misc + 2);
#pragma unroll 1
                    for (int e = tid; e < pbar * pbar; e += GEN_NT) { const int v = e / pbar, w = e - v * pbar; Wb[(size_t)v * pc + w] = -Wb[(size_t)v * pc + w]; }
                    __syncthreads();
                }
                {
                    double* Wh = scratch + oWh + (int64_t)t * pc * pc;
#pragma unroll 1
                    for (int e = tid; e < pbar * pbar; e += GEN_NT) { const int v = e / pbar, w = e - v * pbar; Wh[(size_t)v * pc + w] = Wb[(size_t)v * pc + w]; }
                }
                // S_{t-1,t-1} = A_{t-1} + E W_t E^T
#pragma unroll 1
                for (int e = tid; e < Nb * Nb; e += GEN_NT) {
                    const int r = e / Nb, c2 = e - r * Nb;
                    const int c1c = r / n, i1 = r - c1c * n, c2c = c2 / n, i2 = c2 - c2c * n;
                    Sb[(size_t)r * Nc + c2] += Wb[(size_t)(c1c * p + i1) * pc + (c2c * p + i2)];
                }
                __syncthreads();
            }
            // eliminate block row k = t-1:  S~ = S_kk + (C_k Zb_{k-1}) C_k^T,  Zb_k = sweep(S~)
            const int k = t - 1;
#pragma unroll 1
            for (int i = tid; i < Nb; i += GEN_NT) { const double d = Sb[(size_t)i * Nc + i]; thr[i] = (d > 0.0) ? a.rank_tol * d : SLS_INF_D; }
            __syncthreads();
            if (k >= 1) {
#pragma unroll 1
                for (int e = tid; e < Nb * Nb; e += GEN_NT) {
                    const int r = e / Nb, c2 = e - r * Nb;
                    double s = 0.0;
                    for (int j = 0; j < Nb; j++) s += Cb[(size_t)r * Nc + j] * Zb[(size_t)j * Nc + c2];
                    Tb[(size_t)r * Nc + c2] = s;
                }
                __syncthreads();
#pragma unroll 1
                for (int e = tid; e < Nb * Nb; e += GEN_NT) {
                    const int r = e / Nb, c2 = e - r * Nb;
                    double s = Sb[(size_t)r * Nc + c2];
                    for (int j = 0; j < Nb; j++) s += Tb[(size_t)r * Nc + j] * Cb[(size_t)c2 * Nc + j];
                    Zb[(size_t)r * Nc + c2] = s;     // Zb_{k-1} is dead after Tb
                }
            } else {
#pragma unroll 1
                for (int e = tid; e < Nb * Nb; e += GEN_NT) { const int r = e / Nb, c2 = e - r * Nb; Zb[(size_t)r * Nc + c2] = Sb[(size_t)r * Nc + c2]; }
            }
            __syncthreads();
            gen_sweep(Zb, Nb, Nc, thr, nullptr, colb, misc + 3);
            {
                // redundant rows: a pivot that stayed at or below its threshold zeroed its row (diag 0 after the sweep while thr finite)
                double* Zh = scratch + oZ + (int64_t)k * Nc * Nc;
#pragma unroll 1
                for (int e = tid; e < Nb * Nb; e += GEN_NT) { const int r = e / Nb, c2 = e - r * Nb; Zh[(size_t)r * Nc + c2] = Zb[(size_t)r * Nc + c2]; }
                if (tid == 0) for (int i = 0; i < Nb; i++) if (Zb[(size_t)i * Nc + i] == 0.0 && thr[i] != SLS_INF_D) misc[3]++;
            }
            if (t <= T) {
                // FW = Fb W_t;  A_t = FW Fb^T -> Sb;  C_t = -FW E^T -> Cb
#pragma unroll 1
                for (int e = tid; e < Nb * pbar; e += GEN_NT) {
                    const int r = e / pbar, w = e - r * pbar;
                    const int cc = r / n, i = r - cc * n;
                    double s = 0.0;
                    const double* Fr = F + (size_t)i * p;
                    const double* Wc = Wb + (size_t)(cc * p) * pc + w;
                    for (int aa = 0; aa < p; aa++) s += Fr[aa] * Wc[(size_t)aa * pc];
                    FW[(size_t)r * pc + w] = s;
                }
                __syncthreads();
                double* Ch = scratch + oC + (int64_t)t * Nc * Nc;
#pragma unroll 1
                for (int e = tid; e < Nb * Nb; e += GEN_NT) {
                    const int r = e / Nb, c2 = e - r * Nb;
                    const int c2c = c2 / n, i2 = c2 - c2c * n;
                    double s = 0.0;
                    const double* FWr = FW + (size_t)r * pc + c2c * p;
                    const double* Fr = F + (size_t)i2 * p;
                    for (int bb = 0; bb < p; bb++) s += FWr[bb] * Fr[bb];
                    Sb[(size_t)r * Nc + c2] = s;
                    const double cv = -FWr[i2];
                    Cb[(size_t)r * Nc + c2] = cv;
                    Ch[(size_t)r * Nc + c2] = cv;
                }
            }
            __syncthreads();
        }
        if (misc[2]) st |= SLS_ST_COST_SINGULAR;
        if (misc[3]) st |= SLS_ST_RANK_SKIPPED;
        if (misc[5]) st |= SLS_ST_GROUP_PATTERN;
        // ------------------------------------------------------------------ right-hand side r = h + G W g
        {
            double* r = scratch + oR;
            double* wg = scratch + oDz;                       // W_t gb, t = 1..T (dz is free until the refinement)
            if (has_g) {
#pragma unroll 1
                for (int e = tid; e < T * pbar; e += GEN_NT) {
                    const int t = 1 + e / pbar, v = e % pbar;
                    const double* W = scratch + oWh + (int64_t)t * pc * pc;
                    double s = 0.0;
                    for (int w = 0; w < pbar; w++) s += W[(size_t)v * pc + w] * gb[w];
                    wg[(size_t)t * pc + v] = s;
                }
                __syncthreads();
            }
#pragma unroll 1
            for (int e = tid; e < (T + 1) * Nb; e += GEN_NT) {
                const int k = e / Nb, v = e - k * Nb;
                const int cc = v / n, i = v - cc * n;
                double s = (k == 0) ? e0[v] : 0.0;
                if (has_g) {
                    if (k < T) s += wg[(size_t)(k + 1) * pc + cc * p + i];
                    if (k >= 1) { const double* wk = wg + (size_t)k * pc + cc * p; for (int aa = 0; aa < p; aa++) s -= F[(size_t)i * p + aa] * wk[aa]; }
                }
                r[(size_t)k * Nc + v] = s;
            }
            __syncthreads();
        }
        gen_solve_mu(cx, yv, tv);
        double* z = scratch + oZv;
        gen_primal(cx, z, has_g, avec);
        double res = gen_residual(cx, z, e0, red);
        // ------------------------------------------------------------------ refinement on the stored factors
#pragma unroll 1
        for (int pass = 0; pass < a.max_refine; pass++) {
            if (!(res > a.refine_tol) || res > 1e-2) break;
            gen_solve_mu(cx, yv, tv);
            double* dz = scratch + oDz;
            gen_primal(cx, dz, false, avec);
#pragma unroll 1
            for (int e = tid; e < T * pbar; e += GEN_NT) { const int t = 1 + e / pbar, v = e % pbar; z[(size_t)t * pc + v] += dz[(size_t)t * pc + v]; }
            __syncthreads();
            const double res2 = gen_residual(cx, z, e0, red);
            const bool stop = !(res2 < 0.5 * res);
            if (!(res2 < res)) {                               // no gain: take the step back
#pragma unroll 1
                for (int e = tid; e < T * pbar; e += GEN_NT) { const int t = 1 + e / pbar, v = e % pbar; z[(size_t)t * pc + v] -= dz[(size_t)t * pc + v]; }
                __syncthreads();
            } else res = res2;
            if (stop) break;
        }
        if (res > a.feas_tol || res != res) st |= SLS_ST_RESIDUAL;
        // ------------------------------------------------------------------ values -> nzval of every tap (ascending rows = kernel 3's order)
        if (!misc[5]) {
#pragma unroll 1
            for (int w = tid; w < 2 * T * q; w += GEN_NT) {
                const int cc = w / (2 * T), tt = w - cc * 2 * T;
                const bool isx = tt < T; const int t = isx ? tt : tt - T;
                const int64_t ob = P.colptr[(int64_t)tt * (P.Nx + 1) + jcol[cc]] - 1;
                double* out = a.out_nzval[tt] + ob;
                const unsigned char* fr = freeb + (size_t)(t + 1) * pc + cc * p + (isx ? 0 : n);
                const double* zt = z + (size_t)(t + 1) * pc + cc * p + (isx ? 0 : n);
                const int len = isx ? n : m;
                int rk = 0;
                for (int i = 0; i < len; i++) if (fr[i]) out[rk++] = (st & SLS_ST_NO_IMPULSE) ? 0.0 : zt[i];
            }
        }
        if (tid < q) { a.status[kcol[tid]] = st; a.resid[kcol[tid]] = res; }
    }
}

size_t sls_general_smem(int pc, int Nc, int nb, int pb, int qc, int T, int cap) { return GenLayout(pc, Nc, nb, pb, qc, T, cap).total; }
int64_t sls_general_scratch_doubles(int pc, int Nc, int T) { return gen_scratch_doubles(pc < 1 ? 1 : pc, Nc < 1 ? 1 : Nc, T); }

int sls_launch_general_size(const SlsPlanView& P, const int32_t* grp_ptr, int ngroups, int cap, int32_t* stats, sls_stream_t stream) {
    if (ngroups == 0) return 0;
    auto kp = &sls_general_size_kernel;
    const size_t smem = (size_t)4 * cap * sizeof(int32_t);
    if (smem > 48 * 1024) { if (rt::set_max_smem(kp, (int)smem)) return -1; }
    SLS_LAUNCH(kp, (ngroups + 3) / 4, 128, smem, stream, P, grp_ptr, ngroups, cap, stats);
    return rt::last_error();
}

// resident CTAs of kernel 2g for these sizes (the scratch holds one slot per CTA)
int sls_general_grid(const SlsGeneralArgs& a, int sm_count) {
    auto kp = &sls_solve_general_kernel;
    const size_t smem = GenLayout(a.pc, a.Nc, a.nb, a.pb, a.qc, a.P.T, a.cap).total;
    if (smem > 48 * 1024) { if (rt::set_max_smem(kp, (int)smem)) return -1; }
    int per_sm = rt::max_blocks_per_sm(kp, GEN_NT, smem);
    if (per_sm < 1) per_sm = 1;
    const int64_t g = (int64_t)sm_count * per_sm;
    return (int)(g < a.ngroups ? g : a.ngroups);
}

int sls_launch_solve_general(const SlsGeneralArgs& a, int grid, sls_stream_t stream) {
    if (a.ngroups == 0) return 0;
    auto kp = &sls_solve_general_kernel;
    const size_t smem = GenLayout(a.pc, a.Nc, a.nb, a.pb, a.qc, a.P.T, a.cap).total;
    if (smem > 48 * 1024) { if (rt::set_max_smem(kp, (int)smem)) return -1; }
    SLS_LAUNCH(kp, grid, GEN_NT, smem, stream, a);
    return rt::last_error();
}
