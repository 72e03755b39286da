// sls_solve_wide.cu — kernel 2 for the large size classes (32 < n <= 64, BASELINE config #4): one CTA per column.
//
// Same mathematics as sls_solve.cu (block elimination of the block-tridiagonal G G^T with rank skipping, then back
// substitution; reference src/synthesis.jl:46-62), different mapping.  A 64 x 64 matrix is 128 doubles per lane of one
// warp — it spilled to local memory in the one-warp kernel (278 ms on the 300 x 300 grid).  Here NB = N/8 warps share a
// column: warp w owns block ROW w of the n x n matrices (NB blocks of 8 x 8 in the DMMA accumulator layout: 2 doubles per
// lane per block), so S and T = A~ Pi take 4*NB registers each and nothing spills.  What a warp needs from the other rows
// goes through shared memory in the same blocked layout (one conflict-free LDS.128 per operand fragment):
//     T_w,j   = A~_w,j + sum_l A~_w,l (S_j,l)^T          B fragments = blocks of the previous -Z (full matrix in smem)
//     S_w,j   = D'_w,j + sum_l T_w,l (A~_j,l)^T          B fragments = blocks of A~ diag(mx) (smem)
//     pivot p : every warp posts its block S_w,p (the current column p), all warps invert S_p,p redundantly (8 scalar
//               symmetric sweeps through shuffles, as in the one-warp kernel), then
//               warp w != p:  Tn = -S_w,p P,  S_w,j += Tn (S_j,p)^T (j != p),  S_w,p = -Tn
//               warp p     :  S_p,j = P (S_j,p)^T (j != p),  S_p,p = -P         (one __syncthreads per pivot block)
// -Z_k goes to the column's history slot as full blocks in register order (coalesced 512-byte rows per block); the back
// substitution's mat-vec reads a warp's own block row back.  All sparse / vector phases run one thread per row.
#include "sls_solve_common.h"

struct SlsWideLayout {
    int N, NB, mcap, nnzcap, T, wu;
    int oS, oAblk, oBBt, oY0, oY1, oG, oDref, oLam, oTx, oRhs, oXa, oUc, oValR, oValC;
    int oSx, oSu, oMx, oMu, oPre, oRptr, oRsplit, oColR, oCptr, oRowC, oMisc, total;
    __host__ __device__ SlsWideLayout(int N_, int mcap_, int nnzcap_, int T_, int wu_) {
        N = N_; NB = N_ / 8; mcap = mcap_ < 1 ? 1 : mcap_; nnzcap = nnzcap_ < 1 ? 1 : nnzcap_; T = T_; wu = wu_ < 1 ? 1 : wu_;
        int o = 0;
#define SLS_SEC(name, bytes) name = o; o += (bytes); o = (o + 15) & ~15
        SLS_SEC(oS, 8 * N * N);
        SLS_SEC(oAblk, 8 * N * N);
        SLS_SEC(oBBt, 8 * N * N);
        SLS_SEC(oY0, 8 * N);
        SLS_SEC(oY1, 8 * N);
        SLS_SEC(oG, 8 * N);
        SLS_SEC(oDref, 8 * N);
        SLS_SEC(oLam, 8 * N);
        SLS_SEC(oRhs, 8 * N);
        SLS_SEC(oXa, 8 * N);
        SLS_SEC(oMisc, 128 + 1024);
        SLS_SEC(oSx, 4 * N);
        SLS_SEC(oRptr, 2 * (N + 2));
        SLS_SEC(oRsplit, 2 * (N + 2));
        SLS_SEC(oTx, 8 * (N + mcap));
        SLS_SEC(oUc, 8 * (N + mcap));
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

// resident CTAs per SM the register allocation aims at (shared memory allows about as many)
template <int NB> struct SlsWideCfg { static const int MINB = 2; };

template <int NB>
__global__ void __launch_bounds__(NB * 32, SlsWideCfg<NB>::MINB)
sls_solve_wide_kernel(SlsSolveArgs a) {
    constexpr int N = 8 * NB;
    constexpr int NT = NB * 32;              // threads per CTA = per column
    constexpr int BLK = NB * NB * 64;        // doubles of one full matrix (blocked)
    typedef typename SlsMask<NB>::type mask_t;
    SLS_DYN_SMEM(smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, q = lane & 3;   // accumulator-layout coordinates
    const SlsPlanView& P = a.P;
    const int T = P.T;
    const SlsWideLayout L(N, a.mcap, a.nnzcap, T, P.wu);
    unsigned char* base = smem_raw;
    double* Ssm = reinterpret_cast<double*>(base + L.oS);       // -Z_{k-1} (full, blocked); column panels during the sweep
    double* Am = reinterpret_cast<double*>(base + L.oAblk);     // A~ diag(mx_k), blocked
    double* Dp = reinterpret_cast<double*>(base + L.oBBt);      // diag(mx_{k+1}) + B~ diag(mu_k) B~^T, blocked
    double* ybuf0 = reinterpret_cast<double*>(base + L.oY0);
    double* ybuf1 = reinterpret_cast<double*>(base + L.oY1);
    double* gbuf = reinterpret_cast<double*>(base + L.oG);
    double* thrs = reinterpret_cast<double*>(base + L.oDref);
    double* lam_s = reinterpret_cast<double*>(base + L.oLam);
    double* tx = reinterpret_cast<double*>(base + L.oTx);
    double* rhs_s = reinterpret_cast<double*>(base + L.oRhs);
    double* xa = reinterpret_cast<double*>(base + L.oXa);
    double* ucur = reinterpret_cast<double*>(base + L.oUc);
    double* valr = reinterpret_cast<double*>(base + L.oValR);
    double* valc = reinterpret_cast<double*>(base + L.oValC);
    int32_t* sx = reinterpret_cast<int32_t*>(base + L.oSx);
    int32_t* su = reinterpret_cast<int32_t*>(base + L.oSu);
    mask_t* mxs = reinterpret_cast<mask_t*>(base + L.oMx);
    uint32_t* mus = reinterpret_cast<uint32_t*>(base + L.oMu);
    int64_t* obase = reinterpret_cast<int64_t*>(base + L.oPre);
    uint16_t* rptr = reinterpret_cast<uint16_t*>(base + L.oRptr);
    uint16_t* rsplit = reinterpret_cast<uint16_t*>(base + L.oRsplit);
    uint16_t* colr = reinterpret_cast<uint16_t*>(base + L.oColR);
    uint16_t* cptr = reinterpret_cast<uint16_t*>(base + L.oCptr);
    uint16_t* rowc = reinterpret_cast<uint16_t*>(base + L.oRowC);
    uint64_t* misc64 = reinterpret_cast<uint64_t*>(base + L.oMisc);   // [0] nzA, [1..2] non-empty ballots
    double* miscd = reinterpret_cast<double*>(base + L.oMisc + 32);   // [0..NB) per-warp residual maxima
    int* rankflag = reinterpret_cast<int*>(base + L.oMisc + 104);     // a pivot block dropped a non-empty row
    double* Pbuf = reinterpret_cast<double*>(base + L.oMisc + 128);   // -P of the current pivot block (row-major 8 x 8), two buffers:
                                                                      // the owner of block p+1 may post while block p's is still being read
    const int wu = P.wu;

    double* hist = a.hist + (size_t)blockIdx.x * a.hist_stride;       // [T+1][N] yh history, then [T][BLK] -Z_k in register order
    double* yhist = hist;
    double* zhist = hist + (size_t)(T + 1) * N;
    const int frag = (g << 3) + 2 * q;                                // a lane's double2 inside an 8x8 block

    for (int idx = blockIdx.x; idx < a.count; idx += gridDim.x) {
        const int c = a.list[idx];
        const int n = P.n[c], m = P.m[c], jl = P.jloc[c];
        __syncthreads();
        // ------------------------------------------------------------------ metadata -> shared
#pragma unroll 1
        for (int i = tid; i < n; i += NT) sx[i] = P.sx[(size_t)c * P.nstride + i];
#pragma unroll 1
        for (int l = tid; l < m; l += NT) su[l] = P.su[(size_t)c * P.mstride + l];
#pragma unroll 1
        for (int t = tid; t < T; t += NT) {
            const uint32_t* w = P.maskx + ((size_t)c * T + t) * P.wx;
            mask_t v = (mask_t)w[0];
            if (sizeof(mask_t) == 8 && P.wx > 1) v |= (mask_t)((uint64_t)w[1] << 32);
            mxs[t] = v;
        }
        if (tid == 0) { mxs[T] = (mask_t)0; *rankflag = 0; }
#pragma unroll 1
        for (int t = tid; t < T * wu; t += NT) mus[t] = P.masku[(size_t)c * T * wu + t];
#pragma unroll 1
        for (int t = tid; t < 2 * T; t += NT) obase[t] = P.colptr[(int64_t)t * (P.Nx + 1) + P.cols[c]] - 1;
#pragma unroll 1
        for (int e = tid; e < N * N; e += NT) Am[e] = 0.0;
        __syncthreads();
        // ------------------------------------------------------------------ local CSR / CSC of C~ = [A~ B~]
        // one thread per local row (CSR) / per local column (CSC); offsets by a serial prefix over <= N + mcap counts in
        // shared memory (the column's `tx` vector doubles as integer scratch here)
        {
            int32_t* cnts = reinterpret_cast<int32_t*>(tx);
            if (tid == 0) misc64[0] = 0ull;
            int cntA = 0, cntB = 0;
            int64_t ea0 = 0, ea1 = 0, eb0 = 0, eb1 = 0;
            if (tid < n) {
                const int64_t gr = sx[tid];
                ea0 = a.Ar.colptr[gr] - 1; ea1 = a.Ar.colptr[gr + 1] - 1;
                eb0 = a.B2r.colptr[gr] - 1; eb1 = a.B2r.colptr[gr + 1] - 1;
#pragma unroll 1
                for (int64_t e = ea0; e < ea1; e++) cntA += (sls_find_cold(sx, n, (int32_t)(a.Ar.rowval[e] - 1)) >= 0);
#pragma unroll 1
                for (int64_t e = eb0; e < eb1; e++) cntB += (sls_find_cold(su, m, (int32_t)(a.B2r.rowval[e] - 1)) >= 0);
                cnts[tid] = cntA + cntB;
            }
            __syncthreads();
            if (tid < n) {
                int excl = 0;
#pragma unroll 1
                for (int i = 0; i < tid; i++) excl += cnts[i];
                rptr[tid] = (uint16_t)excl; rsplit[tid] = (uint16_t)(excl + cntA);
                if (tid == n - 1) rptr[n] = (uint16_t)(excl + cntA + cntB);
                int o = excl;
                uint64_t nz = 0ull;
#pragma unroll 1
                for (int64_t e = ea0; e < ea1; e++) {
                    const int l = sls_find_cold(sx, n, (int32_t)(a.Ar.rowval[e] - 1));
                    if (l >= 0) { colr[o] = (uint16_t)l; valr[o] = a.Ar.nzval[e]; o++; nz |= 1ull << ((tid >> 3) * NB + (l >> 3)); }
                }
#pragma unroll 1
                for (int64_t e = eb0; e < eb1; e++) {
                    const int l = sls_find_cold(su, m, (int32_t)(a.B2r.rowval[e] - 1));
                    if (l >= 0) { colr[o] = (uint16_t)(n + l); valr[o] = a.B2r.nzval[e]; o++; }
                }
                if (nz) atomicOr(reinterpret_cast<unsigned long long*>(misc64), (unsigned long long)nz);
            }
            if (n == 0 && tid == 0) rptr[0] = 0;
            __syncthreads();
            // CSC: columns 0..n of A~, then the m columns of B~
#pragma unroll 1
            for (int p0 = 0; p0 < n + m; p0 += NT) {
                const int p = p0 + tid;
                int cnt = 0;
                int64_t e0 = 0, e1 = 0;
                const SlsCsc& M = (p < n) ? a.A : a.B2;
                if (p < n + m) {
                    const int64_t gc = (p < n) ? sx[p] : su[p - n];
                    e0 = M.colptr[gc] - 1; e1 = M.colptr[gc + 1] - 1;
#pragma unroll 1
                    for (int64_t e = e0; e < e1; e++) cnt += (sls_find_cold(sx, n, (int32_t)(M.rowval[e] - 1)) >= 0);
                    cnts[p] = cnt;
                }
            }
            __syncthreads();
#pragma unroll 1
            for (int p0 = 0; p0 < n + m; p0 += NT) {
                const int p = p0 + tid;
                if (p < n + m) {
                    int o = 0;
#pragma unroll 1
                    for (int i = 0; i < p; i++) o += cnts[i];
                    cptr[p] = (uint16_t)o;
                    if (p == n + m - 1) cptr[n + m] = (uint16_t)(o + cnts[p]);
                    const SlsCsc& M = (p < n) ? a.A : a.B2;
                    const int64_t gc = (p < n) ? sx[p] : su[p - n];
                    const int64_t e0 = M.colptr[gc] - 1, e1 = M.colptr[gc + 1] - 1;
#pragma unroll 1
                    for (int64_t e = e0; e < e1; e++) {
                        const int r = sls_find_cold(sx, n, (int32_t)(M.rowval[e] - 1));
                        if (r >= 0) { rowc[o] = (uint16_t)r; valc[o] = M.nzval[e]; o++; }
                    }
                }
            }
            if (n + m == 0 && tid == 0) cptr[0] = 0;
        }
        __syncthreads();
        const uint64_t nzA = misc64[0];

        // ------------------------------------------------------------------ tap 0:  S_0 = diag(mx_1), yh_0 = e
        double Sr[NB][2];                                     // block row `warp` of -Z_k ("swept" S_k), accumulator layout
        const int myrow = 8 * warp + g;
        {
            const mask_t m1 = mxs[0];
#pragma unroll
            for (int bj = 0; bj < NB; bj++) {
                const int c0 = 8 * bj + 2 * q;
                Sr[bj][0] = (myrow == c0 && sls_mbit(m1, myrow)) ? -1.0 : 0.0;
                Sr[bj][1] = (myrow == c0 + 1 && sls_mbit(m1, myrow)) ? -1.0 : 0.0;
                *reinterpret_cast<double2*>(Ssm + ((warp * NB + bj) << 6) + frag) = make_double2(Sr[bj][0], Sr[bj][1]);
            }
        }
        double* yprev = ybuf0;
        double* ycur = ybuf1;
        if (tid < N) { const double y0 = (tid == jl) ? 1.0 : 0.0; yprev[tid] = y0; yhist[tid] = y0; }
        int status = (jl < 0) ? SLS_ST_NO_IMPULSE : 0;
        uint32_t mu_prev[4] = {0u, 0u, 0u, 0u};
        mask_t mk_prev = 0, mk1_prev = 0;
        bool cache_valid = false;
        uint64_t nonempty = 0ull;
        __syncthreads();

        // ------------------------------------------------------------------ forward elimination, taps 1..T
#pragma unroll 1
        for (int k = 1; k <= T; k++) {
            const mask_t mk = mxs[k - 1];
            const mask_t mk1 = mxs[k];
            const uint32_t* muk = mus + (size_t)(k - 1) * wu;
            // (1) g = mx_k o (Z_{k-1} yh_{k-1}) = -(mx_k o (S yprev)): a warp's own rows
            {
                double acc = 0.0;
#pragma unroll
                for (int bj = 0; bj < NB; bj++) {
                    const double2 yv = *reinterpret_cast<const double2*>(yprev + 8 * bj + 2 * q);
                    acc = fma(Sr[bj][0], yv.x, acc);
                    acc = fma(Sr[bj][1], yv.y, acc);
                }
                acc += __shfl_xor_sync(SLS_FULL, acc, 1);
                acc += __shfl_xor_sync(SLS_FULL, acc, 2);
                if (q == 0) gbuf[myrow] = sls_mbit(mk, myrow) ? -acc : 0.0;
            }
            // (2) mask-dependent operands, rebuilt only when a mask changed
            bool same = cache_valid && mk == mk_prev && mk1 == mk1_prev;
#pragma unroll
            for (int w = 0; w < 4; w++) if (w < wu) same = same && (muk[w] == mu_prev[w]);
            if (!same) {
#pragma unroll 1
                for (int e = tid; e < N * N; e += NT) Dp[e] = 0.0;
                __syncthreads();
                double dref = 0.0;
                if (tid < n) {
                    const int i = tid;
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
                if (tid < N) thrs[tid] = (dref != 0.0) ? a.rank_tol * dref : SLS_INF;
                const uint32_t bal = __ballot_sync(SLS_FULL, dref != 0.0);
                if (lane == 0 && warp < 2) misc64[1 + warp] = (tid < N) ? bal : 0u;
#pragma unroll
                for (int w = 0; w < 4; w++) if (w < wu) mu_prev[w] = muk[w];
                mk_prev = mk; mk1_prev = mk1; cache_valid = true;
                __syncthreads();
                nonempty = misc64[1] | ((NB > 4) ? (misc64[2] << 32) : 0ull);
                if (NB == 1) nonempty &= 0xffull;
            } else {
                __syncthreads();
            }
            // (3) yh_k = A~ g
            if (tid < N) {
                double acc = 0.0;
                if (tid < n) {
                    const int r0 = rptr[tid], r1 = rsplit[tid];
#pragma unroll 1
                    for (int e = r0; e < r1; e++) acc = fma(valr[e], gbuf[colr[e]], acc);
                }
                ycur[tid] = acc; yhist[(size_t)k * N + tid] = acc;
            }
            // (5) T = Am (I + S) = Am + Am S^T   (S symmetric; B fragments = blocks of the previous -Z in shared memory)
            double Tr[NB][2];
#pragma unroll
            for (int bj = 0; bj < NB; bj++) {
                const double2 av = *reinterpret_cast<const double2*>(Am + ((warp * NB + bj) << 6) + frag);
                Tr[bj][0] = av.x; Tr[bj][1] = av.y;
            }
#pragma unroll 1
            for (int bl = 0; bl < NB; bl++) {
                if (!((nzA >> (warp * NB + bl)) & 1ull)) continue;
                const double2 av = *reinterpret_cast<const double2*>(Am + ((warp * NB + bl) << 6) + frag);
#pragma unroll
                for (int bj = 0; bj < NB; bj++) {
                    const double2 sv = *reinterpret_cast<const double2*>(Ssm + ((bj * NB + bl) << 6) + frag);
                    sls_dmma(Tr[bj][0], Tr[bj][1], av.x, sv.x);
                    sls_dmma(Tr[bj][0], Tr[bj][1], av.y, sv.y);
                }
            }
            __syncthreads();                                   // every warp is done reading the previous -Z
            // (6) S_k row = Dp row + T row Am^T
#pragma unroll
            for (int bj = 0; bj < NB; bj++) {
                const double2 bv = *reinterpret_cast<const double2*>(Dp + ((warp * NB + bj) << 6) + frag);
                Sr[bj][0] = bv.x; Sr[bj][1] = bv.y;
            }
#pragma unroll
            for (int bl = 0; bl < NB; bl++) {
#pragma unroll
                for (int bj = 0; bj < NB; bj++) {
                    if ((nzA >> (bj * NB + bl)) & 1ull) {      // warp-uniform
                        const double2 av = *reinterpret_cast<const double2*>(Am + ((bj * NB + bl) << 6) + frag);
                        sls_dmma(Sr[bj][0], Sr[bj][1], Tr[bl][0], av.x);
                        sls_dmma(Sr[bj][0], Sr[bj][1], Tr[bl][1], av.y);
                    }
                }
            }
            // (7) block symmetric sweep with rank skipping: S <- -(S^+), with one block of look-ahead.
            //     The owner of pivot block p inverts it ALONE (8 scalar symmetric sweeps through shuffles on its registers)
            //     and posts -P; in step p the owner of block p+1 only brings its diagonal block up to date (its other
            //     blocks are recomputed from the column panel when its row becomes the pivot row) and inverts it while the
            //     other warps run their DMMA updates, so the serial scalar chain overlaps the tensor work.
            auto sweep_post = [&](double d0, double d1, int pb) {      // pb = pivot block whose owner this warp is
                uint32_t kb = 0u;
                const uint32_t actp = (uint32_t)(nonempty >> (8 * pb)) & 0xffu;
                const int npair = (actp == 0u) ? 0 : ((min(8, n - 8 * pb) + 1) >> 1);
#pragma unroll 1
                for (int rp = 0; rp < npair; rp++) {
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int r = 2 * rp + e;
                        const double src = e ? d1 : d0;
                        const double dv = __shfl_sync(SLS_FULL, src, 4 * r + rp);
                        const double cg = __shfl_sync(SLS_FULL, src, (lane & ~3) | rp);
                        const double rr0 = __shfl_sync(SLS_FULL, d0, 4 * r + q);
                        const double rr1 = __shfl_sync(SLS_FULL, d1, 4 * r + q);
                        const bool ok = dv > thrs[8 * pb + r];
                        const double pinv = ok ? sls_rcp(dv) : 0.0;
                        const bool isrow = (g == r);
                        const double cp = cg * pinv;
                        const double f = isrow ? (1.0 - pinv) : cp;
                        double n0 = fma(-f, rr0, d0), n1 = fma(-f, rr1, d1);
                        const double cv = isrow ? -pinv : cp;
                        if (q == rp) { if (e) n1 = cv; else n0 = cv; }
                        d0 = n0; d1 = n1;
                        kb |= (ok ? 1u : 0u) << r;
                    }
                }
                if (lane == 0 && kb != actp) *rankflag = 1;
                *reinterpret_cast<double2*>(Pbuf + 64 * (pb & 1) + frag) = make_double2(d0, d1);   // -P (symmetric), row-major
            };
            *reinterpret_cast<double2*>(Ssm + ((warp * NB) << 6) + frag) = make_double2(Sr[0][0], Sr[0][1]);   // column panel 0
            if (warp == 0) sweep_post(Sr[0][0], Sr[0][1], 0);
            __syncthreads();
#pragma unroll
            for (int p = 0; p < NB; p++) {
                // here: column p of the current S (blocks (j,p), j != p) and -P_p are in shared memory
                const double2 pv2 = *reinterpret_cast<const double2*>(Pbuf + 64 * (p & 1) + frag);
                const double Pn[2] = {pv2.x, pv2.y};           // -P, symmetric
                const int pn = (p + 1 < NB) ? p + 1 : p;       // next pivot block (a compile-time constant after unrolling)
                if (warp == p) {
                    // pivot row: S_pj = P (S_jp)^T from the panel (its own stale blocks are not read)
#pragma unroll
                    for (int bj = 0; bj < NB; bj++) {
                        if (bj == p) continue;
                        const double2 sv = *reinterpret_cast<const double2*>(Ssm + ((bj * NB + p) << 6) + frag);
                        double t0 = 0.0, t1 = 0.0;            // -P S_pj = -(P S_jp^T)
                        sls_dmma(t0, t1, Pn[0], sv.x);
                        sls_dmma(t0, t1, Pn[1], sv.y);
                        Sr[bj][0] = -t0; Sr[bj][1] = -t1;
                    }
                    Sr[p][0] = Pn[0]; Sr[p][1] = Pn[1];
                    if (p + 1 < NB) *reinterpret_cast<double2*>(Ssm + ((warp * NB + p + 1) << 6) + frag) = make_double2(Sr[pn][0], Sr[pn][1]);
                } else {
                    double Tn[2] = {0.0, 0.0};                // -S_wp P
                    sls_dmma(Tn[0], Tn[1], Sr[p][0], Pn[0]);
                    sls_dmma(Tn[0], Tn[1], Sr[p][1], Pn[1]);
                    if (p + 1 < NB && warp == p + 1) {
                        // next pivot owner: diagonal block only, S_ww -= (S_wp P)(S_wp)^T with S_wp still in the registers
                        sls_dmma(Sr[pn][0], Sr[pn][1], Tn[0], Sr[p][0]);
                        sls_dmma(Sr[pn][0], Sr[pn][1], Tn[1], Sr[p][1]);
                        Sr[p][0] = -Tn[0]; Sr[p][1] = -Tn[1];
                        sweep_post(Sr[pn][0], Sr[pn][1], p + 1);
                    } else {
                        if (p + 1 < NB) {                      // the next panel column first: the owner of block p+1 is waiting for nothing else
                            const double2 sv = *reinterpret_cast<const double2*>(Ssm + ((pn * NB + p) << 6) + frag);
                            sls_dmma(Sr[pn][0], Sr[pn][1], Tn[0], sv.x);
                            sls_dmma(Sr[pn][0], Sr[pn][1], Tn[1], sv.y);
                            *reinterpret_cast<double2*>(Ssm + ((warp * NB + p + 1) << 6) + frag) = make_double2(Sr[pn][0], Sr[pn][1]);
                        }
#pragma unroll
                        for (int bj = 0; bj < NB; bj++) {
                            if (bj == p || bj == p + 1) continue;
                            const double2 sv = *reinterpret_cast<const double2*>(Ssm + ((bj * NB + p) << 6) + frag);   // S_jp before the step
                            sls_dmma(Sr[bj][0], Sr[bj][1], Tn[0], sv.x);                                                // S_wj -= (S_wp P) S_jp^T
                            sls_dmma(Sr[bj][0], Sr[bj][1], Tn[1], sv.y);
                        }
                        Sr[p][0] = -Tn[0]; Sr[p][1] = -Tn[1];
                    }
                }
                if (p + 1 < NB) __syncthreads();               // panel p+1 and -P_{p+1} are complete
            }
            __syncthreads();                                   // the last column panel has been read by everybody
            // (8) -Z_k: own block row to shared memory (next tap's B fragments) and to the history (register order);
            //     -Z_T goes to slot 0 (refinement sweeps)
            {
                double* hk = zhist + (size_t)(k < T ? k : 0) * BLK;
#pragma unroll
                for (int bj = 0; bj < NB; bj++) {
                    const double2 v = make_double2(Sr[bj][0], Sr[bj][1]);
                    *reinterpret_cast<double2*>(Ssm + ((warp * NB + bj) << 6) + frag) = v;
                    *reinterpret_cast<double2*>(hk + ((warp * NB + bj) << 6) + 2 * lane) = v;
                }
            }
            double* tsw = yprev; yprev = ycur; ycur = tsw;
        }

        // ------------------------------------------------------------------ lambda_T = Z_T yh_T
        {
            __syncthreads();
            double acc = 0.0;
#pragma unroll
            for (int bj = 0; bj < NB; bj++) {
                const double2 yv = *reinterpret_cast<const double2*>(yprev + 8 * bj + 2 * q);
                acc = fma(Sr[bj][0], yv.x, acc);
                acc = fma(Sr[bj][1], yv.y, acc);
            }
            acc += __shfl_xor_sync(SLS_FULL, acc, 1);
            acc += __shfl_xor_sync(SLS_FULL, acc, 2);
            if (q == 0) lam_s[myrow] = (myrow < n) ? -acc : 0.0;
        }
        __syncthreads();

        // gbuf = (-Z_kk) rhs_s for a stored tap kk (own block row from the history)
        auto hist_matvec = [&](int slot) {
            const double* hk = zhist + (size_t)slot * BLK;
            double acc = 0.0;
#pragma unroll
            for (int bj = 0; bj < NB; bj++) {
                const double2 z = *reinterpret_cast<const double2*>(hk + ((warp * NB + bj) << 6) + 2 * lane);
                const double2 v = *reinterpret_cast<const double2*>(rhs_s + 8 * bj + 2 * q);
                acc = fma(z.x, v.x, acc);
                acc = fma(z.y, v.y, acc);
            }
            acc += __shfl_xor_sync(SLS_FULL, acc, 1);
            acc += __shfl_xor_sync(SLS_FULL, acc, 2);
            if (q == 0) gbuf[myrow] = (myrow < n) ? acc : 0.0;
        };

        // ------------------------------------------------------------------ back substitution (+ refinement sweeps)
        double* const xcur = ucur;                            // [x_k | u_k]
        double resmax = 0.0, resprev = SLS_INF;
#pragma unroll 1
        for (int pass = 0;; pass++) {
            double xnext = 0.0;                               // x_{k+1}[tid]
            resmax = 0.0;
#pragma unroll 1
            for (int k = T; k >= 1; k--) {
                const mask_t mk = mxs[k - 1];
                const uint32_t* muk = mus + (size_t)(k - 1) * wu;
                // t = C~^T lambda_k
                if (tid < n + m) {
                    double acc = 0.0;
                    const int e0 = cptr[tid], e1 = cptr[tid + 1];
#pragma unroll 1
                    for (int e = e0; e < e1; e++) acc = fma(valc[e], lam_s[rowc[e]], acc);
                    tx[tid] = acc;
                }
#pragma unroll 1
                for (int p = tid + NT; p < n + m; p += NT) {  // (n + m > NT only for very wide input sets)
                    double acc = 0.0;
#pragma unroll 1
                    for (int e = cptr[p]; e < cptr[p + 1]; e++) acc = fma(valc[e], lam_s[rowc[e]], acc);
                    tx[p] = acc;
                }
                __syncthreads();
                // u_k = -mu_k o t_u -> output;  rhs = yh_{k-1} + mx_k o t_x
#pragma unroll 1
                for (int l = tid; l < m; l += NT) {
                    const uint32_t word = muk[l >> 5];
                    const bool on = (word >> (l & 31)) & 1u;
                    double u = on ? -tx[n + l] : 0.0;
                    if (on) {
                        int rank = __popc(word & ((1u << (l & 31)) - 1u));
#pragma unroll 1
                        for (int w = 0; w < (l >> 5); w++) rank += __popc(muk[w]);
                        double* dst = a.out_nzval[T + k - 1] + (obase[T + k - 1] + rank);
                        if (pass) u += *dst;
                        *dst = u;
                    }
                    ucur[n + l] = u;
                }
                if (tid < N) rhs_s[tid] = (tid < n) ? (yhist[(size_t)(k - 1) * N + tid] + (sls_mbit(mk, tid) ? tx[tid] : 0.0)) : 0.0;
                __syncthreads();
                // lambda_{k-1} = Z_{k-1} rhs -> gbuf holds -lambda_{k-1}
                if (k - 1 == 0) { if (tid < N) gbuf[tid] = (tid < n && sls_mbit(mxs[0], tid)) ? -rhs_s[tid] : 0.0; }
                else hist_matvec(k - 1);
                __syncthreads();
                double xk = 0.0;
                if (tid < N) {
                    const bool on = (tid < n) && sls_mbit(mk, tid);
                    const double lam_prev = (tid < n) ? -gbuf[tid] : 0.0;
                    double x = on ? (lam_prev - tx[tid]) : 0.0;
                    if (on) {
                        double* dst = a.out_nzval[k - 1] + (obase[k - 1] + sls_mrank(mk, tid));
                        if (pass) x += *dst;
                        *dst = x;
                    }
                    xk = x;
                    if (tid < n) xcur[tid] = x;
                    lam_s[tid] = lam_prev;
                }
                __syncthreads();
                // residual of constraint block k
                if (tid < n) {
                    double r = -xnext;
                    const int r0 = rptr[tid], r2 = rptr[tid + 1];
#pragma unroll 1
                    for (int e = r0; e < r2; e++) r = fma(valr[e], xcur[colr[e]], r);
                    resmax = sls_nanmax(resmax, fabs(r));
                }
                xnext = xk;
                __syncthreads();
            }
            if (tid < n) resmax = sls_nanmax(resmax, fabs(((tid == jl) ? 1.0 : 0.0) - xnext));
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) resmax = sls_nanmax(resmax, __shfl_xor_sync(SLS_FULL, resmax, d));
            if (lane == 0) miscd[warp] = resmax;
            __syncthreads();
            resmax = miscd[0];
#pragma unroll
            for (int w = 1; w < NB; w++) resmax = sls_nanmax(resmax, miscd[w]);
            __syncthreads();
            if (!(resmax > a.refine_tol) || pass >= a.max_refine || !(resmax < 0.5 * resprev) || !(resmax <= 1e-2)) break;
            resprev = resmax;
            if (pass == 0 && tid == 0 && a.refined) atomicAdd(a.refined, 1);
            // ---- (R1) rho_k = h_k - (G v)_k of the stored solution -> yhist[k]
            xnext = 0.0;
#pragma unroll 1
            for (int k = T; k >= 1; k--) {
                const mask_t mk = mxs[k - 1];
                const uint32_t* muk = mus + (size_t)(k - 1) * wu;
#pragma unroll 1
                for (int l = tid; l < m; l += NT) {
                    const uint32_t word = muk[l >> 5];
                    const bool on = (word >> (l & 31)) & 1u;
                    double u = 0.0;
                    if (on) {
                        int rank = __popc(word & ((1u << (l & 31)) - 1u));
#pragma unroll 1
                        for (int w = 0; w < (l >> 5); w++) rank += __popc(muk[w]);
                        u = a.out_nzval[T + k - 1][obase[T + k - 1] + rank];
                    }
                    ucur[n + l] = u;
                }
                double xk = 0.0;
                if (tid < n) {
                    if (sls_mbit(mk, tid)) xk = a.out_nzval[k - 1][obase[k - 1] + sls_mrank(mk, tid)];
                    xcur[tid] = xk;
                }
                __syncthreads();
                if (tid < N) {
                    double r = 0.0;
                    if (tid < n) {
                        r = -xnext;
                        const int r0 = rptr[tid], r2 = rptr[tid + 1];
#pragma unroll 1
                        for (int e = r0; e < r2; e++) r = fma(valr[e], xcur[colr[e]], r);
                    }
                    yhist[(size_t)k * N + tid] = r;
                }
                xnext = xk;
                __syncthreads();
            }
            if (tid < N) { const double r = (tid < n) ? (((tid == jl) ? 1.0 : 0.0) - xnext) : 0.0; yhist[tid] = r; rhs_s[tid] = r; }
            __syncthreads();
            // ---- (R2) forward sweep with the stored factors: yh_k = rho_k + A~ (mx_k o Z_{k-1} yh_{k-1})
#pragma unroll 1
            for (int k = 1; k <= T + 1; k++) {
                if (k == 1) { if (tid < N) gbuf[tid] = (tid < n && sls_mbit(mxs[0], tid)) ? -rhs_s[tid] : 0.0; }
                else hist_matvec(k - 1 < T ? k - 1 : 0);
                __syncthreads();
                if (k == T + 1) {
                    if (tid < N) lam_s[tid] = (tid < n) ? -gbuf[tid] : 0.0;
                    break;
                }
                const mask_t mk = mxs[k - 1];
                if (tid < N) xa[tid] = (tid < n && sls_mbit(mk, tid)) ? -gbuf[tid] : 0.0;
                __syncthreads();
                if (tid < N) {
                    double acc = 0.0;
                    if (tid < n) {
                        acc = yhist[(size_t)k * N + tid];
                        const int r0 = rptr[tid], r1 = rsplit[tid];
#pragma unroll 1
                        for (int e = r0; e < r1; e++) acc = fma(valr[e], xa[colr[e]], acc);
                    }
                    yhist[(size_t)k * N + tid] = acc;
                    rhs_s[tid] = acc;
                }
                __syncthreads();
            }
            __syncthreads();
        }
        if (!(resmax <= a.feas_tol)) status |= SLS_ST_RESIDUAL;
        if (tid == 0) { if (*rankflag) status |= SLS_ST_RANK_SKIPPED; a.status[c] = status; a.resid[c] = resmax; }
    }
}

// ---- host side ------------------------------------------------------------------------------
size_t sls_solve_wide_smem(int N, int mcap, int nnzcap, int T, int wu) { return (size_t)SlsWideLayout(N, mcap, nnzcap, T, wu).total; }
int64_t sls_solve_wide_hist_doubles(int N, int T) { return (int64_t)(T + 1) * N + (int64_t)T * N * N; }

template <int NB>
static int wide_resident_blocks(int mcap, int nnzcap, int T, int wu, int sm_count, int max_smem) {
    if (wu > 4) return -1;
    const size_t smem = sls_solve_wide_smem(8 * NB, mcap, nnzcap, T, wu);
    if (smem > (size_t)max_smem) return -1;
    auto kp = &sls_solve_wide_kernel<NB>;
    if (smem > 48 * 1024) { if (rt::set_max_smem(kp, (int)smem)) return -1; }
    int bps = rt::max_blocks_per_sm(kp, NB * 32, smem);
    if (bps < 1) return -1;
    if (const char* cap = getenv("SLS_MAX_BPS")) { const int c = atoi(cap); if (c >= 1 && c < bps) bps = c; }
    return bps * sm_count;
}

template <int NB>
static int wide_launch(const SlsSolveArgs& a, int sm_count, int max_smem, int* slots_launched, sls_stream_t stream) {
    const size_t smem = sls_solve_wide_smem(8 * NB, a.mcap, a.nnzcap, a.P.T, a.P.wu);
    const int maxb = wide_resident_blocks<NB>(a.mcap, a.nnzcap, a.P.T, a.P.wu, sm_count, max_smem);
    if (maxb <= 0) return -5;
    auto kp = &sls_solve_wide_kernel<NB>;
    int blocks = a.count < maxb ? a.count : maxb;
    if (blocks < 1) blocks = 1;
    *slots_launched = blocks;
    SLS_LAUNCH(kp, blocks, NB * 32, smem, stream, a);
    return rt::last_error();
}

int sls_solve_wide_max_slots(int N, int mcap, int nnzcap, int T, int wu, int sm_count, int max_smem) {
    switch (N / 8) {
        case 5: return wide_resident_blocks<5>(mcap, nnzcap, T, wu, sm_count, max_smem);
        case 6: return wide_resident_blocks<6>(mcap, nnzcap, T, wu, sm_count, max_smem);
        case 7: return wide_resident_blocks<7>(mcap, nnzcap, T, wu, sm_count, max_smem);
        case 8: return wide_resident_blocks<8>(mcap, nnzcap, T, wu, sm_count, max_smem);
        default: return -1;
    }
}

int sls_launch_solve_wide(int N, const SlsSolveArgs& a, int sm_count, int max_smem, int* slots_launched, sls_stream_t stream) {
    *slots_launched = 0;
    if (a.count == 0) return 0;
    switch (N / 8) {
        case 5: return wide_launch<5>(a, sm_count, max_smem, slots_launched, stream);
        case 6: return wide_launch<6>(a, sm_count, max_smem, slots_launched, stream);
        case 7: return wide_launch<7>(a, sm_count, max_smem, slots_launched, stream);
        case 8: return wide_launch<8>(a, sm_count, max_smem, slots_launched, stream);
        default: return -6;
    }
}
