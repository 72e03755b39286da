// sls_kernels.h — host-side launchers of the three kernels (implemented in the .cu files).
#pragma once
#include "sls_common.h"

// ---- kernel 1 (sls_gather.cu) ----
int sls_launch_gather_count(const SlsCsc& A, const SlsSupports& Sx, const SlsSupports& Su, int T, int64_t Nx,
                            const int32_t* cols, int ncols, int32_t* n_out, int32_t* m_out, int cap,
                            int32_t* keep_x, int32_t* keep_u, const int32_t* bad_supports, int32_t* stats, sls_stream_t stream);
// stats (36 ints, zeroed by the caller): [0] max n, [1] max m, [2] overflow, [3] some index set is not a run, [4+k] columns of size class k, [20+k] max m of class k
int sls_launch_class_lists(const int32_t* n, int ncols, const int32_t* stats, int nclass, int32_t* lists, sls_stream_t stream);   // keep_*: [ncols][cap] index sets for kernel 1b, or null
// layout check (rows in 1..nrow, strictly ascending per column) of the support columns [c0, c1) of every tap; sets *flag
int sls_launch_validate_supports(const SlsSupports& S, int T, int64_t Nx, int64_t nrow, int64_t c0, int64_t c1, int32_t* flag,
                                 sls_stream_t stream);
__host__ __device__ int sls_gather_fill_words(int cap, int nstride, int mstride, int wx, int wu, int T);
int sls_launch_gather_fill(const SlsCsc& A, const SlsCsc& B2, const SlsSupports& Sx, const SlsSupports& Su,
                           const SlsPlanView& P, int cap, const int32_t* keep_x, const int32_t* keep_u, sls_stream_t stream);
// scans the window [lo, hi) of every colptr array (counts outside it are known to be zero and are not read)
size_t sls_scan_scratch_bytes(int narr, int64_t span);
int sls_launch_scan(int64_t* colptr, int64_t len, int narr, int64_t* tot, int64_t totlen, int64_t lo, int64_t hi, sls_stream_t stream,
                    int64_t* scratch = nullptr);
int sls_launch_local_dense(const SlsCsc& A, const SlsCsc& B2, const SlsPlanView& P, int c,
                           double* At, double* B2t, uint8_t* Mx, uint8_t* Mu, sls_stream_t stream);
// output-feedback reduction (reduction.jl:16-18): first-occurrence union of q column segments (rows 1-based, seg[q+1]);
// seen = zeroed bitmap of ceil(nrow/32) words; sets *flag on rows outside 1..nrow or not ascending inside a segment
int sls_launch_union_first(const int64_t* rows, const int64_t* seg, int q, int64_t nrow, uint32_t* seen, int64_t* out,
                           int64_t* n_out, int32_t* flag, sls_stream_t stream);

// ---- kernel 2 (sls_solve.cu) ----
struct SlsSolveArgs {
    SlsCsc A, B2;            // CSC (column access)
    SlsCsc Ar, B2r;          // CSC of the transposes == CSR of A, B2 (row access)
    SlsPlanView P;
    const int32_t* list;     // planned-column indices of this size class
    int count;
    int mcap, nnzcap;        // class maxima of m and of the local nnz (shared-memory carve-up)
    double* const* out_nzval;  // [2T] device pointers: nzval arrays of the Φx taps then the Φu taps (written directly)
    double* hist;            // factor history, one slot per resident warp
    int64_t hist_stride;     // doubles per slot
    int32_t* status;         // per planned column
    double* resid;           // per planned column: max-abs achievability residual
    double rank_tol, feas_tol;
    double refine_tol;       // residual above which a column takes a refinement sweep on the stored factors
    int max_refine;          // at most this many sweeps (0 = never)
    int32_t* refined;        // counter of columns that took >= 1 sweep (may be null)
};
// size classes (template instantiations) available for the local state dimension n
int sls_solve_class_of(int n);                 // smallest instantiated N >= n, or -1
int sls_solve_class_count();
int sls_solve_class_N(int k);
// shared memory (bytes per warp) and history doubles per warp for class N
size_t sls_solve_smem_per_warp(int N, int mcap, int nnzcap, int T, int wu);
int64_t sls_solve_hist_doubles(int N, int T, int wu = 1);   // wu: mask words of the plan's inputs (> 4 routes every class to the n > 64 kernel)
// launches one class; *slots_io: in = available history slots (warps), out = warps launched
int sls_launch_solve(int N, const SlsSolveArgs& a, int sm_count, int max_smem, int* warps_launched,
                     sls_stream_t stream);
int sls_solve_max_resident_warps(int N, int mcap, int nnzcap, int T, int wu, int sm_count, int max_smem);
// large size classes (32 < N <= 64): one CTA per column (sls_solve_wide.cu); a history slot per resident CTA
int sls_launch_solve_wide(int N, const SlsSolveArgs& a, int sm_count, int max_smem, int* slots_launched, sls_stream_t stream);
int sls_solve_wide_max_slots(int N, int mcap, int nnzcap, int T, int wu, int sm_count, int max_smem);
int64_t sls_solve_wide_hist_doubles(int N, int T);
size_t sls_solve_wide_smem(int N, int mcap, int nnzcap, int T, int wu);

// 64 < N <= 128: one CTA per column, matrices in the CTA's global scratch slot (sls_solve_huge.cu); a slot per resident CTA
int sls_launch_solve_huge(int N, const SlsSolveArgs& a, int sm_count, int max_smem, int* slots_launched, sls_stream_t stream);
int sls_solve_huge_max_slots(int N, int mcap, int nnzcap, int T, int wx, int wu, int sm_count, int max_smem);
int64_t sls_solve_huge_hist_doubles(int N, int T);

// ---- kernel 2g (sls_solve_general.cu): general cost / coupled column groups, one CTA per group ----
struct SlsGeneralArgs {
    SlsCsc A, B2;                     // CSC
    SlsCsc C1, D12, D11, B1;          // cost matrices by columns (valid when the has_* flag is set)
    int has_cost, has_d11, has_b1;    // has_cost = 0: [C1 D12] = I
    SlsSupports Sx, Su;
    SlsPlanView P;
    const int32_t* grp_ptr;           // [ngroups+1] offsets of each group's members in the planned-column order, or null (singletons)
    int ngroups;
    int pc, Nc, nb, pb, qc, cap;      // launch-wide maxima: q(n+m), q n, n, n+m, q; sort-buffer capacity (power of two)
    double* const* out_nzval;         // [2T]
    double* scratch; int64_t scratch_stride;   // per resident CTA (sls_general_scratch_doubles)
    int32_t* status; double* resid;
    double rank_tol, feas_tol, refine_tol;
    int max_refine;
};
size_t sls_general_smem(int pc, int Nc, int nb, int pb, int qc, int T, int cap);
int64_t sls_general_scratch_doubles(int pc, int Nc, int T);
// stats (8 ints, zeroed by the caller): max q, max n, max n+m, max q n, max q(n+m), overflow
int sls_launch_general_size(const SlsPlanView& P, const int32_t* grp_ptr, int ngroups, int cap, int32_t* stats, sls_stream_t stream);
int sls_general_grid(const SlsGeneralArgs& a, int sm_count);
int sls_launch_solve_general(const SlsGeneralArgs& a, int grid, sls_stream_t stream);

// ---- kernel 3 (sls_scatter.cu) ----
struct SlsScatterArgs {
    SlsPlanView P;
    int c_begin, c_end;      // planned-column range handled by this launch
    // destination arrays: per tap tt in [0,2T): base pointers (device arrays of pointers)
    int64_t* const* rowval;  // [2T] -> int64 rowval buffers (1-based out)
    int64_t row_off_x, row_off_u;   // added to the stored rows (global numbering of a sub-plant window's outputs)
    int sorted;              // the planned columns ascend: 32 consecutive planned columns own one contiguous rowval range
    int runs;                // every index set of the plan is a run of consecutive rows: row = first row + mask bit (no staging)
};
int sls_launch_scatter(const SlsScatterArgs& a, sls_stream_t stream);
int sls_launch_unscale(const SlsPlanView& P, int c_begin, int c_end, const int64_t* const* rowval, double* const* nzval,
                       const double* wx, const double* wu, int64_t row_off_x, int64_t row_off_u, sls_stream_t stream);
int sls_launch_iota(int32_t* dst, int32_t first, int n, sls_stream_t stream);
int sls_kcopy(void* dst, const void* src, size_t bytes, sls_stream_t stream);   // either side may be pinned host memory
int sls_kgather64(int64_t* dst, const int64_t* src, size_t pitch, int count, sls_stream_t stream);
int sls_launch_colptr_rebase(int64_t* dst, const int64_t* src, const int64_t* base, int64_t len, int narr, int64_t j0, int64_t j1,
                             sls_stream_t stream);

// ---- closed-loop FIR simulation (sls_sim.cu) ----
struct SlsSimArgs {
    int64_t Nx, Nu, steps;
    int T, nscen;
    const int64_t* colptr;            // [2T][Nx+1] Julia colptr of Phi_x taps then Phi_u taps
    const int64_t* const* rowval;     // [2T] device pointers
    const double* const* nzval;       // [2T]
    SlsCsc A, Ar, B2r;                // A by columns; A and B2 by rows
    const int64_t* dcol;              // [nscen] 1-based disturbance column j_s
    const int64_t* dtime;             // [nscen] 1-based disturbance time t_s
    double* xnorm; double* unorm;     // [nscen][steps] or null
    double* xout; double* uout;       // dense [nscen][steps][Nx] / [nscen][steps][Nu] or null
    const int32_t* nw; const int32_t* nu;       // window sizes per scenario
    const int32_t* Wg; const int32_t* Wug;      // windows [nscen][wcap], [nscen][ucap] (ascending global indices)
    const int64_t* pk_off;                      // [nscen] first packed entry of the scenario
    const uint32_t* pk_tapcnt;                  // [nscen][2T] packed entries per tap
    const uint16_t* pk_cnt;                     // [nscen][2T][wcap] packed entries per (tap, window column)
    const uint16_t* pk_row; const double* pk_val;   // packed Phi restricted to the window: local row, value
    int* overflow;
    double* cta_vals; uint16_t* cta_meta;       // CTA kernel: per-CTA scratch for the row-ordered window when it does not fit
                                                // shared memory ([grid][nent_cap] each; the wave's windows then live in L2), or null
};
__host__ __device__ size_t sls_sim_smem_per_warp(int cap, int wcap, int ucap, int T);
int sls_launch_sim_size(const SlsSimArgs& a, int32_t* nw_out, int32_t* nu_out, int cap, int32_t* Wg, int32_t* Wug,
                        int wcap, int ucap, sls_stream_t stream);
int sls_launch_sim(const SlsSimArgs& a, int cap, int wcap, int ucap, sls_stream_t stream);
int sls_launch_sim_compact(const int32_t* Wt, const int32_t* Wut, int cap, const int32_t* nw, const int32_t* nu, int nscen,
                           int32_t* W, int32_t* Wu, int wcap, int ucap, sls_stream_t stream);
// fast path: one CTA per scenario, packed Phi window resident in shared memory (needs 2T <= 256, windows <= 256, nent_cap entries fit)
size_t sls_sim_cta_smem(int nent_cap, int wcap, int ucap, int T);   // nent_cap = 0: window in the global scratch
int sls_launch_sim_cta(const SlsSimArgs& a, int wcap, int ucap, int nent_cap, int sm_count, sls_stream_t stream);
int sls_launch_sim_pack(const SlsSimArgs& a, bool fill, int wcap, int ucap, int64_t* nent, uint32_t* tapcnt, const int64_t* pk_off,
                        uint16_t* pk_cnt, uint16_t* pk_row, double* pk_val, sls_stream_t stream);

// ---- support generation (sls_supports.cu) ----
struct SlsSuppArgs {
    SlsCsc A, B2r;                    // A by columns (pattern graph), B2 by rows
    int64_t Nx, Nu;
    int nlev, cap;                    // largest exponent; shared-memory set capacity (power of two)
    const int32_t* ex_slot;           // [nlev+1] pattern slot of exponent s among the distinct x exponents, or -1
    const int32_t* eu_slot;           // same for the u exponents
    int64_t* xcolptr; int64_t* ucolptr;     // [nslots][Nx+1]: counts at [slot][j] (count pass) -> Julia colptr after the scan
    int64_t* xrow; int64_t* urow;           // rowval of all slots back to back (fill pass)
    const int64_t* xoff; const int64_t* uoff;   // [nslots] offsets of each slot inside xrow / urow
    int* overflow;
    int xlast_slot, ulast_slot;       // pattern slots of the last tap: their longest column goes to maxlen[0], maxlen[1]
    int* maxlen;
};
int sls_launch_supports(const SlsSuppArgs& a, bool fill, sls_stream_t stream);
