/*
 * slsb200.h — C ABI of libslsb200: column-wise localized H2 state-feedback SLS synthesis on B200.
 *
 * This is the drop-in boundary for ONE hot path of aaltoKEPO/SystemLevelControl.jl: the body of
 * `_SLS_𝓗₂` (reference src/synthesis.jl:34-72) and the `@distributed` loop around it (:23-27).
 * The Julia host (`SLS_𝓗₂`, src/synthesis.jl:11-32) keeps its signature and calls these symbols
 * through `ccall`; the Python mirror in systemlevelcontrol.jl_b200/ binds the
 * same symbols with ctypes.  No C++ types, no exceptions and no callbacks cross this boundary.
 *
 * Layout contract (Julia `SparseMatrixCSC{Tv,Int64}` fields, passed zero-copy):
 *   colptr : int64[ncol+1], 1-based;  rowval : int64[nnz], 1-based, strictly ascending in a column;
 *   nzval  : double[nnz] (plant / outputs) or uint8 Bool[nnz] (supports).
 * Ownership: the caller owns every host buffer; the library copies inputs to the device during
 * `sls_set_*` and keeps no host pointer after a call returns.  Outputs are sized by `sls_plan`
 * (two-phase sizing) so the library never allocates memory the caller must free.
 * Errors: return 0 = ok, <0 = argument/layout/capacity error, >0 = CUDA error code; text via
 * `sls_last_error`.  Numerical outcomes go to `status[]`, never to the return code.
 * Threading: a handle is not thread-safe; one handle per host thread / per GPU.
 */
#ifndef SLSB200_H
#define SLSB200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SLS_ABI_VERSION 9

/* per-column status bits written by sls_solve* (improves on the reference, which ignores the
 * solver's termination status: src/synthesis.jl:62-65) */
#define SLS_STATUS_OK            0
#define SLS_STATUS_RESIDUAL      1   /* achievability residual above feas_tol after the refinement sweeps: infeasible
                                        localisation, rank-deficient beyond rank_tol, or cond(G) >~ 1e6   */
#define SLS_STATUS_RANK_SKIPPED  2   /* informational: >=1 redundant constraint row was eliminated      */
#define SLS_STATUS_NO_IMPULSE    4   /* column j is not in its own s_x (Ĩ is a zero column) => Φ = 0    */
#define SLS_STATUS_COST_SINGULAR 8  /* general cost only: [C1~ D12~]' [C1~ D12~] is singular on the free entries of a tap (the
                                       reference's minimiser is then not unique); those entries were fixed to 0 */
#define SLS_STATUS_GROUP_PATTERN 16 /* column groups only: a member column's support reaches rows of the group's union index set
                                       that its own index set does not hold; nothing was written for the group */

typedef struct sls_handle sls_handle;

typedef struct sls_options {
    int32_t device;      /* CUDA device ordinal                                                        */
    int32_t flags;       /* bit 0: also validate the supports' rowval (range, strictly ascending) on the host;
                            bit 1: sls_solve copies the outputs back after kernel 2 instead of chunk by chunk;
                            bit 2: no refinement sweeps in kernel 2 (columns whose residual exceeds feas_tol are only flagged);
                            bit 3: sls_synthesize* always copies the rowval outputs from the device (default: a tap whose stored-entry
                                   count equals the support's over a column chunk — then its rows ARE the support's — is filled
                                   from the caller's own S[t].rowval on the host, bit-identical, half the device-to-host bytes) */
    double  rank_tol;    /* pivot <= rank_tol * diag(D_k) => constraint row treated as redundant; 0 -> 1e-10 */
    double  feas_tol;    /* max-abs achievability residual above which SLS_STATUS_RESIDUAL is set; 0 -> 1e-12.  Columns above
                            feas_tol / 10 first take up to 3 refinement sweeps on the stored factors (flags bit 2: none) */
} sls_options;

typedef struct sls_timing {      /* device times (CUDA events on the handle's stream) of the last plan+solve */
    float gather_ms;             /* kernel 1: index sets, masks, counts (+ scans)                       */
    float solve_ms;              /* kernel 2: batched FP64 equality-constrained LS                      */
    float scatter_ms;            /* kernel 3: CSC scatter (runs alone, before kernel 2)                 */
    float h2d_ms, d2h_ms;        /* host<->device copies done inside sls_set_* / sls_solve              */
    int64_t solve_launches, gather_launches, scatter_launches;
    int64_t ncols, nnz_out;      /* columns synthesised, total stored entries of Φx+Φu                 */
    double  alg_flops;           /* sum_cols n^2*sum_t(nx_t+nu_t) + (7/3)*T*n^3            */
    double  alg_bytes;           /* B_col summed over columns                              */
    float sim_ms;                /* sls_simulate: device time of the simulation kernel                  */
    int32_t sim_window, sim_uwindow;   /* largest state / input window over the scenarios             */
    float gather_gpu_ms;         /* kernel 1 launches alone (gather_ms also holds the host round trips between them) */
    float sim_pack_ms;           /* sls_simulate: window + pack passes before the simulation kernel     */
    int64_t refined;             /* columns of the last solve that took >= 1 refinement sweep of kernel 2 */
    double  sim_bytes;           /* sls_simulate: packed Phi bytes the recursion consumed (sum over scenarios and steps) */
    int64_t host_copied_bytes;   /* sls_synthesize*: bytes of rowval outputs filled from the caller's own S[t].rowval on the host
                                    (taps whose stored pattern the device proved equal to the support's) instead of device-to-host copies */
} sls_timing;

int         sls_abi_version(void);
/* number of visible CUDA devices, or <0 when the CUDA runtime/driver is unusable (never a CPU fallback) */
int         sls_device_count(void);

int         sls_create(const sls_options* opts, sls_handle** out);
void        sls_destroy(sls_handle* h);
const char* sls_last_error(const sls_handle* h);   /* h may be NULL: returns the last create() error   */

/* Plant blocks read on the path: A (Nx x Nx) and B2 (Nx x Nu) of a StateFeedback plant built by the
 * 3-argument constructor `Plant(A,B1,B2)` with B1 = I (GeneralizedPlant.jl:105-110): [C1 D12] = I,
 * D11 = 0, so the cost is the squared Frobenius norm of the free entries (synthesis.jl:50-52,76-86). */
int sls_set_plant(sls_handle* h, int64_t Nx, int64_t Nu,
                  const int64_t* A_colptr, const int64_t* A_rowval, const double* A_nzval,
                  const int64_t* B2_colptr, const int64_t* B2_rowval, const double* B2_nzval);

/* Diagonal cost weights — the 6-argument `Plant(A,B1,B2,C1,D11,D12)` (GeneralizedPlant.jl:101-103) with
 * C1 = [diag(wx); 0], D12 = [0; diag(wu)], D11 = 0:  cost = sum_t ||wx .* Φx[t][:,j]||^2 + ||wu .* Φu[t][:,j]||^2
 * (synthesis.jl:50,76-77).  wx (Nx) / wu (Nu) positive; NULL = ones; both NULL clears the weights.  Must be called
 * BEFORE sls_set_plant (it invalidates the plant on the device: the weighted problem is solved as the unweighted one
 * of the similarity-scaled plant Wx A Wx^-1, Wx B2 Wu^-1 and Φ is scaled back on the device).  resid[] then refers
 * to the scaled variables.  A diagonal B1 only rescales the cost of each column and does not change Φ. */
int sls_set_weights(sls_handle* h, int64_t Nx, int64_t Nu, const double* wx, const double* wu);

/* General cost — the 6-argument `Plant(A,B1,B2,C1,D11,D12)` (GeneralizedPlant.jl:101-103) with arbitrary sparse matrices,
 * Julia CSC like the plant: C1 (Nz x Nx), D12 (Nz x Nu), D11 (Nz x Nw; NULL colptr = zero), B1 (Nx x Nw; NULL colptr = I).
 * Replaces, per column group c_j, the objective of synthesis.jl:42-43,50,76-77:
 *     sum_t || [C1~ D12~] [Φx~[t]; Φu~[t]] B1~[ii_x,:] + D11~ ||_F^2,   [C1~ D12~] = rows [s_x; Nx+s_u], columns s_x / s_u
 * (reduction.jl:15, GeneralizedPlant.jl:266-271).  Call after sls_set_plant; all four colptr NULL clears the cost (back to
 * the identity).  The columns are then solved by the general kernel (one CTA per group, q(n+m) bounded by shared memory:
 * about 64 for q n <= 48) through sls_set_supports + sls_plan + sls_solve; sls_synthesize refuses.  Not combinable with
 * sls_set_weights.  Where the cost is singular on a tap's free entries the column is flagged SLS_STATUS_COST_SINGULAR. */
int sls_set_cost(sls_handle* h, int64_t Nz, int64_t Nw,
                 const int64_t* C1_colptr, const int64_t* C1_rowval, const double* C1_nzval,
                 const int64_t* D12_colptr, const int64_t* D12_rowval, const double* D12_nzval,
                 const int64_t* D11_colptr, const int64_t* D11_rowval, const double* D11_nzval,
                 const int64_t* B1_colptr, const int64_t* B1_rowval, const double* B1_nzval);

/* Column groups — the `𝓘` keyword of SLS_𝓗₂ (synthesis.jl:15) with q > 1 columns per group: group g holds the planned
 * columns grp_ptr[g] .. grp_ptr[g+1]-1 (0-based offsets into the column list of the NEXT sls_plan / sls_solve;
 * grp_ptr[ngroups] must equal its ncols; the columns of a group ascend).  A group is solved as ONE problem on the union of
 * its columns' index sets (reduction.jl:14), coupled through B1~[ii_x,:] (synthesis.jl:42) when B1 is not diagonal.
 * grp_ptr NULL clears the groups (single columns). */
int sls_set_groups(sls_handle* h, int64_t ngroups, const int64_t* grp_ptr);

/* 𝓢ₓ[t] (Nx x Nx) and 𝓢ᵤ[t] (Nu x Nx), t = 1..T: arrays of T pointers.  *_nzval (Bool) may be NULL
 * (= all true).  Index sets use the stored pattern of 𝓢[T] (reduction.jl:14); masks use nzval
 * (synthesis.jl:58-59). */
int sls_set_supports(sls_handle* h, int64_t T,
                     const int64_t* const* Sx_colptr, const int64_t* const* Sx_rowval, const uint8_t* const* Sx_nzval,
                     const int64_t* const* Su_colptr, const int64_t* const* Su_rowval, const uint8_t* const* Su_nzval);

/* Alternative to sls_set_supports: build the (d,T)-localized supports of the README on the device from the plant
 * already set (README.md:52-54):  𝓢ₓ[t] = (A .≠ 0)^min(d, floor(α(t-1))) .≠ 0,
 *                                 𝓢ᵤ[t] = (B₂' .≠ 0)*(A .≠ 0)^min(d+1, floor(α(t-1))) .≠ 0,   t = 1..T.
 * Nothing but d, T, α crosses the bus.  nnz_x / nnz_u (T each, may be NULL) return the stored entries per tap so
 * that sls_get_supports can copy a tap back in Julia layout (all stored values are `true`). */
int sls_make_supports(sls_handle* h, int64_t d, int64_t T, double alpha, int64_t* nnz_x, int64_t* nnz_u);
int sls_get_supports(sls_handle* h, int which /*0=𝓢ₓ,1=𝓢ᵤ*/, int64_t t /*0-based tap*/,
                     int64_t* colptr, int64_t* rowval, uint8_t* nzval);

/* Kernel 1 on the column list `cols` (1-based disturbance columns, the reference's 𝓘 keyword with
 * single-column groups, synthesis.jl:11,15; NULL = 1..Nx).  Writes per-tap stored-entry counts of
 * Φx[t] / Φu[t] (T each) so the caller can allocate exact-size outputs, and optionally n=|s_x|,
 * m=|s_u| per listed column. */
int sls_plan(sls_handle* h, const int64_t* cols, int64_t ncols,
             int64_t* nnz_x, int64_t* nnz_u, int32_t* n_out, int32_t* m_out);

/* sparsity_dim_reduction (reduction.jl:11-27) for one column group cj (1-based, q >= 1):
 * s_x, s_u in the reference's first-occurrence order (1-based), ii_x Bool per s_x entry.
 * Ĩ is I(|s_x|)[:, ii_x] (reduction.jl:23) and is implied by ii_x. */
int sls_reduce(sls_handle* h, const int64_t* cj, int64_t q,
               int64_t* sx_out, int64_t sx_cap, int64_t* n_sx,
               int64_t* su_out, int64_t su_cap, int64_t* n_su,
               uint8_t* iix_out);

/* Output-feedback branch of sparsity_dim_reduction (reduction.jl:16-18):
 *     s_x, s_u, s_y = (unique(findnz(𝓢ⱼ[end][:,cⱼ])[1]) for 𝓢ⱼ in 𝓢),   𝓢 = [𝓢ₓ, 𝓢ᵤ, 𝓢ᵧ]
 * — the stored rows of the listed columns of each family's LAST tap, first occurrences in CSC order (no product with the
 * pattern of A on this branch).  The caller passes that last tap per family (Julia colptr of ncols_S+1 entries, rowval;
 * nrows[f] rows); s_out[f] (capacity s_cap[f], 1-based) may be NULL to query n_s[f].  iix_out (n_s[0] Bools, may be NULL)
 * is indexin(s_x, cⱼ) .≠ nothing (reduction.jl:22); Ĩ = [I(|s_x|)[:,ii_x] 0] (reduction.jl:23) is implied by it.
 * No output-feedback synthesis exists in the reference (SLS_𝓗₂ returns nothing for such plants, synthesis.jl:13,29-31) and
 * its L⁺ regulariser hook is identically 0 (synthesis.jl:21,52), so nothing else of that branch crosses the ABI. */
int sls_reduce_of(sls_handle* h, int64_t nfam, const int64_t* const* S_colptr, const int64_t* const* S_rowval,
                  const int64_t* nrows, int64_t ncols_S, const int64_t* cj, int64_t q,
                  int64_t* const* s_out, const int64_t* s_cap, int64_t* n_s, uint8_t* iix_out);

/* After sls_plan: the local sub-plant blocks of listed column `k` (0-based position in the plan's
 * list): dense column-major A~ = A[s_x,s_x] (n x n), B2~ = B2[s_x,s_u] (n x m), and the free-variable
 * masks Mx (T x n), Mu (T x m) row-major Bool.  Test/inspection entry point for kernel 1. */
int sls_local_problem(sls_handle* h, int64_t k, double* At, double* B2t, uint8_t* Mx, uint8_t* Mu);

/* Kernels 2+3 on the planned columns.  Outputs: T CSC matrices Φx[t] (Nx x Nx) and Φu[t] (Nu x Nx);
 * colptr arrays have Nx+1 entries (columns outside the plan are empty), rowval/nzval have the sizes
 * returned by sls_plan.  status (ncols, may be NULL) and resid (ncols max-abs achievability
 * residual, may be NULL) are per planned column.  Stored pattern = full in-mask pattern (explicit
 * zeros kept; the reference's `+=` at synthesis.jl:67 would drop exact 0.0). */
int sls_solve(sls_handle* h,
              int64_t* const* Px_colptr, int64_t* const* Px_rowval, double* const* Px_nzval,
              int64_t* const* Pu_colptr, int64_t* const* Pu_rowval, double* const* Pu_nzval,
              int32_t* status, double* resid);

/* One-shot, pipelined form of sls_set_supports + sls_plan(all columns) + sls_solve — what `SLS_𝓗₂(P, [𝓢ₓ,𝓢ᵤ])` with the
 * default 𝓘 = 1..Nx does (src/synthesis.jl:11-32).  The columns are cut into contiguous chunks and the host->device
 * copy of chunk c+1's support columns, kernels 1-3 of chunk c and the device->host copy of chunk c-1's CSC segments
 * overlap; results are bit-for-bit those of the three-call path.  Because the stored pattern of Φ[t] is a subset of
 * the stored pattern of 𝓢[t], the caller sizes rowval/nzval of tap t by cap_x[t] / cap_u[t] (NULL = nnz(𝓢[t])) and
 * receives the stored-entry counts in nnz_x / nnz_u (T each; Julia: `resize!`).  colptr arrays have Nx+1 entries.
 * Pinned (page-locked) host buffers make the copies truly asynchronous; pageable buffers work but serialise.
 * first_col (1-based), ncols: a contiguous column range — one GPU's shard of the `@distributed` loop (synthesis.jl:16,
 * 23-27); ncols <= 0 means all of 1..Nx.  Only the support columns of the range and its halo (the columns reached
 * through nz(A[:,j])) are read and uploaded; status / resid have ncols entries; columns outside the range are empty.
 * Afterwards device-resident accessors (sls_device_output, sls_simulate) and plans of other columns need the
 * three-call path (sls_set_supports, sls_plan, sls_solve_device). */
int sls_synthesize(sls_handle* h, int64_t T,
                   const int64_t* const* Sx_colptr, const int64_t* const* Sx_rowval, const uint8_t* const* Sx_nzval,
                   const int64_t* const* Su_colptr, const int64_t* const* Su_rowval, const uint8_t* const* Su_nzval,
                   int64_t first_col, int64_t ncols, const int64_t* cap_x, const int64_t* cap_u,
                   int64_t* const* Px_colptr, int64_t* const* Px_rowval, double* const* Px_nzval,
                   int64_t* const* Pu_colptr, int64_t* const* Pu_rowval, double* const* Pu_nzval,
                   int64_t* nnz_x, int64_t* nnz_u, int32_t* status, double* resid);

/* Shard form of sls_synthesize — the shape of one `@distributed` chunk (src/synthesis.jl:16,23-27) without any array that
 * grows with the total Nx.  The supports are given as a column WINDOW [win_first, win_first + win_ncols) (1-based) that
 * must cover the range and its halo (the rows of nz(A[:,j]) for j in the range): every colptr array has win_ncols + 1
 * entries whose values are 1-based offsets into the rowval / nzval pointers beside it (Julia: `pointer(𝓢[t].colptr,
 * win_first)` with the full rowval, or the fields of the slice `𝓢[t][:, win]`).  The outputs are the column SLAB
 * Φ[t][:, first_col : first_col+ncols-1]: colptr arrays of ncols + 1 entries starting at 1, rowval (global rows) and
 * nzval sized by cap_x / cap_u (NULL = nnz of the window).  Everything else is as in sls_synthesize.
 * Sub-plant shards: when the handle's plant is itself the sub-plant P[pw, pw] of a larger plant (a rank that owns the
 * columns pw ⊇ range ∪ every s_x of the range sets A[pw,pw], B2[pw,uw] with sls_set_plant), the supports may keep the
 * row numbers of the large plant: row_off_x / row_off_u (>= 0) are subtracted from every support row on the device
 * (local row = stored row - row_off; a row that falls outside 1..Nx / 1..Nu of the handle's plant is a layout error)
 * and added to every output row, so the slabs of all ranks concatenate into the global Φ[t] without a host pass over
 * rowval.  Column numbers (win_first, first_col) are those of the handle's plant.  Per-GPU host work, device memory and
 * bytes over PCIe then depend on the shard only, not on the global Nx. */
int sls_synthesize_window(sls_handle* h, int64_t T, int64_t win_first, int64_t win_ncols, int64_t row_off_x, int64_t row_off_u,
                          const int64_t* const* Sx_colptr, const int64_t* const* Sx_rowval, const uint8_t* const* Sx_nzval,
                          const int64_t* const* Su_colptr, const int64_t* const* Su_rowval, const uint8_t* const* Su_nzval,
                          int64_t first_col, int64_t ncols, const int64_t* cap_x, const int64_t* cap_u,
                          int64_t* const* Px_colptr, int64_t* const* Px_rowval, double* const* Px_nzval,
                          int64_t* const* Pu_colptr, int64_t* const* Pu_rowval, double* const* Pu_nzval,
                          int64_t* nnz_x, int64_t* nnz_u, int32_t* status, double* resid);

/* Page-lock (cudaHostRegister) / release a caller-owned host range: the copy pipelines of sls_solve and sls_synthesize*
 * only overlap with the kernels when the host buffers are page-locked; pageable buffers are correct but serialise.
 * Ranges still registered are released by sls_destroy. */
int sls_host_register(sls_handle* h, void* p, size_t bytes);
int sls_host_unregister(sls_handle* h, void* p);

/* Device-resident variants: run kernels 1-3 without any host<->device payload copy (inputs must
 * already be set).  sls_solve_device leaves Φ in handle-owned device buffers; sls_device_output
 * returns their device addresses (for NCCL all-gather / the FIR simulation / torch wrapping). */
int sls_solve_device(sls_handle* h);
int sls_device_output(sls_handle* h, int64_t t, int which /*0=Φx,1=Φu*/,
                      const int64_t** d_colptr, const int64_t** d_rowval, const double** d_nzval, int64_t* nnz);
int sls_device_status(sls_handle* h, const int32_t** d_status, const double** d_resid);

int sls_get_timing(const sls_handle* h, sls_timing* out);

/* Batched closed-loop FIR simulation (README.md:62-71) of the controller left on the device by the last
 * sls_solve / sls_solve_device: scenario s applies w(t) = delta(t - dist_time[s]) e_{dist_col[s]} (both 1-based)
 * and runs x, beta, u for `steps` steps from rest.  Outputs (host, any may be NULL): xnorm[s*steps + t-1] = ||x(t)||_2,
 * unorm likewise for u(t); x_out / u_out are dense [nscen][steps][Nx] / [nscen][steps][Nu] trajectories (small cases).
 * Every disturbance column's neighbourhood must be among the planned columns (unplanned columns have Phi = 0). */
int sls_simulate(sls_handle* h, int64_t nscen, int64_t steps, const int64_t* dist_col, const int64_t* dist_time,
                 double* xnorm, double* unorm, double* x_out, double* u_out);

/* Device-side stopwatch on the handle's stream (CUDA events): sls_mark(h,0) ... sls_mark(h,1), then
 * sls_elapsed_ms waits for mark 1 and returns the device time between the two marks.  bench.py brackets a
 * whole step (plan + solve, with or without the host copies) with it. */
int sls_mark(sls_handle* h, int which);
int sls_elapsed_ms(sls_handle* h, float* ms);

/* Calibration for the roofline of kernel 2: measured FP64 throughput of this device in TFLOP/s — a
 * register-resident DFMA chain (vector pipe) and a DMMA m8n8k4 chain (FP64 tensor path).  Not used by
 * any synthesis call. */
int sls_fp64_peak(int device, int iters, double* dfma_tflops, double* dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif /* SLSB200_H */
