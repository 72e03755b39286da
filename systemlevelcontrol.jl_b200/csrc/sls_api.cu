// sls_api.cu — host side of the C ABI declared in include/slsb200.h.
//
// Owns the device copies of the plant (A, B2 and their transposes), of the supports 𝓢ₓ[t], 𝓢ᵤ[t], and the
// per-plan metadata; sequences kernel 1 (sls_gather.cu), kernel 2 (sls_solve.cu) and kernel 3
// (sls_scatter.cu) on one stream and times them with CUDA events.  It replaces the per-chunk worker
// `_SLS_𝓗₂` of the reference (src/synthesis.jl:34-72); the Julia-side `SLS_𝓗₂` (src/synthesis.jl:11-32)
// keeps its signature and calls these entry points.
// There is no CPU path here: without a usable CUDA device every entry point fails with an error.
#include "../../include/slsb200.h"
#include "sls_kernels.h"
#include <vector>
#include <string>
#include <algorithm>
#include <new>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <atomic>
#include <mutex>
#include <condition_variable>
#include <deque>
#include <cstdint>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t n) {
        if (n <= bytes && p) return 0;
        rt::dfree(p); p = nullptr; bytes = 0;
        int e = rt::alloc(&p, n);
        if (!e) bytes = n ? n : 1;
        return e;
    }
    void release() { rt::dfree(p); p = nullptr; bytes = 0; }
    template <class T> T* as() const { return reinterpret_cast<T*>(p); }
};

template <class T> struct PinVec {     // pinned host staging array (fast, truly asynchronous D2H)
    T* p = nullptr;
    size_t cap = 0, len = 0;
    int resize(size_t n) {
        len = n;
        if (n <= cap && p) return 0;
        rt::host_free(p); p = nullptr; cap = 0;
        int e = rt::host_alloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
        if (!e) cap = std::max<size_t>(n, 1);
        return e;
    }
    void release() { rt::host_free(p); p = nullptr; cap = len = 0; }
    T& operator[](size_t i) { return p[i]; }
    const T& operator[](size_t i) const { return p[i]; }
    T* data() { return p; }
};

struct DevCsc {                 // device copy of one Julia SparseMatrixCSC{Float64,Int64}
    DevBuf colptr, rowval, nzval;
    int64_t ncol = 0, nnz = 0;
    SlsCsc view() const { return SlsCsc{colptr.as<int64_t>(), rowval.as<int64_t>(), nzval.as<double>()}; }
    void release() { colptr.release(); rowval.release(); nzval.release(); }
};

struct DevSupports {            // T Bool CSC matrices, taps concatenated
    DevBuf colptr, rowval, nzval, tap_off;
    bool has_nzval = false;
    int64_t nnz = 0;
    int64_t row_off = 0;        // stored row - row_off = plant row (sls_synthesize_window on a sub-plant)
    int maxlen_last = 0;        // longest column of the last tap (bounds kernel 1's candidate list)
    std::vector<int64_t> h_tap_off;   // host copy of tap_off (sls_get_supports)
    SlsSupports view() const {
        return SlsSupports{colptr.as<int64_t>(), rowval.as<int64_t>(), has_nzval ? nzval.as<uint8_t>() : nullptr,
                           tap_off.as<int64_t>(), row_off};
    }
    void release() { colptr.release(); rowval.release(); nzval.release(); tap_off.release(); }
};

struct SolveClass { int N, count, mcap, nnzcap; size_t list_off; };

struct Plan {
    bool valid = false;
    int ncols = 0, nstride = 1, mstride = 1, wx = 1, wu = 1, nmax = 0, mmax = 0, cap = 32;
    DevBuf cols, n, m, sx, su, maskx, masku, jloc, lnnz, lnnzA, lnnzB, pre, tot, colptr, lists, keepx, keepu, stats, scanws;
    PinVec<int32_t> h_stats;            // [36] plan statistics computed by kernel 1a (sls_gather.cu)
    bool sizes_on_host = false, lists_on_host = false;   // h_n / h_m and h_lists are fetched on demand
    PinVec<int32_t> h_cols, h_lists;    // pinned: read by kernel copies (sls_kcopy)
    bool cols_sorted = true;            // planned columns strictly ascending => a planned-index range is a CSC column range
    bool cols_contiguous = true;        // planned columns are first, first+1, ...: the list is generated on the device
    bool sets_are_runs = false;         // every s_x / s_u of the plan is a run of consecutive rows (kernel 1a, stats[3])
    PinVec<int32_t> h_status; PinVec<double> h_resid;
    PinVec<int64_t> h_seg;              // [(nchunk+1)][2T] output offsets of the D2H pipeline's chunk boundaries
    PinVec<int32_t> h_n, h_m, h_lnnz, h_lA, h_lB;
    PinVec<int64_t> h_tot, h_last;
    std::vector<int64_t> nnz_tap;       // 2T stored-entry counts (Φx taps then Φu taps)
    std::vector<int64_t> out_off;       // 2T+1 offsets of each tap inside the output buffers
    std::vector<SolveClass> classes;
    int64_t tot_packed = 0;
    double alg_flops = 0, alg_bytes = 0;
    bool stats_valid = false;           // alg_flops / alg_bytes are computed on demand (plan_stats)
    float gpu_ms = 0;                   // device time of kernels 1a + 1b + scan (without the host round trips between them)
    int64_t col_lo = 0, col_hi = 0;     // planned columns lie in [col_lo, col_hi)
    int identity_cols = 0;              // h_cols currently holds 0..identity_cols-1 (the cols == NULL plan)
    void release() {
        DevBuf* all[] = {&cols, &n, &m, &sx, &su, &maskx, &masku, &jloc, &lnnz, &lnnzA, &lnnzB, &pre, &tot, &colptr, &lists, &keepx, &keepu, &stats, &scanws};
        h_stats.release();
        for (DevBuf* b : all) b->release();
        h_seg.release(); h_status.release(); h_resid.release(); h_cols.release(); h_lists.release();
        h_n.release(); h_m.release(); h_lnnz.release(); h_lA.release(); h_lB.release(); h_tot.release(); h_last.release();
        valid = false;
    }
};

// one column chunk of the pipelined one-shot path (sls_synthesize): its own plan and output buffers, so that the
// chunk's D2H copies never alias a later chunk's kernels
struct PipeCtx {
    Plan plan;
    DevBuf out_rowval, out_nzval, out_ptrs, status, resid, base;
    PinVec<void*> h_ptrs; PinVec<int64_t> h_base;
    rt::Event ev_plan, ev_scat, ev_solved, ev_out;
    int create() { int e = ev_plan.create(); if (!e) e = ev_scat.create(); if (!e) e = ev_solved.create(); if (!e) e = ev_out.create(); return e; }
    void release() {
        plan.release();
        DevBuf* all[] = {&out_rowval, &out_nzval, &out_ptrs, &status, &resid, &base};
        for (DevBuf* b : all) b->release();
        h_ptrs.release(); h_base.release();
        ev_plan.destroy(); ev_scat.destroy(); ev_solved.destroy(); ev_out.destroy();
    }
};
#define SLS_PIPE_MAX 64

}  // namespace

struct sls_handle {
    sls_options opt{};
    int device = 0, sm_count = 0, max_smem = 0;
    sls_stream_t stream{}, copy_stream{};
    rt::Event chunk_ev[16];
    std::string err;
    int64_t Nx = 0, Nu = 0, T = 0;
    bool have_plant = false, have_supports = false, solved = false;
    DevCsc A, B2, Ar, B2r;
    int maxcolA = 0, maxcolB2 = 0;
    DevSupports Sx, Su;
    Plan plan;
    DevBuf xw, hist, status, resid, out_rowval, out_nzval, out_ptrs;
    rt::Event ev[8];
    rt::Event marks[2];
    sls_timing tm{};
    int pipe_chunks = 0;                // chunks of the last sls_synthesize (their plans hold its statistics)
    bool tm_from_plan = true;           // alg_flops / alg_bytes of tm come from h->plan (lazily) rather than from sls_synthesize
    // pipelined one-shot path
    sls_stream_t in_stream{}, scat_stream{}, out2_stream{};
    rt::Event pipe_done2;
    rt::Event pipe_in[SLS_PIPE_MAX];
    rt::Event pipe_done;
    std::vector<PipeCtx*> pipe;
    DevBuf pipe_colptr;
    PinVec<int32_t> pipe_status; PinVec<double> pipe_resid;
    PinVec<int64_t> t_cp[2], t_rv[2]; PinVec<double> t_nz[2]; std::vector<int64_t> t_fill;   // transposes of A, B2 (set_plant)
    bool have_weights = false;          // diagonal cost weights (sls_set_weights)
    std::vector<double> h_wx, h_wu;
    DevBuf d_wx, d_wu;
    std::vector<int32_t> h_Akmax, h_Akmin;   // largest / smallest row index (0-based) stored in each column of A
    int64_t supp_lo = 0, supp_hi = 0;        // support columns [supp_lo, supp_hi) (0-based) are valid on the device
    DevBuf d_flag; PinVec<int32_t> h_flag;   // set by sls_validate_supports_kernel (malformed support rows)
    rt::Event gev[4];                        // around kernel 1a and kernels 1b + scan (gather_gpu_ms)
    std::vector<void*> registered;           // host ranges page-locked through sls_host_register
    DevBuf supp[8]; PinVec<int64_t> supp_h; int supp_cap = 0;   // scratch of sls_make_supports (kept across calls)
    DevBuf simb[19]; int sim_cap = 0;                           // scratch of sls_simulate (kept across calls)
    // general cost / column groups (sls_set_cost, sls_set_groups -> kernel 2g)
    DevCsc C1, D12, D11, B1;
    bool have_cost = false, have_d11 = false, have_b1 = false, have_groups = false;
    int64_t Nz = 0, Nw = 0;
    std::vector<int32_t> h_grp; DevBuf d_grp, gen_scratch, gen_stats; PinVec<int32_t> h_gen_stats;
};

static thread_local std::string g_create_err;

static int fail(sls_handle* h, int code, const char* what) {
    std::string msg = std::string(what);
    if (code > 0) { msg += ": "; msg += rt::err_str(code); }
    if (h) h->err = msg; else g_create_err = msg;
    return code;
}
#define SLS_CK(expr) do { int e_ = (expr); if (e_) return fail(h, e_, #expr); } while (0)
#define SLS_ARG(cond, msg) do { if (!(cond)) return fail(h, -1, msg); } while (0)

// ---------------------------------------------------------------------------------------------
extern "C" int sls_abi_version(void) { return SLS_ABI_VERSION; }
extern "C" int sls_device_count(void) { return rt::device_count(); }

extern "C" const char* sls_last_error(const sls_handle* h) { return h ? h->err.c_str() : g_create_err.c_str(); }

extern "C" int sls_create(const sls_options* opts, sls_handle** out) {
    sls_handle* h = nullptr;
    if (!out) return fail(nullptr, -1, "sls_create: out is NULL");
    *out = nullptr;
    const int ndev = rt::device_count();
    if (ndev <= 0) return fail(nullptr, ndev < 0 ? -ndev : 100, "sls_create: no usable CUDA device (this library has no CPU path)");
    h = new (std::nothrow) sls_handle();
    if (!h) return fail(nullptr, -2, "sls_create: out of host memory");
    if (opts) h->opt = *opts;
    if (h->opt.rank_tol <= 0) h->opt.rank_tol = 1e-10;
    // kernel 2 takes refinement sweeps while the residual exceeds feas_tol / 10 (1e-13: 15x the largest residual of the
    // BASELINE configs, so those never refine); what is still above feas_tol afterwards is infeasible, rank-deficient beyond
    // rank_tol or has cond(G) >~ 1e6 and is flagged (tests/parity_cases.py::check_random_banded)
    if (h->opt.feas_tol <= 0) h->opt.feas_tol = 1e-12;
    h->device = h->opt.device;
    if (h->device < 0 || h->device >= ndev) { delete h; return fail(nullptr, -1, "sls_create: device ordinal out of range"); }
    int e = rt::set_device(h->device);
    if (!e) e = rt::stream_create(&h->stream);
    if (!e) e = rt::stream_create(&h->copy_stream);
    if (!e) e = rt::stream_create(&h->in_stream);
    if (!e) e = rt::stream_create(&h->scat_stream);
    if (!e) e = rt::stream_create(&h->out2_stream);
    if (!e) e = h->pipe_done2.create();
    for (int i = 0; i < SLS_PIPE_MAX && !e; i++) e = h->pipe_in[i].create();
    if (!e) e = h->pipe_done.create();
    for (int i = 0; i < 16 && !e; i++) e = h->chunk_ev[i].create();
    for (int i = 0; i < 8 && !e; i++) e = h->ev[i].create();
    for (int i = 0; i < 2 && !e; i++) e = h->marks[i].create();
    for (int i = 0; i < 4 && !e; i++) e = h->gev[i].create();
    if (!e) e = h->d_flag.ensure(2 * sizeof(int32_t));          // [0] malformed-supports flag, [1] refined-columns counter
    if (!e) e = h->h_flag.resize(2);
    if (!e) e = rt::zero(h->d_flag.p, 2 * sizeof(int32_t), h->stream);
    if (!e) e = rt::sync(h->stream);
    if (e) { delete h; return fail(nullptr, e, "sls_create: CUDA initialisation failed"); }
    h->sm_count = rt::sm_count(h->device);
    h->max_smem = rt::max_smem_optin(h->device);
    *out = h;
    return 0;
}

extern "C" void sls_destroy(sls_handle* h) {
    if (!h) return;
    rt::set_device(h->device);
    rt::sync(h->stream);
    rt::sync(h->copy_stream);
    rt::sync(h->in_stream);
    rt::sync(h->scat_stream);
    rt::sync(h->out2_stream);
    for (PipeCtx* x : h->pipe) { x->release(); delete x; }
    h->pipe.clear();
    h->pipe_colptr.release(); h->pipe_status.release(); h->pipe_resid.release();
    h->d_wx.release(); h->d_wu.release(); h->d_flag.release(); h->h_flag.release();
    for (DevBuf& b : h->supp) b.release();
    for (DevBuf& b : h->simb) b.release();
    h->supp_h.release();
    for (void* q : h->registered) rt::host_unregister(q);
    h->registered.clear();
    for (int i = 0; i < 4; i++) h->gev[i].destroy();
    for (int i = 0; i < 2; i++) { h->t_cp[i].release(); h->t_rv[i].release(); h->t_nz[i].release(); }
    for (int i = 0; i < SLS_PIPE_MAX; i++) h->pipe_in[i].destroy();
    h->pipe_done.destroy();
    rt::stream_destroy(h->in_stream);
    rt::stream_destroy(h->scat_stream);
    rt::stream_destroy(h->out2_stream);
    h->pipe_done2.destroy();
    h->A.release(); h->B2.release(); h->Ar.release(); h->B2r.release();
    h->Sx.release(); h->Su.release(); h->plan.release();
    DevBuf* all[] = {&h->xw, &h->hist, &h->status, &h->resid, &h->out_rowval, &h->out_nzval, &h->out_ptrs};
    for (DevBuf* b : all) b->release();
    for (int i = 0; i < 8; i++) h->ev[i].destroy();
    for (int i = 0; i < 2; i++) h->marks[i].destroy();
    for (int i = 0; i < 16; i++) h->chunk_ev[i].destroy();
    h->C1.release(); h->D12.release(); h->D11.release(); h->B1.release();
    h->d_grp.release(); h->gen_scratch.release(); h->gen_stats.release(); h->h_gen_stats.release();
    rt::stream_destroy(h->copy_stream);
    rt::stream_destroy(h->stream);
    delete h;
}

// ---- plant ----------------------------------------------------------------------------------
static const char* check_csc(int64_t nrow, int64_t ncol, const int64_t* cp, const int64_t* rv, bool rows) {
    if (!cp) return "colptr is NULL";
    if (cp[0] != 1) return "colptr[1] must be 1 (Julia 1-based CSC)";
    for (int64_t c = 0; c < ncol; c++) if (cp[c + 1] < cp[c]) return "colptr is not monotone";
    const int64_t nnz = cp[ncol] - 1;
    if (nnz > 0 && !rv) return "rowval is NULL";
    if (rows) {
        for (int64_t c = 0; c < ncol; c++) {
            int64_t prev = 0;
            for (int64_t e = cp[c] - 1; e < cp[c + 1] - 1; e++) {
                if (rv[e] < 1 || rv[e] > nrow) return "rowval out of range";
                if (rv[e] <= prev) return "rowval not strictly ascending inside a column";
                prev = rv[e];
            }
        }
    }
    return nullptr;
}

// CSC of the transpose (== CSR of M), 1-based like the input
// (into pinned scratch owned by the handle: the upload that follows is then truly asynchronous and needs no
// intermediate synchronisation; `fill` is reusable host scratch)
static int transpose_csc(int64_t nrow, int64_t ncol, const int64_t* cp, const int64_t* rv, const double* nz,
                         PinVec<int64_t>& tcp, PinVec<int64_t>& trv, PinVec<double>& tnz, std::vector<int64_t>& fill) {
    const int64_t nnz = cp[ncol] - 1;
    int e_ = tcp.resize(nrow + 1); if (!e_) e_ = trv.resize(std::max<int64_t>(nnz, 1)); if (!e_) e_ = tnz.resize(std::max<int64_t>(nnz, 1));
    if (e_) return e_;
    fill.assign(nrow + 1, 0);
    for (int64_t e = 0; e < nnz; e++) fill[rv[e]]++;          // counts at index row (1-based) -> shifted prefix below
    tcp[0] = 1;
    for (int64_t r = 0; r < nrow; r++) { tcp[r + 1] = tcp[r] + fill[r + 1]; fill[r] = tcp[r]; }
    for (int64_t c = 0; c < ncol; c++)
        for (int64_t e = cp[c] - 1; e < cp[c + 1] - 1; e++) {
            const int64_t dst = fill[rv[e] - 1]++ - 1;
            trv[dst] = c + 1; tnz[dst] = nz[e];
        }
    return 0;
}

static int upload_csc(sls_handle* h, DevCsc& d, int64_t ncol, const int64_t* cp, const int64_t* rv, const double* nz) {
    const int64_t nnz = cp[ncol] - 1;
    d.ncol = ncol; d.nnz = nnz;
    SLS_CK(d.colptr.ensure((ncol + 1) * sizeof(int64_t)));
    SLS_CK(d.rowval.ensure(std::max<int64_t>(nnz, 1) * sizeof(int64_t)));
    SLS_CK(d.nzval.ensure(std::max<int64_t>(nnz, 1) * sizeof(double)));
    SLS_CK(rt::h2d(d.colptr.p, cp, (ncol + 1) * sizeof(int64_t), h->stream));
    SLS_CK(rt::h2d(d.rowval.p, rv, nnz * sizeof(int64_t), h->stream));
    SLS_CK(rt::h2d(d.nzval.p, nz, nnz * sizeof(double), h->stream));
    return 0;
}

extern "C" int sls_set_plant(sls_handle* h, int64_t Nx, int64_t Nu,
                             const int64_t* A_colptr, const int64_t* A_rowval, const double* A_nzval,
                             const int64_t* B2_colptr, const int64_t* B2_rowval, const double* B2_nzval) {
    if (!h) return -1;
    SLS_ARG(Nx >= 1 && Nu >= 0 && Nx < (int64_t)1 << 31 && Nu < (int64_t)1 << 31, "sls_set_plant: bad dimensions");
    SLS_CK(rt::set_device(h->device));
    const char* m = check_csc(Nx, Nx, A_colptr, A_rowval, true);
    if (m) return fail(h, -2, (std::string("sls_set_plant: A: ") + m).c_str());
    m = check_csc(Nx, Nu, B2_colptr, B2_rowval, true);
    if (m) return fail(h, -2, (std::string("sls_set_plant: B2: ") + m).c_str());
    SLS_ARG(A_nzval || A_colptr[Nx] == 1, "sls_set_plant: A nzval is NULL");
    SLS_ARG(B2_nzval || B2_colptr[Nu] == 1, "sls_set_plant: B2 nzval is NULL");
    for (int64_t e = 0; e < A_colptr[Nx] - 1; e++) if (!std::isfinite(A_nzval[e])) return fail(h, -3, "sls_set_plant: A holds a non-finite value");
    for (int64_t e = 0; e < B2_colptr[Nu] - 1; e++) if (!std::isfinite(B2_nzval[e])) return fail(h, -3, "sls_set_plant: B2 holds a non-finite value");
    std::vector<double> As, Bs;                                // similarity-scaled values when cost weights are set
    if (h->have_weights) {
        if ((int64_t)h->h_wx.size() != Nx || (int64_t)h->h_wu.size() != Nu)
            return fail(h, -1, "sls_set_plant: the cost weights were set for other dimensions (call sls_set_weights again)");
        As.resize(A_colptr[Nx] - 1); Bs.resize(B2_colptr[Nu] - 1);
        for (int64_t c = 0; c < Nx; c++)
            for (int64_t e = A_colptr[c] - 1; e < A_colptr[c + 1] - 1; e++) As[e] = A_nzval[e] * h->h_wx[A_rowval[e] - 1] / h->h_wx[c];
        for (int64_t c = 0; c < Nu; c++)
            for (int64_t e = B2_colptr[c] - 1; e < B2_colptr[c + 1] - 1; e++) Bs[e] = B2_nzval[e] * h->h_wx[B2_rowval[e] - 1] / h->h_wu[c];
        A_nzval = As.data(); B2_nzval = Bs.data();
    }
    h->ev[0].record(h->stream);
    SLS_CK(upload_csc(h, h->A, Nx, A_colptr, A_rowval, A_nzval));
    SLS_CK(upload_csc(h, h->B2, Nu, B2_colptr, B2_rowval, B2_nzval));
    SLS_CK(transpose_csc(Nx, Nx, A_colptr, A_rowval, A_nzval, h->t_cp[0], h->t_rv[0], h->t_nz[0], h->t_fill));
    SLS_CK(upload_csc(h, h->Ar, Nx, h->t_cp[0].data(), h->t_rv[0].data(), h->t_nz[0].data()));
    SLS_CK(transpose_csc(Nx, Nu, B2_colptr, B2_rowval, B2_nzval, h->t_cp[1], h->t_rv[1], h->t_nz[1], h->t_fill));
    SLS_CK(upload_csc(h, h->B2r, Nx, h->t_cp[1].data(), h->t_rv[1].data(), h->t_nz[1].data()));
    h->ev[1].record(h->stream);
    int mc = 0;
    for (int64_t c = 0; c < Nx; c++) mc = std::max<int>(mc, (int)(A_colptr[c + 1] - A_colptr[c]));
    h->maxcolA = mc;
    mc = 0;
    for (int64_t c = 0; c < Nu; c++) mc = std::max<int>(mc, (int)(B2_colptr[c + 1] - B2_colptr[c]));
    h->maxcolB2 = mc;
    h->h_Akmax.assign(Nx, 0); h->h_Akmin.assign(Nx, 0);
    for (int64_t c = 0; c < Nx; c++) {
        const bool any = A_colptr[c + 1] > A_colptr[c];
        h->h_Akmax[c] = any ? (int32_t)(A_rowval[A_colptr[c + 1] - 2] - 1) : (int32_t)c;
        h->h_Akmin[c] = any ? (int32_t)(A_rowval[A_colptr[c] - 1] - 1) : (int32_t)c;
    }
    SLS_CK(rt::sync(h->stream));                              // As/Bs and the caller's arrays must outlive the copies
    h->tm.h2d_ms = rt::elapsed_ms(h->ev[0], h->ev[1]);
    if (h->have_supports && (h->Nx != Nx || h->Nu != Nu)) { h->have_supports = false; }
    h->Nx = Nx; h->Nu = Nu; h->have_plant = true; h->plan.valid = false; h->solved = false;
    return 0;
}

// Diagonal cost weights: cost = sum_t ||wx .* x_t||^2 + ||wu .* u_t||^2, i.e. C1 = [diag(wx); 0], D12 = [0; diag(wu)]
// (GeneralizedPlant.jl:101-103; synthesis.jl:50,76-77).  Handled by similarity scaling (sls_scatter.cu: sls_unscale_kernel).
extern "C" int sls_set_weights(sls_handle* h, int64_t Nx, int64_t Nu, const double* wx, const double* wu) {
    if (!h) return -1;
    SLS_CK(rt::set_device(h->device));
    h->have_plant = false; h->plan.valid = false; h->solved = false;   // the plant values on the device depend on the weights
    if (!wx && !wu) { h->have_weights = false; return 0; }
    SLS_ARG(Nx >= 1 && Nu >= 0, "sls_set_weights: bad dimensions");
    h->h_wx.assign(Nx, 1.0); h->h_wu.assign(std::max<int64_t>(Nu, 1), 1.0);
    if (wx) for (int64_t i = 0; i < Nx; i++) { if (!(wx[i] > 0) || !std::isfinite(wx[i])) return fail(h, -3, "sls_set_weights: weights must be positive and finite"); h->h_wx[i] = wx[i]; }
    if (wu) for (int64_t i = 0; i < Nu; i++) { if (!(wu[i] > 0) || !std::isfinite(wu[i])) return fail(h, -3, "sls_set_weights: weights must be positive and finite"); h->h_wu[i] = wu[i]; }
    h->h_wu.resize(Nu);
    SLS_CK(h->d_wx.ensure(Nx * sizeof(double))); SLS_CK(h->d_wu.ensure(std::max<int64_t>(Nu, 1) * sizeof(double)));
    SLS_CK(rt::h2d(h->d_wx.p, h->h_wx.data(), Nx * sizeof(double), h->stream));
    SLS_CK(rt::h2d(h->d_wu.p, h->h_wu.data(), Nu * sizeof(double), h->stream));
    SLS_CK(rt::sync(h->stream));
    h->have_weights = true;
    return 0;
}

// General cost: the 6-argument Plant(A,B1,B2,C1,D11,D12) (GeneralizedPlant.jl:101-103) with arbitrary sparse
// C1 (Nz x Nx), D12 (Nz x Nu), D11 (Nz x Nw, NULL = 0) and B1 (Nx x Nw, NULL = I) in the objective
// || [C1~ D12~] [Φx~;Φu~] B1~[ii_x,:] + D11~ ||  (synthesis.jl:42-43,50,76-77).  Solved by kernel 2g (sls_solve_general.cu).
static int upload_cost_csc(sls_handle* h, DevCsc& d, const char* name, int64_t nrow, int64_t ncol,
                           const int64_t* cp, const int64_t* rv, const double* nz) {
    const char* m = check_csc(nrow, ncol, cp, rv, true);
    if (m) return fail(h, -2, (std::string("sls_set_cost: ") + name + ": " + m).c_str());
    if (!nz && cp[ncol] != 1) return fail(h, -1, (std::string("sls_set_cost: ") + name + " nzval is NULL").c_str());
    for (int64_t e = 0; e < cp[ncol] - 1; e++)
        if (!std::isfinite(nz[e])) return fail(h, -3, (std::string("sls_set_cost: ") + name + " holds a non-finite value").c_str());
    return upload_csc(h, d, ncol, cp, rv, nz);
}

extern "C" int sls_set_cost(sls_handle* h, int64_t Nz, int64_t Nw,
                            const int64_t* C1_colptr, const int64_t* C1_rowval, const double* C1_nzval,
                            const int64_t* D12_colptr, const int64_t* D12_rowval, const double* D12_nzval,
                            const int64_t* D11_colptr, const int64_t* D11_rowval, const double* D11_nzval,
                            const int64_t* B1_colptr, const int64_t* B1_rowval, const double* B1_nzval) {
    if (!h) return -1;
    SLS_CK(rt::set_device(h->device));
    h->solved = false;
    if (!C1_colptr && !D12_colptr && !D11_colptr && !B1_colptr) { h->have_cost = h->have_d11 = h->have_b1 = false; return 0; }
    SLS_ARG(h->have_plant, "sls_set_cost: call sls_set_plant first (the cost matrices are sized by Nx, Nu)");
    SLS_ARG(!h->have_weights, "sls_set_cost: diagonal weights (sls_set_weights) are installed; put them into C1 / D12 instead");
    SLS_ARG((C1_colptr != nullptr) == (D12_colptr != nullptr), "sls_set_cost: C1 and D12 must be given together");
    SLS_ARG(Nz >= 0 && Nw >= 0 && Nz < (int64_t)1 << 31 && Nw < (int64_t)1 << 31, "sls_set_cost: bad dimensions");
    h->have_cost = h->have_d11 = h->have_b1 = false;
    if (C1_colptr) {
        if (int e = upload_cost_csc(h, h->C1, "C1", Nz, h->Nx, C1_colptr, C1_rowval, C1_nzval)) return e;
        if (int e = upload_cost_csc(h, h->D12, "D12", Nz, h->Nu, D12_colptr, D12_rowval, D12_nzval)) return e;
    }
    if (D11_colptr) { if (int e = upload_cost_csc(h, h->D11, "D11", Nz, Nw, D11_colptr, D11_rowval, D11_nzval)) return e; }
    if (B1_colptr) { if (int e = upload_cost_csc(h, h->B1, "B1", h->Nx, Nw, B1_colptr, B1_rowval, B1_nzval)) return e; }
    SLS_CK(rt::sync(h->stream));
    h->have_cost = C1_colptr != nullptr; h->have_d11 = D11_colptr != nullptr && D11_colptr[Nw] > 1; h->have_b1 = B1_colptr != nullptr;
    h->Nz = Nz; h->Nw = Nw;
    return 0;
}

// Column groups (the 𝓘 keyword, synthesis.jl:15): group g = planned columns grp_ptr[g] .. grp_ptr[g+1]-1 of the NEXT
// sls_solve (0-based offsets into the column list given to sls_plan; each group's columns ascending).  NULL clears.
extern "C" int sls_set_groups(sls_handle* h, int64_t ngroups, const int64_t* grp_ptr) {
    if (!h) return -1;
    SLS_CK(rt::set_device(h->device));
    h->solved = false;
    if (!grp_ptr) { h->have_groups = false; h->h_grp.clear(); return 0; }
    SLS_ARG(ngroups >= 0 && ngroups < (int64_t)1 << 31, "sls_set_groups: bad group count");
    SLS_ARG(grp_ptr[0] == 0, "sls_set_groups: grp_ptr[0] must be 0");
    for (int64_t g = 0; g < ngroups; g++)
        SLS_ARG(grp_ptr[g + 1] > grp_ptr[g] && grp_ptr[g + 1] < (int64_t)1 << 31, "sls_set_groups: every group needs at least one column");
    h->h_grp.assign(grp_ptr, grp_ptr + ngroups + 1);
    SLS_CK(h->d_grp.ensure((size_t)(ngroups + 1) * sizeof(int32_t)));
    SLS_CK(rt::h2d(h->d_grp.p, h->h_grp.data(), (size_t)(ngroups + 1) * sizeof(int32_t), h->stream));
    SLS_CK(rt::sync(h->stream));
    h->have_groups = true;
    return 0;
}

// ---- supports -------------------------------------------------------------------------------
static int upload_supports(sls_handle* h, DevSupports& d, int64_t T, int64_t nrow, int64_t ncol,
                           const int64_t* const* cps, const int64_t* const* rvs, const uint8_t* const* nzs,
                           const char* name) {
    std::vector<int64_t> off(T + 1, 0);
    bool any_nz = false, all_nz = true;
    for (int64_t t = 0; t < T; t++) {                          // cheap checks first: enough to size and start the copies
        if (!cps[t]) return fail(h, -2, (std::string("sls_set_supports: ") + name + ": colptr is NULL").c_str());
        if (cps[t][0] != 1 || cps[t][ncol] < 1) return fail(h, -2, (std::string("sls_set_supports: ") + name + ": colptr[1] must be 1 (Julia 1-based CSC)").c_str());
        if (cps[t][ncol] > 1 && !(rvs && rvs[t])) return fail(h, -2, (std::string("sls_set_supports: ") + name + ": rowval is NULL").c_str());
        off[t + 1] = off[t] + cps[t][ncol] - 1;
        if (nzs && nzs[t]) any_nz = true; else all_nz = false;
    }
    if (any_nz && !all_nz) return fail(h, -2, "sls_set_supports: nzval must be given for every tap or for none");
    d.nnz = off[T]; d.has_nzval = any_nz;
    SLS_CK(d.colptr.ensure(T * (ncol + 1) * sizeof(int64_t)));
    SLS_CK(d.rowval.ensure(std::max<int64_t>(d.nnz, 1) * sizeof(int64_t)));
    SLS_CK(d.tap_off.ensure((T + 1) * sizeof(int64_t)));
    if (any_nz) SLS_CK(d.nzval.ensure(std::max<int64_t>(d.nnz, 1)));
    for (int64_t t = 0; t < T; t++) {
        const int64_t nnz = off[t + 1] - off[t];
        SLS_CK(rt::h2d(d.colptr.as<int64_t>() + t * (ncol + 1), cps[t], (ncol + 1) * sizeof(int64_t), h->stream));
        SLS_CK(rt::h2d(d.rowval.as<int64_t>() + off[t], rvs[t], nnz * sizeof(int64_t), h->stream));
        if (any_nz) SLS_CK(rt::h2d(d.nzval.as<uint8_t>() + off[t], nzs[t], nnz, h->stream));
    }
    SLS_CK(rt::h2d(d.tap_off.p, off.data(), (T + 1) * sizeof(int64_t), h->stream));
    d.h_tap_off = off;
    // full layout validation on the host while the copies are in flight (a failure leaves have_supports = false)
    const char* bad = nullptr;
    for (int64_t t = 0; t < T && !bad; t++) bad = check_csc(nrow, ncol, cps[t], rvs ? rvs[t] : nullptr, (h->opt.flags & 1) != 0);
    int ml = 0;
    if (!bad) for (int64_t c = 0; c < ncol; c++) ml = std::max<int>(ml, (int)(cps[T - 1][c + 1] - cps[T - 1][c]));
    d.maxlen_last = ml;
    if (bad) { rt::sync(h->stream); return fail(h, -2, (std::string("sls_set_supports: ") + name + ": " + bad).c_str()); }
    SLS_CK(rt::sync(h->stream));                              // `off` is a local
    return 0;
}

extern "C" int sls_set_supports(sls_handle* h, int64_t T,
                                const int64_t* const* Sx_colptr, const int64_t* const* Sx_rowval, const uint8_t* const* Sx_nzval,
                                const int64_t* const* Su_colptr, const int64_t* const* Su_rowval, const uint8_t* const* Su_nzval) {
    if (!h) return -1;
    SLS_ARG(h->have_plant, "sls_set_supports: call sls_set_plant first");
    SLS_ARG(T >= 1 && T < 4096, "sls_set_supports: bad T");
    SLS_ARG(Sx_colptr && Sx_rowval && Su_colptr && Su_rowval, "sls_set_supports: NULL array of pointers");
    SLS_CK(rt::set_device(h->device));
    h->have_supports = false; h->plan.valid = false; h->solved = false;
    h->Sx.row_off = h->Su.row_off = 0;
    h->ev[0].record(h->stream);
    SLS_CK(rt::zero(h->d_flag.p, sizeof(int32_t), h->stream));
    int e = upload_supports(h, h->Sx, T, h->Nx, h->Nx, Sx_colptr, Sx_rowval, Sx_nzval, "Sx");
    if (e) return e;
    e = upload_supports(h, h->Su, T, h->Nu, h->Nx, Su_colptr, Su_rowval, Su_nzval, "Su");
    if (e) return e;
    h->ev[1].record(h->stream);
    // row indices are checked on the device (range, strictly ascending): ~0.1 ms per 100 MB instead of a host pass
    SLS_CK(sls_launch_validate_supports(h->Sx.view(), (int)T, h->Nx, h->Nx, 0, h->Nx, h->d_flag.as<int32_t>(), h->stream));
    SLS_CK(sls_launch_validate_supports(h->Su.view(), (int)T, h->Nx, h->Nu, 0, h->Nx, h->d_flag.as<int32_t>(), h->stream));
    SLS_CK(sls_kcopy(h->h_flag.data(), h->d_flag.p, sizeof(int32_t), h->stream));
    SLS_CK(rt::sync(h->stream));
    if (h->h_flag[0]) return fail(h, -2, "sls_set_supports: rowval out of range or not strictly ascending inside a column");
    h->tm.h2d_ms = rt::elapsed_ms(h->ev[0], h->ev[1]);
    h->T = T; h->have_supports = true; h->supp_lo = 0; h->supp_hi = h->Nx;
    return 0;
}

// ---- supports generated on the device (README.md:52-54) ----------------------------------------
// Scratch of sls_make_supports lives in the handle (no cudaMalloc / cudaFree per call, which also synchronise the device).
static int make_both_families(sls_handle* h, int64_t T, const std::vector<int>& kx, const std::vector<int>& ku,
                              int nlev, int64_t* nnz_x, int64_t* nnz_u) {
    const int64_t Nx = h->Nx;
    auto distinct_of = [](const std::vector<int>& v) {
        std::vector<int> d(v.begin(), v.end());
        std::sort(d.begin(), d.end());
        d.erase(std::unique(d.begin(), d.end()), d.end());
        return d;
    };
    const std::vector<int> dx = distinct_of(kx), du = distinct_of(ku);
    const int nsx = (int)dx.size(), nsu = (int)du.size();
    std::vector<int32_t> sx(nlev + 1, -1), su(nlev + 1, -1);
    for (int k = 0; k < nsx; k++) sx[dx[k]] = k;
    for (int k = 0; k < nsu; k++) su[du[k]] = k;
    DevBuf& dsx = h->supp[0]; DevBuf& dsu = h->supp[1]; DevBuf& dcpx = h->supp[2]; DevBuf& dcpu = h->supp[3];
    DevBuf& dtot = h->supp[4]; DevBuf& doffx = h->supp[5]; DevBuf& doffu = h->supp[6]; DevBuf& dflags = h->supp[7];
#define MS_CK(expr) do { int e_ = (expr); if (e_) return fail(h, e_, #expr); } while (0)
    MS_CK(dsx.ensure((nlev + 1) * 4)); MS_CK(dsu.ensure((nlev + 1) * 4));
    MS_CK(dcpx.ensure((size_t)nsx * (Nx + 1) * 8)); MS_CK(dcpu.ensure((size_t)nsu * (Nx + 1) * 8));
    MS_CK(dtot.ensure(16)); MS_CK(doffx.ensure((size_t)nsx * 8)); MS_CK(doffu.ensure((size_t)nsu * 8)); MS_CK(dflags.ensure(16));
    MS_CK(h->supp_h.resize(4 + nsx + nsu));
    MS_CK(rt::h2d(dsx.p, sx.data(), (nlev + 1) * 4, h->stream));
    MS_CK(rt::h2d(dsu.p, su.data(), (nlev + 1) * 4, h->stream));
    MS_CK(rt::sync(h->stream));                                // sx / su are locals
    SlsSuppArgs a{};
    a.A = h->A.view(); a.B2r = h->B2r.view(); a.Nx = Nx; a.Nu = h->Nu; a.nlev = nlev;
    a.ex_slot = dsx.as<int32_t>(); a.eu_slot = dsu.as<int32_t>();
    a.xcolptr = dcpx.as<int64_t>(); a.ucolptr = dcpu.as<int64_t>();
    a.overflow = dflags.as<int>(); a.maxlen = dflags.as<int>() + 1;
    a.xlast_slot = sx[kx[T - 1]]; a.ulast_slot = su[ku[T - 1]];
    // set capacity of the per-warp scratch (3 x cap ints): start small for occupancy, double while a set overflows
    std::vector<int64_t> lastx(nsx), lastu(nsu), offx(nsx + 1, 0), offu(nsu + 1, 0);
    int64_t* hs = h->supp_h.data();                            // pinned: [0] overflow|maxlen x|maxlen u (ints), [2..] last colptr entries
    int cap = h->supp_cap > 0 ? h->supp_cap : 128;
    for (;;) {
        a.cap = cap;
        MS_CK(rt::zero(dcpx.p, (size_t)nsx * (Nx + 1) * 8, h->stream));
        MS_CK(rt::zero(dcpu.p, (size_t)nsu * (Nx + 1) * 8, h->stream));
        MS_CK(rt::zero(dtot.p, 16, h->stream)); MS_CK(rt::zero(dflags.p, 16, h->stream));
        MS_CK(sls_launch_supports(a, false, h->stream));
        MS_CK(sls_launch_scan(dcpx.as<int64_t>(), Nx + 1, nsx, dtot.as<int64_t>(), 1, 0, Nx + 1, h->stream));
        MS_CK(sls_launch_scan(dcpu.as<int64_t>(), Nx + 1, nsu, dtot.as<int64_t>(), 1, 0, Nx + 1, h->stream));
        MS_CK(sls_kcopy(hs, dflags.p, 16, h->stream));
        MS_CK(sls_kgather64(hs + 2, dcpx.as<int64_t>() + Nx, (size_t)(Nx + 1), nsx, h->stream));
        MS_CK(sls_kgather64(hs + 2 + nsx, dcpu.as<int64_t>() + Nx, (size_t)(Nx + 1), nsu, h->stream));
        MS_CK(rt::sync(h->stream));
        const int32_t* fl = reinterpret_cast<const int32_t*>(hs);
        if (!fl[0]) break;
        if ((size_t)4 * 3 * (cap * 2) * sizeof(int32_t) > (size_t)h->max_smem)
            return fail(h, -4, "sls_make_supports: a reachability set does not fit shared memory");
        cap *= 2;
    }
    h->supp_cap = cap;
    const int32_t* fl = reinterpret_cast<const int32_t*>(hs);
    for (int k = 0; k < nsx; k++) { lastx[k] = hs[2 + k]; offx[k + 1] = offx[k] + lastx[k] - 1; }
    for (int k = 0; k < nsu; k++) { lastu[k] = hs[2 + nsx + k]; offu[k + 1] = offu[k] + lastu[k] - 1; }
    DevSupports& X = h->Sx; DevSupports& U = h->Su;
    X.row_off = U.row_off = 0;
    X.nnz = offx[nsx]; U.nnz = offu[nsu]; X.has_nzval = U.has_nzval = false;
    X.maxlen_last = fl[1]; U.maxlen_last = fl[2];
    MS_CK(X.rowval.ensure(std::max<int64_t>(X.nnz, 1) * 8));
    MS_CK(U.rowval.ensure(std::max<int64_t>(U.nnz, 1) * 8));
    MS_CK(rt::h2d(doffx.p, offx.data(), (size_t)nsx * 8, h->stream));
    MS_CK(rt::h2d(doffu.p, offu.data(), (size_t)nsu * 8, h->stream));
    a.xrow = X.rowval.as<int64_t>(); a.urow = U.rowval.as<int64_t>();
    a.xoff = doffx.as<int64_t>(); a.uoff = doffu.as<int64_t>();
    MS_CK(sls_launch_supports(a, true, h->stream));
    // per-tap views: colptr copies, shared rowval through tap_off
    MS_CK(X.colptr.ensure((size_t)T * (Nx + 1) * 8)); MS_CK(U.colptr.ensure((size_t)T * (Nx + 1) * 8));
    MS_CK(X.tap_off.ensure((T + 1) * 8)); MS_CK(U.tap_off.ensure((T + 1) * 8));
    X.h_tap_off.assign(T + 1, 0); U.h_tap_off.assign(T + 1, 0);
    for (int64_t t = 0; t < T; t++) {
        const int k = sx[kx[t]], l = su[ku[t]];
        MS_CK(rt::d2d(X.colptr.as<int64_t>() + t * (Nx + 1), dcpx.as<int64_t>() + (int64_t)k * (Nx + 1), (Nx + 1) * 8, h->stream));
        MS_CK(rt::d2d(U.colptr.as<int64_t>() + t * (Nx + 1), dcpu.as<int64_t>() + (int64_t)l * (Nx + 1), (Nx + 1) * 8, h->stream));
        X.h_tap_off[t] = offx[k]; U.h_tap_off[t] = offu[l];
        if (nnz_x) nnz_x[t] = lastx[k] - 1;
        if (nnz_u) nnz_u[t] = lastu[l] - 1;
    }
    X.h_tap_off[T] = X.nnz; U.h_tap_off[T] = U.nnz;
    MS_CK(rt::h2d(X.tap_off.p, X.h_tap_off.data(), (T + 1) * 8, h->stream));
    MS_CK(rt::h2d(U.tap_off.p, U.h_tap_off.data(), (T + 1) * 8, h->stream));
    MS_CK(rt::sync(h->stream));                                // offx / offu are locals
#undef MS_CK
    return 0;
}

extern "C" int sls_make_supports(sls_handle* h, int64_t d, int64_t T, double alpha, int64_t* nnz_x, int64_t* nnz_u) {
    if (!h) return -1;
    SLS_ARG(h->have_plant, "sls_make_supports: call sls_set_plant first");
    SLS_ARG(T >= 1 && T < 4096 && d >= 0 && d < 4096 && alpha >= 0, "sls_make_supports: bad d, T or alpha");
    SLS_CK(rt::set_device(h->device));
    h->have_supports = false; h->plan.valid = false; h->solved = false;
    std::vector<int> kx(T), ku(T);
    int nlev = 0;
    for (int64_t t = 1; t <= T; t++) {
        const double k = std::floor(alpha * (double)(t - 1));              // Julia: floor(α*(t-1)) in Float64
        kx[t - 1] = (int)std::min<double>((double)d, k);
        ku[t - 1] = (int)std::min<double>((double)(d + 1), k);
        nlev = std::max(nlev, std::max(kx[t - 1], ku[t - 1]));
    }
    h->ev[0].record(h->stream);
    SLS_CK(rt::zero(h->d_flag.p, sizeof(int32_t), h->stream));
    int e = make_both_families(h, T, kx, ku, nlev, nnz_x, nnz_u);
    if (e) return e;
    h->ev[1].record(h->stream);
    SLS_CK(rt::sync(h->stream));
    h->tm.h2d_ms = rt::elapsed_ms(h->ev[0], h->ev[1]);
    h->T = T; h->have_supports = true; h->supp_lo = 0; h->supp_hi = h->Nx;
    return 0;
}

extern "C" int sls_get_supports(sls_handle* h, int which, int64_t t, int64_t* colptr, int64_t* rowval, uint8_t* nzval) {
    if (!h) return -1;
    SLS_ARG(h->have_supports, "sls_get_supports: no supports on the device");
    SLS_ARG(h->supp_lo == 0 && h->supp_hi == h->Nx, "sls_get_supports: only a column window of the supports is on the device (after sls_synthesize); call sls_set_supports");
    SLS_ARG((which == 0 || which == 1) && t >= 0 && t < h->T, "sls_get_supports: bad tap");
    SLS_CK(rt::set_device(h->device));
    const DevSupports& d = which ? h->Su : h->Sx;
    const int64_t Nx = h->Nx;
    std::vector<int64_t> cp(Nx + 1);
    SLS_CK(rt::d2h(cp.data(), d.colptr.as<int64_t>() + t * (Nx + 1), (Nx + 1) * 8, h->stream));
    SLS_CK(rt::sync(h->stream));
    const int64_t nnz = cp[Nx] - 1;
    if (colptr) std::copy(cp.begin(), cp.end(), colptr);
    if (rowval && nnz) SLS_CK(rt::d2h(rowval, d.rowval.as<int64_t>() + d.h_tap_off[t], nnz * 8, h->stream));
    if (nzval && nnz) {
        if (d.has_nzval) SLS_CK(rt::d2h(nzval, d.nzval.as<uint8_t>() + d.h_tap_off[t], nnz, h->stream));
        else memset(nzval, 1, nnz);
    }
    SLS_CK(rt::sync(h->stream));
    return 0;
}

// ---- plan (kernel 1) ------------------------------------------------------------------------
static SlsPlanView plan_view(const sls_handle* h, const Plan& p) {
    SlsPlanView v{};
    v.Nx = h->Nx; v.Nu = h->Nu; v.T = (int32_t)h->T; v.ncols = p.ncols;
    v.nstride = p.nstride; v.mstride = p.mstride; v.wx = p.wx; v.wu = p.wu;
    v.cols = p.cols.as<int32_t>(); v.n = p.n.as<int32_t>(); v.m = p.m.as<int32_t>();
    v.sx = p.sx.as<int32_t>(); v.su = p.su.as<int32_t>();
    v.maskx = p.maskx.as<uint32_t>(); v.masku = p.masku.as<uint32_t>();
    v.jloc = p.jloc.as<int32_t>(); v.lnnz = p.lnnz.as<int32_t>(); v.lnnzA = p.lnnzA.as<int32_t>();
    v.lnnzB = p.lnnzB.as<int32_t>(); v.pre = p.pre.as<uint32_t>(); v.tot = p.tot.as<int64_t>();
    v.colptr = p.colptr.as<int64_t>();
    return v;
}

static int build_plan(sls_handle* h, Plan& p, const int64_t* cols, int64_t ncols, bool unique_cols) {
    const int64_t Nx = h->Nx, T = h->T;
    p.valid = false;
    static const bool trace = getenv("SLS_PLAN_TRACE") != nullptr;
    auto now = []() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t0 = trace ? now() : 0; double t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
    SLS_ARG(h->have_plant && h->have_supports, "plan: plant and supports must be set first");
    SLS_ARG(ncols >= 0 && ncols < (int64_t)1 << 31, "plan: bad ncols");
    if (!cols) ncols = Nx;
    p.ncols = (int)ncols;
    SLS_CK(p.h_cols.resize(std::max<int64_t>(ncols, 1)));
    p.col_lo = 0; p.col_hi = cols ? 0 : ncols;
    if (!cols) {
        if (p.identity_cols != ncols) { for (int64_t i = 0; i < ncols; i++) p.h_cols[i] = (int32_t)i; p.identity_cols = (int)ncols; }
    } else {
        // An explicit column list is what every rank of a multi-GPU job passes (its range of a sub-plant).  Fast path: one pass
        // that only tests cols[i] - cols[i-1] == 1 (no early exit, 64-bit equality: vectorises with baseline SSE2) proves the list
        // contiguous and ascending, its ends give the extent and the range check, and the 32-bit copy is an iota.  Anything else
        // takes the general passes: convert + range + extent + contiguity, then strict ascent; the uniqueness bitmap is only
        // needed for a list that does not ascend.  (The branchy one-pass form cost ~0.2 ms per 100 000 columns.)
        p.identity_cols = 0;
        int64_t lo = Nx, hi = -1;
        int bad = 0, notcontig = 0, notasc = 0;
        const int64_t first = ncols > 0 ? cols[0] - 1 : 0;
        int32_t* hc = p.h_cols.data();
        for (int64_t i = 1; i < ncols; i++) notcontig |= (int)(((uint64_t)cols[i] - (uint64_t)cols[i - 1]) != 1ull);
        if (!notcontig && ncols > 0) {
            lo = first; hi = first + ncols - 1;
            bad = (int)((lo < 0) | (hi >= Nx));
            if (!bad) { const int32_t f32 = (int32_t)first; for (int64_t i = 0; i < ncols; i++) hc[i] = f32 + (int32_t)i; }
        } else {
            for (int64_t i = 0; i < ncols; i++) {
                const int64_t j = cols[i] - 1;
                bad |= (int)((j < 0) | (j >= Nx));
                lo = std::min(lo, j); hi = std::max(hi, j);
                hc[i] = (int32_t)j;
            }
            for (int64_t i = 1; i < ncols; i++) notasc |= (int)(cols[i] <= cols[i - 1]);
        }
        if (bad) return fail(h, -3, "plan: column index out of range (1-based, 1..Nx)");
        if (ncols > 0) { p.col_lo = lo; p.col_hi = hi + 1; }
        if (unique_cols && notasc) {
            std::vector<uint8_t> seen(Nx, 0);
            for (int64_t i = 0; i < ncols; i++) { if (seen[p.h_cols[i]]) return fail(h, -3, "plan: duplicate column"); seen[p.h_cols[i]] = 1; }
        }
        p.cols_sorted = !notasc;
        p.cols_contiguous = !notcontig;
    }
    // every support column kernel 1 reads (the rows of nz(A[:,j])) must be on the device: after sls_synthesize of a
    // column range only that range and its halo are
    if (!cols) { p.cols_sorted = true; p.cols_contiguous = true; }
    if ((h->supp_lo > 0 || h->supp_hi < Nx) && ncols > 0) {
        int64_t kmin = Nx, kmax = -1, jmin = Nx, jmax = -1;
        if (p.cols_contiguous) {                              // a column range: two reductions over slices (no early exit)
            const int64_t j0 = p.h_cols[0];
            for (int64_t i = 0; i < ncols; i++) { kmin = std::min<int64_t>(kmin, h->h_Akmin[j0 + i]); kmax = std::max<int64_t>(kmax, h->h_Akmax[j0 + i]); }
            jmin = j0; jmax = j0 + ncols - 1;
        } else {
            for (int64_t i = 0; i < ncols; i++) {
                const int64_t j = p.h_cols[i];
                kmin = std::min<int64_t>(kmin, h->h_Akmin[j]); kmax = std::max<int64_t>(kmax, h->h_Akmax[j]);
                jmin = std::min(jmin, j); jmax = std::max(jmax, j);
            }
        }
        if (kmin < h->supp_lo || kmax >= h->supp_hi || jmin < h->supp_lo || jmax >= h->supp_hi)
            return fail(h, -3, "plan: a column needs support columns that are not on the device (only the window of the last "
                               "sls_synthesize is): call sls_set_supports");
    }
    // candidate-list capacity of kernel 1 (power of two, >= 32)
    int64_t cand = (int64_t)h->maxcolA * std::max(h->Sx.maxlen_last, h->Su.maxlen_last);
    int cap = 32; while (cap < cand) cap <<= 1;
    p.cap = cap;
    const int nc1 = std::max(p.ncols, 1);
    SLS_CK(p.cols.ensure(nc1 * sizeof(int32_t)));
    SLS_CK(p.n.ensure(nc1 * sizeof(int32_t)));
    SLS_CK(p.m.ensure(nc1 * sizeof(int32_t)));
    {
        // a contiguous ascending list (the all-columns plan, a shard, a pipeline chunk) is generated on the device
        const bool contiguous = p.ncols > 0 && p.cols_contiguous;
        const int32_t first = p.ncols > 0 ? p.h_cols[0] : 0;
        if (contiguous) SLS_CK(sls_launch_iota(p.cols.as<int32_t>(), first, p.ncols, h->stream));
        else SLS_CK(sls_kcopy(p.cols.p, p.h_cols.data(), (size_t)p.ncols * sizeof(int32_t), h->stream));
    }
    if ((size_t)cap * 4 * sizeof(int32_t) + 4096 > (size_t)h->max_smem)
        return fail(h, -4, "plan: candidate index list of a column does not fit shared memory");
    // kernel 1a leaves the index sets for kernel 1b (row stride cap) unless that would take more than 1 GB
    const bool keep = (size_t)nc1 * cap * 2 * sizeof(int32_t) <= ((size_t)1 << 30);
    if (keep) { SLS_CK(p.keepx.ensure((size_t)nc1 * cap * sizeof(int32_t))); SLS_CK(p.keepu.ensure((size_t)nc1 * cap * sizeof(int32_t))); }
    SLS_CK(p.stats.ensure(36 * sizeof(int32_t))); SLS_CK(p.h_stats.resize(36));
    SLS_CK(p.lists.ensure(nc1 * sizeof(int32_t)));
    SLS_CK(rt::zero(p.stats.p, 36 * sizeof(int32_t), h->stream));
    h->gev[0].record(h->stream);
    SLS_CK(sls_launch_gather_count(h->A.view(), h->Sx.view(), h->Su.view(), (int)T, Nx, p.cols.as<int32_t>(), p.ncols,
                                   p.n.as<int32_t>(), p.m.as<int32_t>(), cap, keep ? p.keepx.as<int32_t>() : nullptr,
                                   keep ? p.keepu.as<int32_t>() : nullptr, h->d_flag.as<int32_t>(), p.stats.as<int32_t>(), h->stream));
    h->gev[1].record(h->stream);
    SLS_CK(sls_kcopy(h->h_flag.data(), h->d_flag.p, sizeof(int32_t), h->stream));
    // sizes, class histogram and the kernel-2 work lists stay on the device; the host reads 36 integers
    SLS_CK(sls_kcopy(p.h_stats.data(), p.stats.p, 36 * sizeof(int32_t), h->stream));
    SLS_CK(sls_launch_class_lists(p.n.as<int32_t>(), p.ncols, p.stats.as<int32_t>(), sls_solve_class_count(), p.lists.as<int32_t>(), h->stream));
    p.sizes_on_host = false; p.lists_on_host = false;
    if (trace) t1 = now();
    SLS_CK(rt::sync(h->stream));
    if (trace) t2 = now();
    if (h->h_flag[0]) return fail(h, -2, "plan: support rowval out of range or not strictly ascending inside a column");
    const int nk_ = sls_solve_class_count();
    const int32_t* st = p.h_stats.data();
    if (st[2]) return fail(h, -4, "plan: candidate index list overflow in kernel 1");
    const int nmax = st[0], mmax = st[1];
    p.sets_are_runs = (st[3] == 0);
    int cls_count[16] = {0};
    for (int k = 0; k < 16; k++) cls_count[k < nk_ ? k : nk_] += st[4 + k];
    p.nmax = nmax; p.mmax = mmax;
    p.nstride = std::max(nmax, 1); p.mstride = std::max(mmax, 1);
    p.wx = (p.nstride + 31) / 32; p.wu = (p.mstride + 31) / 32;
    SLS_CK(p.sx.ensure((size_t)nc1 * p.nstride * sizeof(int32_t)));
    SLS_CK(p.su.ensure((size_t)nc1 * p.mstride * sizeof(int32_t)));
    SLS_CK(p.maskx.ensure((size_t)nc1 * T * p.wx * sizeof(uint32_t)));
    SLS_CK(p.masku.ensure((size_t)nc1 * T * p.wu * sizeof(uint32_t)));
    SLS_CK(p.jloc.ensure(nc1 * sizeof(int32_t)));
    SLS_CK(p.lnnz.ensure(nc1 * sizeof(int32_t)));
    SLS_CK(p.lnnzA.ensure(nc1 * sizeof(int32_t)));
    SLS_CK(p.lnnzB.ensure(nc1 * sizeof(int32_t)));
    SLS_CK(p.pre.ensure((size_t)nc1 * 2 * T * sizeof(uint32_t)));
    SLS_CK(p.tot.ensure((size_t)(p.ncols + 1) * sizeof(int64_t)));
    SLS_CK(p.colptr.ensure((size_t)2 * T * (Nx + 1) * sizeof(int64_t)));
    // counts are written for the planned columns only: zero (and later scan) just the window that holds them
    SLS_CK(rt::zero2d(p.colptr.as<int64_t>() + p.col_lo, (size_t)(Nx + 1) * sizeof(int64_t), (size_t)(p.col_hi - p.col_lo) * sizeof(int64_t), (size_t)(2 * T), h->stream));
    SLS_CK(rt::zero(p.tot.p, (size_t)(p.ncols + 1) * sizeof(int64_t), h->stream));
    const SlsPlanView v = plan_view(h, p);
    const size_t fill_smem = (size_t)4 * sls_gather_fill_words(cap, p.nstride, p.mstride, p.wx, p.wu, (int)T) * sizeof(int32_t);
    if (fill_smem > (size_t)h->max_smem) return fail(h, -4, "plan: index sets do not fit shared memory");
    if (trace) t3 = now();
    h->gev[2].record(h->stream);
    SLS_CK(sls_launch_gather_fill(h->A.view(), h->B2.view(), h->Sx.view(), h->Su.view(), v, cap, keep ? p.keepx.as<int32_t>() : nullptr,
                                  keep ? p.keepu.as<int32_t>() : nullptr, h->stream));
    SLS_CK(p.scanws.ensure(sls_scan_scratch_bytes((int)(2 * T), std::max<int64_t>(p.col_hi - p.col_lo, p.ncols + 1))));
    SLS_CK(sls_launch_scan(v.colptr, Nx + 1, (int)(2 * T), v.tot, p.ncols + 1, p.col_lo, p.col_hi, h->stream, p.scanws.as<int64_t>()));
    h->gev[3].record(h->stream);
    // sizes back to the host
    SLS_CK(p.h_lnnz.resize(nc1)); SLS_CK(p.h_lA.resize(nc1)); SLS_CK(p.h_lB.resize(nc1));
    SLS_CK(p.h_tot.resize(p.ncols + 1)); SLS_CK(p.h_last.resize(2 * T));
    PinVec<int32_t>& lA = p.h_lA; PinVec<int32_t>& lB = p.h_lB; PinVec<int64_t>& tot = p.h_tot;
    p.nnz_tap.assign(2 * T, 0);
    // only the totals cross the bus here; the per-column arrays (lnnz, lA, lB, tot) are statistics and are fetched by
    // plan_stats when somebody asks for them
    SLS_CK(sls_kcopy(tot.data() + p.ncols, p.tot.as<int64_t>() + p.ncols, sizeof(int64_t), h->stream));
    SLS_CK(sls_kgather64(p.h_last.data(), v.colptr + Nx, (size_t)(Nx + 1), (int)(2 * T), h->stream));
    (void)lA; (void)lB;
    // size classes of kernel 2 (by n): counts and largest m per class from kernel 1a's statistics; the lists were written
    // by sls_class_lists_kernel (ascending inside a class, classes back to back)
    p.classes.clear();
    {
        size_t off = 0;
        for (int k = 0; k < nk_; k++) {
            SolveClass c_{sls_solve_class_N(k), cls_count[k], st[20 + k], 0, off};
            off += cls_count[k];
            if (c_.count) p.classes.push_back(c_);
        }
    }
    if (trace) t4 = now();
    SLS_CK(rt::sync(h->stream));
    if (trace) t5 = now();
    for (int64_t tt = 0; tt < 2 * T; tt++) p.nnz_tap[tt] = p.h_last[tt];
    p.out_off.assign(2 * T + 1, 0);
    for (int64_t tt = 0; tt < 2 * T; tt++) { p.nnz_tap[tt] -= 1; p.out_off[tt + 1] = p.out_off[tt] + p.nnz_tap[tt]; }
    p.tot_packed = tot[p.ncols];
    if (p.tot_packed != p.out_off[2 * T]) return fail(h, -5, "plan: internal count mismatch");
    // local-nnz capacity of each size class of kernel 2: nnz(A[sx,sx]) + nnz(B2[sx,su]) <= n*maxcol(A) + m*maxcol(B2)
    // (a bound instead of the exact class maximum keeps a 100k-entry host loop off the critical path)
    for (SolveClass& c : p.classes)
        c.nnzcap = (int)std::min<int64_t>((int64_t)c.N * h->maxcolA + (int64_t)c.mcap * h->maxcolB2, (int64_t)c.N * (c.N + c.mcap));
    p.stats_valid = false;
    p.gpu_ms = rt::elapsed_ms(h->gev[0], h->gev[1]) + rt::elapsed_ms(h->gev[2], h->gev[3]);
    if (trace) fprintf(stderr, "[plan] to-sync1 %.3f  sync1 %.3f  host->fill %.3f  fill..sync2 %.3f  sync2 %.3f  tail %.3f  total %.3f ms\n", t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, now() - t5, now() - t0);
    p.valid = true;
    return 0;
}

// per-column sizes / kernel-2 lists on the host: only for callers that ask (sls_plan's n_out / m_out, sls_reduce,
// sls_local_problem, the statistics, the chunk boundaries of sls_solve's D2H pipeline)
static int plan_fetch_sizes(sls_handle* h, Plan& p) {
    if (p.sizes_on_host) return 0;
    const int nc1 = std::max(p.ncols, 1);
    SLS_CK(p.h_n.resize(nc1)); SLS_CK(p.h_m.resize(nc1));
    p.h_n[0] = 0; p.h_m[0] = 0;
    SLS_CK(sls_kcopy(p.h_n.data(), p.n.p, (size_t)p.ncols * sizeof(int32_t), h->stream));
    SLS_CK(sls_kcopy(p.h_m.data(), p.m.p, (size_t)p.ncols * sizeof(int32_t), h->stream));
    SLS_CK(rt::sync(h->stream));
    p.sizes_on_host = true;
    return 0;
}
static int plan_fetch_lists(sls_handle* h, Plan& p) {
    if (p.lists_on_host) return 0;
    const int nc1 = std::max(p.ncols, 1);
    SLS_CK(p.h_lists.resize(nc1));
    SLS_CK(sls_kcopy(p.h_lists.data(), p.lists.p, (size_t)p.ncols * sizeof(int32_t), h->stream));
    SLS_CK(rt::sync(h->stream));
    p.lists_on_host = true;
    return 0;
}

// SURVEY §8d algorithmic work of a plan (reported by sls_get_timing; not needed to run the kernels)
static void plan_stats(sls_handle* h, Plan& p, int64_t T) {
    if (p.stats_valid) return;
    if (plan_fetch_sizes(h, p)) return;
    p.alg_flops = 0; p.alg_bytes = 0;
    if (p.ncols > 0) {
        int e = sls_kcopy(p.h_lnnz.data(), p.lnnz.p, (size_t)p.ncols * sizeof(int32_t), h->stream);
        if (!e) e = sls_kcopy(p.h_lA.data(), p.lnnzA.p, (size_t)p.ncols * sizeof(int32_t), h->stream);
        if (!e) e = sls_kcopy(p.h_lB.data(), p.lnnzB.p, (size_t)p.ncols * sizeof(int32_t), h->stream);
        if (!e) e = sls_kcopy(p.h_tot.data(), p.tot.p, (size_t)(p.ncols + 1) * sizeof(int64_t), h->stream);
        if (!e) e = rt::sync(h->stream);
        if (e) return;
    }
    for (int i = 0; i < p.ncols; i++) {
        const double n = p.h_n[i], m = p.h_m[i], v_ = (double)(p.h_tot[i + 1] - p.h_tot[i]);
        p.alg_flops += n * n * v_ + (7.0 / 3.0) * (double)T * n * n * n;
        p.alg_bytes += v_ * 25.0 + 2.0 * T * 24.0 + 16.0 * ((double)p.h_lA[i] + n + (double)p.h_lB[i] + m);
    }
    p.stats_valid = true;
}

extern "C" int sls_plan(sls_handle* h, const int64_t* cols, int64_t ncols,
                        int64_t* nnz_x, int64_t* nnz_u, int32_t* n_out, int32_t* m_out) {
    if (!h) return -1;
    SLS_CK(rt::set_device(h->device));
    h->solved = false;
    h->ev[2].record(h->stream);
    int e = build_plan(h, h->plan, cols, ncols, true);
    if (e) return e;
    h->ev[3].record(h->stream);
    SLS_CK(rt::sync(h->stream));
    h->tm.gather_ms = rt::elapsed_ms(h->ev[2], h->ev[3]);
    h->tm.gather_gpu_ms = h->plan.gpu_ms;
    h->tm.gather_launches = 3;
    h->tm.ncols = h->plan.ncols; h->tm.nnz_out = h->plan.tot_packed; h->tm_from_plan = true;
    const int64_t T = h->T;
    for (int64_t t = 0; t < T; t++) {
        if (nnz_x) nnz_x[t] = h->plan.nnz_tap[t];
        if (nnz_u) nnz_u[t] = h->plan.nnz_tap[T + t];
    }
    if (n_out || m_out) {
        e = plan_fetch_sizes(h, h->plan);
        if (e) return e;
        for (int i = 0; i < h->plan.ncols; i++) {
            if (n_out) n_out[i] = h->plan.h_n[i];
            if (m_out) m_out[i] = h->plan.h_m[i];
        }
    }
    return 0;
}

extern "C" int sls_reduce(sls_handle* h, const int64_t* cj, int64_t q,
                          int64_t* sx_out, int64_t sx_cap, int64_t* n_sx,
                          int64_t* su_out, int64_t su_cap, int64_t* n_su, uint8_t* iix_out) {
    if (!h) return -1;
    SLS_ARG(cj && q >= 1, "sls_reduce: empty column group");
    SLS_ARG(n_sx && n_su, "sls_reduce: n_sx / n_su is NULL");
    SLS_CK(rt::set_device(h->device));
    // unique(findnz(M[:,cj])[1]) only sees first occurrences: a column repeated in cj adds nothing (reduction.jl:14)
    std::vector<int64_t> cju;
    {
        std::vector<uint8_t> seen(h->Nx, 0);
        for (int64_t i = 0; i < q; i++) {
            if (cj[i] < 1 || cj[i] > h->Nx) return fail(h, -3, "sls_reduce: column index out of range (1-based, 1..Nx)");
            if (!seen[cj[i] - 1]) { seen[cj[i] - 1] = 1; cju.push_back(cj[i]); }
        }
    }
    Plan p;
    int e = build_plan(h, p, cju.data(), (int64_t)cju.size(), false);
    if (e) { p.release(); return e; }
    e = plan_fetch_sizes(h, p);
    if (e) { p.release(); return e; }
    std::vector<int32_t> sx((size_t)p.ncols * p.nstride), su((size_t)p.ncols * p.mstride);
    e = rt::d2h(sx.data(), p.sx.p, sx.size() * sizeof(int32_t), h->stream);
    if (!e) e = rt::d2h(su.data(), p.su.p, su.size() * sizeof(int32_t), h->stream);
    if (!e) e = rt::sync(h->stream);
    if (e) { p.release(); return fail(h, e, "sls_reduce: device to host copy"); }
    // unique(findnz(M[:,cj])[1]): CSC order = cj order, rows ascending inside a column, first occurrences
    // kept (reduction.jl:14).  Kernel 1 returns each column's rows ascending, so this is a stable merge.
    auto merge = [&](const std::vector<int32_t>& s, const PinVec<int32_t>& len, int stride, int64_t nrows,
                     std::vector<int64_t>& out) {
        std::vector<uint8_t> seen(nrows, 0);
        for (int c = 0; c < p.ncols; c++)
            for (int i = 0; i < len[c]; i++) {
                const int32_t r = s[(size_t)c * stride + i];
                if (!seen[r]) { seen[r] = 1; out.push_back((int64_t)r + 1); }
            }
    };
    std::vector<int64_t> gx, gu;
    merge(sx, p.h_n, p.nstride, h->Nx, gx);
    merge(su, p.h_m, p.mstride, h->Nu, gu);
    p.release();
    *n_sx = (int64_t)gx.size(); *n_su = (int64_t)gu.size();
    if ((sx_out && (int64_t)gx.size() > sx_cap) || (su_out && (int64_t)gu.size() > su_cap))
        return fail(h, -4, "sls_reduce: output capacity too small (sizes returned in n_sx / n_su)");
    if (sx_out) std::copy(gx.begin(), gx.end(), sx_out);
    if (su_out) std::copy(gu.begin(), gu.end(), su_out);
    if (iix_out) {                                            // indexin(sₓ, cⱼ) .≠ nothing  (reduction.jl:22)
        std::vector<uint8_t> incj(h->Nx, 0);
        for (int64_t i = 0; i < q; i++) incj[cj[i] - 1] = 1;
        for (size_t i = 0; i < gx.size(); i++) iix_out[i] = incj[gx[i] - 1];
    }
    return 0;
}

// Output-feedback branch of sparsity_dim_reduction (reduction.jl:16-18): s_f = unique(findnz(𝓢_f[end][:,cⱼ])[1]) for every
// support family f (𝓢 = [𝓢ₓ,𝓢ᵤ,𝓢ᵧ]); the caller passes the LAST tap of each family.  Only the listed columns' segments are
// uploaded; the union runs on the device (sls_union_first_kernel).
extern "C" int sls_reduce_of(sls_handle* h, int64_t nfam, const int64_t* const* S_colptr, const int64_t* const* S_rowval,
                             const int64_t* nrows, int64_t ncols_S, const int64_t* cj, int64_t q,
                             int64_t* const* s_out, const int64_t* s_cap, int64_t* n_s, uint8_t* iix_out) {
    if (!h) return -1;
    SLS_ARG(nfam >= 1 && nfam <= 8, "sls_reduce_of: nfam must be 1..8 (the reference uses 3: S_x, S_u, S_y)");
    SLS_ARG(S_colptr && S_rowval && nrows && n_s, "sls_reduce_of: NULL argument");
    SLS_ARG(cj && q >= 1, "sls_reduce_of: empty column group");
    SLS_ARG(ncols_S >= 1, "sls_reduce_of: supports without columns");
    for (int64_t i = 0; i < q; i++)
        if (cj[i] < 1 || cj[i] > ncols_S) return fail(h, -3, "sls_reduce_of: column index out of range (1-based, 1..ncols)");
    SLS_CK(rt::set_device(h->device));
    std::vector<int64_t> first;                               // s of family 0 (for ii_x)
    for (int64_t f = 0; f < nfam; f++) {
        SLS_ARG(S_colptr[f] && S_rowval[f] && nrows[f] >= 0, "sls_reduce_of: NULL support arrays");
        const int64_t* cp = S_colptr[f];
        std::vector<int64_t> seg((size_t)q + 1, 0), rows;
        for (int64_t i = 0; i < q; i++) {
            const int64_t e0 = cp[cj[i] - 1], e1 = cp[cj[i]];
            if (e0 < 1 || e1 < e0) return fail(h, -2, "sls_reduce_of: colptr is not a 1-based non-decreasing Julia colptr");
            rows.insert(rows.end(), S_rowval[f] + (e0 - 1), S_rowval[f] + (e1 - 1));
            seg[(size_t)i + 1] = (int64_t)rows.size();
        }
        const int64_t C = (int64_t)rows.size();
        const size_t words = (size_t)((nrows[f] + 31) / 32) + 1;
        DevBuf d_rows, d_seg, d_seen, d_out, d_meta;
        auto drop = [&]() { d_rows.release(); d_seg.release(); d_seen.release(); d_out.release(); d_meta.release(); };
        int e = d_rows.ensure((size_t)std::max<int64_t>(C, 1) * 8);
        if (!e) e = d_seg.ensure(seg.size() * 8);
        if (!e) e = d_seen.ensure(words * 4);
        if (!e) e = d_out.ensure((size_t)std::max<int64_t>(C, 1) * 8);
        if (!e) e = d_meta.ensure(16);                        // [0..7] n_out (int64), [8..11] flag
        if (!e && C) e = rt::h2d(d_rows.p, rows.data(), (size_t)C * 8, h->stream);
        if (!e) e = rt::h2d(d_seg.p, seg.data(), seg.size() * 8, h->stream);
        if (!e) e = rt::zero(d_seen.p, words * 4, h->stream);
        if (!e) e = rt::zero(d_meta.p, 16, h->stream);
        if (!e) e = sls_launch_union_first(d_rows.as<int64_t>(), d_seg.as<int64_t>(), (int)q, nrows[f], d_seen.as<uint32_t>(),
                                           d_out.as<int64_t>(), d_meta.as<int64_t>(), reinterpret_cast<int32_t*>(d_meta.as<char>() + 8), h->stream);
        int64_t meta[2] = {0, 0};
        if (!e) e = rt::d2h(meta, d_meta.p, 16, h->stream);
        if (!e) e = rt::sync(h->stream);
        if (e) { drop(); return fail(h, e, "sls_reduce_of: device work failed"); }
        if ((int32_t)(meta[1] & 0xffffffff)) { drop(); return fail(h, -2, "sls_reduce_of: support rows out of range or not strictly ascending inside a column"); }
        const int64_t ns = meta[0];
        n_s[f] = ns;
        std::vector<int64_t> res((size_t)ns);
        if (ns) { e = rt::d2h(res.data(), d_out.p, (size_t)ns * 8, h->stream); if (!e) e = rt::sync(h->stream); }
        drop();
        if (e) return fail(h, e, "sls_reduce_of: device to host copy");
        if (s_out && s_out[f]) {
            if (!s_cap || ns > s_cap[f]) return fail(h, -4, "sls_reduce_of: output capacity too small (sizes returned in n_s)");
            std::copy(res.begin(), res.end(), s_out[f]);
        }
        if (f == 0) first.swap(res);
    }
    if (iix_out) {                                            // indexin(sₓ, cⱼ) .≠ nothing  (reduction.jl:22)
        std::vector<int64_t> sorted(cj, cj + q);
        std::sort(sorted.begin(), sorted.end());
        for (size_t i = 0; i < first.size(); i++) iix_out[i] = std::binary_search(sorted.begin(), sorted.end(), first[i]) ? 1 : 0;
    }
    return 0;
}

extern "C" int sls_local_problem(sls_handle* h, int64_t k, double* At, double* B2t, uint8_t* Mx, uint8_t* Mu) {
    if (!h) return -1;
    SLS_ARG(h->plan.valid, "sls_local_problem: no plan");
    SLS_ARG(k >= 0 && k < h->plan.ncols, "sls_local_problem: column position out of range");
    SLS_CK(rt::set_device(h->device));
    SLS_CK(plan_fetch_sizes(h, h->plan));
    const int n = h->plan.h_n[k], m = h->plan.h_m[k];
    const int64_t T = h->T;
    DevBuf dA, dB, dMx, dMu;
    int e = dA.ensure((size_t)std::max(n * n, 1) * 8);
    if (!e) e = dB.ensure((size_t)std::max(n * m, 1) * 8);
    if (!e) e = dMx.ensure((size_t)std::max<int64_t>(T * n, 1));
    if (!e) e = dMu.ensure((size_t)std::max<int64_t>(T * m, 1));
    if (!e) e = sls_launch_local_dense(h->A.view(), h->B2.view(), plan_view(h, h->plan), (int)k, dA.as<double>(),
                                       dB.as<double>(), dMx.as<uint8_t>(), dMu.as<uint8_t>(), h->stream);
    if (!e && At) e = rt::d2h(At, dA.p, (size_t)n * n * 8, h->stream);
    if (!e && B2t) e = rt::d2h(B2t, dB.p, (size_t)n * m * 8, h->stream);
    if (!e && Mx) e = rt::d2h(Mx, dMx.p, (size_t)T * n, h->stream);
    if (!e && Mu) e = rt::d2h(Mu, dMu.p, (size_t)T * m, h->stream);
    if (!e) e = rt::sync(h->stream);
    dA.release(); dB.release(); dMx.release(); dMu.release();
    if (e) return fail(h, e, "sls_local_problem");
    return 0;
}

// ---- solve (kernels 2 + 3) ------------------------------------------------------------------
namespace {
struct HostOut {
    int64_t* const* cps[2]; int64_t* const* rvs[2]; double* const* nzs[2];
    int32_t* status; double* resid;
};
}

// Runs kernel 2 and kernel 3 on the planned columns.  With host outputs (`ho`) and ascending columns the
// work is cut into contiguous column chunks: while chunk c+1 is being solved, chunk c's CSC segments of all
// 2T taps travel device->host on a second stream, so the D2H time hides behind kernel 2.
static int run_solve(sls_handle* h, const HostOut* ho) {
    Plan& p = h->plan;
    SLS_ARG(p.valid, "sls_solve: call sls_plan first");
    SLS_CK(rt::set_device(h->device));
    const int64_t T = h->T, Nx = h->Nx;
    const int nc1 = std::max(p.ncols, 1);
    h->solved = false;
    const bool general = h->have_cost || h->have_groups;       // kernel 2g instead of the size-class kernels
    if (general) {
        SLS_ARG(!h->have_weights, "sls_solve: diagonal weights and a general cost / column groups cannot be combined");
        if (h->have_groups) SLS_ARG(h->h_grp.back() == p.ncols, "sls_solve: the column groups (sls_set_groups) do not cover the planned columns");
    }
    if (!general && p.nmax > sls_solve_class_N(sls_solve_class_count() - 1))
        return fail(h, -6, "sls_solve: a column's local state dimension |s_x| exceeds the largest solver size class");
    SLS_CK(h->status.ensure(nc1 * sizeof(int32_t)));
    SLS_CK(h->resid.ensure(nc1 * sizeof(double)));
    SLS_CK(h->out_rowval.ensure((size_t)std::max<int64_t>(p.tot_packed, 1) * sizeof(int64_t)));
    SLS_CK(h->out_nzval.ensure((size_t)std::max<int64_t>(p.tot_packed, 1) * sizeof(double)));
    SLS_CK(h->out_ptrs.ensure((size_t)4 * T * sizeof(void*)));
    int nchunk = 1;
    if (ho && p.cols_sorted && p.ncols >= 16384 && !general) {
        nchunk = (h->opt.flags & 2) ? 1 : 4;                   // flags bit 1: no D2H pipeline
        if (const char* ev = getenv("SLS_D2H_CHUNKS")) { const int c = atoi(ev); if (c >= 1 && c <= 16) nchunk = c; }
    }
    std::vector<int> cb(nchunk + 1);
    for (int c = 0; c <= nchunk; c++) cb[c] = (int)((int64_t)p.ncols * c / nchunk);
    if (nchunk > 1) SLS_CK(plan_fetch_lists(h, p));
    {
        std::vector<void*> ptrs(4 * T);
        for (int64_t tt = 0; tt < 2 * T; tt++) {
            ptrs[tt] = h->out_rowval.as<int64_t>() + p.out_off[tt];
            ptrs[2 * T + tt] = h->out_nzval.as<double>() + p.out_off[tt];
        }
        SLS_CK(rt::h2d(h->out_ptrs.p, ptrs.data(), ptrs.size() * sizeof(void*), h->stream));
        if (nchunk > 1) {
            // 0-based position, inside every tap's arrays, of the first stored entry of each chunk boundary column
            SLS_CK(p.h_seg.resize((size_t)(nchunk + 1) * 2 * T));
            for (int c = 0; c <= nchunk; c++) {
                const int64_t col = (cb[c] < p.ncols) ? p.h_cols[cb[c]] : Nx;
                SLS_CK(rt::d2h_strided(p.h_seg.data() + (size_t)c * 2 * T, p.colptr.as<int64_t>() + col, sizeof(int64_t),
                                       (size_t)(Nx + 1) * sizeof(int64_t), (size_t)(2 * T), h->stream));
            }
        }
        SLS_CK(rt::sync(h->stream));
    }
    // history slots: sized for the most demanding class
    size_t hist_bytes = 8;
    const SlsPlanView v = plan_view(h, p);
    if (!general) for (const SolveClass& c : p.classes) {
        const int maxw = sls_solve_max_resident_warps(c.N, c.mcap, c.nnzcap, (int)T, p.wu, h->sm_count, h->max_smem);
        if (maxw <= 0) return fail(h, -4, "sls_solve: local problem does not fit shared memory");
        const int64_t warps = std::min<int64_t>(maxw, ((int64_t)c.count + 3) / 4 * 4);
        hist_bytes = std::max(hist_bytes, (size_t)warps * sls_solve_hist_doubles(c.N, (int)T, p.wu) * sizeof(double));
    }
    SLS_CK(h->hist.ensure(hist_bytes));
    SLS_CK(rt::zero(h->d_flag.as<int32_t>() + 1, sizeof(int32_t), h->stream));
    h->ev[4].record(h->stream);
    if (ho) {                                                  // colptr arrays are final since sls_plan
        SLS_CK(rt::stream_wait(h->copy_stream, h->ev[4]));
        for (int64_t tt = 0; tt < 2 * T; tt++) {
            const int x = tt >= T; const int64_t t = x ? tt - T : tt;
            if (ho->cps[x] && ho->cps[x][t])
                SLS_CK(rt::d2h(ho->cps[x][t], p.colptr.as<int64_t>() + tt * (Nx + 1), (Nx + 1) * sizeof(int64_t), h->copy_stream));
        }
    }
    int64_t launches = 0, scatters = 0;
    {
        // kernel 3 (rowval) is independent of kernel 2 but runs BEFORE it on the same stream (0.2 ms at config #3): launched
        // beside the persistent kernel 2 its blocks take shared memory on every SM first, a third of kernel 2's CTAs then
        // only become resident when the first wave retires, and kernel 2 takes 18.7 instead of 12.3 ms (measured whenever
        // a plan has a single size class, i.e. no small launch between the two: every interior rank of a multi-GPU job)
        SLS_CK(rt::stream_wait(h->copy_stream, h->ev[4]));
        SlsScatterArgs s{};
        s.P = v; s.c_begin = 0; s.c_end = p.ncols;
        s.rowval = reinterpret_cast<int64_t* const*>(h->out_ptrs.as<void*>());
        s.row_off_x = h->Sx.row_off; s.row_off_u = h->Su.row_off; s.sorted = p.cols_sorted ? 1 : 0; s.runs = p.sets_are_runs ? 1 : 0;
        h->ev[5].record(h->stream);
        SLS_CK(sls_launch_scatter(s, h->stream));
        h->chunk_ev[15].record(h->stream);
        SLS_CK(rt::stream_wait(h->copy_stream, h->chunk_ev[15]));
        if (p.ncols) scatters++;
        // rowval is final now (it does not depend on kernel 2): its device->host copies start at once, whole taps, so the
        // D2H engine works from the first millisecond of kernel 2 and only the values trail the solve chunks
        if (ho)
            for (int64_t tt = 0; tt < 2 * T; tt++) {
                const int x = tt >= T; const int64_t t = x ? tt - T : tt;
                if (ho->rvs[x] && ho->rvs[x][t] && p.nnz_tap[tt])
                    SLS_CK(rt::d2h(ho->rvs[x][t], h->out_rowval.as<int64_t>() + p.out_off[tt], (size_t)p.nnz_tap[tt] * sizeof(int64_t), h->copy_stream));
            }
    }
    if (general && p.ncols) {
        // one CTA per column group: sizes of the union index sets first (device), then the shared-memory carve-up
        const int ngroups = h->have_groups ? (int)h->h_grp.size() - 1 : p.ncols;
        int qmax = 1;
        if (h->have_groups) for (int g = 0; g < ngroups; g++) qmax = std::max(qmax, h->h_grp[g + 1] - h->h_grp[g]);
        int cap = 32;
        while (cap < qmax * std::max(p.nstride, p.mstride)) cap <<= 1;
        if (cap > 8192) return fail(h, -4, "sls_solve: a column group's index sets exceed the general solver's capacity");
        SLS_CK(h->gen_stats.ensure(8 * sizeof(int32_t))); SLS_CK(h->h_gen_stats.resize(8));
        SLS_CK(rt::zero(h->gen_stats.p, 8 * sizeof(int32_t), h->stream));
        const int32_t* gp = h->have_groups ? h->d_grp.as<int32_t>() : nullptr;
        SLS_CK(sls_launch_general_size(v, gp, ngroups, cap, h->gen_stats.as<int32_t>(), h->stream));
        SLS_CK(sls_kcopy(h->h_gen_stats.data(), h->gen_stats.p, 8 * sizeof(int32_t), h->stream));
        SLS_CK(rt::sync(h->stream));
        const int32_t* gs = h->h_gen_stats.data();
        if (gs[5]) return fail(h, -4, "sls_solve: a column group's index sets exceed the general solver's capacity");
        SlsGeneralArgs a{};
        a.A = h->A.view(); a.B2 = h->B2.view();
        a.has_cost = h->have_cost; a.has_d11 = h->have_d11; a.has_b1 = h->have_b1;
        if (h->have_cost) { a.C1 = h->C1.view(); a.D12 = h->D12.view(); }
        if (h->have_d11) a.D11 = h->D11.view();
        if (h->have_b1) a.B1 = h->B1.view();
        a.Sx = h->Sx.view(); a.Su = h->Su.view(); a.P = v; a.grp_ptr = gp; a.ngroups = ngroups;
        a.qc = gs[0]; a.nb = gs[1]; a.pb = gs[2]; a.Nc = gs[3]; a.pc = gs[4]; a.cap = cap;
        const size_t smem = sls_general_smem(a.pc, a.Nc, a.nb, a.pb, a.qc, (int)T, cap);
        if (smem > (size_t)h->max_smem)
            return fail(h, -4, "sls_solve: a group's local problem does not fit the general solver's shared memory (q(n+m) too large)");
        if (h->have_b1 || h->have_d11) {
            // B1 / D11 are indexed by the group's columns: they must exist as columns of these matrices
            if (p.col_hi > h->Nw) return fail(h, -1, "sls_solve: a planned column exceeds Nw (B1 / D11 have no such column)");
        }
        const int grid = sls_general_grid(a, h->sm_count);
        if (grid <= 0) return fail(h, -4, "sls_solve: the general solver cannot be launched with this shared-memory size");
        a.scratch_stride = sls_general_scratch_doubles(a.pc, a.Nc, (int)T);
        SLS_CK(h->gen_scratch.ensure((size_t)grid * a.scratch_stride * sizeof(double)));
        a.scratch = h->gen_scratch.as<double>();
        a.out_nzval = reinterpret_cast<double* const*>(h->out_ptrs.as<void*>() + 2 * T);
        a.status = h->status.as<int32_t>(); a.resid = h->resid.as<double>();
        a.rank_tol = h->opt.rank_tol; a.feas_tol = h->opt.feas_tol; a.refine_tol = 0.1 * h->opt.feas_tol;
        a.max_refine = (h->opt.flags & 4) ? 0 : 3;
        SLS_CK(sls_launch_solve_general(a, grid, h->stream));
        launches++;
    }
    for (int ch = 0; ch < nchunk; ch++) {
        for (const SolveClass& c : p.classes) {
            if (general) break;
            int lo = 0, hi = c.count;
            if (nchunk > 1) {
                const int32_t* lb = p.h_lists.data() + c.list_off;
                lo = (int)(std::lower_bound(lb, lb + c.count, cb[ch]) - lb);
                hi = (int)(std::lower_bound(lb, lb + c.count, cb[ch + 1]) - lb);
            }
            if (hi <= lo) continue;
            SlsSolveArgs a{};
            a.A = h->A.view(); a.B2 = h->B2.view(); a.Ar = h->Ar.view(); a.B2r = h->B2r.view();
            a.P = v; a.list = p.lists.as<int32_t>() + c.list_off + lo; a.count = hi - lo;
            a.mcap = c.mcap; a.nnzcap = c.nnzcap;
            a.out_nzval = reinterpret_cast<double* const*>(h->out_ptrs.as<void*>() + 2 * T); a.hist = h->hist.as<double>();
            a.hist_stride = sls_solve_hist_doubles(c.N, (int)T, p.wu);
            a.status = h->status.as<int32_t>(); a.resid = h->resid.as<double>();
            a.rank_tol = h->opt.rank_tol; a.feas_tol = h->opt.feas_tol;
            a.refine_tol = 0.1 * h->opt.feas_tol; a.max_refine = (h->opt.flags & 4) ? 0 : 3; a.refined = h->d_flag.as<int32_t>() + 1;
            int warps = 0;
            SLS_CK(sls_launch_solve(c.N, a, h->sm_count, h->max_smem, &warps, h->stream));
            launches++;
        }
        if (h->have_weights) {                                  // back-scaling needs kernel 3's rowval as well
            SLS_CK(rt::stream_wait(h->stream, h->chunk_ev[15]));
            SLS_CK(sls_launch_unscale(v, cb[ch], cb[ch + 1], reinterpret_cast<const int64_t* const*>(h->out_ptrs.as<void*>()),
                                      reinterpret_cast<double* const*>(h->out_ptrs.as<void*>() + 2 * T), h->d_wx.as<double>(), h->d_wu.as<double>(), h->Sx.row_off, h->Su.row_off, h->stream));
        }
        if (ho) {
            h->chunk_ev[ch].record(h->stream);
            SLS_CK(rt::stream_wait(h->copy_stream, h->chunk_ev[ch]));
            for (int64_t tt = 0; tt < 2 * T; tt++) {
                const int x = tt >= T; const int64_t t = x ? tt - T : tt;
                int64_t o0 = 0, o1 = p.nnz_tap[tt];
                if (nchunk > 1) { o0 = p.h_seg[(size_t)ch * 2 * T + tt] - 1; o1 = p.h_seg[(size_t)(ch + 1) * 2 * T + tt] - 1; }
                if (o1 <= o0) continue;
                if (ho->nzs[x] && ho->nzs[x][t])
                    SLS_CK(rt::d2h(ho->nzs[x][t] + o0, h->out_nzval.as<double>() + p.out_off[tt] + o0, (o1 - o0) * sizeof(double), h->copy_stream));
            }
        }
    }
    h->ev[6].record(h->stream);
    SLS_CK(sls_kcopy(h->h_flag.data() + 1, h->d_flag.as<int32_t>() + 1, sizeof(int32_t), h->stream));
    SLS_CK(rt::sync(h->stream));
    h->tm.refined = h->h_flag[1];
    if (ho) {
        // per-column outcomes go through pinned staging: the caller's arrays may be pageable, and a pageable
        // cudaMemcpyAsync would block the host in the middle of the pipeline
        SLS_CK(p.h_status.resize(nc1)); SLS_CK(p.h_resid.resize(nc1));
        if (ho->status) SLS_CK(rt::d2h(p.h_status.data(), h->status.p, (size_t)p.ncols * sizeof(int32_t), h->copy_stream));
        if (ho->resid) SLS_CK(rt::d2h(p.h_resid.data(), h->resid.p, (size_t)p.ncols * sizeof(double), h->copy_stream));
        h->ev[7].record(h->copy_stream);
        SLS_CK(rt::sync(h->copy_stream));
        if (ho->status) memcpy(ho->status, p.h_status.data(), (size_t)p.ncols * sizeof(int32_t));
        if (ho->resid) memcpy(ho->resid, p.h_resid.data(), (size_t)p.ncols * sizeof(double));
        SLS_CK(rt::stream_wait(h->stream, h->ev[7]));          // later marks on the main stream see the copies
        h->tm.d2h_ms = rt::elapsed_ms(h->ev[6], h->ev[7]);     // D2H tail left exposed after the last kernel
    }
    h->tm.solve_ms = rt::elapsed_ms(h->chunk_ev[15], h->ev[6]);   // kernel 2 (all classes / chunks)
    h->tm.scatter_ms = rt::elapsed_ms(h->ev[5], h->chunk_ev[15]);   // kernel 3, alone before kernel 2
    h->tm.solve_launches = launches; h->tm.scatter_launches = scatters;
    h->solved = true;
    return 0;
}

extern "C" int sls_solve_device(sls_handle* h) {
    if (!h) return -1;
    return run_solve(h, nullptr);
}

extern "C" int sls_solve(sls_handle* h,
                         int64_t* const* Px_colptr, int64_t* const* Px_rowval, double* const* Px_nzval,
                         int64_t* const* Pu_colptr, int64_t* const* Pu_rowval, double* const* Pu_nzval,
                         int32_t* status, double* resid) {
    if (!h) return -1;
    HostOut ho{{Px_colptr, Pu_colptr}, {Px_rowval, Pu_rowval}, {Px_nzval, Pu_nzval}, status, resid};
    return run_solve(h, &ho);
}


// ---- pipelined one-shot synthesis -------------------------------------------------------------
// sls_set_supports + sls_plan + sls_solve for ALL columns 1..Nx in one call, cut into contiguous column chunks so
// that the three PCIe/compute stages overlap:   H2D of chunk c+1's support columns (in_stream)
//                                             | kernels 1-3 of chunk c (stream; kernel 3 on scat_stream)
//                                             | D2H of chunk c-1's CSC segments (copy_stream).
// Every chunk is an ordinary plan on a contiguous column list over the shared device copy of the supports (whose
// rowval/nzval arrive slice by slice, in final position), with its own output buffers.  Results are bit-for-bit
// those of the three-call path.
namespace {
// Host-side copies of the pipelined call, off the calling thread: a few workers take (dst, src, bytes) pieces from a queue.
// Used for the rowval outputs of sls_synthesize*: Phi[t][:,j] stores a SUBSET of the rows of S[t][:,j] (kernel 1 keeps the
// support entries that are true and whose row lies in s_x), so when a tap's stored-entry count over a chunk's columns equals
// the support's, its row indices ARE the caller's own S[t].rowval slice — they are copied host to host instead of coming back
// over PCIe (on these hosts the device-to-host direction is the scarce one: 57 GB/s for one GPU, 71-92 GB/s for all eight
// together, profiles/r02_e2e_pipeline_trace.txt).
// memcpy with non-temporal stores: the destination (an output array of hundreds of MB) is not read again by this call, so
// the lines need not be fetched for ownership nor kept in cache — a third less DRAM traffic than memcpy, which on these
// hosts is what the copy engines are competing for (glibc only switches to streaming stores far above the 2 MB pieces)
static void stream_copy(char* d, const char* s, size_t n) {
#if defined(__SSE2__)
    static const bool nt = !getenv("SLS_HOST_COPY_NO_NT");
    if (nt && n >= 256) {
        size_t head = (16 - ((uintptr_t)d & 15)) & 15;
        memcpy(d, s, head); d += head; s += head; n -= head;
        const size_t blocks = n / 64;
        for (size_t i = 0; i < blocks; i++) {
            const __m128i a = _mm_loadu_si128((const __m128i*)(s)), b = _mm_loadu_si128((const __m128i*)(s + 16));
            const __m128i c = _mm_loadu_si128((const __m128i*)(s + 32)), e = _mm_loadu_si128((const __m128i*)(s + 48));
            _mm_stream_si128((__m128i*)(d), a); _mm_stream_si128((__m128i*)(d + 16), b);
            _mm_stream_si128((__m128i*)(d + 32), c); _mm_stream_si128((__m128i*)(d + 48), e);
            s += 64; d += 64;
        }
        n -= blocks * 64;
        memcpy(d, s, n);
        _mm_sfence();
        return;
    }
#endif
    memcpy(d, s, n);
}
struct HostCopier {
    struct Job { char* dst; const char* src; size_t n; };
    std::deque<Job> q;
    std::mutex mu;
    std::condition_variable cv;
    bool closing = false;
    std::vector<std::thread> th;
    int64_t bytes = 0;
    static int default_threads() {
        if (const char* ev = getenv("SLS_HOST_COPY_THREADS")) { const int c = atoi(ev); if (c >= 1) return std::min(c, 64); }
        const int hc = (int)std::thread::hardware_concurrency();
        return std::max(2, std::min(4, hc / 4));   // streaming stores: two threads already keep up with the pipeline on config #3
    }
    void start() {
        const int nt = default_threads();
        for (int i = 0; i < nt; i++)
            th.emplace_back([this]() {
                for (;;) {
                    Job j;
                    {
                        std::unique_lock<std::mutex> lk(mu);
                        cv.wait(lk, [this]() { return closing || !q.empty(); });
                        if (q.empty()) return;
                        j = q.front(); q.pop_front();
                    }
                    stream_copy(j.dst, j.src, j.n);
                }
            });
    }
    void add(void* dst, const void* src, size_t n) {            // pieces of at most 2 MB so that the workers stay balanced
        if (!n || dst == src) { bytes += (int64_t)n; return; }
        if (th.empty()) start();
        const size_t piece = (size_t)2 << 20;
        {
            std::lock_guard<std::mutex> lk(mu);
            for (size_t o = 0; o < n; o += piece) q.push_back(Job{(char*)dst + o, (const char*)src + o, std::min(piece, n - o)});
        }
        cv.notify_all();
        bytes += (int64_t)n;
    }
    void finish() {                                             // drains the queue; the calling thread helps
        for (;;) {
            Job j;
            {
                std::lock_guard<std::mutex> lk(mu);
                if (q.empty()) break;
                j = q.front(); q.pop_front();
            }
            stream_copy(j.dst, j.src, j.n);
        }
        { std::lock_guard<std::mutex> lk(mu); closing = true; }
        cv.notify_all();
        for (auto& t : th) t.join();
        th.clear();
    }
    ~HostCopier() { finish(); }
};
}

namespace {
struct CopyBatch {                                          // slice copies of one chunk and direction, issued as one call
    std::vector<void*> dst, src; std::vector<size_t> len;
    void add(void* d, const void* s_, size_t n) { if (n) { dst.push_back(d); src.push_back(const_cast<void*>(s_)); len.push_back(n); } }
    int flush(sls_stream_t st) { int e = rt::copy_batch(dst.data(), src.data(), len.data(), len.size(), st); dst.clear(); src.clear(); len.clear(); return e; }
};
struct PipeSupp {
    DevSupports* d; const int64_t* const* cps; const int64_t* const* rvs; const uint8_t* const* nzs;
    int64_t nrow; const char* name;
};
}

// The caller passes a column WINDOW of every support tap: colptr arrays of wn + 1 entries for the columns [w0, w0 + wn)
// whose values are 1-based offsets into the rowval / nzval pointers given beside them (a full matrix is the window
// w0 = 0, wn = Nx; a view `pointer(S.colptr, w0 + 1)` into a full matrix and a slice `S[:, w0+1:w0+wn]` both qualify).
// On the device the taps' window slices lie back to back; tap_off[t] is shifted so that kernels keep indexing
// rowval[tap_off[t] + colptr[t][j] - 1] with the caller's colptr values.
static int pipe_prepare(sls_handle* h, PipeSupp& f, int64_t T, int64_t Nx, int64_t w0, int64_t wn) {
    DevSupports& d = *f.d;
    std::vector<int64_t> slot(T + 1, 0), off(T + 1, 0);
    bool any_nz = false, all_nz = true;
    for (int64_t t = 0; t < T; t++) {
        if (!f.cps[t]) return fail(h, -2, (std::string("sls_synthesize: ") + f.name + ": colptr is NULL").c_str());
        if (f.cps[t][0] < 1 || f.cps[t][wn] < f.cps[t][0] || (w0 == 0 && wn == Nx && f.cps[t][0] != 1))
            return fail(h, -2, (std::string("sls_synthesize: ") + f.name + ": colptr[1] must be 1 (Julia 1-based CSC)").c_str());
        if (f.cps[t][wn] > f.cps[t][0] && !(f.rvs && f.rvs[t])) return fail(h, -2, (std::string("sls_synthesize: ") + f.name + ": rowval is NULL").c_str());
        slot[t + 1] = slot[t] + f.cps[t][wn] - f.cps[t][0];
        off[t] = slot[t] - (f.cps[t][0] - 1);
        if (f.nzs && f.nzs[t]) any_nz = true; else all_nz = false;
    }
    off[T] = slot[T];
    if (any_nz && !all_nz) return fail(h, -2, "sls_synthesize: nzval must be given for every tap or for none");
    d.nnz = slot[T]; d.has_nzval = any_nz;
    SLS_CK(d.colptr.ensure(T * (Nx + 1) * sizeof(int64_t)));
    SLS_CK(d.rowval.ensure(std::max<int64_t>(d.nnz, 1) * sizeof(int64_t)));
    SLS_CK(d.tap_off.ensure((T + 1) * sizeof(int64_t)));
    if (any_nz) SLS_CK(d.nzval.ensure(std::max<int64_t>(d.nnz, 1)));
    d.h_tap_off = off;
    int ml = 0;
    const int64_t* cl = f.cps[T - 1];
    for (int64_t c = 0; c < wn; c++) ml = std::max<int>(ml, (int)std::min<int64_t>(std::max<int64_t>(cl[c + 1] - cl[c], 0), 1 << 30));
    d.maxlen_last = ml;
    SLS_CK(rt::h2d(d.tap_off.p, d.h_tap_off.data(), (T + 1) * sizeof(int64_t), h->in_stream));   // pageable: before the big copies
    for (int64_t t = 0; t < T; t++)
        SLS_CK(rt::h2d(d.colptr.as<int64_t>() + t * (Nx + 1) + w0, f.cps[t], (wn + 1) * sizeof(int64_t), h->in_stream));
    // the Bool values are 1/9 of the bytes but would be half of the slice copies (each copy costs 10-14 us on the copy-in
    // engine of this box): they go up whole, once per tap, ahead of the column chunks.  Measured the other way (slices
    // travelling with their chunk, so that chunk 0 lands 0.9 ms earlier): the copy-in of the call ends at 21.9 instead of
    // 17.2 ms and the call takes the same 32.1 ms — it is bound by the copy-out engine (profiles/r02_e2e_pipeline_trace.txt).
    if (any_nz)
        for (int64_t t = 0; t < T; t++)
            SLS_CK(rt::h2d(d.nzval.as<uint8_t>() + slot[t], f.nzs[t] + (f.cps[t][0] - 1), (size_t)(slot[t + 1] - slot[t]), h->in_stream));
    return 0;
}

// validates and enqueues the support columns [c0, c1) of every tap of one family
// (runs on the feeder thread: errors are returned as a message, never through the handle)
static int pipe_enqueue_slice(sls_handle* h, PipeSupp& f, int64_t T, int64_t w0, int64_t wn, int64_t c0, int64_t c1, CopyBatch& cbatch,
                              std::string& msg) {
    DevSupports& d = *f.d;
    const bool rows = (h->opt.flags & 1) != 0;
    for (int64_t t = 0; t < T; t++) {
        const int64_t* cp = f.cps[t] - w0;                      // indexed by the global column
        for (int64_t c = c0; c < c1; c++)
            if (cp[c + 1] < cp[c]) { msg = std::string("sls_synthesize: ") + f.name + ": colptr is not monotone"; return -2; }
        if (cp[c0] < cp[w0] || cp[c1] > cp[w0 + wn]) { msg = std::string("sls_synthesize: ") + f.name + ": colptr is not monotone"; return -2; }
        const int64_t s0 = cp[c0] - 1, s1 = cp[c1] - 1;
        if (rows)
            for (int64_t c = c0; c < c1; c++) {
                int64_t prev = 0;
                for (int64_t e = cp[c] - 1; e < cp[c + 1] - 1; e++) {
                    if (f.rvs[t][e] - d.row_off < 1 || f.rvs[t][e] - d.row_off > f.nrow || f.rvs[t][e] <= prev) {
                        msg = std::string("sls_synthesize: ") + f.name + ": rowval out of range or not strictly ascending"; return -2;
                    }
                    prev = f.rvs[t][e];
                }
            }
        if (s1 > s0) {
            cbatch.add(d.rowval.as<int64_t>() + d.h_tap_off[t] + s0, f.rvs[t] + s0, (s1 - s0) * sizeof(int64_t));
        }
    }
    return 0;
}

static void pipe_drain(sls_handle* h) {
    rt::sync(h->in_stream); rt::sync(h->stream); rt::sync(h->scat_stream); rt::sync(h->copy_stream); rt::sync(h->out2_stream);
}

// `slab`: the supports are a column window [w0, w0 + wn) and the outputs are the column slab of the range
// (colptr arrays of ncols + 1 entries starting at 1) — sls_synthesize_window; otherwise full-width arrays.
static int synthesize_impl(sls_handle* h, int64_t T, bool slab, int64_t w0, int64_t wn, int64_t row_off_x, int64_t row_off_u,
                           const int64_t* const* Sx_colptr, const int64_t* const* Sx_rowval, const uint8_t* const* Sx_nzval,
                           const int64_t* const* Su_colptr, const int64_t* const* Su_rowval, const uint8_t* const* Su_nzval,
                           int64_t first_col, int64_t ncols, const int64_t* cap_x, const int64_t* cap_u,
                           int64_t* const* Px_colptr, int64_t* const* Px_rowval, double* const* Px_nzval,
                           int64_t* const* Pu_colptr, int64_t* const* Pu_rowval, double* const* Pu_nzval,
                           int64_t* nnz_x, int64_t* nnz_u, int32_t* status, double* resid) {
    if (!h) return -1;
    SLS_ARG(h->have_plant, "sls_synthesize: call sls_set_plant first");
    SLS_ARG(!h->have_cost && !h->have_groups, "sls_synthesize: a general cost / column groups are installed: use sls_set_supports + sls_plan + sls_solve");
    SLS_ARG(T >= 1 && T < 4096, "sls_synthesize: bad T");
    SLS_ARG(Sx_colptr && Sx_rowval && Su_colptr && Su_rowval, "sls_synthesize: NULL array of pointers");
    SLS_CK(rt::set_device(h->device));
    const int64_t Nx = h->Nx, Nu = h->Nu;
    if (!slab) { w0 = 0; wn = Nx; }
    if (ncols <= 0) { first_col = 1; ncols = Nx; }
    SLS_ARG(first_col >= 1 && first_col + ncols - 1 <= Nx, "sls_synthesize: column range out of 1..Nx");
    SLS_ARG(w0 >= 0 && wn >= 1 && w0 + wn <= Nx, "sls_synthesize_window: support window out of 1..Nx");
    const int64_t r0 = first_col - 1, r1 = r0 + ncols;          // 0-based half-open column range
    h->have_supports = false; h->plan.valid = false; h->solved = false;
    // chunk count: every chunk costs ~230 slice copies of ~10 us driver/engine time each way, against the pipeline's
    // fill and drain time (measured on config #3: 4-5 chunks best, 8 chunks +2 ms)
    int K = (ncols >= 16384) ? 4 : 1;
    if (const char* ev = getenv("SLS_PIPE_CHUNKS")) { const int c = atoi(ev); if (c >= 1) K = c; }
    // relative chunk sizes (tuning knob, "w0,w1,...").  Measured on config #3: no split beats equal chunks by more than 0.3 ms
    // (small first chunks start the download earlier but add ~116 copies of 5-14 us each per chunk and direction)
    std::vector<double> fr;
    if (const char* ev = getenv("SLS_PIPE_FRACS")) {
        for (const char* s = ev; *s;) {
            char* end = nullptr;
            const double w = strtod(s, &end);
            if (end == s) break;
            if (w > 0) fr.push_back(w);
            s = (*end == ',') ? end + 1 : end;
        }
        if (!fr.empty() && ncols >= 16384) K = (int)fr.size(); else fr.clear();
    }
    K = (int)std::min<int64_t>(std::min<int64_t>(K, SLS_PIPE_MAX), ncols);
    std::vector<int64_t> cb(K + 1), ub(K + 1);
    for (int c = 0; c <= K; c++) cb[c] = r0 + ncols * c / K;
    if ((int)fr.size() == K && K > 1) {
        double tot = 0, run = 0;
        for (double w : fr) tot += w;
        for (int c = 1; c < K; c++) {
            run += fr[c - 1];
            cb[c] = std::min(r1 - (K - c), std::max(cb[c - 1] + 1, r0 + (int64_t)llround((double)ncols * run / tot)));
        }
    }
    // support columns to upload: the range itself plus the columns kernel 1 reaches through nz(A[:,j]) (the halo)
    // upload chunk c ends where the reach of plan chunk c ends, so that chunk c never waits for chunk c+1's copy
    ub = cb;
    for (int c = 0; c < K; c++) {
        int64_t hi = cb[c + 1];
        for (int64_t j = cb[c]; j < cb[c + 1]; j++) { ub[0] = std::min<int64_t>(ub[0], h->h_Akmin[j]); hi = std::max<int64_t>(hi, (int64_t)h->h_Akmax[j] + 1); }
        ub[c + 1] = std::max(ub[c], hi);
    }
    for (int c = 1; c <= K; c++) ub[c] = std::max(ub[c], ub[c - 1]);
    if (ub[0] < w0 || ub[K] > w0 + wn)
        return fail(h, -3, "sls_synthesize_window: the support window must cover the column range and its halo (the rows of nz(A[:,j]))");
    while ((int)h->pipe.size() < K) {
        PipeCtx* x = new (std::nothrow) PipeCtx();
        if (!x) return fail(h, -2, "sls_synthesize: out of host memory");
        int e = x->create();
        if (e) { x->release(); delete x; return fail(h, e, "sls_synthesize: event creation"); }
        h->pipe.push_back(x);
    }
    SLS_CK(rt::zero(h->d_flag.as<int32_t>() + 1, sizeof(int32_t), h->stream));
    h->ev[0].record(h->stream);
    SLS_CK(rt::stream_wait(h->in_stream, h->ev[0]));            // earlier work on the main stream may still read the supports
    PipeSupp fam[2] = {{&h->Sx, Sx_colptr, Sx_rowval, Sx_nzval, Nx, "Sx"}, {&h->Su, Su_colptr, Su_rowval, Su_nzval, Nu, "Su"}};
    SLS_CK(rt::zero(h->d_flag.p, sizeof(int32_t), h->in_stream));
    h->Sx.row_off = row_off_x; h->Su.row_off = row_off_u;
    for (int f = 0; f < 2; f++) { int e = pipe_prepare(h, fam[f], T, Nx, w0, wn); if (e) { pipe_drain(h); return e; } }
    h->T = T; h->supp_lo = ub[0]; h->supp_hi = ub[K];
    SLS_CK(h->pipe_colptr.ensure((size_t)2 * T * (Nx + 1) * sizeof(int64_t)));
    SLS_CK(h->pipe_status.resize(ncols)); SLS_CK(h->pipe_resid.resize(ncols));
    // Feeder thread: validates and enqueues the support slices of every chunk on the copy-in stream as fast as the
    // driver accepts them (~230 copies per chunk), so that their per-call cost stays off the thread that launches
    // the kernels.  `fed` = chunks whose copy-in event has been recorded.
    std::atomic<int> fed{0}, feed_err{0};
    std::atomic<bool> feed_stop{false};
    std::string feed_msg;
    std::thread feeder([&]() {
        rt::set_device(h->device);
        CopyBatch cbatch;
        for (int c = 0; c < K && !feed_stop.load(); c++) {
            int e = 0;
            for (int f = 0; f < 2 && !e; f++) e = pipe_enqueue_slice(h, fam[f], T, w0, wn, ub[c], ub[c + 1], cbatch, feed_msg);
            if (!e) { e = cbatch.flush(h->in_stream); if (e) feed_msg = "sls_synthesize: host to device copy"; }
            // row indices of the slice are checked on the device behind its copies; kernel 1a indexes nothing once the flag is set
            if (!e) e = sls_launch_validate_supports(h->Sx.view(), (int)T, Nx, Nx, ub[c], ub[c + 1], h->d_flag.as<int32_t>(), h->in_stream);
            if (!e) e = sls_launch_validate_supports(h->Su.view(), (int)T, Nx, Nu, ub[c], ub[c + 1], h->d_flag.as<int32_t>(), h->in_stream);
            if (e && feed_msg.empty()) feed_msg = "sls_synthesize: support validation launch";
            if (e) { feed_err.store(e); return; }
            h->pipe_in[c].record(h->in_stream);
            fed.store(c + 1, std::memory_order_release);
        }
    });
    auto wait_fed = [&](int c) -> int {                         // host-side wait until chunk c's copy-in event exists
        while (fed.load(std::memory_order_acquire) <= c) {
            if (feed_err.load()) return feed_err.load();
            std::this_thread::yield();
        }
        return 0;
    };
    CopyBatch obatch;
    HostCopier hcopy;
    const bool host_rows = !(h->opt.flags & 8) && !getenv("SLS_PIPE_ROWS_FROM_DEVICE");   // flags bit 3: rowval always from the device
    auto chunk_of = [&](int64_t col) { int c = 0; while (c < K - 1 && col >= ub[c + 1]) c++; return c; };
    std::vector<int64_t> base(2 * T, 0);
    std::vector<int64_t> cols;
    int64_t launches = 0, nnz_total = 0;
    double flops = 0, bytes = 0;
    int e = 0;
#define PIPE_CK(expr) { e = (expr); if (e) { fail(h, e, #expr); break; } }
    h->have_supports = true;                                    // build_plan's precondition; reset on failure below
    for (int c = 0; c < K && !e; c++) {
        const int64_t c0 = cb[c], c1 = cb[c + 1];
        // kernel 1 of column j reads the support columns k in nz(A[:,j]): wait for the chunk holding the largest one
        int32_t kmax = (int32_t)(c1 - 1);
        for (int64_t j = c0; j < c1; j++) kmax = std::max(kmax, h->h_Akmax[j]);
        const int need = chunk_of(kmax);
        e = wait_fed(need);
        if (e) { fail(h, e, feed_msg.c_str()); break; }
        PIPE_CK(rt::stream_wait(h->stream, h->pipe_in[need]));
        PipeCtx& x = *h->pipe[c];
        Plan& p = x.plan;
        cols.resize(c1 - c0);
        for (int64_t j = c0; j < c1; j++) cols[j - c0] = j + 1;
        e = build_plan(h, p, cols.data(), c1 - c0, false);
        if (e) break;
        if (p.nmax > sls_solve_class_N(sls_solve_class_count() - 1)) { e = fail(h, -6, "sls_synthesize: a column's local state dimension |s_x| exceeds the largest solver size class"); break; }
        for (int64_t tt = 0; tt < 2 * T && !e; tt++) {
            const int x_ = tt >= T; const int64_t t = x_ ? tt - T : tt;
            const int64_t* capv = x_ ? cap_u : cap_x;
            const int64_t cap = capv ? capv[t] : (fam[x_].cps[t][wn] - fam[x_].cps[t][0]);
            if (base[tt] + p.nnz_tap[tt] > cap) e = fail(h, -4, "sls_synthesize: output capacity of a tap is too small");
        }
        if (e) break;
        const int nc1 = std::max(p.ncols, 1);
        PIPE_CK(x.status.ensure(nc1 * sizeof(int32_t)));
        PIPE_CK(x.resid.ensure(nc1 * sizeof(double)));
        PIPE_CK(x.out_rowval.ensure((size_t)std::max<int64_t>(p.tot_packed, 1) * sizeof(int64_t)));
        PIPE_CK(x.out_nzval.ensure((size_t)std::max<int64_t>(p.tot_packed, 1) * sizeof(double)));
        PIPE_CK(x.out_ptrs.ensure((size_t)4 * T * sizeof(void*)));
        PIPE_CK(x.base.ensure((size_t)2 * T * sizeof(int64_t)));
        PIPE_CK(x.h_ptrs.resize(4 * T)); PIPE_CK(x.h_base.resize(2 * T));
        for (int64_t tt = 0; tt < 2 * T; tt++) {
            x.h_ptrs[tt] = x.out_rowval.as<int64_t>() + p.out_off[tt];
            x.h_ptrs[2 * T + tt] = x.out_nzval.as<double>() + p.out_off[tt];
            x.h_base[tt] = base[tt];
        }
        PIPE_CK(sls_kcopy(x.out_ptrs.p, x.h_ptrs.data(), (size_t)4 * T * sizeof(void*), h->stream));
        PIPE_CK(sls_kcopy(x.base.p, x.h_base.data(), (size_t)2 * T * sizeof(int64_t), h->stream));
        const SlsPlanView v = plan_view(h, p);
        size_t hist_bytes = 8;
        for (const SolveClass& cl : p.classes) {
            const int maxw = sls_solve_max_resident_warps(cl.N, cl.mcap, cl.nnzcap, (int)T, p.wu, h->sm_count, h->max_smem);
            if (maxw <= 0) { e = fail(h, -4, "sls_synthesize: local problem does not fit shared memory"); break; }
            const int64_t warps = std::min<int64_t>(maxw, ((int64_t)cl.count + 3) / 4 * 4);
            hist_bytes = std::max(hist_bytes, (size_t)warps * sls_solve_hist_doubles(cl.N, (int)T, p.wu) * sizeof(double));
        }
        if (e) break;
        PIPE_CK(h->hist.ensure(hist_bytes));
        // the caller's colptr = the chunk's colptr + entries of the earlier chunks (columns c0+1..c1; chunk 0 also writes [0])
        const int64_t j0 = (c == 0) ? (slab ? r0 : 0) : c0 + 1, j1 = (c == K - 1) ? (slab ? r1 : Nx) : c1;   // columns outside the range are empty
        PIPE_CK(sls_launch_colptr_rebase(h->pipe_colptr.as<int64_t>(), p.colptr.as<int64_t>(), x.base.as<int64_t>(), Nx + 1, (int)(2 * T), j0, j1, h->stream));
        x.ev_plan.record(h->stream);
        // kernel 3 before kernel 2 on the same stream (see run_solve)
        {
            SlsScatterArgs sa{};
            sa.P = v; sa.c_begin = 0; sa.c_end = p.ncols;
            sa.rowval = reinterpret_cast<int64_t* const*>(x.out_ptrs.as<void*>());
            sa.row_off_x = h->Sx.row_off; sa.row_off_u = h->Su.row_off; sa.sorted = p.cols_sorted ? 1 : 0; sa.runs = p.sets_are_runs ? 1 : 0;
            PIPE_CK(sls_launch_scatter(sa, h->stream));
        }
        x.ev_scat.record(h->stream);
        for (const SolveClass& cl : p.classes) {
            SlsSolveArgs a{};
            a.A = h->A.view(); a.B2 = h->B2.view(); a.Ar = h->Ar.view(); a.B2r = h->B2r.view();
            a.P = v; a.list = p.lists.as<int32_t>() + cl.list_off; a.count = cl.count;
            a.mcap = cl.mcap; a.nnzcap = cl.nnzcap;
            a.out_nzval = reinterpret_cast<double* const*>(x.out_ptrs.as<void*>() + 2 * T); a.hist = h->hist.as<double>();
            a.hist_stride = sls_solve_hist_doubles(cl.N, (int)T, p.wu);
            a.status = x.status.as<int32_t>(); a.resid = x.resid.as<double>();
            a.rank_tol = h->opt.rank_tol; a.feas_tol = h->opt.feas_tol;
            a.refine_tol = 0.1 * h->opt.feas_tol; a.max_refine = (h->opt.flags & 4) ? 0 : 3; a.refined = h->d_flag.as<int32_t>() + 1;
            int warps = 0;
            PIPE_CK(sls_launch_solve(cl.N, a, h->sm_count, h->max_smem, &warps, h->stream));
            launches++;
        }
        if (e) break;
        if (h->have_weights) {
            PIPE_CK(rt::stream_wait(h->stream, x.ev_scat));
            PIPE_CK(sls_launch_unscale(v, 0, p.ncols, reinterpret_cast<const int64_t* const*>(x.out_ptrs.as<void*>()),
                                       reinterpret_cast<double* const*>(x.out_ptrs.as<void*>() + 2 * T), h->d_wx.as<double>(), h->d_wu.as<double>(),
                                       h->Sx.row_off, h->Su.row_off, h->stream));
        }
        x.ev_solved.record(h->stream);
        // device -> host, in the order the data becomes final: colptr, rowval, then nzval / status / residual
        PIPE_CK(rt::stream_wait(h->copy_stream, x.ev_plan));
        if (c == K - 1) {                                       // colptr: whole arrays once (58 copies instead of 58 per chunk)
            for (int64_t tt = 0; tt < 2 * T; tt++) {
                const int x_ = tt >= T; const int64_t t = x_ ? tt - T : tt;
                int64_t* const* hc = x_ ? Pu_colptr : Px_colptr;
                if (!(hc && hc[t])) continue;
                // slab form: only the range's ncols + 1 entries (they start at 1: nothing is stored before column r0)
                if (slab) obatch.add(hc[t], h->pipe_colptr.as<int64_t>() + tt * (Nx + 1) + r0, (size_t)(ncols + 1) * sizeof(int64_t));
                else obatch.add(hc[t], h->pipe_colptr.as<int64_t>() + tt * (Nx + 1), (size_t)(Nx + 1) * sizeof(int64_t));
            }
            PIPE_CK(obatch.flush(h->copy_stream));
        }
        PIPE_CK(rt::stream_wait(h->copy_stream, x.ev_scat));
        for (int64_t tt = 0; tt < 2 * T; tt++) {
            const int x_ = tt >= T; const int64_t t = x_ ? tt - T : tt;
            int64_t* const* hr = x_ ? Pu_rowval : Px_rowval;
            if (!(hr && hr[t])) continue;
            // stored rows of a column are a subset of the support column's: equal counts over the chunk <=> the very same
            // row indices, which the caller already holds (HostCopier above); otherwise they come from kernel 3's output
            const int64_t* cp = fam[x_].cps[t] - w0;            // indexed by the global column
            if (host_rows && p.nnz_tap[tt] == cp[c1] - cp[c0])
                hcopy.add(hr[t] + base[tt], fam[x_].rvs[t] + (cp[c0] - 1), (size_t)p.nnz_tap[tt] * sizeof(int64_t));
            else
                obatch.add(hr[t] + base[tt], x.out_rowval.as<int64_t>() + p.out_off[tt], (size_t)p.nnz_tap[tt] * sizeof(int64_t));
        }
        PIPE_CK(obatch.flush(h->copy_stream));
        sls_stream_t vstream = getenv("SLS_PIPE_ONE_OUT") ? h->copy_stream : h->out2_stream;   // values on their own stream
        PIPE_CK(rt::stream_wait(vstream, x.ev_solved));
        for (int64_t tt = 0; tt < 2 * T; tt++) {
            const int x_ = tt >= T; const int64_t t = x_ ? tt - T : tt;
            double* const* hz = x_ ? Pu_nzval : Px_nzval;
            if (hz && hz[t]) obatch.add(hz[t] + base[tt], x.out_nzval.as<double>() + p.out_off[tt], (size_t)p.nnz_tap[tt] * sizeof(double));
        }
        if (status) obatch.add(h->pipe_status.data() + (c0 - r0), x.status.p, (size_t)p.ncols * sizeof(int32_t));
        if (resid) obatch.add(h->pipe_resid.data() + (c0 - r0), x.resid.p, (size_t)p.ncols * sizeof(double));
        PIPE_CK(obatch.flush(vstream));
        x.ev_out.record(vstream);
        for (int64_t tt = 0; tt < 2 * T; tt++) base[tt] += p.nnz_tap[tt];
        nnz_total += p.tot_packed;
    }
#undef PIPE_CK
    if (e) feed_stop.store(true);
    feeder.join();
    hcopy.finish();
    if (!e && feed_err.load()) e = fail(h, feed_err.load(), feed_msg.c_str());
    if (e) { pipe_drain(h); h->have_supports = false; return e; }
    h->pipe_done.record(h->copy_stream);
    h->pipe_done2.record(h->out2_stream);
    SLS_CK(rt::stream_wait(h->stream, h->pipe_done));           // later marks on the main stream see the copies
    SLS_CK(rt::stream_wait(h->stream, h->pipe_done2));
    SLS_CK(rt::stream_wait(h->stream, h->pipe_in[K - 1]));
    h->ev[1].record(h->stream);
    SLS_CK(sls_kcopy(h->h_flag.data(), h->d_flag.p, 2 * sizeof(int32_t), h->stream));
    SLS_CK(rt::sync(h->stream));
    SLS_CK(rt::sync(h->scat_stream));
    if (h->h_flag[0]) { h->have_supports = false; return fail(h, -2, "sls_synthesize: support rowval out of range or not strictly ascending inside a column"); }
    if (status) memcpy(status, h->pipe_status.data(), (size_t)ncols * sizeof(int32_t));
    if (resid) memcpy(resid, h->pipe_resid.data(), (size_t)ncols * sizeof(double));
    for (int64_t t = 0; t < T; t++) {
        if (nnz_x) nnz_x[t] = base[t];
        if (nnz_u) nnz_u[t] = base[T + t];
    }
    if (getenv("SLS_PIPE_TRACE")) {                            // per-chunk completion times on the device clock (ms since the call)
        for (int c = 0; c < K; c++)
            fprintf(stderr, "[pipe] chunk %d: in %.2f plan %.2f scat %.2f solved %.2f out %.2f\n", c,
                    rt::elapsed_ms(h->ev[0], h->pipe_in[c]), rt::elapsed_ms(h->ev[0], h->pipe[c]->ev_plan),
                    rt::elapsed_ms(h->ev[0], h->pipe[c]->ev_scat), rt::elapsed_ms(h->ev[0], h->pipe[c]->ev_solved),
                    rt::elapsed_ms(h->ev[0], h->pipe[c]->ev_out));
        fprintf(stderr, "[pipe] total %.2f (rowval filled host-side: %.1f MB)\n", rt::elapsed_ms(h->ev[0], h->ev[1]), hcopy.bytes / 1e6);
    }
    h->tm = sls_timing{};
    h->tm.solve_ms = rt::elapsed_ms(h->ev[0], h->ev[1]);       // whole pipelined call (copies, kernels 1-3) on the device clock
    h->tm.solve_launches = launches; h->tm.gather_launches = 3 * K; h->tm.scatter_launches = 2 * K;
    h->pipe_chunks = K;                                        // alg_flops / alg_bytes: summed over the chunks' plans on demand (sls_get_timing)
    h->tm.ncols = ncols; h->tm.nnz_out = nnz_total; h->tm.alg_flops = flops; h->tm.alg_bytes = bytes;
    h->tm.refined = h->h_flag[1];
    h->tm.host_copied_bytes = hcopy.bytes;
    h->tm_from_plan = false;
    // the support columns of the range (+halo) stay on the device; a later sls_plan of other columns needs
    // sls_set_supports.  The outputs live in per-chunk buffers, so the
    // device-resident accessors (sls_device_output, sls_simulate) need a sls_plan + sls_solve_device first
    h->plan.valid = false; h->solved = false;
    return 0;
}

extern "C" int sls_synthesize(sls_handle* h, int64_t T,
                              const int64_t* const* Sx_colptr, const int64_t* const* Sx_rowval, const uint8_t* const* Sx_nzval,
                              const int64_t* const* Su_colptr, const int64_t* const* Su_rowval, const uint8_t* const* Su_nzval,
                              int64_t first_col, int64_t ncols, const int64_t* cap_x, const int64_t* cap_u,
                              int64_t* const* Px_colptr, int64_t* const* Px_rowval, double* const* Px_nzval,
                              int64_t* const* Pu_colptr, int64_t* const* Pu_rowval, double* const* Pu_nzval,
                              int64_t* nnz_x, int64_t* nnz_u, int32_t* status, double* resid) {
    return synthesize_impl(h, T, false, 0, 0, 0, 0, Sx_colptr, Sx_rowval, Sx_nzval, Su_colptr, Su_rowval, Su_nzval, first_col, ncols, cap_x, cap_u,
                           Px_colptr, Px_rowval, Px_nzval, Pu_colptr, Pu_rowval, Pu_nzval, nnz_x, nnz_u, status, resid);
}

extern "C" int sls_synthesize_window(sls_handle* h, int64_t T, int64_t win_first, int64_t win_ncols, int64_t row_off_x, int64_t row_off_u,
                                     const int64_t* const* Sx_colptr, const int64_t* const* Sx_rowval, const uint8_t* const* Sx_nzval,
                                     const int64_t* const* Su_colptr, const int64_t* const* Su_rowval, const uint8_t* const* Su_nzval,
                                     int64_t first_col, int64_t ncols, const int64_t* cap_x, const int64_t* cap_u,
                                     int64_t* const* Px_colptr, int64_t* const* Px_rowval, double* const* Px_nzval,
                                     int64_t* const* Pu_colptr, int64_t* const* Pu_rowval, double* const* Pu_nzval,
                                     int64_t* nnz_x, int64_t* nnz_u, int32_t* status, double* resid) {
    if (!h) return -1;
    SLS_ARG(ncols >= 1, "sls_synthesize_window: ncols must be >= 1");
    return synthesize_impl(h, T, true, win_first - 1, win_ncols, row_off_x, row_off_u, Sx_colptr, Sx_rowval, Sx_nzval, Su_colptr, Su_rowval, Su_nzval, first_col, ncols,
                           cap_x, cap_u, Px_colptr, Px_rowval, Px_nzval, Pu_colptr, Pu_rowval, Pu_nzval, nnz_x, nnz_u, status, resid);
}

// Page-locks a caller-owned host range so that copies to / from it are truly asynchronous (the D2H / H2D pipelines of
// sls_solve and sls_synthesize serialise on pageable memory).  Julia: call on the `Vector`s behind 𝓢[t].rowval / Φ[t].nzval
// once, under GC.@preserve; ranges still registered are released by sls_destroy.
extern "C" int sls_host_register(sls_handle* h, void* p, size_t bytes) {
    if (!h) return -1;
    SLS_ARG(p && bytes > 0, "sls_host_register: empty range");
    SLS_CK(rt::set_device(h->device));
    int e = rt::host_register(p, bytes);
    if (e) { rt::last_error(); return fail(h, e, "sls_host_register"); }
    h->registered.push_back(p);
    return 0;
}

extern "C" int sls_host_unregister(sls_handle* h, void* p) {
    if (!h) return -1;
    auto it = std::find(h->registered.begin(), h->registered.end(), p);
    SLS_ARG(it != h->registered.end(), "sls_host_unregister: range was not registered through this handle");
    h->registered.erase(it);
    int e = rt::host_unregister(p);
    if (e) { rt::last_error(); return fail(h, e, "sls_host_unregister"); }
    return 0;
}

// ---- closed-loop FIR simulation --------------------------------------------------------------
extern "C" int sls_simulate(sls_handle* h, int64_t nscen, int64_t steps, const int64_t* dist_col, const int64_t* dist_time,
                            double* xnorm, double* unorm, double* x_out, double* u_out) {
    if (!h) return -1;
    SLS_ARG(h->solved, "sls_simulate: no synthesised controller on the device (call sls_solve first)");
    SLS_ARG(nscen >= 0 && nscen < (int64_t)1 << 30 && steps >= 2 && dist_col && dist_time, "sls_simulate: bad arguments");
    SLS_CK(rt::set_device(h->device));
    const int64_t Nx = h->Nx, Nu = h->Nu, T = h->T;
    for (int64_t s = 0; s < nscen; s++)
        if (dist_col[s] < 1 || dist_col[s] > Nx || dist_time[s] < 1) return fail(h, -3, "sls_simulate: scenario out of range");
    if (h->have_weights) return fail(h, -1, "sls_simulate: not available with cost weights (the plant on the device is the scaled one)");
    if (nscen == 0) return 0;
    const Plan& p = h->plan;
    // scratch lives in the handle and is reused by later calls (no cudaMalloc / cudaFree — and the device synchronisation
    // they imply — per call); the dense trajectory buffers (x_out / u_out: test-sized) are the exception
    DevBuf& dcol = h->simb[0]; DevBuf& dtime = h->simb[1]; DevBuf& dnw = h->simb[2]; DevBuf& dnu = h->simb[3]; DevBuf& dxn = h->simb[4];
    DevBuf& dun = h->simb[5]; DevBuf& dovf = h->simb[6]; DevBuf& dW = h->simb[7]; DevBuf& dWu = h->simb[8]; DevBuf& dnent = h->simb[9];
    DevBuf& dtapcnt = h->simb[10]; DevBuf& dpkoff = h->simb[11]; DevBuf& dpkcnt = h->simb[12]; DevBuf& dpkrow = h->simb[13];
    DevBuf& dpkval = h->simb[14]; DevBuf& dctav = h->simb[15]; DevBuf& dctam = h->simb[16];
    DevBuf dxo, duo;
    int e = 0;
    auto cleanup = [&]() { dxo.release(); duo.release(); };
#define SIM_CK(expr) do { e = (expr); if (e) { cleanup(); return fail(h, e, #expr); } } while (0)
    h->ev[2].record(h->stream);
    SIM_CK(dcol.ensure(nscen * 8)); SIM_CK(dtime.ensure(nscen * 8));
    SIM_CK(dnw.ensure(nscen * 4)); SIM_CK(dnu.ensure(nscen * 4)); SIM_CK(dovf.ensure(4));
    SIM_CK(rt::h2d(dcol.p, dist_col, nscen * 8, h->stream));
    SIM_CK(rt::h2d(dtime.p, dist_time, nscen * 8, h->stream));
    SIM_CK(rt::zero(dovf.p, 4, h->stream));
    SlsSimArgs a{};
    a.Nx = Nx; a.Nu = Nu; a.steps = steps; a.T = (int)T; a.nscen = (int)nscen;
    a.colptr = p.colptr.as<int64_t>();
    a.rowval = reinterpret_cast<const int64_t* const*>(h->out_ptrs.as<void*>());
    a.nzval = reinterpret_cast<const double* const*>(h->out_ptrs.as<void*>() + 2 * T);
    a.A = h->A.view(); a.Ar = h->Ar.view(); a.B2r = h->B2r.view();
    a.dcol = dcol.as<int64_t>(); a.dtime = dtime.as<int64_t>(); a.overflow = dovf.as<int>();
    // pass 1: window sizes (scratch capacity = as much as 4 warps can have)
    // scratch lists of the window pass: 3 x cap ints per warp.  Start small (occupancy: 4 warps x 3 x 512 ints = 24 KB per
    // block) and grow only when a scenario overflows (the kernel reports -1), up to what shared memory allows.
    // The windows are stored by the same launch at stride cap and compacted to their exact strides afterwards.
    int cap = h->sim_cap > 0 ? h->sim_cap : 512;
    std::vector<int32_t> nw(nscen), nu(nscen);
    int wmax = 1, umax = 1;
    DevBuf& dWt = h->simb[17]; DevBuf& dWut = h->simb[18];
    for (;;) {
        SIM_CK(dWt.ensure((size_t)nscen * cap * 4)); SIM_CK(dWut.ensure((size_t)nscen * cap * 4));
        SIM_CK(sls_launch_sim_size(a, dnw.as<int32_t>(), dnu.as<int32_t>(), cap, dWt.as<int32_t>(), dWut.as<int32_t>(), cap, cap, h->stream));
        SIM_CK(rt::d2h(nw.data(), dnw.p, nscen * 4, h->stream));
        SIM_CK(rt::d2h(nu.data(), dnu.p, nscen * 4, h->stream));
        SIM_CK(rt::sync(h->stream));
        bool ovf_ = false;
        wmax = 1; umax = 1;
        for (int64_t s = 0; s < nscen; s++) {
            if (nw[s] < 0 || nu[s] < 0) ovf_ = true;
            wmax = std::max(wmax, nw[s]); umax = std::max(umax, nu[s]);
        }
        if (!ovf_) break;
        if ((size_t)4 * 3 * (cap * 2) * sizeof(int32_t) > (size_t)h->max_smem || cap >= 8192) {
            cleanup(); return fail(h, -4, "sls_simulate: a scenario's window does not fit shared memory");
        }
        cap *= 2;
    }
    h->sim_cap = cap;
    const int wcap = (wmax + 1) & ~1, ucap = (umax + 1) & ~1;
    SIM_CK(dW.ensure((size_t)nscen * wcap * 4)); SIM_CK(dWu.ensure((size_t)nscen * ucap * 4));
    SIM_CK(sls_launch_sim_compact(dWt.as<int32_t>(), dWut.as<int32_t>(), cap, dnw.as<int32_t>(), dnu.as<int32_t>(), (int)nscen,
                                  dW.as<int32_t>(), dWu.as<int32_t>(), wcap, ucap, h->stream));
    a.nw = dnw.as<int32_t>(); a.nu = dnu.as<int32_t>(); a.Wg = dW.as<int32_t>(); a.Wug = dWu.as<int32_t>();
    // pack pass: Phi restricted to every scenario's window, local row indices, in consumption order
    if (wcap > 65535 || ucap > 65535) { cleanup(); return fail(h, -4, "sls_simulate: window too large"); }
    SIM_CK(dnent.ensure(nscen * 8)); SIM_CK(dtapcnt.ensure((size_t)nscen * 2 * T * 4)); SIM_CK(dpkoff.ensure(nscen * 8));
    SIM_CK(sls_launch_sim_pack(a, false, wcap, ucap, dnent.as<int64_t>(), dtapcnt.as<uint32_t>(), nullptr, nullptr, nullptr, nullptr, h->stream));
    double sim_bytes = 0;
    int64_t nent_max = 0;
    {
        std::vector<int64_t> nent(nscen), off(nscen);
        std::vector<uint32_t> tapcnt((size_t)nscen * 2 * T);
        SIM_CK(rt::d2h(nent.data(), dnent.p, nscen * 8, h->stream));
        SIM_CK(rt::d2h(tapcnt.data(), dtapcnt.p, tapcnt.size() * 4, h->stream));
        SIM_CK(rt::sync(h->stream));
        // packed entries (8 B value + 2 B local row) the recursion multiplies: tap tau of scenario s is applied at every
        // step t with t_s + tau <= t <= steps - 1 (what(t+1-tau) is exactly zero before that)
        for (int64_t s = 0; s < nscen; s++)
            for (int64_t tau = 1; tau <= T; tau++) {
                const double reps = (double)std::max<int64_t>(0, steps - std::max<int64_t>(dist_time[s], 1) - tau);
                const double ent = (tau <= T - 1 ? (double)tapcnt[(size_t)s * 2 * T + tau] : 0.0) + (double)tapcnt[(size_t)s * 2 * T + T + tau - 1];
                sim_bytes += reps * ent * 10.0;
            }
        int64_t run = 0;
        for (int64_t s = 0; s < nscen; s++) { off[s] = run; run += nent[s]; nent_max = std::max(nent_max, nent[s]); }
        SIM_CK(rt::h2d(dpkoff.p, off.data(), nscen * 8, h->stream));
        SIM_CK(dpkcnt.ensure((size_t)nscen * 2 * T * wcap * 2));
        SIM_CK(dpkrow.ensure((size_t)std::max<int64_t>(run, 1) * 2));
        SIM_CK(dpkval.ensure((size_t)std::max<int64_t>(run, 1) * 8));
        SIM_CK(sls_launch_sim_pack(a, true, wcap, ucap, dnent.as<int64_t>(), dtapcnt.as<uint32_t>(), dpkoff.as<int64_t>(),
                                   dpkcnt.as<uint16_t>(), dpkrow.as<uint16_t>(), dpkval.as<double>(), h->stream));
        SIM_CK(rt::sync(h->stream));                           // `off` is a local
    }
    a.pk_off = dpkoff.as<int64_t>(); a.pk_tapcnt = dtapcnt.as<uint32_t>(); a.pk_cnt = dpkcnt.as<uint16_t>();
    a.pk_row = dpkrow.as<uint16_t>(); a.pk_val = dpkval.as<double>();
    const int cap2 = 0;
    if ((size_t)4 * sls_sim_smem_per_warp(cap2, wcap, ucap, (int)T) > (size_t)h->max_smem) {
        cleanup(); return fail(h, -4, "sls_simulate: windows do not fit shared memory");
    }
    const size_t nn = (size_t)nscen * steps;
    SIM_CK(dxn.ensure(nn * 8)); SIM_CK(dun.ensure(nn * 8));
    SIM_CK(rt::zero(dxn.p, nn * 8, h->stream)); SIM_CK(rt::zero(dun.p, nn * 8, h->stream));
    a.xnorm = dxn.as<double>(); a.unorm = dun.as<double>();
    if (x_out) { SIM_CK(dxo.ensure(nn * Nx * 8)); SIM_CK(rt::zero(dxo.p, nn * Nx * 8, h->stream)); a.xout = dxo.as<double>(); }
    if (u_out) { SIM_CK(duo.ensure(nn * std::max<int64_t>(Nu, 1) * 8)); SIM_CK(rt::zero(duo.p, nn * std::max<int64_t>(Nu, 1) * 8, h->stream)); a.uout = duo.as<double>(); }
    // fast path: one CTA per scenario, the packed window read from HBM once (then shared-memory or L2 resident)
    const int nent_cap = (int)((nent_max + 1) & ~(int64_t)1);
    const bool cta_path = !getenv("SLS_SIM_STREAM") && 2 * T <= 256 && wcap <= 256 && ucap <= 256 && nent_max < ((int64_t)1 << 24) &&
                          sls_sim_cta_smem(0, wcap, ucap, (int)T) <= (size_t)h->max_smem;
    if (cta_path) {
        if (getenv("SLS_SIM_GLOBAL") || sls_sim_cta_smem(nent_cap, wcap, ucap, (int)T) > (size_t)h->max_smem) {
            const size_t grid = (size_t)std::min<int64_t>(nscen, 2 * h->sm_count);
            SIM_CK(dctav.ensure(grid * nent_cap * 8)); SIM_CK(dctam.ensure(grid * nent_cap * 2));
            a.cta_vals = dctav.as<double>(); a.cta_meta = dctam.as<uint16_t>();
        }
        h->ev[0].record(h->stream);
        SIM_CK(sls_launch_sim_cta(a, wcap, ucap, nent_cap, h->sm_count, h->stream));
    } else {
        h->ev[0].record(h->stream);
        SIM_CK(sls_launch_sim(a, cap2, wcap, ucap, h->stream));
    }
    h->ev[1].record(h->stream);
    int ovf = 0;
    SIM_CK(rt::d2h(&ovf, dovf.p, 4, h->stream));
    if (xnorm) SIM_CK(rt::d2h(xnorm, dxn.p, nn * 8, h->stream));
    if (unorm) SIM_CK(rt::d2h(unorm, dun.p, nn * 8, h->stream));
    if (x_out) SIM_CK(rt::d2h(x_out, dxo.p, nn * Nx * 8, h->stream));
    if (u_out) SIM_CK(rt::d2h(u_out, duo.p, nn * Nu * 8, h->stream));
    SIM_CK(rt::sync(h->stream));
#undef SIM_CK
    h->tm.sim_ms = rt::elapsed_ms(h->ev[0], h->ev[1]);
    h->tm.sim_pack_ms = rt::elapsed_ms(h->ev[2], h->ev[0]);
    h->tm.sim_bytes = sim_bytes;
    h->tm.sim_window = wmax; h->tm.sim_uwindow = umax;
    cleanup();
    if (ovf) return fail(h, -4, "sls_simulate: window scratch overflow");
    return 0;
}

extern "C" int sls_device_output(sls_handle* h, int64_t t, int which,
                                 const int64_t** d_colptr, const int64_t** d_rowval, const double** d_nzval, int64_t* nnz) {
    if (!h) return -1;
    SLS_ARG(h->solved, "sls_device_output: no solved plan");
    SLS_ARG(t >= 0 && t < h->T && (which == 0 || which == 1), "sls_device_output: bad tap");
    const int64_t tt = which ? h->T + t : t;
    if (d_colptr) *d_colptr = h->plan.colptr.as<int64_t>() + tt * (h->Nx + 1);
    if (d_rowval) *d_rowval = h->out_rowval.as<int64_t>() + h->plan.out_off[tt];
    if (d_nzval) *d_nzval = h->out_nzval.as<double>() + h->plan.out_off[tt];
    if (nnz) *nnz = h->plan.nnz_tap[tt];
    return 0;
}

extern "C" int sls_device_status(sls_handle* h, const int32_t** d_status, const double** d_resid) {
    if (!h) return -1;
    SLS_ARG(h->solved, "sls_device_status: no solved plan");
    if (d_status) *d_status = h->status.as<int32_t>();
    if (d_resid) *d_resid = h->resid.as<double>();
    return 0;
}

extern "C" int sls_get_timing(const sls_handle* h, sls_timing* out) {
    if (!h || !out) return -1;
    sls_handle* hm = const_cast<sls_handle*>(h);
    if (hm->tm_from_plan && hm->plan.valid) {
        plan_stats(hm, hm->plan, hm->T);
        hm->tm.alg_flops = hm->plan.alg_flops; hm->tm.alg_bytes = hm->plan.alg_bytes;
    } else if (!hm->tm_from_plan && hm->pipe_chunks > 0) {
        double flops = 0, bytes = 0;
        for (int c = 0; c < hm->pipe_chunks && c < (int)hm->pipe.size(); c++) {
            plan_stats(hm, hm->pipe[c]->plan, hm->T);
            flops += hm->pipe[c]->plan.alg_flops; bytes += hm->pipe[c]->plan.alg_bytes;
        }
        hm->tm.alg_flops = flops; hm->tm.alg_bytes = bytes;
    }
    *out = h->tm;
    return 0;
}

extern "C" int sls_mark(sls_handle* h, int which) {
    if (!h || which < 0 || which > 1) return -1;
    SLS_CK(rt::set_device(h->device));
    h->marks[which].record(h->stream);
    return 0;
}

extern "C" int sls_elapsed_ms(sls_handle* h, float* ms) {
    if (!h || !ms) return -1;
    SLS_CK(rt::sync(h->stream));
    *ms = rt::elapsed_ms(h->marks[0], h->marks[1]);
    return 0;
}
