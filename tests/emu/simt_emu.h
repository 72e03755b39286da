// simt_emu.h — TEST-ONLY SIMT emulator.  NOT part of the product, never loaded by the package.
//
// The development container has nvcc but no GPU.  To debug the *logic* of the CUDA kernels before
// spending GPU minutes, tests/emu builds the very same .cu sources with g++ (-DSLS_EMU) on top of this
// header: every CUDA thread of a block becomes a ucontext fiber, scheduled round-robin on one OS
// thread; __syncthreads/__syncwarp are cooperative barriers; warp shuffles/ballots exchange values
// through a per-warp slot array between two barriers.  Blocks run one after another.
// It is a debugging aid like compute-sanitizer: results from it are never reported as product
// results, and systemlevelcontrol.jl_b200/_capi.py refuses to load anything but the CUDA library.
#pragma once
#ifndef SLS_EMU
#error "simt_emu.h is only for -DSLS_EMU test builds"
#endif

#include <ucontext.h>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cmath>
#include <vector>
#include <mutex>
#include <functional>
#include <algorithm>
#include <chrono>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct uint3emu { unsigned x, y, z; };

namespace emu {

struct Warp {
    int count = 0; unsigned gen = 0;
    uint64_t slot[32];
    double fa[32], fb[32];
    int expected = 32;
};
struct Fiber {
    ucontext_t ctx; char* stack = nullptr; bool done = false; uint3emu tid;
};
struct Block {
    std::vector<Fiber> fibers; std::vector<Warp> warps;
    ucontext_t sched; int cur = -1;
    int bar_count = 0; unsigned bar_gen = 0; int nthreads = 0;
    unsigned char* smem = nullptr;
    std::function<void()> body;
};
inline Block*& blk() { static Block* b = nullptr; return b; }
inline uint3emu& tidx() { static uint3emu t; return t; }
inline uint3emu& bidx() { static uint3emu t; return t; }
inline dim3& bdim() { static dim3 t; return t; }
inline dim3& gdim() { static dim3 t; return t; }

inline void yield() {
    Block* b = blk(); int me = b->cur;
    swapcontext(&b->fibers[me].ctx, &b->sched);
}
inline int linear_tid() { return tidx().x + bdim().x * (tidx().y + bdim().y * tidx().z); }
inline Warp& my_warp() { return blk()->warps[linear_tid() / 32]; }
inline void warp_barrier() {
    Warp& w = my_warp(); unsigned g = w.gen;
    if (++w.count == w.expected) { w.count = 0; w.gen++; } else { while (w.gen == g) yield(); }
}
inline void block_barrier() {
    Block* b = blk(); unsigned g = b->bar_gen;
    if (++b->bar_count == b->nthreads) { b->bar_count = 0; b->bar_gen++; } else { while (b->bar_gen == g) yield(); }
}
inline void trampoline() {
    Block* b = blk();
    b->body();
    b->fibers[b->cur].done = true;
    swapcontext(&b->fibers[b->cur].ctx, &b->sched);
}
// the emulator's state is global: launches from different host threads (the feeder thread of sls_synthesize launches
// the support validation kernels) are serialised
inline std::mutex& launch_mutex() { static std::mutex m; return m; }
template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F&& f) {
    std::lock_guard<std::mutex> lock(launch_mutex());
    const size_t STACK = 256 * 1024;
    int nt = block.x * block.y * block.z;
    gdim() = grid; bdim() = block;
    std::vector<unsigned char> smem(smem_bytes + 64);
    std::vector<char*> stacks(nt);
    for (int i = 0; i < nt; i++) stacks[i] = (char*)malloc(STACK);
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        Block b; blk() = &b;
        b.nthreads = nt; b.fibers.resize(nt); b.warps.resize((nt + 31) / 32);
        for (size_t w = 0; w < b.warps.size(); w++) b.warps[w].expected = std::min(32, nt - (int)w * 32);
        uintptr_t p = (uintptr_t)smem.data(); p = (p + 63) & ~(uintptr_t)63; b.smem = (unsigned char*)p;
        b.body = f;
        bidx() = {bx, by, bz};
        for (int i = 0; i < nt; i++) {
            Fiber& fb = b.fibers[i];
            fb.tid = {(unsigned)(i % block.x), (unsigned)((i / block.x) % block.y), (unsigned)(i / (block.x * block.y))};
            getcontext(&fb.ctx);
            fb.ctx.uc_stack.ss_sp = stacks[i]; fb.ctx.uc_stack.ss_size = STACK; fb.ctx.uc_link = &b.sched;
            makecontext(&fb.ctx, (void (*)())trampoline, 0);
        }
        int live = nt;
        while (live > 0) {
            for (int i = 0; i < nt; i++) {
                Fiber& fb = b.fibers[i];
                if (fb.done) continue;
                b.cur = i; tidx() = fb.tid;
                swapcontext(&b.sched, &fb.ctx);
                if (fb.done) live--;
            }
        }
    }
    for (int i = 0; i < nt; i++) free(stacks[i]);
    blk() = nullptr;
}
// mma.sync.aligned.m8n8k4.row.col.f64: A[row=lane>>2][k=lane&3], B[k=lane&3][col=lane>>2],
// C/D[row=lane>>2][col=2*(lane&3)+{0,1}]
inline void dmma(double& c0, double& c1, double a, double b) {
    Warp& w = my_warp(); int lane = linear_tid() & 31;
    w.fa[lane] = a; w.fb[lane] = b;
    warp_barrier();
    int g = lane >> 2, q = lane & 3;
    double s0 = c0, s1 = c1;
    for (int k = 0; k < 4; k++) {
        s0 = std::fma(w.fa[4 * g + k], w.fb[4 * (2 * q) + k], s0);
        s1 = std::fma(w.fa[4 * g + k], w.fb[4 * (2 * q + 1) + k], s1);
    }
    c0 = s0; c1 = s1;
    warp_barrier();
}
}  // namespace emu

#define threadIdx (emu::tidx())
#define blockIdx (emu::bidx())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())

inline void __syncthreads() { emu::block_barrier(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
inline void __threadfence() {}
inline void __threadfence_block() {}

template <class T> inline T emu_xchg(T v, int src) {
    static_assert(sizeof(T) <= 8, "shuffle payload");
    emu::Warp& w = emu::my_warp(); int lane = emu::linear_tid() & 31;
    uint64_t raw = 0; memcpy(&raw, &v, sizeof(T)); w.slot[lane] = raw;
    emu::warp_barrier();
    T out = v;
    if (src >= 0 && src < w.expected) { uint64_t r = w.slot[src]; memcpy(&out, &r, sizeof(T)); }
    emu::warp_barrier();
    return out;
}
template <class T> inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    int lane = emu::linear_tid() & 31; int base = lane & ~(width - 1);
    return emu_xchg(v, base + (src & (width - 1)));
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32) {
    int lane = emu::linear_tid() & 31; (void)width; return emu_xchg(v, lane ^ m);
}
template <class T> inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu::linear_tid() & 31; int base = lane & ~(width - 1);
    int src = lane - (int)d; return emu_xchg(v, src < base ? lane : src);
}
template <class T> inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32) {
    int lane = emu::linear_tid() & 31; int base = lane & ~(width - 1);
    int src = lane + (int)d; return emu_xchg(v, src >= base + width ? lane : src);
}
inline unsigned __ballot_sync(unsigned, int pred) {
    emu::Warp& w = emu::my_warp(); int lane = emu::linear_tid() & 31;
    w.slot[lane] = pred ? 1 : 0; emu::warp_barrier();
    unsigned r = 0; for (int i = 0; i < w.expected; i++) r |= (unsigned)(w.slot[i] & 1) << i;
    emu::warp_barrier(); return r;
}
inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0; }
inline int __all_sync(unsigned m, int p) { return __ballot_sync(m, !p) == 0; }
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    emu::Warp& w = emu::my_warp(); int lane = emu::linear_tid() & 31;
    w.slot[lane] = v; emu::warp_barrier();
    unsigned r = 0; for (int i = 0; i < w.expected; i++) r |= (unsigned)w.slot[i];
    emu::warp_barrier(); return r;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v) {
    emu::Warp& w = emu::my_warp(); int lane = emu::linear_tid() & 31;
    w.slot[lane] = v; emu::warp_barrier();
    unsigned r = 0; for (int i = 0; i < w.expected; i++) r += (unsigned)w.slot[i];
    emu::warp_barrier(); return r;
}
inline int __reduce_add_sync(unsigned m, int v) { return (int)__reduce_add_sync(m, (unsigned)v); }
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    emu::Warp& w = emu::my_warp(); int lane = emu::linear_tid() & 31;
    w.slot[lane] = v; emu::warp_barrier();
    unsigned r = 0; for (int i = 0; i < w.expected; i++) r = std::max(r, (unsigned)w.slot[i]);
    emu::warp_barrier(); return r;
}
inline int __reduce_max_sync(unsigned m, int v) {
    return (int)(__reduce_max_sync(m, (unsigned)(v ^ 0x80000000)) ^ 0x80000000);
}

inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    // position of the offset-th set bit at or above `base` (offset >= 1), 0xffffffff if none
    int seen = 0;
    for (unsigned b = base; b < 32; b++) if ((mask >> b) & 1u) { if (++seen == offset) return b; }
    return 0xffffffffu;
}
inline double __drcp_rn(double x) { return 1.0 / x; }
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
template <class T> inline T __ldg(const T* p) { return *p; }
template <class T> inline T __ldcg(const T* p) { return *p; }
template <class T> inline void __stcg(T* p, T v) { *p = v; }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
using std::max; using std::min; using std::sqrt;

struct uint4 { unsigned int x, y, z, w; };
struct double2 { double x, y; };
inline double2 make_double2(double a, double b) { return double2{a, b}; }
