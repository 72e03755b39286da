// sls_rt.h — thin runtime layer shared by the kernels and the C-ABI host code.
// Product build: nvcc, CUDA runtime.  The -DSLS_EMU branch exists only for tests/emu (a g++ build of
// the same sources on a SIMT emulator, used to debug kernel logic in the GPU-less dev container).
#pragma once
#include <cstdint>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <cstdlib>

#ifdef SLS_EMU
#include "simt_emu.h"
typedef int sls_stream_t;
namespace rt {
inline int alloc(void** p, size_t n) { *p = calloc(n ? n : 1, 1); return *p ? 0 : 2; }
inline void dfree(void* p) { free(p); }
inline int h2d(void* d, const void* h, size_t n, sls_stream_t) { memcpy(d, h, n); return 0; }
inline int d2h(void* h, const void* d, size_t n, sls_stream_t) { memcpy(h, d, n); return 0; }
inline int d2d(void* d, const void* s, size_t n, sls_stream_t) { memcpy(d, s, n); return 0; }
inline int zero(void* d, size_t n, sls_stream_t) { memset(d, 0, n); return 0; }
inline int zero2d(void* d, size_t pitch, size_t width, size_t height, sls_stream_t) {
    for (size_t i = 0; i < height; i++) memset((char*)d + i * pitch, 0, width);
    return 0;
}
inline int d2h_strided(void* h, const void* d, size_t elem, size_t pitch, size_t count, sls_stream_t) {
    for (size_t i = 0; i < count; i++) memcpy((char*)h + i * elem, (const char*)d + i * pitch, elem);
    return 0;
}
inline int copy_batch(void** dsts, void** srcs, size_t* sizes, size_t n, sls_stream_t) {
    for (size_t i = 0; i < n; i++) memcpy(dsts[i], srcs[i], sizes[i]);
    return 0;
}
inline int host_alloc(void** p, size_t n) { *p = malloc(n ? n : 1); return *p ? 0 : 2; }
inline void host_free(void* p) { free(p); }
inline int host_register(void*, size_t) { return 0; }
inline int host_unregister(void*) { return 0; }
inline int sync(sls_stream_t) { return 0; }
inline int last_error() { return 0; }
inline const char* err_str(int) { return "emu"; }
inline int set_device(int) { return 0; }
inline int device_count() { return 1; }
inline int sm_count(int) { return 4; }
inline int max_smem_optin(int) { return 227 * 1024; }
inline int stream_create(sls_stream_t* s) { *s = 0; return 0; }
inline void stream_destroy(sls_stream_t) {}
struct Event {
    std::chrono::steady_clock::time_point t;
    int create() { return 0; }
    void destroy() {}
    void record(sls_stream_t) { t = std::chrono::steady_clock::now(); }
};
inline float elapsed_ms(const Event& a, const Event& b) {
    return std::chrono::duration<float, std::milli>(b.t - a.t).count();
}
inline int stream_wait(sls_stream_t, const Event&) { return 0; }
template <class K> inline int set_max_smem(K, int) { return 0; }
template <class K> inline int max_blocks_per_sm(K, int, size_t) { return 2; }
}  // namespace rt
#define SLS_LAUNCH(kp, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kp(__VA_ARGS__); })
#define SLS_DYN_SMEM(name) unsigned char* name = emu::blk()->smem
#define SLS_UNROLL _Pragma("GCC unroll 1")
#else
#include <cuda_runtime.h>
typedef cudaStream_t sls_stream_t;
namespace rt {
inline int alloc(void** p, size_t n) { return (int)cudaMalloc(p, n ? n : 1); }
inline void dfree(void* p) { if (p) cudaFree(p); }
inline int h2d(void* d, const void* h, size_t n, sls_stream_t s) { return n ? (int)cudaMemcpyAsync(d, h, n, cudaMemcpyHostToDevice, s) : 0; }
inline int d2h(void* h, const void* d, size_t n, sls_stream_t s) { return n ? (int)cudaMemcpyAsync(h, d, n, cudaMemcpyDeviceToHost, s) : 0; }
inline int d2d(void* d, const void* s_, size_t n, sls_stream_t s) { return n ? (int)cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, s) : 0; }
inline int zero(void* d, size_t n, sls_stream_t s) { return n ? (int)cudaMemsetAsync(d, 0, n, s) : 0; }
inline int zero2d(void* d, size_t pitch, size_t width, size_t height, sls_stream_t s) {
    return (width && height) ? (int)cudaMemset2DAsync(d, pitch, 0, width, height, s) : 0;
}
inline int d2h_strided(void* h, const void* d, size_t elem, size_t pitch, size_t count, sls_stream_t s) {
    return count ? (int)cudaMemcpy2DAsync(h, elem, d, pitch, elem, count, cudaMemcpyDeviceToHost, s) : 0;
}
// a list of independent copies (any direction), in stream order
inline int copy_batch(void** dsts, void** srcs, size_t* sizes, size_t n, sls_stream_t s) {
    for (size_t i = 0; i < n; i++) {
        cudaError_t e = cudaMemcpyAsync(dsts[i], srcs[i], sizes[i], cudaMemcpyDefault, s);
        if (e != cudaSuccess) return (int)e;
    }
    return 0;
}
inline int host_alloc(void** p, size_t n) { return (int)cudaMallocHost(p, n ? n : 1); }
inline void host_free(void* p) { if (p) cudaFreeHost(p); }
inline int host_register(void* p, size_t n) { return (int)cudaHostRegister(p, n, cudaHostRegisterPortable); }
inline int host_unregister(void* p) { return (int)cudaHostUnregister(p); }
inline int sync(sls_stream_t s) { return (int)cudaStreamSynchronize(s); }
inline int last_error() { return (int)cudaGetLastError(); }
inline const char* err_str(int e) { return cudaGetErrorString((cudaError_t)e); }
inline int set_device(int d) { return (int)cudaSetDevice(d); }
inline int device_count() { int n = 0; cudaError_t e = cudaGetDeviceCount(&n); return e == cudaSuccess ? n : -(int)e; }
inline int sm_count(int dev) { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev); return v; }
inline int max_smem_optin(int dev) { int v = 0; cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev); return v; }
inline int stream_create(sls_stream_t* s) { return (int)cudaStreamCreateWithFlags(s, cudaStreamNonBlocking); }
inline void stream_destroy(sls_stream_t s) { if (s) cudaStreamDestroy(s); }
struct Event {
    cudaEvent_t e = nullptr;
    int create() { return (int)cudaEventCreate(&e); }
    void destroy() { if (e) cudaEventDestroy(e); e = nullptr; }
    void record(sls_stream_t s) { cudaEventRecord(e, s); }
};
inline float elapsed_ms(const Event& a, const Event& b) { float ms = 0; cudaEventElapsedTime(&ms, a.e, b.e); return ms; }
inline int stream_wait(sls_stream_t s, const Event& ev) { return (int)cudaStreamWaitEvent(s, ev.e, 0); }
template <class K> inline int set_max_smem(K k, int bytes) {
    return (int)cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
template <class K> inline int max_blocks_per_sm(K k, int threads, size_t smem) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k, threads, smem) != cudaSuccess) return 0;
    return nb;
}
}  // namespace rt
#define SLS_LAUNCH(kp, grid, block, smem, stream, ...) kp<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define SLS_DYN_SMEM(name) extern __shared__ __align__(16) unsigned char name[]
#define SLS_UNROLL _Pragma("unroll")
#endif
