// sls_peak.cu — measured FP64 roofline denominators for kernel 2 (MEASURED_PEAKS.json has no FP64 entry,
// SURVEY §7.2): a register-resident DFMA chain kernel (vector FP64 pipe) and a DMMA m8n8k4 chain kernel
// (FP64 tensor path).  Calibration only — never part of a synthesis call.
#include "../../include/slsb200.h"
#include "sls_rt.h"

#ifndef SLS_EMU
__global__ void __launch_bounds__(512) sls_dfma_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
            x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
        }
    }
    const double s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456) out[0] = s;
}

__global__ void __launch_bounds__(512) sls_dmma_peak_kernel(double* out, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int k = 0; k < 8; k++) { c[k][0] = threadIdx.x * 1e-3 + k; c[k][1] = c[k][0] + 0.5; }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c[k][0]), "+d"(c[k][1]) : "d"(a), "d"(b));
        }
    }
    double s = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) s += c[k][0] + c[k][1];
    if (s == 123.456) out[0] = s;
}

extern "C" int sls_fp64_peak(int device, int iters, double* dfma_tflops, double* dmma_tflops) {
    if (iters < 1) iters = 2000;
    int e = rt::set_device(device);
    if (e) return e;
    const int sms = rt::sm_count(device);
    double* out = nullptr;
    if ((e = rt::alloc((void**)&out, 8))) return e;
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    const int grid = sms * 4, block = 512;
    float best_f = 1e30f, best_m = 1e30f;
    for (int rep = 0; rep < 5; rep++) {
        cudaEventRecord(t0);
        sls_dfma_peak_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-6);
        cudaEventRecord(t1); cudaEventSynchronize(t1);
        float ms; cudaEventElapsedTime(&ms, t0, t1); if (rep && ms < best_f) best_f = ms;
        cudaEventRecord(t0);
        sls_dmma_peak_kernel<<<grid, block>>>(out, iters, 0.999999, 1e-6);
        cudaEventRecord(t1); cudaEventSynchronize(t1);
        cudaEventElapsedTime(&ms, t0, t1); if (rep && ms < best_m) best_m = ms;
    }
    e = (int)cudaGetLastError();
    cudaEventDestroy(t0); cudaEventDestroy(t1);
    rt::dfree(out);
    const double threads = (double)grid * block;
    if (dfma_tflops) *dfma_tflops = threads * iters * 64.0 * 2.0 / (best_f * 1e-3) / 1e12;
    if (dmma_tflops) *dmma_tflops = (threads / 32.0) * iters * 8.0 * 512.0 / (best_m * 1e-3) / 1e12;
    return e;
}
#else
extern "C" int sls_fp64_peak(int, int, double*, double*) { return -1; }
#endif
