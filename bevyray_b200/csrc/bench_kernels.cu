// bench_kernels.cu — measurement helper, not on the render path: an FP32 FMA throughput probe that
// gives the roofline denominator MEASURED_PEAKS.json lacks (it has HBM and bf16 tensor peaks only).

#include "../../include/bevyray_b200.h"
#include <cuda_runtime.h>

namespace {

// 16 independent FFMA chains per thread: enough ILP to saturate the FMA pipe at 4 cycles latency
__global__ void __launch_bounds__(256) fma_probe_kernel(float* out, int iters, float a, float b) {
    float x[16];
#pragma unroll
    for (int k = 0; k < 16; k++) x[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) x[k] = __fmaf_rn(x[k], a, b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; k++) s += x[k];
    if (s == 123456.789f) out[0] = s;   // never true; keeps the chains alive
}

}  // namespace

extern "C" int bvr_bench_fp32_peak(int device, float* tflops_out) {
    if (!tflops_out) return BVR_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) { cudaGetLastError(); return BVR_ERR_NO_DEVICE; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { cudaGetLastError(); return BVR_ERR_CUDA; }
    float* d = nullptr;
    cudaEvent_t e0, e1;
    if (cudaMalloc(&d, 4) != cudaSuccess) { cudaGetLastError(); return BVR_ERR_CUDA; }
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 14;
    float best = 0.0f;
    for (int rep = 0; rep < 6; rep++) {
        cudaEventRecord(e0);
        fma_probe_kernel<<<blocks, threads>>>(d, iters, 0.999f, 1e-4f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaGetLastError(); cudaFree(d); return BVR_ERR_CUDA; }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 16.0 * (double)iters * (double)blocks * (double)threads;
        const float tf = (float)(flops / (ms * 1e-3) / 1e12);
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *tflops_out = best;
    return BVR_OK;
}
