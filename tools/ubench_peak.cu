// Measured FP32 CUDA-core peak of the device bench.py runs on: dependent-free scalar FFMA chains on every SM, timed with
// CUDA events.  bench.py loads this through ctypes (tools/libwlm_ubench.so, built by __graft_entry__.build()) and reports
// the FP32 side of the roofline "of measured" instead of "of nominal" (BASELINE.md 3).  Not part of the
// product library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -shared -Xcompiler -fPIC -o tools/libwlm_ubench.so tools/ubench_peak.cu
#include <cuda_runtime.h>

namespace {
constexpr int kIters = 4096, kIlp = 8, kThreads = 1024;

__global__ void __launch_bounds__(kThreads) ffma_kernel(float* out, float m, float c) {
    float acc[kIlp];
#pragma unroll
    for (int j = 0; j < kIlp; ++j) acc[j] = 1.0f + j + threadIdx.x * 1e-3f;
#pragma unroll 1
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int j = 0; j < kIlp; ++j) acc[j] = fmaf(acc[j], m, c);
    }
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kIlp; ++j) s += acc[j];
    if (s == 123.456f) out[0] = s;
}
}  // namespace

// returns the best of `reps` runs in TFLOP/s (2 flop per FFMA), or a negative CUDA error code
extern "C" double wlm_ubench_ffma_tflops(int device, int reps) {
    if (cudaSetDevice(device) != cudaSuccess) return -1.0;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -2.0;
    const int blocks = prop.multiProcessorCount * 2 * 8;       // 2 resident CTAs of 1024 threads per SM, 8 waves
    float* out = nullptr;
    if (cudaMalloc(&out, 4) != cudaSuccess) return -3.0;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int r = 0; r < reps + 1; ++r) {
        cudaEventRecord(e0);
        ffma_kernel<<<blocks, kThreads>>>(out, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { best = -4.0; break; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * kIlp * kIters * (double)kThreads * blocks;
        if (r > 0 && ms > 0.f && flop / (ms * 1e-3) / 1e12 > best) best = flop / (ms * 1e-3) / 1e12;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}
