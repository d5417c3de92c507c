// How many warps per scheduler does the register-resident FFT code (fft_pfa.cuh, packed f32x2) need to saturate the
// FMA pipe?  Each warp runs rfft25 (+ the 50 window multiplies) or cfft16 (+ power) back to back on registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I whisper_context_biasing_b200/csrc -I include -o tools/ubench_fft tools/ubench_fft.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "logmel_fused.cuh"
using namespace wlm;
using namespace wlm::fused;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int ITERS = 256;

template <int WHICH>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cycles, float seed) {
    float acc = 0.f;
    long long t0 = clock64();
    if (WHICH == 0) {
        V2 y[25], o[25];
        float wv[25];
        for (int t = 0; t < 25; ++t) { y[t] = mk(seed + t + threadIdx.x, seed - t); wv[t] = 0.5f + 0.01f * t + seed; }
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
#pragma unroll
            for (int t = 0; t < 25; ++t) y[t] = mk(y[t].v.x * wv[t], y[t].v.y * wv[t]);
            fft::rfft25<V2>(y, o);
#pragma unroll
            for (int t = 0; t < 25; ++t) y[t] = o[t];
        }
        for (int t = 0; t < 25; ++t) acc += y[t].v.x + y[t].v.y;
    } else {
        V2 xr[16], xi[16];
        for (int t = 0; t < 16; ++t) { xr[t] = mk(seed + t + threadIdx.x, seed - t); xi[t] = mk(seed * t, 1.f + t); }
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            fft::cfft16<V2>(xr, xi);
#pragma unroll
            for (int t = 0; t < 16; ++t) { const V2 p = vfma(xr[t], xr[t], vmul(xi[t], xi[t])); xr[t] = vmulc(p, 1e-3f); }
        }
        for (int t = 0; t < 16; ++t) acc += xr[t].v.x + xi[t].v.y;
    }
    long long t1 = clock64();
    if (acc == 123.456f) out[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int WHICH>
void run(const char* name, int sms, int threads, double fp2_per_iter) {
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    k<WHICH><<<sms, threads>>>(out, cyc, 0.f);
    k<WHICH><<<sms, threads>>>(out, cyc, 0.f);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms); CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : h) avg += (double)v; avg /= sms;
    const double wps = threads / 128.0;
    const double per_task = avg / ITERS;                 // cycles one warp needs per transform
    const double pipe = fp2_per_iter * 2.0 * wps;        // FMA-pipe cycles per scheduler for one round of all its warps
    printf("%-8s warps/scheduler %1.0f: %7.1f cycles per transform per warp, FMA pipe busy %5.1f %%\n", name, wps, per_task, 100.0 * pipe / per_task);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    for (int threads : {128, 256, 384, 512}) {
        run<0>("rfft25", sms, threads, 188 + 25);   // 188 packed ops + 50 scalar window multiplies (= 25 packed)
        run<1>("cfft16", sms, threads, 160 + 48);   // 160 packed ops + 16 x (FMUL2, FFMA2, FMUL2)
    }
    return 0;
}
