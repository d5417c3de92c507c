// Ceiling of the FFT half of the kernel WITHOUT any synchronisation between warps: every warp runs stage 1 (raw loads, Hann
// window, 25-point real DFTs, Y stores) and stage 2 (Y loads, 16-point complex DFTs, |X|^2, P stores) of the real kernel
// (logmel_fused.cuh, same shared-memory layout and footprint) back to back on a raw buffer that is filled once -- no TMA,
// no mbarriers, no mel stage, no output.  If even this loop takes about as long per 64 frames as the real kernel's step,
// the hand-overs are not what limits the kernel; if it is much faster, they are.
//   MODE 0: stage 1 + stage 2     MODE 1: stage 1 only     MODE 2: stage 2 only
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I whisper_context_biasing_b200/csrc -I include -o tools/ubench_fe tools/ubench_fe.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "logmel_fused.cuh"
using namespace wlm;
using namespace wlm::fused;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int ITERS = 200;

template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) fe_loop(const float* __restrict__ win_lane, float* out, long long* cycles, int nwarps) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);
    const int grp = warp / kGroupWarps, wg = warp % kGroupWarps;
    unsigned char* gbase = smem + grp * kSmemGroup;
    float* raw = reinterpret_cast<float*>(gbase);
    float2* Y = reinterpret_cast<float2*>(gbase + kSmemRaw) + wg * kYWarpFloat2;
    float* P = reinterpret_cast<float*>(gbase + kSmemRaw + kSmemY);
    for (int i = tid; i < kGroups * kSmemGroup / 4; i += blockDim.x) reinterpret_cast<float*>(smem)[i] = 1e-3f * (i % 977);
    __syncthreads();
    float wv[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) wv[t] = win_lane[(lane & 15) * 25 + t];
    long long t0 = clock64();
    if (warp < nwarps) {
#pragma unroll 1
        for (int it = 0; it < ITERS; ++it) {
            if (MODE != 2) stage1(stage1_base(raw, wg, lane), Y, wv, wg, lane, [&]() {}, [&]() {});
            __syncwarp();
            if (MODE != 1) stage2(Y, P, wg, lane, [&]() {});
            __syncwarp();
        }
    }
    long long t1 = clock64();
    __syncthreads();
    if (P[tid] == 123.456f) out[0] = P[tid];
    if (tid == 0) cycles[blockIdx.x] = t1 - t0;
    if (lane == 0 && warp < nwarps && warp != 0 && t1 - t0 > 0) atomicMax(reinterpret_cast<unsigned long long*>(cycles + gridDim.x + blockIdx.x), (unsigned long long)(t1 - t0));
}

template <int MODE>
void run(const char* name, int sms, int nwarps, const float* d_win) {
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, sizeof(long long) * 2 * sms));
    CK(cudaMemset(cyc, 0, sizeof(long long) * 2 * sms));
    CK(cudaFuncSetAttribute(fe_loop<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    fe_loop<MODE><<<sms, kThreads, kSmemBytes>>>(d_win, out, cyc, nwarps);
    fe_loop<MODE><<<sms, kThreads, kSmemBytes>>>(d_win, out, cyc, nwarps);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(2 * sms); CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * 2 * sms, cudaMemcpyDeviceToHost));
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)(h[sms + i] > h[i] ? h[sms + i] : h[i]); avg /= sms;
    // one iteration of one warp = 4 frames; nwarps warps -> 4 * nwarps frames per iteration
    const double per64 = avg / ITERS * 64.0 / (4.0 * nwarps);
    printf("%-18s warps %2d: %8.1f cycles per iteration, %8.1f cycles per 64 frames per SM\n", name, nwarps, avg / ITERS, per64);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    std::vector<float> win(16 * 25, 0.5f);
    float* d_win; CK(cudaMalloc(&d_win, win.size() * 4)); CK(cudaMemcpy(d_win, win.data(), win.size() * 4, cudaMemcpyHostToDevice));
    for (int nw : {4, 8, 12, 16}) {
        run<0>("stage 1 + stage 2", sms, nw, d_win);
        run<1>("stage 1 only", sms, nw, d_win);
        run<2>("stage 2 only", sms, nw, d_win);
    }
    return 0;
}
