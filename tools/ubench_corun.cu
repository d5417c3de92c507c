// Can a second kernel use the 16 SMs that the 22 six-CTA clusters of the log-mel kernel leave idle?
// Runs wlm_logmel (256 clips, 80 mels) alone, a 16-CTA dummy kernel (203 KB shared memory, spins ~0.3 ms) alone, and both
// on two streams; prints the three times and the SM ids the dummy's CTAs ran on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I include -o tools/ubench_corun tools/ubench_corun.cu -ldl
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "wlm.h"
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__global__ void __launch_bounds__(512, 1) dummy(long long spin_cycles, int* smids) {
    extern __shared__ float sm[];
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) smids[blockIdx.x] = (int)smid;
    const long long t0 = clock64();
    float acc = threadIdx.x;
    while (clock64() - t0 < spin_cycles) acc = acc * 1.0001f + 1.0f;
    sm[threadIdx.x] = acc;
}

int main() {
    void* h = dlopen("whisper_context_biasing_b200/lib/libwlm.so", RTLD_NOW);
    if (!h) { printf("dlopen: %s\n", dlerror()); return 1; }
    auto plan_create = (int (*)(int, int, const float*, wlm_plan**))dlsym(h, "wlm_plan_create");
    auto logmel = (int (*)(wlm_plan*, const void*, int, const int64_t*, const int32_t*, int64_t, int, float*, float*, void*, size_t, void*))dlsym(h, "wlm_logmel");
    auto last_error = (const char* (*)())dlsym(h, "wlm_last_error");
    const int M = 80, B = 256;
    // any triangular two-adjacent table will do for timing
    std::vector<float> mel(201 * M, 0.f);
    for (int k = 1; k < 200; ++k) { const float pos = k * (M - 1) / 200.0f; const int m = (int)pos; const float f = pos - m; mel[k * M + m] = 1.f - f; if (m + 1 < M) mel[k * M + m + 1] = f; }
    wlm_plan* plan = nullptr;
    if (plan_create(0, M, mel.data(), &plan) != 0) { printf("plan: %s\n", last_error()); return 1; }
    float *pcm, *out; int* smids; void* ws;
    CK(cudaMalloc(&pcm, (size_t)B * 480000 * 4)); CK(cudaMalloc(&out, (size_t)B * M * 3000 * 4)); CK(cudaMalloc(&smids, 64 * 4)); CK(cudaMalloc(&ws, 1 << 20));
    CK(cudaMemset(pcm, 0, (size_t)B * 480000 * 4));
    cudaStream_t s1, s2; CK(cudaStreamCreate(&s1)); CK(cudaStreamCreate(&s2));
    const int smem = 203 * 1024;
    CK(cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](bool a, bool b, const char* name) {
        float best = 1e9f;
        for (int it = 0; it < 5; ++it) {
            CK(cudaDeviceSynchronize());
            CK(cudaEventRecord(e0, s1));
            CK(cudaStreamWaitEvent(s2, e0, 0));
            if (a && logmel(plan, pcm, 0, nullptr, nullptr, 480000, B, out, nullptr, ws, 1 << 20, s1) != 0) { printf("logmel: %s\n", last_error()); exit(1); }
            if (b) dummy<<<16, 512, smem, s2>>>(600000, smids);
            cudaEvent_t e2; CK(cudaEventCreate(&e2)); CK(cudaEventRecord(e2, s2)); CK(cudaStreamWaitEvent(s1, e2, 0));
            CK(cudaEventRecord(e1, s1));
            CK(cudaDeviceSynchronize());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            if (ms < best) best = ms;
            CK(cudaEventDestroy(e2));
        }
        printf("%-28s %.3f ms\n", name, best);
    };
    run(true, false, "log-mel alone");
    run(false, true, "dummy (16 CTAs) alone");
    run(true, true, "both, two streams");
    std::vector<int> hs(16); CK(cudaMemcpy(hs.data(), smids, 64, cudaMemcpyDeviceToHost));
    printf("dummy CTAs ran on SMs:"); for (int i = 0; i < 16; ++i) printf(" %d", hs[i]); printf("\n");
    return 0;
}
