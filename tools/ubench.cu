// Micro-benchmarks that fix the design constants of the fused log-mel kernel on B200:
// scalar vs packed (f32x2) FP32 issue/throughput, the cost of interleaved LDS, MUFU.LG2 rate,
// and how many thread-block clusters of a given size/footprint are co-resident.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu && tools/ubench
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

typedef unsigned long long u64;

__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) { u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) { u64 d; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

constexpr int ITERS = 2048;
constexpr int ILP = 8;

// MODE 0: scalar FFMA   1: FFMA2   2: FADD2   3: FMUL2   4: scalar FADD   5: MUFU.LG2
// 6: FFMA2 + LDS.64 every 4th   7: scalar FFMA + LDS.32 every 4th   8: FFMA2 x4 + (LDS.64 + STS.64)
// 9: alternate FADD2 / FFMA2
template <int MODE>
__global__ void __launch_bounds__(512) k_fp(float* out, long long* cycles) {
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(1.0f + i * 1e-6f, 1.0f - i * 1e-6f);
    __syncthreads();
    float2 acc[ILP];
    for (int j = 0; j < ILP; ++j) acc[j] = make_float2(1.0f + j + threadIdx.x * 1e-3f, 0.5f + j);
    const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
    const u64 um = *reinterpret_cast<const u64*>(&m), uc = *reinterpret_cast<const u64*>(&c);
    int idx = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            u64& a = *reinterpret_cast<u64*>(&acc[j]);
            if (MODE == 0) { acc[j].x = fmaf(acc[j].x, m.x, c.x); acc[j].y = fmaf(acc[j].y, m.y, c.y); }
            if (MODE == 1) a = ffma2(a, um, uc);
            if (MODE == 2) a = fadd2(a, uc);
            if (MODE == 3) a = fmul2(a, um);
            if (MODE == 4) { acc[j].x = acc[j].x + c.x; acc[j].y = acc[j].y + c.y; }
            if (MODE == 5) { acc[j].x = __log2f(acc[j].x); acc[j].y = __log2f(acc[j].y); }
            if (MODE == 6) { a = ffma2(a, um, uc); if ((j & 3) == 3) { float2 v = sm[(idx + j * 32) & 2047]; acc[j].x += v.x; } }
            if (MODE == 7) { acc[j].x = fmaf(acc[j].x, m.x, c.x); acc[j].y = fmaf(acc[j].y, m.y, c.y); if ((j & 3) == 3) { float v = reinterpret_cast<float*>(sm)[(idx + j * 32) & 4095]; acc[j].x += v; } }
            if (MODE == 8) { a = ffma2(a, um, uc); if ((j & 3) == 3) { float2 v = sm[(idx + j * 32) & 2047]; acc[j].x += v.x; sm[(idx + j * 32 + 1024) & 2047] = acc[j]; } }
            if (MODE == 9) { if (j & 1) a = ffma2(a, um, uc); else a = fadd2(a, uc); }
        }
        idx += 7;
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < ILP; ++j) s += acc[j].x + acc[j].y;
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run_fp(const char* name, int sms, double ops_per_iter_per_thread, int threads) {
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    k_fp<MODE><<<sms, threads>>>(out, cyc);
    CK(cudaEventRecord(e0));
    k_fp<MODE><<<sms, threads>>>(out, cyc);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<long long> h(sms); CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : h) avg += (double)v; avg /= sms;
    const double lane_ops = (double)threads * ITERS * ops_per_iter_per_thread;   // scalar-equivalent fp ops per SM
    printf("%-34s threads/SM %4d  cycles %9.0f  fp32-lane-ops/clk/SM %7.1f  warp-instr/clk/SM %5.2f  (%.3f ms, %.0f MHz eff)\n",
           name, threads, avg, lane_ops / avg, (double)threads / 32 * ITERS * ILP / avg, ms, avg / (ms * 1e3));
    cudaFree(out); cudaFree(cyc);
}

__global__ void dummy_cluster_kernel(float* p) { extern __shared__ float s[]; if (p) p[0] = s[0]; }

int main() {
    int dev = 0; CK(cudaSetDevice(dev));
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    printf("device %s  SMs %d  clock %d kHz  smem/SM %zu  smem/block optin %zu\n", prop.name, sms, prop.clockRate,
           prop.sharedMemPerMultiprocessor, prop.sharedMemPerBlockOptin);
    for (int threads : {512}) {
      if (getenv("UBENCH_FP")) {
        run_fp<0>("scalar FFMA (2 per slot pair)", sms, 2.0 * ILP, threads);
        run_fp<1>("FFMA2", sms, 2.0 * ILP, threads);
        run_fp<2>("FADD2", sms, 2.0 * ILP, threads);
        run_fp<3>("FMUL2", sms, 2.0 * ILP, threads);
        run_fp<4>("scalar FADD", sms, 2.0 * ILP, threads);
        run_fp<9>("FADD2/FFMA2 alternating", sms, 2.0 * ILP, threads);
        run_fp<5>("MUFU.LG2 (2 per j)", sms, 2.0 * ILP, threads);
        run_fp<6>("FFMA2 + LDS.64 per 4", sms, 2.0 * ILP, threads);
        run_fp<7>("scalar FFMA + LDS.32 per 4 pairs", sms, 2.0 * ILP, threads);
        run_fp<8>("FFMA2 + (LDS.64+STS.64) per 4", sms, 2.0 * ILP, threads);
      }
    }
    // co-resident clusters
    for (int cs : {1, 2, 3, 4, 5, 6, 7, 8, 16}) {
        for (int smem_kb : {100, 197}) {
            for (int threads : {512}) {
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem_kb * 1024;
                cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaFuncSetAttribute(dummy_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_kb * 1024);
                if (cs > 8) cudaFuncSetAttribute(dummy_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
                int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy_cluster_kernel, &cfg);
                printf("clusters: size %2d smem %3d KB threads %3d -> max active clusters %d (CTAs %d)%s\n", cs, smem_kb, threads, n, n * cs,
                       e == cudaSuccess ? "" : cudaGetErrorString(e));
                cudaGetLastError();
            }
        }
    }
    return 0;
}
