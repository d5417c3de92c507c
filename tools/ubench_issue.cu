// Does a packed FP32 instruction (FADD2/FFMA2, 2 cycles on the FMA pipe) block the scheduler's issue port
// for its second cycle, or can another warp / another pipe issue in its shadow?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench_issue tools/ubench_issue.cu
// Body per iteration (fully unrolled x REP): NF packed FP ops on 8 independent chains interleaved with NI "other" ops
// (OTHER 0: IADD3 on the ALU pipe, 1: LDS.64 conflict-free, 2: scalar FMUL on the FMA pipe, 3: LOP3).
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
typedef unsigned long long u64;
constexpr int ITERS = 512, REP = 8;

template <int NF, int NI, int OTHER, int FPKIND>
__global__ void __launch_bounds__(512) k(float* out, long long* cycles, int seed) {
    __shared__ float2 sm[2048];
    for (int i = threadIdx.x; i < 2048; i += blockDim.x) sm[i] = make_float2(1.f, 2.f);
    __syncthreads();
    u64 acc[8];
    unsigned ia[8];
    float fa[8];
    for (int j = 0; j < 8; ++j) { float2 v = make_float2(1.0f + j + threadIdx.x * 1e-3f, 0.5f + j); acc[j] = *reinterpret_cast<u64*>(&v); ia[j] = seed + j * 7 + threadIdx.x; fa[j] = 1.0f + j; }
    const float2 cc = make_float2(1e-3f, -1e-3f);
    const u64 uc = *reinterpret_cast<const u64*>(&cc);
    const unsigned sbase = (unsigned)__cvta_generic_to_shared(sm) + (threadIdx.x & 31) * 8;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int r = 0; r < REP; ++r) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (j < NF) {
                    if (FPKIND == 0) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(acc[j]) : "l"(uc));
                    if (FPKIND == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(acc[j]) : "l"(uc));
                    if (FPKIND == 2) { float2& v = *reinterpret_cast<float2*>(&acc[j]); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v.x) : "f"(cc.x)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(v.y) : "f"(cc.y)); }
                }
                if (j < NI) {
                    if (OTHER == 0) asm volatile("add.u32 %0, %0, %1;" : "+r"(ia[j]) : "r"(ia[(j + 3) & 7]));
                    if (OTHER == 1) { u64 v; asm volatile("ld.shared.b64 %0, [%1];" : "=l"(v) : "r"(sbase + (r * 8 + j) * 256)); ia[j] ^= (unsigned)v; }
                    if (OTHER == 2) asm volatile("mul.rn.f32 %0, %0, 0f3F7FF972;" : "+f"(fa[j]));
                    if (OTHER == 3) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(ia[j]) : "r"(ia[(j + 3) & 7]), "r"(ia[(j + 5) & 7]));
                }
            }
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < 8; ++j) { float2 v = *reinterpret_cast<float2*>(&acc[j]); s += v.x + v.y + (float)ia[j] + fa[j]; }
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int NF, int NI, int OTHER, int FPKIND>
void run(const char* name, int sms, int threads) {
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    k<NF, NI, OTHER, FPKIND><<<sms, threads>>>(out, cyc, 1);
    k<NF, NI, OTHER, FPKIND><<<sms, threads>>>(out, cyc, 1);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms); CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : h) avg += (double)v; avg /= sms;
    const double warps_per_smsp = threads / 128.0;
    const double per_iter = avg / ((double)ITERS * REP) / warps_per_smsp;    // scheduler cycles per (NF fp + NI other) group of one warp
    printf("%-44s warps/SMSP %2.0f: %6.2f cyc per group  (fp %d x%s, other %d)\n", name, warps_per_smsp, per_iter, NF,
           FPKIND == 2 ? "2 scalar" : "1 packed", NI);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    for (int threads : {128, 256, 512}) {
        run<8, 0, 0, 0>("8 FADD2", sms, threads);
        run<8, 0, 0, 1>("8 FFMA2", sms, threads);
        run<8, 0, 0, 2>("16 scalar FADD", sms, threads);
        run<0, 8, 0, 0>("8 IADD", sms, threads);
        run<0, 8, 1, 0>("8 LDS.64", sms, threads);
        run<0, 8, 2, 0>("8 FMUL", sms, threads);
        run<8, 8, 0, 0>("8 FADD2 + 8 IADD", sms, threads);
        run<8, 4, 0, 0>("8 FADD2 + 4 IADD", sms, threads);
        run<8, 8, 3, 0>("8 FADD2 + 8 LOP3", sms, threads);
        run<8, 8, 1, 0>("8 FADD2 + 8 LDS.64", sms, threads);
        run<8, 4, 1, 0>("8 FADD2 + 4 LDS.64", sms, threads);
        run<8, 8, 2, 0>("8 FADD2 + 8 FMUL", sms, threads);
        run<8, 8, 0, 1>("8 FFMA2 + 8 IADD", sms, threads);
        run<8, 8, 0, 2>("16 scalar FADD + 8 IADD", sms, threads);
        run<8, 8, 1, 2>("16 scalar FADD + 8 LDS.64", sms, threads);
    }
    return 0;
}
