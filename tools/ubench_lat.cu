// Dependent-issue latency and throughput of packed FP32 ops vs ILP and warps per SM.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)
constexpr int ITERS = 4096;
// OP 0: FFMA2 reg*imm+reg  1: FADD2 reg+reg  2: FFMA2 reg*reg+reg (3 distinct)  3: scalar FFMA  4: FMUL2 imm
template <int OP, int ILP>
__global__ void k(float* out, long long* cycles, float seed) {
    float2 acc[ILP], b[ILP];
    for (int j = 0; j < ILP; ++j) { acc[j] = make_float2(1.0f + j + threadIdx.x * 1e-3f + seed, 0.5f + j); b[j] = make_float2(0.999f + j * 1e-4f + seed, 1.001f); }
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) {
            if (OP == 0) acc[j] = __ffma2_rn(acc[j], make_float2(0.999f, 0.999f), b[j]);
            if (OP == 1) acc[j] = __fadd2_rn(acc[j], b[j]);
            if (OP == 2) acc[j] = __ffma2_rn(acc[j], b[j], b[(j + 1) % ILP]);
            if (OP == 3) { acc[j].x = fmaf(acc[j].x, 0.999f, b[j].x); }
            if (OP == 4) acc[j] = __fmul2_rn(acc[j], make_float2(0.9999f, 0.9999f));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int j = 0; j < ILP; ++j) s += acc[j].x + acc[j].y;
    if (s == 123.456f) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int OP, int ILP>
void run(const char* name, int sms, int threads) {
    float* out; long long* cyc;
    CK(cudaMalloc(&out, 4)); CK(cudaMalloc(&cyc, sizeof(long long) * sms));
    k<OP, ILP><<<sms, threads>>>(out, cyc, 0.f);
    k<OP, ILP><<<sms, threads>>>(out, cyc, 0.f);
    CK(cudaDeviceSynchronize());
    std::vector<long long> h(sms); CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost));
    double avg = 0; for (auto v : h) avg += (double)v; avg /= sms;
    const double per_warp = avg / ((double)ITERS * ILP);                 // cycles per instruction as seen by one warp
    const double per_smsp = per_warp / (threads / 32 / 4.0);             // cycles per instruction per scheduler
    printf("%-22s ILP %d warps/SMSP %d: %6.2f cyc/instr/warp  %5.2f cyc/instr/SMSP\n", name, ILP, threads / 128, per_warp, per_smsp);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    for (int threads : {128, 256, 512}) {
        run<0, 1>("FFMA2 r*imm+r", sms, threads); run<0, 2>("FFMA2 r*imm+r", sms, threads); run<0, 4>("FFMA2 r*imm+r", sms, threads); run<0, 8>("FFMA2 r*imm+r", sms, threads);
        run<1, 1>("FADD2 r+r", sms, threads); run<1, 2>("FADD2 r+r", sms, threads); run<1, 4>("FADD2 r+r", sms, threads); run<1, 8>("FADD2 r+r", sms, threads);
        run<2, 2>("FFMA2 r*r+r", sms, threads); run<2, 4>("FFMA2 r*r+r", sms, threads); run<2, 8>("FFMA2 r*r+r", sms, threads);
        run<3, 1>("FFMA scalar", sms, threads); run<3, 4>("FFMA scalar", sms, threads); run<3, 8>("FFMA scalar", sms, threads);
        run<4, 1>("FMUL2 imm", sms, threads); run<4, 4>("FMUL2 imm", sms, threads);
    }
    return 0;
}
