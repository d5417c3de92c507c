// v0: correctness-first two-kernel form.
//   K1  reflect-pad/frame/Hann/direct DFT-400/|X|^2/sparse mel/log10 -> unclamped log-mel in
//       `out`, per-clip max via an ordered-int atomic.
//   K2  max(x, gmax-8), (x+4)/4 in place.
// Kept only as the bring-up path and as an on-device cross-check of the fused kernel
// (tools/crosscheck.py); the product entry point does not dispatch to it.
#pragma once
#include "wlm_common.cuh"

namespace wlm {
namespace v0 {

constexpr int kFramesPerCta = 16;
constexpr int kThreads = 256;
constexpr int kSlice = (kFramesPerCta - 1) * kHop + kNfft;  // 2800 samples

__device__ __forceinline__ float load_sample(const ClipArgs& a, int64_t base, int len, int s) {
    // s indexes the zero-padded/truncated 480000 buffer after torch.stft's reflect padding:
    // reflect is applied to the *padded* buffer (TF-FE:149), so mirror first, then length-test.
    if (s < 0) s = -s;
    if (s >= kNSamples) s = 2 * (kNSamples - 1) - s;
    if (s >= len) return 0.0f;
    if (a.pcm_format == WLM_PCM_I16)
        return static_cast<float>(reinterpret_cast<const int16_t*>(a.pcm)[base + s]) * (1.0f / 32768.0f);
    return reinterpret_cast<const float*>(a.pcm)[base + s];
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.0f)
        atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

__global__ void init_gmax(float* gmax, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) gmax[i] = __int_as_float(0xff800000);  // -inf
}

// tw: [400] float2 (cos, -sin)(2 pi j / 400); win: [400]; mel_dense: [n_mels][201]
__global__ void __launch_bounds__(kThreads)
logmel_k1(ClipArgs a, const float2* __restrict__ tw, const float* __restrict__ win,
          const float* __restrict__ mel_dense, const int16_t* __restrict__ mel_klo,
          const int16_t* __restrict__ mel_khi) {
    __shared__ float s_pcm[kSlice];
    __shared__ float2 s_tw[kNfft];
    __shared__ float s_win[kNfft];
    __shared__ float s_pow[kFramesPerCta][kNFreq + 1];

    const int b = blockIdx.y;
    const int f0 = blockIdx.x * kFramesPerCta;
    const int64_t base = a.offsets ? a.offsets[b] : static_cast<int64_t>(b) * a.row_stride;
    int len = a.lengths ? a.lengths[b] : a.dense_len;
    len = max(0, min(len, kNSamples));

    for (int i = threadIdx.x; i < kSlice; i += kThreads)
        s_pcm[i] = load_sample(a, base, len, f0 * kHop - kNfft / 2 + i);
    for (int i = threadIdx.x; i < kNfft; i += kThreads) {
        s_tw[i] = tw[i];
        s_win[i] = win[i];
    }
    __syncthreads();

    for (int w = threadIdx.x; w < kFramesPerCta * kNFreq; w += kThreads) {
        const int fr = w / kNFreq, k = w - fr * kNFreq;
        const float* x = s_pcm + fr * kHop;
        float re = 0.f, im = 0.f;
        int j = 0;
        for (int n = 0; n < kNfft; ++n) {
            const float v = x[n] * s_win[n];
            const float2 t = s_tw[j];
            re = fmaf(v, t.x, re);
            im = fmaf(v, t.y, im);
            j += k;
            if (j >= kNfft) j -= kNfft;
        }
        s_pow[fr][k] = re * re + im * im;
    }
    __syncthreads();

    float local_max = __int_as_float(0xff800000);
    for (int w = threadIdx.x; w < kFramesPerCta * a.n_mels; w += kThreads) {
        const int m = w / kFramesPerCta, fr = w - m * kFramesPerCta;
        const int t = f0 + fr;
        if (t >= kNFrames) continue;
        float acc = 0.f;
        const float* wrow = mel_dense + m * kNFreq;
        for (int k = mel_klo[m]; k <= mel_khi[m]; ++k) acc = fmaf(wrow[k], s_pow[fr][k], acc);
        const float lg = log10f(fmaxf(acc, 1e-10f));
        a.out[(static_cast<int64_t>(b) * a.n_mels + m) * kNFrames + t] = lg;
        local_max = fmaxf(local_max, lg);
    }
    for (int o = 16; o > 0; o >>= 1) local_max = fmaxf(local_max, __shfl_xor_sync(0xffffffffu, local_max, o));
    if ((threadIdx.x & 31) == 0) atomic_max_float(a.gmax + b, local_max);
}

__global__ void logmel_k2(float* __restrict__ out, const float* __restrict__ gmax, int n_mels, int B) {
    const int64_t per_clip = static_cast<int64_t>(n_mels) * kNFrames;
    const int64_t total = per_clip * B;
    for (int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) * 4; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x * 4) {
        const float floor_v = gmax[i / per_clip] - 8.0f;  // per_clip % 4 == 0: one clip per float4
        float4 v = *reinterpret_cast<float4*>(out + i);
        v.x = (fmaxf(v.x, floor_v) + 4.0f) * 0.25f;
        v.y = (fmaxf(v.y, floor_v) + 4.0f) * 0.25f;
        v.z = (fmaxf(v.z, floor_v) + 4.0f) * 0.25f;
        v.w = (fmaxf(v.w, floor_v) + 4.0f) * 0.25f;
        *reinterpret_cast<float4*>(out + i) = v;
    }
}

}  // namespace v0
}  // namespace wlm
