// Fused log-mel kernel for sm_100a.
//
// One CTA (16 warps) owns a tile of 64 consecutive frames of one clip and walks it through
//
//   TMA (cp.async.bulk, mbarrier)  raw PCM  ->  shared memory, two regions 16 banks apart
//   stage 1  per warp: 4 frames x 16 sub-transforms; lane = (n1, frame group); 25-point real DFT of
//            the Hann-windowed samples n = (25 n1 + 16 n2) mod 400, packed f32x2 over two frames
//   stage 2  per warp: one k2 slot for 32 frame pairs; 16-point complex DFT over n1, |X|^2
//   mel      per warp: a run of filters for 32 frame pairs; sparse gather, log10, running max
//
// with the prime-factor index maps of fft_pfa.cuh (no twiddles between the stages).  All arithmetic
// on the data path is FADD2 / FMUL2 / FFMA2 on (frame a, frame b) pairs with immediate constants.
//
// Shared memory (bytes):  raw 42,880 | Y 102,528 | P 51,456 | mbarrier 16
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include "wlm_common.cuh"

namespace wlm {
namespace fused {

// ---- packed two-frame value --------------------------------------------------------------------
struct V2 {
    float2 v;
};
__device__ __forceinline__ V2 mk(float a, float b) { V2 r; r.v = make_float2(a, b); return r; }
__device__ __forceinline__ V2 vadd(V2 a, V2 b) { V2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ V2 vsub(V2 a, V2 b) { V2 r; r.v = __fadd2_rn(a.v, make_float2(-b.v.x, -b.v.y)); return r; }
__device__ __forceinline__ V2 vmul(V2 a, V2 b) { V2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ V2 vfma(V2 a, V2 b, V2 c) { V2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
__device__ __forceinline__ V2 vmulc(V2 a, float s) { V2 r; r.v = __fmul2_rn(a.v, make_float2(s, s)); return r; }
__device__ __forceinline__ V2 vfmac(V2 a, float s, V2 c) { V2 r; r.v = __ffma2_rn(a.v, make_float2(s, s), c.v); return r; }

}  // namespace fused
namespace fft {
using fused::vadd; using fused::vsub; using fused::vmul; using fused::vfma; using fused::vmulc; using fused::vfmac;
}
}  // namespace wlm

#include "fft_pfa.cuh"

namespace wlm {
namespace fused {

constexpr int kTile = 64;                 // frames per tile
constexpr int kWarps = 16;
constexpr int kThreads = kWarps * 32;
constexpr int kTilesPerClip = (kNFrames + kTile - 1) / kTile;  // 47
constexpr int kRegion = 31 * kHop + kNfft;                     // 5360 samples: frames 0..31 of a half tile
constexpr int kRegionStep = 32 * kHop;                         // 5120: region B starts 32 frames later
constexpr int kRawFloats = 2 * kRegion;                        // 10720 (5360 = 16 mod 32: regions 16 banks apart)
constexpr int kTileSamples = 63 * kHop + kNfft;                // 10480
constexpr int kYStride = 801;                                  // float2 per n1 row (25*32 + 1: odd)
constexpr int kYFloat2 = 16 * kYStride;
constexpr int kPFloat2 = kNFreq * 32;

constexpr int kSmemRaw = kRawFloats * 4;        // 42,880
constexpr int kSmemY = kYFloat2 * 8;            // 102,528
constexpr int kSmemP = kPFloat2 * 8;            // 51,456
constexpr int kSmemBytes = kSmemRaw + kSmemY + kSmemP + 64;

// Everything the kernel reads with warp-uniform indices, passed by value (constant bank).
//
// Mel projection in streaming form: FFT bin k adds w_lo[k]*P[k] to filter lo[k] and w_hi[k]*P[k] to
// filter lo[k]+1 (host-built from the caller's dense table, weights bit-identical).  A warp walks its
// bin range with two accumulators; shift[k] = lo[k] - lo[k-1] says how many filters complete before
// bin k is consumed.
struct KernelTables {
    float4 w4[kNFreq + 3];          // (w_lo, w_lo, w_hi, w_hi) per bin
    uint8_t shift[kNFreq + 3];      // filters completed before bin k
    int16_t warp_m0[kWarps + 1];    // filters [warp_m0[w], warp_m0[w+1]) belong to warp w
    int16_t warp_kb[kWarps];        // first bin of warp w
    int16_t warp_ke[kWarps];        // one past the last bin of warp w
    int16_t warp_cur[kWarps];       // filter fed by w_lo at bin warp_kb[w]  (may be warp_m0[w]-1)
    // stage 2: per k2 slot, float2 offsets into Y (component) and into P (output bin) per FFT16 output
    int16_t slot_comp_off[16];      // comp * 32
    int16_t slot_pbin_off[13][16];  // output_bin(k1, k2) * 32, indexed by cfft16 array position
    int16_t n_mels;
};
using MelParams = KernelTables;

// Host-visible tables
struct Tables {
    float win_lane[16 * 25];     // Hann window at n = (25 n1 + 16 t) mod 400
    KernelTables mel;
};

// returns 0, or -1 if a filter has no bins (not supported by the fused path)
inline int build_tables(const MelSparse& sp, int n_mels, Tables* t) {
    for (int n1 = 0; n1 < 16; ++n1)
        for (int tt = 0; tt < 25; ++tt) {
            const int n = (25 * n1 + 16 * tt) % 400;
            t->win_lane[n1 * 25 + tt] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / 400.0));
        }
    KernelTables& mp = t->mel;
    memset(&mp, 0, sizeof(mp));
    mp.n_mels = (int16_t)n_mels;
    int klo[kMaxMels], khi[kMaxMels];
    for (int m = 0; m < n_mels; ++m) { klo[m] = -1; khi[m] = -1; }
    for (int k = 0; k < kNFreq; ++k) {
        mp.w4[k] = make_float4(sp.w_lo[k], sp.w_lo[k], sp.w_hi[k], sp.w_hi[k]);
        mp.shift[k] = (uint8_t)(k == 0 ? 0 : sp.lo[k] - sp.lo[k - 1]);
        if (sp.w_lo[k] != 0.0f && sp.lo[k] >= 0) { const int m = sp.lo[k]; if (klo[m] < 0) klo[m] = k; khi[m] = k; }
        if (sp.w_hi[k] != 0.0f && sp.lo[k] + 1 < n_mels) { const int m = sp.lo[k] + 1; if (klo[m] < 0) klo[m] = k; khi[m] = k; }
    }
    for (int m = 0; m < n_mels; ++m)
        if (klo[m] < 0) return -1;
    // contiguous filter runs per warp, balanced on (14 + 4.5 * bins) per filter
    double total = 0;
    for (int m = 0; m < n_mels; ++m) total += 14.0 + 4.5 * (khi[m] - klo[m] + 1);
    int m = 0;
    double acc = 0;
    for (int w = 0; w < kWarps; ++w) {
        mp.warp_m0[w] = (int16_t)m;
        const double target = total * (w + 1) / kWarps;
        while (m < n_mels) {
            const double c = 14.0 + 4.5 * (khi[m] - klo[m] + 1);
            if (acc + 0.5 * c > target && n_mels - m <= (kWarps - 1 - w) * kMaxMels) break;
            acc += c;
            ++m;
        }
    }
    mp.warp_m0[kWarps] = (int16_t)n_mels;
    if (m < n_mels) return -1;
    for (int w = 0; w < kWarps; ++w) {
        const int ma = mp.warp_m0[w], mb = mp.warp_m0[w + 1];
        if (ma >= mb) { mp.warp_kb[w] = 0; mp.warp_ke[w] = 0; mp.warp_cur[w] = 0; continue; }
        mp.warp_kb[w] = (int16_t)klo[ma];
        mp.warp_ke[w] = (int16_t)(khi[mb - 1] + 1);
        mp.warp_cur[w] = (int16_t)sp.lo[klo[ma]];
    }
    for (int s2 = 0; s2 < fft::kNumSlots; ++s2) {
        mp.slot_comp_off[s2] = (int16_t)(fft::kSlotComp[s2] * 32);
        for (int k1 = 0; k1 < 16; ++k1)
            mp.slot_pbin_off[s2][fft::fft16_slot_of_k1(k1)] = (int16_t)(fft::output_bin(k1, fft::kSlotK2[s2]) * 32);
    }
    return 0;
}

// ---- PTX helpers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- tile bookkeeping (all values CTA-uniform) ----------------------------------------------------
struct TileCtx {
    int b, tile, f0, s0, len;
    int64_t base;
    bool active, needs_fix;
};

__device__ __forceinline__ TileCtx tile_ctx(const ClipArgs& a, int b, int tile) {
    TileCtx c;
    c.b = b;
    c.tile = tile;
    c.f0 = tile * kTile;
    c.s0 = c.f0 * kHop - kNfft / 2;
    c.base = a.offsets ? a.offsets[b] : static_cast<int64_t>(b) * a.row_stride;
    int len = a.lengths ? a.lengths[b] : a.dense_len;
    if (!a.offsets) len = static_cast<int>(min(static_cast<int64_t>(len), a.row_stride));
    c.len = max(0, min(len, kNSamples));
    c.active = c.s0 < c.len;
    c.needs_fix = (c.s0 < 0) || (c.s0 + kTileSamples > c.len);
    return c;
}

// issued by one thread: both regions of the tile, valid sample range only
__device__ __forceinline__ void tile_issue_tma(const ClipArgs& a, const TileCtx& c, float* raw, uint32_t bar) {
    const int esz = a.pcm_format == WLM_PCM_I16 ? 2 : 4;
    const int gran = 16 / esz;
    const int len_up = min((c.len + gran - 1) / gran * gran, kNSamples);
    uint32_t total = 0;
    int lo[2], n[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int s_lo = c.s0 + r * kRegionStep;
        lo[r] = max(s_lo, 0);
        const int hi = min(s_lo + kRegion, len_up);
        n[r] = max(hi - lo[r], 0);
        total += static_cast<uint32_t>(n[r]) * esz;
    }
    mbar_expect_tx(bar, total);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (n[r] <= 0) continue;
        const int s_lo = c.s0 + r * kRegionStep;
        const char* src = static_cast<const char*>(a.pcm) + (c.base + lo[r]) * esz;
        uint32_t dst;
        if (esz == 4) dst = smem_u32(raw) + static_cast<uint32_t>(r * kRegion + (lo[r] - s_lo)) * 4u;
        else dst = smem_u32(raw) + static_cast<uint32_t>(kRegion) * 4u + static_cast<uint32_t>(r * kRegion + (lo[r] - s_lo)) * 2u;
        tma_bulk_g2s(dst, src, static_cast<uint32_t>(n[r]) * esz, bar);
    }
}

// int16 -> float32 expansion in place (staging sits in the byte range of region B) + reflect /
// zero-fill patching of every position outside [0, len).  Only edge tiles and int16 input pay.
__device__ __forceinline__ void tile_fixup(const ClipArgs& a, const TileCtx& c, float* raw) {
    const int tid = threadIdx.x;
    if (a.pcm_format == WLM_PCM_I16) {
        const int16_t* st = reinterpret_cast<const int16_t*>(raw + kRegion);
        constexpr float kScale = 1.0f / 32768.0f;
        for (int i = tid; i < kRegion; i += kThreads) raw[i] = static_cast<float>(st[i]) * kScale;
        float tmp[(kRegion + kThreads - 1) / kThreads];
#pragma unroll
        for (int j = 0; j < (kRegion + kThreads - 1) / kThreads; ++j) {
            const int i = tid + j * kThreads;
            tmp[j] = i < kRegion ? static_cast<float>(st[kRegion + i]) * kScale : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < (kRegion + kThreads - 1) / kThreads; ++j) {
            const int i = tid + j * kThreads;
            if (i < kRegion) raw[kRegion + i] = tmp[j];
        }
        __syncthreads();
    }
    if (c.needs_fix) {
        for (int idx = tid; idx < kRawFloats; idx += kThreads) {
            const int r = idx >= kRegion ? 1 : 0;
            const int s = c.s0 + r * kRegionStep + (idx - r * kRegion);
            if (s >= 0 && s < c.len) continue;
            // reflect of the zero-padded 480000 buffer (torch.stft center=True, TF-FE:149)
            int sr = s < 0 ? -s : (s >= kNSamples ? 2 * (kNSamples - 1) - s : s);
            float v = 0.f;
            if (sr >= 0 && sr < c.len) {
                const int u = sr - c.s0;
                if (u >= 0 && u < kTileSamples) v = raw[u < kRegion ? u : kRegion + (u - kRegionStep)];
            }
            raw[idx] = v;
        }
        __syncthreads();
    }
}

// ---- stage 1 ----------------------------------------------------------------------------------
// warp w, lane (n1 = lane & 15, g = lane >> 4): frames (32 g + w, 32 g + w + 16) of the tile.
__device__ __forceinline__ void stage1(const float* raw, float2* Y, const float (&wv)[25], int tw, int warp, int lane) {
    const int n1 = lane & 15, g = lane >> 4;
    const float* p0 = raw + g * kRegion + kHop * warp + 25 * n1;
    const float* p1 = p0 - kNfft;
    V2 y[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) {
        const float* p = (t >= tw) ? p1 : p0;
        const float xa = p[16 * t], xb = p[16 * t + 16 * kHop];
        y[t] = mk(xa * wv[t], xb * wv[t]);
    }
    V2 out[25];
    fft::rfft25<V2>(y, out);
    float2* yo = Y + n1 * kYStride + (warp + 16 * g);
#pragma unroll
    for (int c = 0; c < 25; ++c) yo[c * 32] = out[c].v;
}

// ---- stage 2 ----------------------------------------------------------------------------------
// warp = k2 slot (uniform), lane = frame pair.  One code path for all 13 slots: the slot only selects
// table offsets, so every warp runs the same instructions (the 13-way templated version thrashed the
// instruction cache: 28 % of issue stalls were "no instruction").
__device__ __forceinline__ void stage2(const KernelTables& kt, const float2* Y, float2* P, int slot, int lane) {
    const float2* yl = Y + kt.slot_comp_off[slot] + lane;
    V2 xr[16], xi[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) xr[n1].v = yl[n1 * kYStride];
    if (slot != 0) {
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) xi[n1].v = yl[n1 * kYStride + 32];
    } else {  // k2 = 0: Y is purely real
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) xi[n1] = mk(0.f, 0.f);
    }
    fft::cfft16<V2>(xr, xi);
    float2* pl = P + lane;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        // slot 0 writes bins 25 j twice (k1 and 16-k1 are conjugates): same thread, same value class
        const V2 pw = vfma(xr[i], xr[i], vmul(xi[i], xi[i]));
        pl[kt.slot_pbin_off[slot][i]] = pw.v;
    }
}

// lane = frame pair P: frames (P, P+16) for P < 16, (P+16, P+32) for P >= 16
__device__ __forceinline__ int pair_frame_a(int lane) { return lane < 16 ? lane : lane + 16; }

// ---- mel + log10 for one warp's run of filters, 32 frame pairs --------------------------------------
// Emit(m, log10 pair) consumes the unclamped log-mel values of filter m (called in filter order).
template <class Emit>
__device__ __forceinline__ float2 mel_stage(const KernelTables& kt, const float2* P, int warp, int lane, Emit emit) {
    constexpr float kLog10_2 = 0.30102999566398120f;
    float2 mx = make_float2(-INFINITY, -INFINITY);
    const int ma = kt.warp_m0[warp], mb = kt.warp_m0[warp + 1];
    const int kb = kt.warp_kb[warp], ke = kt.warp_ke[warp];
    int cur = kt.warp_cur[warp];
    float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
    auto finish = [&](float2 acc) {
        if (cur >= ma && cur < mb) {
            // log10(max(acc, 1e-10)) == max(log10(acc), -10): exact -10 for silence (TF-FE:155)
            float2 lg = __fmul2_rn(make_float2(lg2_approx(acc.x), lg2_approx(acc.y)), make_float2(kLog10_2, kLog10_2));
            lg.x = fmaxf(lg.x, -10.0f);
            lg.y = fmaxf(lg.y, -10.0f);
            emit(cur, lg);
            mx.x = fmaxf(mx.x, lg.x);
            mx.y = fmaxf(mx.y, lg.y);
        }
        ++cur;
    };
    // Two virtual bins past the end (shift 1, no data) flush both accumulators, so `finish` is
    // instantiated once (code size) and the loop has a single back edge.
    const float2* p = P + lane;
    const int k_stop = ma < mb ? ke + 2 : kb;
#pragma unroll 1
    for (int k = kb; k < k_stop; ++k) {
        const bool real = k < ke;
        int sh = real ? (k == kb ? 0 : kt.shift[k]) : 1;
#pragma unroll 1
        while (sh > 0) {
            finish(acc0);
            acc0 = acc1;
            acc1 = make_float2(0.f, 0.f);
            --sh;
        }
        if (real) {
            const float4 w = kt.w4[k];
            const float2 pv = p[k * 32];
            acc0 = __ffma2_rn(pv, make_float2(w.x, w.y), acc0);
            acc1 = __ffma2_rn(pv, make_float2(w.z, w.w), acc1);
        }
    }
    return mx;
}

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.0f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// ================================================================================================
// Step-1 kernel: persistent CTAs over (clip, tile) work items; unclamped log10 mel to `out`, per-clip
// max through an ordered-int atomic; a second small kernel applies max-8 and (x+4)/4.
// ================================================================================================
__global__ void __launch_bounds__(kThreads, 1)
logmel_tiles_kernel(const ClipArgs a, const __grid_constant__ MelParams mp, const float* __restrict__ win_lane,
                    int total_items) {
    extern __shared__ __align__(128) unsigned char smem[];
    float* raw = reinterpret_cast<float*>(smem);
    float2* Y = reinterpret_cast<float2*>(smem + kSmemRaw);
    float2* P = reinterpret_cast<float2*>(smem + kSmemRaw + kSmemY);
    const uint32_t bar = smem_u32(smem + kSmemRaw + kSmemY + kSmemP);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // tells the compiler it is warp-uniform
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // per-lane stage-1 constants
    const int n1 = lane & 15;
    float wv[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) wv[t] = win_lane[n1 * 25 + t];
    const int tw = n1 == 0 ? 25 : (kNfft - 25 * n1 + 15) / 16;
    __syncthreads();

    int item = blockIdx.x;
    if (item >= total_items) return;
    TileCtx cur = tile_ctx(a, item / kTilesPerClip, item % kTilesPerClip);
    if (tid == 0 && cur.active) tile_issue_tma(a, cur, raw, bar);
    uint32_t parity = 0;
    bool have_cur = true, have_prev = false;
    TileCtx prev = cur;

    // One extra trip at the end runs only the mel stage of the last tile, so every stage appears
    // exactly once in the instruction stream (code size = instruction-cache footprint).
    while (true) {
        // ---- phase X: stage 1 of `cur` (needs raw) + mel of `prev` (needs P) ------------------------
        if (have_cur && cur.active) {
            mbar_wait(bar, parity);
            parity ^= 1;
            if (cur.needs_fix || a.pcm_format == WLM_PCM_I16) tile_fixup(a, cur, raw);
            stage1(raw, Y, wv, tw, warp, lane);
        }
        if (have_prev) {
            const TileCtx& c = prev;
            const int fa = c.f0 + pair_frame_a(lane), fb = fa + 16;
            float* ob = a.out + static_cast<int64_t>(c.b) * a.n_mels * kNFrames + fa;
            float2 mx;
            if (c.active) {
                mx = mel_stage(mp, P, warp, lane, [&](int m, float2 lg) {
                    float* row = ob + m * kNFrames;
                    if (fa < kNFrames) row[0] = lg.x;
                    if (fb < kNFrames) row[16] = lg.y;
                });
                if (fa >= kNFrames) mx.x = -INFINITY;
                if (fb >= kNFrames) mx.y = -INFINITY;
            } else {
                // every frame of the tile is digital silence: log10(1e-10) = -10 exactly
                for (int m = mp.warp_m0[warp]; m < mp.warp_m0[warp + 1]; ++m) {
                    float* row = ob + m * kNFrames;
                    if (fa < kNFrames) row[0] = -10.0f;
                    if (fb < kNFrames) row[16] = -10.0f;
                }
                mx = make_float2(-10.0f, -10.0f);
            }
            float v = fmaxf(mx.x, mx.y);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
            if (lane == 0 && mp.warp_m0[warp] < mp.warp_m0[warp + 1]) atomic_max_float(a.gmax + c.b, v);
        }
        if (!have_cur) break;
        __syncthreads();
        // raw is free: prefetch the next work item
        const int next_item = item + gridDim.x;
        TileCtx nxt = cur;
        const bool have_next = next_item < total_items;
        if (have_next) {
            nxt = tile_ctx(a, next_item / kTilesPerClip, next_item % kTilesPerClip);
            if (tid == 0 && nxt.active) tile_issue_tma(a, nxt, raw, bar);
        }
        // ---- phase Y: stage 2 of `cur` -----------------------------------------------------------
        if (cur.active && warp < fft::kNumSlots) stage2(mp, Y, P, warp, lane);
        __syncthreads();
        prev = cur;
        have_prev = true;
        have_cur = have_next;
        cur = nxt;
        item = next_item;
    }
}

__global__ void init_gmax_kernel(float* gmax, int B) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) gmax[i] = __int_as_float(0xff800000);
}

__global__ void clamp_scale_kernel(float* __restrict__ out, const float* __restrict__ gmax, int n_mels, int B) {
    const int64_t per_clip = static_cast<int64_t>(n_mels) * kNFrames;
    const int64_t total = per_clip * B;
    for (int64_t i = (blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x) * 4; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x * 4) {
        const float floor_v = gmax[i / per_clip] - 8.0f;
        float4 v = *reinterpret_cast<float4*>(out + i);
        v.x = (fmaxf(v.x, floor_v) + 4.0f) * 0.25f;
        v.y = (fmaxf(v.y, floor_v) + 4.0f) * 0.25f;
        v.z = (fmaxf(v.z, floor_v) + 4.0f) * 0.25f;
        v.w = (fmaxf(v.w, floor_v) + 4.0f) * 0.25f;
        *reinterpret_cast<float4*>(out + i) = v;
    }
}

// ---- host side -----------------------------------------------------------------------------------
inline cudaError_t configure(int /*n_mels*/, int* max_clusters) {
    *max_clusters = 0;
    return cudaFuncSetAttribute(logmel_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
}

inline cudaError_t launch(const ClipArgs& a, const Tables* d_tables, const Tables& h_tables, int sm_count,
                          int /*max_clusters*/, cudaStream_t st, int* n_launches) {
    const int total = a.B * kTilesPerClip;
    init_gmax_kernel<<<(a.B + 255) / 256, 256, 0, st>>>(a.gmax, a.B);
    const int grid = total < sm_count ? total : sm_count;
    logmel_tiles_kernel<<<grid, kThreads, kSmemBytes, st>>>(a, h_tables.mel, d_tables->win_lane, total);
    const int64_t total4 = static_cast<int64_t>(a.B) * a.n_mels * kNFrames / 4;
    int64_t blocks = (total4 + 255) / 256;
    if (blocks > static_cast<int64_t>(sm_count) * 16) blocks = static_cast<int64_t>(sm_count) * 16;
    clamp_scale_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(a.out, a.gmax, a.n_mels, a.B);
    *n_launches = 3;
    return cudaGetLastError();
}

}  // namespace fused
}  // namespace wlm
