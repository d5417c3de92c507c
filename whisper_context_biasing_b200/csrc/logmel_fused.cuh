// Fused log-mel kernel for sm_100a.
//
// A thread-block CLUSTER of 6 CTAs owns one clip (3000 frames = 47 tiles of 64 frames, tile t goes to
// CTA t mod 6).  Each CTA (16 warps) walks its tiles through
//
//   TMA (cp.async.bulk, mbarrier)  raw PCM  ->  shared memory, two regions 16 banks apart
//   stage 1  per warp: 4 frames x 16 sub-transforms; lane = (n1, frame group); 25-point real DFT of
//            the Hann-windowed samples n = (25 n1 + 16 n2) mod 400, packed f32x2 over two frames
//   stage 2  per warp: one k2 slot for 32 frame pairs; 16-point complex DFT over n1, |X|^2
//   mel      per warp: a run of filters for 32 frame pairs; sparse gather; mel POWER retained in
//            TENSOR MEMORY (tcgen05.st), running max in registers
//
// with the prime-factor index maps of fft_pfa.cuh (no twiddles between the stages).  All arithmetic on
// the data path is FADD2 / FMUL2 / FFMA2 on (frame a, frame b) pairs with immediate constants.
// When the clip is done the 6 CTAs exchange their maxima through distributed shared memory
// (one cluster barrier), and a single pass reads the retained mel power back (tcgen05.ld) and writes
// (max(log10(max(p,1e-10)), gmax - 8) + 4) / 4 -- the features touch HBM exactly once.
//
// Shared memory (bytes):  raw 42,880 | Y 102,528 | P 51,456 | mbarrier + scratch 256
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
#include "mel_structure.inc"

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
constexpr int kSmemBytes = kSmemRaw + kSmemY + kSmemP + 256;

constexpr int kCluster = 6;                     // CTAs per clip: 22 co-resident clusters = 132 of 148 SMs (size 8: 15 = 120)
constexpr int kMaxTilesPerCta = (kTilesPerClip + kCluster - 1) / kCluster;   // 8
constexpr int kMaxFiltersPerWarp = 8;           // 16 warps x 8 >= 128 mels
constexpr int kMaxGroupBins = 16;               // bins between two adjacent filter centres
constexpr int kTmemColsPerTile = 2 * kMaxFiltersPerWarp;                      // 16 (two frames per filter)
constexpr int kTmemColsPerWarp = kMaxTilesPerCta * kTmemColsPerTile;          // 128; 4 warps per lane quarter = 512 columns
static_assert(4 * kTmemColsPerWarp <= 512, "the retained mel power must fit the 512 TMEM columns");
static_assert(kCluster <= 8, "max reduction over the cluster uses 8 lanes");

// Everything the kernel reads with warp-uniform indices, passed by value (constant bank).
//
// Mel projection: FFT bin k adds w_lo[k] P[k] to filter lo[k] and w_hi[k] P[k] to filter lo[k]+1
// (host-built from the caller's dense table, weights bit-identical).  Bins with the same lo[k] form
// a "group" (the bins between two adjacent filter centres).  Warp w owns filters
// [m0, m0+nf) and walks groups g = 0..nf, group g = bins [gb[g], gb[g+1]) with lo = m0-1+g:
//     filter m0+q  =  sum_{k in group q} w_hi[k] P[k]  +  sum_{k in group q+1} w_lo[k] P[k]
struct KernelTables {
    float4 w4[kNFreq + 3];                          // (w_lo, w_lo, w_hi, w_hi) per bin
    int16_t gb[kWarps][kMaxFiltersPerWarp + 2];     // group boundaries (bin indices)
    int16_t m0[kWarps];                             // first filter of warp w
    int16_t nf[kWarps];                             // number of filters of warp w (<= 8)
    // stage 2: per k2 slot, float2 offsets into Y (component) and into P (output bin) per FFT16 output
    int32_t slot_comp_off[16];                      // comp * 32
    int32_t slot_pbin_off[13][16];                  // BYTE offset of output_bin(k1, k2) in P, indexed by cfft16 array position
    int16_t n_mels;
};
using MelParams = KernelTables;

// Host-visible tables
struct Tables {
    float win_lane[16 * 25];     // Hann window at n = (25 n1 + 16 t) mod 400
    KernelTables mel;
};

// returns 0, or -1 if the table does not fit the fused path (a group longer than 16 bins)
// *variant receives 80 / 128 when the structure equals the baked Whisper bank (unrolled kernel), else 0
inline int build_tables(const MelSparse& sp, int n_mels, Tables* t, int* variant) {
    *variant = 0;
    for (int n1 = 0; n1 < 16; ++n1)
        for (int tt = 0; tt < 25; ++tt) {
            const int n = (25 * n1 + 16 * tt) % 400;
            t->win_lane[n1 * 25 + tt] = (float)(0.5 - 0.5 * cos(2.0 * M_PI * n / 400.0));
        }
    KernelTables& mp = t->mel;
    memset(&mp, 0, sizeof(mp));
    mp.n_mels = (int16_t)n_mels;
    for (int k = 0; k < kNFreq; ++k) mp.w4[k] = make_float4(sp.w_lo[k], sp.w_lo[k], sp.w_hi[k], sp.w_hi[k]);
    // first bin of every group: gstart[v] = first k with lo[k] >= v - 1   (v = lo + 1 in 0..n_mels)
    int gstart[kMaxMels + 2];
    {
        int k = 0;
        for (int v = 0; v <= n_mels + 1; ++v) {
            while (k < kNFreq && sp.lo[k] + 1 < v) ++k;
            gstart[v] = k;
        }
    }
    for (int v = 0; v <= n_mels; ++v)
        if (gstart[v + 1] - gstart[v] > kMaxGroupBins) return -1;
    const short* baked_g = n_mels == 80 ? kMelGstartHost80 : (n_mels == 128 ? kMelGstartHost128 : nullptr);
    const short* baked_m0 = n_mels == 80 ? kMelM0Host80 : kMelM0Host128;
    const short* baked_nf = n_mels == 80 ? kMelNfHost80 : kMelNfHost128;
    bool same = baked_g != nullptr;
    for (int v = 0; same && v <= n_mels + 1; ++v) same = gstart[v] == baked_g[v];
    if (same) *variant = n_mels;
    // contiguous filter runs per warp, balanced on issue slots: ~4 per bin of the two groups a filter
    // touches (shared with its neighbour) + ~12 per filter
    auto cost = [&](int m) { return 12.0 + 2.0 * (gstart[m + 2] - gstart[m]); };
    double total = 0;
    for (int m = 0; m < n_mels; ++m) total += cost(m);
    int m = 0;
    double acc = 0;
    for (int w = 0; w < kWarps; ++w) {
        const double target = total * (w + 1) / kWarps;
        int cnt = 0;
        mp.m0[w] = (int16_t)m;
        if (same) {   // partition baked into the unrolled kernel
            mp.m0[w] = baked_m0[w];
            cnt = baked_nf[w];
            m = mp.m0[w] + cnt;
        }
        while (!same && m < n_mels && cnt < kMaxFiltersPerWarp) {
            const bool must_take = n_mels - m > (kWarps - 1 - w) * kMaxFiltersPerWarp;   // the rest could not hold them
            if (!must_take && cnt > 0 && acc + 0.5 * cost(m) > target) break;
            acc += cost(m);
            ++m;
            ++cnt;
        }
        mp.nf[w] = (int16_t)cnt;
        // groups g = 0..cnt: lo = m0 - 1 + g  ->  v = m0 + g
        for (int g = 0; g <= kMaxFiltersPerWarp + 1; ++g) {
            const int v = mp.m0[w] + (g <= cnt + 1 ? g : cnt + 1);
            mp.gb[w][g] = (int16_t)gstart[v <= n_mels + 1 ? v : n_mels + 1];
        }
    }
    if (m < n_mels) return -1;
    for (int s2 = 0; s2 < fft::kNumSlots; ++s2) {
        mp.slot_comp_off[s2] = fft::kSlotComp[s2] * 32;
        for (int k1 = 0; k1 < 16; ++k1)
            mp.slot_pbin_off[s2][fft::fft16_slot_of_k1(k1)] = fft::output_bin(k1, fft::kSlotK2[s2]) * 32 * 8;
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
#ifdef WLM_OPT_HINT
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, 0x989680;\n"   /* suspend-time hint: sleep, do not spin */
#else
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
#endif
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
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

// ---- tensor memory (tcgen05) -------------------------------------------------------------------
// The retained mel power never needs a tensor core; TMEM is used as 256 KB of per-SM scratch so the
// clip's features can wait on-chip for the cluster-wide max.  Warp w may only touch TMEM lanes
// [32 (w & 3), +32): lane i of the warp <-> TMEM lane 32 (w & 3) + i, i.e. one frame pair per lane.
__device__ __forceinline__ void tmem_alloc_512(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_dst) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_512(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st_x2(uint32_t taddr, float a, float b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void tmem_st_x16(uint32_t taddr, const float2 (&v)[8]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "f"(v[0].x), "f"(v[0].y), "f"(v[1].x), "f"(v[1].y), "f"(v[2].x), "f"(v[2].y), "f"(v[3].x), "f"(v[3].y),
          "f"(v[4].x), "f"(v[4].y), "f"(v[5].x), "f"(v[5].y), "f"(v[6].x), "f"(v[6].y), "f"(v[7].x), "f"(v[7].y) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld_x16(uint32_t taddr, float (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7]),
          "=f"(r[8]), "=f"(r[9]), "=f"(r[10]), "=f"(r[11]), "=f"(r[12]), "=f"(r[13]), "=f"(r[14]), "=f"(r[15])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- clip / tile bookkeeping (all values CTA-uniform) ---------------------------------------------
struct ClipCtx {
    int b, len, n_act;     // clip index, valid samples (<= 480000), tiles that contain any real sample
    int64_t base;          // element offset of the clip in the PCM buffer
};

__device__ __forceinline__ ClipCtx clip_ctx(const ClipArgs& a, int b) {
    ClipCtx c;
    c.b = b;
    c.base = a.offsets ? a.offsets[b] : static_cast<int64_t>(b) * a.row_stride;
    int len = a.lengths ? a.lengths[b] : a.dense_len;
    if (!a.offsets) len = static_cast<int>(min(static_cast<int64_t>(len), a.row_stride));
    c.len = max(0, min(len, kNSamples));
    // tile t starts at sample 10240 t - 200: active iff that is < len
    c.n_act = min(kTilesPerClip, (c.len + kNfft / 2 + kTile * kHop - 1) / (kTile * kHop));
    return c;
}
__device__ __forceinline__ int tile_s0(int tile) { return tile * (kTile * kHop) - kNfft / 2; }

// issued by one thread: both regions of the tile, valid sample range only
__device__ __forceinline__ void tile_issue_tma(const ClipArgs& a, const ClipCtx& c, int tile, float* raw, uint32_t bar) {
    const int s0 = tile_s0(tile);
    const int esz = a.pcm_format == WLM_PCM_I16 ? 2 : 4;
    const int gran = 16 / esz;
    const int len_up = min((c.len + gran - 1) / gran * gran, kNSamples);
    uint32_t total = 0;
    int lo[2], n[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        const int s_lo = s0 + r * kRegionStep;
        lo[r] = max(s_lo, 0);
        const int hi = min(s_lo + kRegion, len_up);
        n[r] = max(hi - lo[r], 0);
        total += static_cast<uint32_t>(n[r]) * esz;
    }
    mbar_expect_tx(bar, total);
#pragma unroll
    for (int r = 0; r < 2; ++r) {
        if (n[r] <= 0) continue;
        const int s_lo = s0 + r * kRegionStep;
        const char* src = static_cast<const char*>(a.pcm) + (c.base + lo[r]) * esz;
        uint32_t dst;
        if (esz == 4) dst = smem_u32(raw) + static_cast<uint32_t>(r * kRegion + (lo[r] - s_lo)) * 4u;
        else dst = smem_u32(raw) + static_cast<uint32_t>(kRegion) * 4u + static_cast<uint32_t>(r * kRegion + (lo[r] - s_lo)) * 2u;
        tma_bulk_g2s(dst, src, static_cast<uint32_t>(n[r]) * esz, bar);
    }
}

// int16 -> float32 expansion in place (staging sits in the byte range of region B) + reflect /
// zero-fill patching of every position outside [0, len).  Only edge tiles and int16 input pay.
__device__ __forceinline__ void tile_fixup(const ClipArgs& a, const ClipCtx& c, int tile, float* raw) {
    const int tid = threadIdx.x;
    const int s0 = tile_s0(tile);
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
    if (s0 < 0 || s0 + kTileSamples > c.len) {
        for (int idx = tid; idx < kRawFloats; idx += kThreads) {
            const int r = idx >= kRegion ? 1 : 0;
            const int s = s0 + r * kRegionStep + (idx - r * kRegion);
            if (s >= 0 && s < c.len) continue;
            // reflect of the zero-padded 480000 buffer (torch.stft center=True, TF-FE:149)
            const int sr = s < 0 ? -s : (s >= kNSamples ? 2 * (kNSamples - 1) - s : s);
            float v = 0.f;
            if (sr >= 0 && sr < c.len) {
                const int u = sr - s0;
                if (u >= 0 && u < kTileSamples) v = raw[u < kRegion ? u : kRegion + (u - kRegionStep)];
            }
            raw[idx] = v;
        }
        __syncthreads();
    }
}

// ---- stage 1 ----------------------------------------------------------------------------------
// warp w, lane (n1 = lane & 15, g = lane >> 4): the two ADJACENT frames 32 g + 2 w, 32 g + 2 w + 1 of the tile
// (pair index 16 g + w = lane of the later stages, whose outputs are then one float2 per lane).
// `loaded()` runs once the warp no longer needs the raw buffer, `before_store()` just before Y is written.
template <class Loaded, class BeforeStore>
__device__ __forceinline__ void stage1(const float* raw, float2* Y, const float (&wv)[25], int tw, int warp, int lane,
                                       Loaded loaded, BeforeStore before_store) {
    const int n1 = lane & 15, g = lane >> 4;
    const float* p0 = raw + g * kRegion + 2 * kHop * warp + 25 * n1;
    const float* p1 = p0 - kNfft;
    V2 y[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) {
        const float* p = (t >= tw) ? p1 : p0;
        const float xa = p[16 * t], xb = p[16 * t + kHop];
        y[t] = mk(xa * wv[t], xb * wv[t]);
    }
    loaded();      // (fence inside: every LDS above has been performed)
    V2 out[25];
    fft::rfft25<V2>(y, out);
    before_store();
    float2* yo = Y + n1 * kYStride + (warp + 16 * g);
#pragma unroll
    for (int c = 0; c < 25; ++c) yo[c * 32] = out[c].v;
}

// ---- stage 2 ----------------------------------------------------------------------------------
// warp = k2 slot (uniform), lane = frame pair.  One code path for all 13 slots: the slot only selects
// table offsets, so every warp runs the same instructions (a 13-way templated version thrashed the
// instruction cache: 28 % of issue stalls were "no instruction").
// `loaded()` runs once Y has been read, `before_store()` just before P is written.
template <class Loaded, class BeforeStore>
__device__ __forceinline__ void stage2(const KernelTables& kt, const float2* Y, float2* P, int slot, int lane,
                                       Loaded loaded, BeforeStore before_store) {
    const float2* yl = Y + kt.slot_comp_off[slot] + lane;
    V2 xr[16], xi[16];
#pragma unroll
    for (int n1 = 0; n1 < 16; ++n1) {
        xr[n1].v = yl[n1 * kYStride];
        xi[n1].v = yl[n1 * kYStride + 32];      // slot 0 (k2 = 0, purely real Y): fetches component 1, discarded below
    }
    loaded();
    if (slot == 0) {
#pragma unroll
        for (int n1 = 0; n1 < 16; ++n1) xi[n1] = mk(0.f, 0.f);
    }
    fft::cfft16<V2>(xr, xi);
    before_store();
    char* pl = reinterpret_cast<char*>(P + lane);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        // slot 0 writes bins 25 j twice (k1 and 16-k1 are conjugates): same thread, same value class
        const V2 pw = vfma(xr[i], xr[i], vmul(xi[i], xi[i]));
        *reinterpret_cast<float2*>(pl + kt.slot_pbin_off[slot][i]) = pw.v;
    }
}

// lane = frame pair of the tile: frames (2 lane, 2 lane + 1)
__device__ __forceinline__ int pair_frame_a(int lane) { return 2 * lane; }

// ---- mel stage ------------------------------------------------------------------------------------
// One warp, its run of <= 8 filters, 32 frame pairs.  Groups of bins between adjacent filter centres
// are walked once: every P value is loaded once and feeds the falling side of one filter and the
// rising side of the next (two independent FFMA2 chains).  The group loop is unrolled (static
// register indices for the 8 outputs); inside, a fall-through switch on the group length gives one
// straight-line copy of the 16 possible terms.  The POWER goes to tensor memory (16 columns:
// 8 filters x two frames); log10 is monotone, so the running max is kept on the power.
#define WLM_MEL_TERM(i)                                                            \
    case (i) + 1: {                                                                \
        const float4 w = ww[i];                                                    \
        const float2 pv = pp[(i) * 32];                                            \
        A = __ffma2_rn(pv, make_float2(w.z, w.w), A);                              \
        Bq = __ffma2_rn(pv, make_float2(w.x, w.y), Bq);                            \
    }

__device__ __forceinline__ float2 mel_stage(const KernelTables& kt, const float2* P, int warp, int lane, uint32_t tcol) {
    const int nf = kt.nf[warp];
    const float2* pl = P + lane;
    float2 out[kMaxFiltersPerWarp];
#pragma unroll
    for (int q = 0; q < kMaxFiltersPerWarp; ++q) out[q] = make_float2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g <= kMaxFiltersPerWarp; ++g) {
        if (g <= nf) {
            const int k0 = kt.gb[warp][g];
            const int n = kt.gb[warp][g + 1] - k0;
            const float2* pp = pl + k0 * 32;
            const float4* ww = kt.w4 + k0;
            float2 A = make_float2(0.f, 0.f), Bq = make_float2(0.f, 0.f);
            switch (n) {
                WLM_MEL_TERM(15) WLM_MEL_TERM(14) WLM_MEL_TERM(13) WLM_MEL_TERM(12)
                WLM_MEL_TERM(11) WLM_MEL_TERM(10) WLM_MEL_TERM(9) WLM_MEL_TERM(8)
                WLM_MEL_TERM(7) WLM_MEL_TERM(6) WLM_MEL_TERM(5) WLM_MEL_TERM(4)
                WLM_MEL_TERM(3) WLM_MEL_TERM(2) WLM_MEL_TERM(1) WLM_MEL_TERM(0)
                default: break;
            }
            if (g < kMaxFiltersPerWarp) out[g] = A;                      // rising side of filter m0+g
            if (g > 0) out[g - 1] = __fadd2_rn(out[g - 1], Bq);          // falling side of filter m0+g-1
        }
    }
    tmem_st_x16(tcol, out);
    float2 mx = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < kMaxFiltersPerWarp; ++q)
        if (q < nf) {
            mx.x = fmaxf(mx.x, out[q].x);
            mx.y = fmaxf(mx.y, out[q].y);
        }
    return mx;
}
#undef WLM_MEL_TERM

// Unrolled variant for the two Whisper banks: group boundaries and the warp's filter run are
// compile-time constants (mel_structure.inc), so the stage is straight-line code -- one LDS.64 and
// one 16-byte constant load per bin, FFMA2 straight into statically indexed accumulators.
template <int NMELS> struct MelFixed;
template <> struct MelFixed<80> {
    static __device__ __forceinline__ constexpr int gstart(int v) { return kMelGstart80[v]; }
    static __device__ __forceinline__ constexpr int m0(int w) { return kMelM0_80[w]; }
    static __device__ __forceinline__ constexpr int nf(int w) { return kMelNf_80[w]; }
};
template <> struct MelFixed<128> {
    static __device__ __forceinline__ constexpr int gstart(int v) { return kMelGstart128[v]; }
    static __device__ __forceinline__ constexpr int m0(int w) { return kMelM0_128[w]; }
    static __device__ __forceinline__ constexpr int nf(int w) { return kMelNf_128[w]; }
};

template <int NMELS, int W>
__device__ __forceinline__ float2 mel_fixed_warp(const KernelTables& kt, const float2* P, int lane, uint32_t tcol) {
    using S = MelFixed<NMELS>;
    constexpr int nf = S::nf(W), m0 = S::m0(W);
    const float2* pl = P + lane;
    float2 out[kMaxFiltersPerWarp];
#pragma unroll
    for (int q = 0; q < kMaxFiltersPerWarp; ++q) out[q] = make_float2(0.f, 0.f);
#pragma unroll
    for (int g = 0; g <= nf; ++g) {
#pragma unroll
        for (int k = S::gstart(m0 + g); k < S::gstart(m0 + g + 1); ++k) {
            const float4 w = kt.w4[k];
            const float2 pv = pl[k * 32];
            if (g < nf) out[g] = __ffma2_rn(pv, make_float2(w.z, w.w), out[g]);            // rising side of m0+g
            if (g > 0) out[g - 1] = __ffma2_rn(pv, make_float2(w.x, w.y), out[g - 1]);     // falling side of m0+g-1
        }
    }
    tmem_st_x16(tcol, out);
    float2 mx = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < nf; ++q) {
        mx.x = fmaxf(mx.x, out[q].x);
        mx.y = fmaxf(mx.y, out[q].y);
    }
    return mx;
}

template <int NMELS>
__device__ __forceinline__ float2 mel_fixed(const KernelTables& kt, const float2* P, int warp, int lane, uint32_t tcol) {
    switch (warp) {
        case 0: return mel_fixed_warp<NMELS, 0>(kt, P, lane, tcol);
        case 1: return mel_fixed_warp<NMELS, 1>(kt, P, lane, tcol);
        case 2: return mel_fixed_warp<NMELS, 2>(kt, P, lane, tcol);
        case 3: return mel_fixed_warp<NMELS, 3>(kt, P, lane, tcol);
        case 4: return mel_fixed_warp<NMELS, 4>(kt, P, lane, tcol);
        case 5: return mel_fixed_warp<NMELS, 5>(kt, P, lane, tcol);
        case 6: return mel_fixed_warp<NMELS, 6>(kt, P, lane, tcol);
        case 7: return mel_fixed_warp<NMELS, 7>(kt, P, lane, tcol);
        case 8: return mel_fixed_warp<NMELS, 8>(kt, P, lane, tcol);
        case 9: return mel_fixed_warp<NMELS, 9>(kt, P, lane, tcol);
        case 10: return mel_fixed_warp<NMELS, 10>(kt, P, lane, tcol);
        case 11: return mel_fixed_warp<NMELS, 11>(kt, P, lane, tcol);
        case 12: return mel_fixed_warp<NMELS, 12>(kt, P, lane, tcol);
        case 13: return mel_fixed_warp<NMELS, 13>(kt, P, lane, tcol);
        case 14: return mel_fixed_warp<NMELS, 14>(kt, P, lane, tcol);
        default: return mel_fixed_warp<NMELS, 15>(kt, P, lane, tcol);
    }
}

// log10(max(p, 1e-10)) == max(log10 p, -10): exactly -10 for silence (TF-FE:155); p = 0 -> -inf -> -10
__device__ __forceinline__ float log10_floor(float p) {
    constexpr float kLog10_2 = 0.30102999566398120f;
    return fmaxf(lg2_approx(p) * kLog10_2, -10.0f);
}

// ================================================================================================
// The kernel: persistent clusters of 8 CTAs, one clip per cluster at a time.
// ================================================================================================
// NMELS = 80 / 128: unrolled mel stage for the Whisper banks; NMELS = 0: table-driven mel stage.
template <int NMELS>
__global__ void __launch_bounds__(kThreads, 1)
logmel_cluster_kernel(const ClipArgs a, const __grid_constant__ KernelTables kt, const float* __restrict__ win_lane) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(128) unsigned char smem[];
    float* raw = reinterpret_cast<float*>(smem);
    float2* Y = reinterpret_cast<float2*>(smem + kSmemRaw);
    float2* P = reinterpret_cast<float2*>(smem + kSmemRaw + kSmemY);
    unsigned char* misc = smem + kSmemRaw + kSmemY + kSmemP;
    // mbarriers (8 B each).  No CTA-wide barrier separates the stages of a tile: every hand-over between
    // warps is one of these, so warps drift apart and FMA-bound, load-bound and idle phases overlap.
    const uint32_t bar_raw = smem_u32(misc);         // TMA landed the tile's PCM            (tx, 1 arrival)
    const uint32_t bar_yfull = smem_u32(misc + 8);   // all 16 warps stored stage-1 output    (16)
    const uint32_t bar_yfree = smem_u32(misc + 16);  // all 13 stage-2 warps have read Y      (13)
    const uint32_t bar_pfull = smem_u32(misc + 24);  // all 13 stage-2 warps stored the power (13)
    const uint32_t bar_pfree = smem_u32(misc + 32);  // all 16 warps finished the mel stage   (16)
    uint32_t* raw_readers = reinterpret_cast<uint32_t*>(misc + 40);   // warps done with the raw buffer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 48);     // TMEM base address
    float* warp_max = reinterpret_cast<float*>(misc + 64);            // [2][16] (clip parity)
    float* cta_max = reinterpret_cast<float*>(misc + 192);            // [2] (clip parity), read by the peers

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    const int cluster_id = blockIdx.x / kCluster;
    const int n_clusters = gridDim.x / kCluster;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
    if (tid == 0) {
        mbar_init(bar_raw, 1);
        mbar_init(bar_yfull, kWarps);
        mbar_init(bar_yfree, fft::kNumSlots);
        mbar_init(bar_pfull, fft::kNumSlots);
        mbar_init(bar_pfree, kWarps);
        *raw_readers = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc_512(smem_u32(tmem_slot));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // this warp's TMEM window: lane quarter (warp & 3), 96 columns at (warp >> 2) * 96
    const uint32_t twin = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) +
                          static_cast<uint32_t>((warp >> 2) * kTmemColsPerWarp);

    // per-lane stage-1 constants
    const int n1 = lane & 15;
    float wv[25];
#pragma unroll
    for (int t = 0; t < 25; ++t) wv[t] = win_lane[n1 * 25 + t];
    const int tw = n1 == 0 ? 25 : (kNfft - 25 * n1 + 15) / 16;

    // The CTA's work is a stream of steps, one per tile it owns (a clip in which it owns no active
    // tile still contributes one empty step so that it takes part in that clip's cluster barrier).
    // Program order of every warp in step i (tile t_i):
    //   A  stage 1 of t_i            wait raw | load | last warp re-arms TMA | FFT | wait Y free | store | arrive Y full
    //   F  output pass of the clip that ended one step ago   (cluster barrier WAIT, TMEM read-back, stores)
    //   B  mel stage of t_{i-1}      wait P full | ... | arrive P free
    //   D  if t_{i-1} ended a clip:  wait P free | CTA max -> cta_max | cluster barrier ARRIVE
    //   C  stage 2 of t_i (13 warps) wait Y full | load | arrive Y free | FFT | wait P free | store | arrive P full
    // Iteration state kept in plain scalars (a struct-per-step version spent ~100 instructions per warp and
    // step on copies): `c*` = the step whose tile is in stage 1 / stage 2, `p*` = the previous step (mel stage).
    auto my_tiles = [&](int n_act) { return n_act > rank ? (n_act - rank + kCluster - 1) / kCluster : 0; };
    int cb = cluster_id, cj = 0, cn_my = 0;          // clip, step inside the clip, tiles of mine in the clip
    bool cvalid = cb < a.B;
    ClipCtx cc;
    cc.b = cb; cc.len = 0; cc.n_act = 0; cc.base = 0;
    if (cvalid) {
        cc = clip_ctx(a, cb);
        cn_my = my_tiles(cc.n_act);
    }
    bool pvalid = false, phas = false, plast = false;
    int pb = 0, pj = 0, ptile = 0, pn_my = 0;
    // TMA target after tile (clip cb0, step j0): the next tile of the same clip, else the first tile of the
    // next clip in which this CTA owns one.  Executed by ONE lane (the last warp to finish reading raw).
    auto issue_next_tile = [&](int cb0, const ClipCtx& c0, int n_my0, int j0) {
        if (j0 + 1 < n_my0) {
            tile_issue_tma(a, c0, rank + (j0 + 1) * kCluster, raw, bar_raw);
            return;
        }
        for (int nb = cb0 + n_clusters; nb < a.B; nb += n_clusters) {
            const ClipCtx c2 = clip_ctx(a, nb);
            if (c2.n_act > rank) {
                tile_issue_tma(a, c2, rank, raw, bar_raw);
                return;
            }
        }
    };

    // Phase bookkeeping: the n-th tile this CTA processes (n = 0, 1, ...) uses phase n of every barrier,
    // i.e. parity n & 1.  A wait for phase n is only issued by a warp that has already arrived on phase n
    // or whose own later work is needed to complete phase n+1, so the barrier is never more than one
    // phase ahead of a waiter.
    int fin_parity = 0;                      // parity of the clip whose max is exchanged next
    int tnum = 0, prev_tnum = 0;             // ordinal of cur's / prev's tile among the tiles of this CTA
    if (tid == 0 && cvalid) {
        if (cn_my > 0) tile_issue_tma(a, cc, rank, raw, bar_raw);
        else issue_next_tile(cb, cc, 0, 0);
    }
    float2 mx = make_float2(0.f, 0.f);       // running max of the mel power of the clip in flight (>= 0)
    bool pend = false;                       // an output pass is owed (cluster barrier arrived, not yet waited)
    int pend_b = 0, pend_n_my = 0;

    while (cvalid || pvalid || pend) {
        const bool do_tile = cvalid && cj < cn_my;
        const int ctile = rank + cj * kCluster;
        // ---- A: stage 1 ----------------------------------------------------------------------------
        if (do_tile) {
            mbar_wait(bar_raw, tnum & 1);
            tile_fixup(a, cc, ctile, raw);
            stage1(raw, Y, wv, tw, warp, lane,
                   [&]() {   // this warp is done with raw: the last of the 16 re-arms the TMA for the next tile
                       __syncwarp();
                       if (lane == 0) {
                           __threadfence_block();
                           const uint32_t old = atomicAdd(raw_readers, 1u);
                           if (old == kWarps - 1) {
                               *raw_readers = 0;
                               __threadfence_block();
                               asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                               issue_next_tile(cb, cc, cn_my, cj);
                           }
                       }
                   },
                   [&]() {   // stage 2 of the previous tile must have read Y
                       if (tnum > 0) mbar_wait(bar_yfree, (tnum - 1) & 1);
                   });
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_yfull);
        }
        // ---- F: output pass of the clip that ended one step ago -----------------------------------------
        if (pend) {
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
            float pmax = 0.f;
            if (lane < kCluster) pmax = *cluster.map_shared_rank(cta_max + fin_parity, lane);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
            pmax = __shfl_sync(0xffffffffu, pmax, 0);
            fin_parity ^= 1;
            const float gmax = log10_floor(pmax);                 // TF-FE:157
            const float floor_v = fmaxf(gmax - 8.0f, -10.0f);     // TF-FE:158 (log-mel is never below -10)
            if (rank == 0 && tid == 0 && a.gmax) a.gmax[pend_b] = gmax;

            // single pass: TMEM -> (max(log10, gmax-8)+4)/4 -> HBM
            tmem_wait_st();
            constexpr float kLog10_2 = 0.30102999566398120f;
            const int nf = kt.nf[warp];
            float* ob = a.out + (static_cast<int64_t>(pend_b) * a.n_mels + kt.m0[warp]) * kNFrames + pair_frame_a(lane);
            for (int j = 0; j < pend_n_my; ++j) {
                float r[16];
                tmem_ld_x16(twin + j * kTmemColsPerTile, r);
                const int f0 = (rank + j * kCluster) * kTile;
                const int fa = f0 + pair_frame_a(lane);
                float* of = ob + f0;
                const bool va = fa < kNFrames;      // 3000 is even: both frames of the pair or none
#define WLM_OUT_ROW(q)                                                                                         \
    case (q) + 1: {                                                                                            \
        float2 lg = __fmul2_rn(make_float2(lg2_approx(r[2 * (q)]), lg2_approx(r[2 * (q) + 1])),                \
                               make_float2(kLog10_2, kLog10_2));                                               \
        lg.x = fmaxf(lg.x, floor_v);                                                                           \
        lg.y = fmaxf(lg.y, floor_v);                                                                           \
        lg = __ffma2_rn(lg, make_float2(0.25f, 0.25f), make_float2(1.0f, 1.0f)); /* (x+4)/4, TF-FE:161 */       \
        if (va) *reinterpret_cast<float2*>(of + (q) * kNFrames) = lg;                                          \
    }
                switch (nf) {   // fall-through: exactly nf rows, static register indices
                    WLM_OUT_ROW(7) WLM_OUT_ROW(6) WLM_OUT_ROW(5) WLM_OUT_ROW(4)
                    WLM_OUT_ROW(3) WLM_OUT_ROW(2) WLM_OUT_ROW(1) WLM_OUT_ROW(0)
                    default: break;
                }
#undef WLM_OUT_ROW
            }
            // tiles of mine that hold no real sample: log-mel is exactly -10 everywhere
            const float silent = (floor_v + 4.0f) * 0.25f;
            for (int tile = rank + pend_n_my * kCluster; tile < kTilesPerClip; tile += kCluster) {
                const int fa = tile * kTile + pair_frame_a(lane);
                float* of = ob + tile * kTile;
                if (fa < kNFrames)
                    for (int q = 0; q < nf; ++q) *reinterpret_cast<float2*>(of + q * kNFrames) = make_float2(silent, silent);
            }
            pend = false;
        }
        // ---- B: mel stage of the previous tile -------------------------------------------------------------
        const bool clip_ends = pvalid && plast;
        const bool mel_tile = pvalid && phas;
        const int cpar = fin_parity;             // F has run: this is the parity of the clip ending now
        if (mel_tile) {
            mbar_wait(bar_pfull, prev_tnum & 1);
            const uint32_t tcol = twin + pj * kTmemColsPerTile;
            const float2 m2 = NMELS == 0 ? mel_stage(kt, P, warp, lane, tcol) : mel_fixed<NMELS == 0 ? 80 : NMELS>(kt, P, warp, lane, tcol);
            if (ptile * kTile + pair_frame_a(lane) < kNFrames) {     // frames past 3000 do not exist
                mx.x = fmaxf(mx.x, m2.x);
                mx.y = fmaxf(mx.y, m2.y);
            }
            if (clip_ends) {
                float v = fmaxf(mx.x, mx.y);
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
                if (lane == 0) warp_max[cpar * kWarps + warp] = v;
                mx = make_float2(0.f, 0.f);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_pfree);   // phase prev_tnum
        }
        // ---- D: the clip ended: CTA max -> cta_max (every warp writes the same value), cluster ARRIVE ---------
        // Done before stage 2 so that the peers get a whole step of slack before anyone WAITs (F, next step).
        if (clip_ends) {
            if (mel_tile) mbar_wait(bar_pfree, prev_tnum & 1);   // every warp's warp_max is visible
            float c = (mel_tile && lane < kWarps) ? warp_max[cpar * kWarps + lane] : 0.f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, o));
            if (lane == 0) cta_max[cpar] = c;
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            pend = true;
            pend_b = pb;
            pend_n_my = pn_my;
        }
        // ---- C: stage 2 ----------------------------------------------------------------------------------
        // 13 slots on 16 warps: one scheduler gets 4 tasks, the others 3.  Rotating the assignment by one
        // warp per tile moves the extra task around, so over four tiles every scheduler issues the same work.
#ifdef WLM_OPT_ROT
        const int slot = (warp + tnum) & 15;
#else
        const int slot = warp;
#endif
        if (do_tile && slot < fft::kNumSlots) {
            mbar_wait(bar_yfull, tnum & 1);
            stage2(kt, Y, P, slot, lane,
                   [&]() {
                       __syncwarp();
                       if (lane == 0) mbar_arrive(bar_yfree);   // phase tnum
                   },
                   [&]() {   // the mel stage of the previous tile must have read P (all 16 warps)
                       if (tnum > 0) mbar_wait(bar_pfree, (tnum - 1) & 1);
                   });
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_pfull);   // phase tnum
        }
        // this step becomes the previous one; advance to the next step of the stream
        const int steps = cn_my > 0 ? cn_my : 1;
        pvalid = cvalid; phas = do_tile; plast = cvalid && cj + 1 >= steps;
        pb = cb; pj = cj; ptile = ctile; pn_my = cn_my;
        prev_tnum = tnum;
        if (do_tile) ++tnum;
        if (cvalid) {
            if (cj + 1 < steps) {
                ++cj;
            } else {
                cb += n_clusters;
                cj = 0;
                cn_my = 0;
                cvalid = cb < a.B;
                if (cvalid) {
                    cc = clip_ctx(a, cb);
                    cn_my = my_tiles(cc.n_act);
                }
            }
        }
    }
    // all TMEM reads are complete (tcgen05.wait::ld inside tmem_ld_x16); release the allocation
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();   // also keeps every CTA's shared memory alive until its peers have read cta_max
    if (warp == 0) tmem_dealloc_512(tmem_base);
}

// ---- host side -----------------------------------------------------------------------------------
// variant: 80 / 128 when the table's structure equals the baked one (and the partition was taken from it)
typedef void (*KernelFn)(const ClipArgs, const KernelTables, const float*);
inline KernelFn kernel_for(int variant) {
    if (variant == 80) return logmel_cluster_kernel<80>;
    if (variant == 128) return logmel_cluster_kernel<128>;
    return logmel_cluster_kernel<0>;
}

inline void fill_launch_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, int n_clusters, cudaStream_t st) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->gridDim = dim3(kCluster * n_clusters);
    cfg->blockDim = dim3(kThreads);
    cfg->dynamicSmemBytes = kSmemBytes;
    cfg->stream = st;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kCluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg->attrs = at;
    cfg->numAttrs = 1;
}

inline cudaError_t configure(int variant, int* max_clusters) {
    KernelFn fn = kernel_for(variant);
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    fill_launch_config(&cfg, at, 148, nullptr);
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, fn, &cfg);
    if (e != cudaSuccess) return e;
    if (n < 1) return cudaErrorLaunchOutOfResources;
    *max_clusters = n;
    return cudaSuccess;
}

inline cudaError_t launch(const ClipArgs& a, const Tables* d_tables, const Tables& h_tables, int variant,
                          int max_clusters, cudaStream_t st, int* n_launches) {
    const int n_clusters = a.B < max_clusters ? a.B : max_clusters;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    fill_launch_config(&cfg, at, n_clusters, st);
    *n_launches = 1;
    return cudaLaunchKernelEx(&cfg, kernel_for(variant), a, h_tables.mel, static_cast<const float*>(d_tables->win_lane));
}

}  // namespace fused
}  // namespace wlm
