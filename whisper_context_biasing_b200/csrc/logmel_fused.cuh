// Fused log-mel kernel for sm_100a.
//
// A thread-block CLUSTER of 8 CTAs owns one clip (3000 frames = 47 tiles of 64 frames, tile t goes to
// CTA t mod 8).  Each CTA (16 warps) walks its tiles through
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
// When the clip is done the 8 CTAs exchange their maxima through distributed shared memory
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

constexpr int kCluster = 8;                     // CTAs per clip
constexpr int kMaxTilesPerCta = (kTilesPerClip + kCluster - 1) / kCluster;   // 6
constexpr int kMaxFiltersPerWarp = 8;           // 16 warps x 8 >= 128 mels
constexpr int kMaxBinsPerFilter = 16;
constexpr int kTmemColsPerTile = 2 * kMaxFiltersPerWarp;                      // 16 (two frames per filter)
constexpr int kTmemColsPerWarp = kMaxTilesPerCta * kTmemColsPerTile;          // 96; 4 warps per lane quarter = 384 <= 512
constexpr int kMaxEntries = 448;

// One (warp, filter) visit of the mel stage.
struct FilterRef {
    int16_t poff;   // k0 * 32: float2 offset of the filter's first bin in P
    int16_t n;      // number of bins (1..16)
    int16_t e0;     // first weight entry
    int16_t m;      // filter index (row of the output)
};

// Everything the kernel reads with warp-uniform indices, passed by value (constant bank).
struct KernelTables {
    float2 w2[kMaxEntries];                         // (w, w) per (filter, bin), bit-identical to the caller's table
    FilterRef fref[kWarps][kMaxFiltersPerWarp];     // filters of warp w
    int16_t nf[kWarps];                             // how many
    // stage 2: per k2 slot, float2 offsets into Y (component) and into P (output bin) per FFT16 output
    int16_t slot_comp_off[16];                      // comp * 32
    int16_t slot_pbin_off[13][16];                  // output_bin(k1, k2) * 32, indexed by cfft16 array position
    int16_t n_mels;
};
using MelParams = KernelTables;

// Host-visible tables
struct Tables {
    float win_lane[16 * 25];     // Hann window at n = (25 n1 + 16 t) mod 400
    KernelTables mel;
};

// returns 0, or -1 if the table does not fit the fused path (empty filter, support > 16 bins)
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
        if (sp.w_lo[k] != 0.0f && sp.lo[k] >= 0) { const int m = sp.lo[k]; if (klo[m] < 0) klo[m] = k; khi[m] = k; }
        if (sp.w_hi[k] != 0.0f && sp.lo[k] + 1 < n_mels) { const int m = sp.lo[k] + 1; if (klo[m] < 0) klo[m] = k; khi[m] = k; }
    }
    int e0[kMaxMels], e = 0;
    for (int m = 0; m < n_mels; ++m) {
        if (klo[m] < 0 || khi[m] - klo[m] + 1 > kMaxBinsPerFilter) return -1;
        e0[m] = e;
        for (int k = klo[m]; k <= khi[m]; ++k) {
            float w = 0.0f;
            if (sp.lo[k] == m) w = sp.w_lo[k];
            else if (sp.lo[k] + 1 == m) w = sp.w_hi[k];
            if (e >= kMaxEntries) return -1;
            mp.w2[e++] = make_float2(w, w);
        }
    }
    // contiguous filter runs per warp, balanced on (18 + 3 * bins) issue slots per filter, at most 8 each
    auto cost = [&](int m) { return 18.0 + 3.0 * (khi[m] - klo[m] + 1); };
    double total = 0;
    for (int m = 0; m < n_mels; ++m) total += cost(m);
    int m = 0;
    double acc = 0;
    for (int w = 0; w < kWarps; ++w) {
        const double target = total * (w + 1) / kWarps;
        int cnt = 0;
        while (m < n_mels && cnt < kMaxFiltersPerWarp) {
            const int left_after = n_mels - (m + 1);
            const bool must_take = n_mels - m > (kWarps - 1 - w) * kMaxFiltersPerWarp;   // the rest could not hold them
            if (!must_take && cnt > 0 && acc + 0.5 * cost(m) > target) break;
            (void)left_after;
            FilterRef& r = mp.fref[w][cnt];
            r.poff = (int16_t)(klo[m] * 32);
            r.n = (int16_t)(khi[m] - klo[m] + 1);
            r.e0 = (int16_t)e0[m];
            r.m = (int16_t)m;
            acc += cost(m);
            ++m;
            ++cnt;
        }
        mp.nf[w] = (int16_t)cnt;
    }
    if (m < n_mels) return -1;
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
// table offsets, so every warp runs the same instructions (a 13-way templated version thrashed the
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

// lane = frame pair: frames (lane, lane+16) for lane < 16, (lane+16, lane+32) for lane >= 16
__device__ __forceinline__ int pair_frame_a(int lane) { return lane < 16 ? lane : lane + 16; }

// ---- mel stage ------------------------------------------------------------------------------------
// One warp, its run of filters, 32 frame pairs.  Per filter: gather-sum over the filter's bins
// (fall-through switch = one straight-line copy of the 16 possible terms), keep the POWER in tensor
// memory (two columns: frame a, frame b), track the running max.  log10 is monotone, so the max of
// the log-mel is the log of this max.
__device__ __forceinline__ float2 mel_stage(const KernelTables& kt, const float2* P, int warp, int lane,
                                            uint32_t tcol, float2 mx) {
    const int nf = kt.nf[warp];
    const float2* pl = P + lane;
#pragma unroll 1
    for (int j = 0; j < nf; ++j) {
        const FilterRef fr = kt.fref[warp][j];
        const float2* pp = pl + fr.poff;
        const float2* ww = kt.w2 + fr.e0;
        float2 acc = make_float2(0.f, 0.f);
        switch (fr.n) {
            case 16: acc = __ffma2_rn(pp[15 * 32], ww[15], acc);
            case 15: acc = __ffma2_rn(pp[14 * 32], ww[14], acc);
            case 14: acc = __ffma2_rn(pp[13 * 32], ww[13], acc);
            case 13: acc = __ffma2_rn(pp[12 * 32], ww[12], acc);
            case 12: acc = __ffma2_rn(pp[11 * 32], ww[11], acc);
            case 11: acc = __ffma2_rn(pp[10 * 32], ww[10], acc);
            case 10: acc = __ffma2_rn(pp[9 * 32], ww[9], acc);
            case 9: acc = __ffma2_rn(pp[8 * 32], ww[8], acc);
            case 8: acc = __ffma2_rn(pp[7 * 32], ww[7], acc);
            case 7: acc = __ffma2_rn(pp[6 * 32], ww[6], acc);
            case 6: acc = __ffma2_rn(pp[5 * 32], ww[5], acc);
            case 5: acc = __ffma2_rn(pp[4 * 32], ww[4], acc);
            case 4: acc = __ffma2_rn(pp[3 * 32], ww[3], acc);
            case 3: acc = __ffma2_rn(pp[2 * 32], ww[2], acc);
            case 2: acc = __ffma2_rn(pp[1 * 32], ww[1], acc);
            case 1: acc = __ffma2_rn(pp[0], ww[0], acc);
            default: break;
        }
        tmem_st_x2(tcol + 2 * j, acc.x, acc.y);
        mx.x = fmaxf(mx.x, acc.x);
        mx.y = fmaxf(mx.y, acc.y);
    }
    return mx;
}

// log10(max(p, 1e-10)) == max(log10 p, -10): exactly -10 for silence (TF-FE:155); p = 0 -> -inf -> -10
__device__ __forceinline__ float log10_floor(float p) {
    constexpr float kLog10_2 = 0.30102999566398120f;
    return fmaxf(lg2_approx(p) * kLog10_2, -10.0f);
}

// ================================================================================================
// The kernel: persistent clusters of 8 CTAs, one clip per cluster at a time.
// ================================================================================================
__global__ void __launch_bounds__(kThreads, 1)
logmel_cluster_kernel(const ClipArgs a, const __grid_constant__ KernelTables kt, const float* __restrict__ win_lane) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(128) unsigned char smem[];
    float* raw = reinterpret_cast<float*>(smem);
    float2* Y = reinterpret_cast<float2*>(smem + kSmemRaw);
    float2* P = reinterpret_cast<float2*>(smem + kSmemRaw + kSmemY);
    unsigned char* misc = smem + kSmemRaw + kSmemY + kSmemP;
    const uint32_t bar = smem_u32(misc);                               // mbarrier (8 B)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 16);      // TMEM base address
    float* warp_max = reinterpret_cast<float*>(misc + 32);             // [16]
    float* cta_max = reinterpret_cast<float*>(misc + 96);              // [2] (clip parity), read by the peers

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    const int cluster_id = blockIdx.x / kCluster;
    const int n_clusters = gridDim.x / kCluster;

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler
    if (tid == 0) {
        mbar_init(bar, 1);
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

    uint32_t parity = 0;
    int clip_parity = 0;
    int b = cluster_id;
    ClipCtx cc;
    if (b < a.B) {
        cc = clip_ctx(a, b);
        if (tid == 0 && rank < cc.n_act) tile_issue_tma(a, cc, rank, raw, bar);
    }
    for (; b < a.B; b += n_clusters, clip_parity ^= 1) {
        const int n_my = cc.n_act > rank ? (cc.n_act - rank + kCluster - 1) / kCluster : 0;
        const int b_next = b + n_clusters;
        ClipCtx cn = cc;
        if (b_next < a.B) cn = clip_ctx(a, b_next);
        float2 mx = make_float2(0.f, 0.f);   // powers are >= 0

        // One extra trip at the end runs only the mel stage of the last tile.
        for (int j = 0; j <= n_my; ++j) {
            const int tile = rank + j * kCluster;
            // ---- phase X: stage 1 of tile j (needs raw) + mel of tile j-1 (needs P) --------------------
            if (j < n_my) {
                mbar_wait(bar, parity);
                parity ^= 1;
                tile_fixup(a, cc, tile, raw);
                stage1(raw, Y, wv, tw, warp, lane);
            }
            if (j > 0) {
                float2 m2 = mel_stage(kt, P, warp, lane, twin + (j - 1) * kTmemColsPerTile, make_float2(0.f, 0.f));
                // frames past 3000 (last tile only) do not exist
                const int fa = (tile - kCluster) * kTile + pair_frame_a(lane);
                if (fa < kNFrames) mx.x = fmaxf(mx.x, m2.x);
                if (fa + 16 < kNFrames) mx.y = fmaxf(mx.y, m2.y);
            }
            if (j == n_my) break;
            __syncthreads();
            // raw is free: prefetch the next tile of this clip, or the first tile of the next clip
            if (tid == 0) {
                if (j + 1 < n_my) tile_issue_tma(a, cc, tile + kCluster, raw, bar);
                else if (b_next < a.B && rank < cn.n_act) tile_issue_tma(a, cn, rank, raw, bar);
            }
            // ---- phase Y: stage 2 of tile j ------------------------------------------------------------
            if (warp < fft::kNumSlots) stage2(kt, Y, P, warp, lane);
            __syncthreads();
        }
        if (n_my == 0 && tid == 0 && b_next < a.B && rank < cn.n_act) tile_issue_tma(a, cn, rank, raw, bar);
        tmem_wait_st();

        // ---- per-clip max: warp -> CTA -> cluster (distributed shared memory) -----------------------------
        float v = fmaxf(mx.x, mx.y);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) warp_max[warp] = v;
        __syncthreads();
        if (warp == 0) {
            float c = lane < kWarps ? warp_max[lane] : 0.f;
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) c = fmaxf(c, __shfl_xor_sync(0xffffffffu, c, o));
            if (lane == 0) cta_max[clip_parity] = c;
        }
        cluster.sync();   // release/acquire: every CTA's cta_max[clip_parity] is visible cluster-wide
        float pmax = 0.f;
        if (lane < kCluster) pmax = *cluster.map_shared_rank(cta_max + clip_parity, lane);
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
        pmax = __shfl_sync(0xffffffffu, pmax, 0);
        const float gmax = log10_floor(pmax);                 // TF-FE:157
        const float floor_v = gmax - 8.0f;                    // TF-FE:158
        if (rank == 0 && tid == 0 && a.gmax) a.gmax[b] = gmax;

        // ---- single output pass: TMEM -> (max(log10, gmax-8)+4)/4 -> HBM -----------------------------------
        {
            float* ob = a.out + static_cast<int64_t>(b) * a.n_mels * kNFrames;
            const int nf = kt.nf[warp];
            const int pa = pair_frame_a(lane);
            for (int j = 0; j < n_my; ++j) {
                float r[16];
                tmem_ld_x16(twin + j * kTmemColsPerTile, r);
                const int fa = (rank + j * kCluster) * kTile + pa;
                float* of = ob + fa;
#pragma unroll
                for (int q = 0; q < kMaxFiltersPerWarp; ++q) {
                    if (q < nf) {
                        float* row = of + kt.fref[warp][q].m * kNFrames;
                        const float oa = (fmaxf(log10_floor(r[2 * q]), floor_v) + 4.0f) * 0.25f;      // TF-FE:161
                        const float obv = (fmaxf(log10_floor(r[2 * q + 1]), floor_v) + 4.0f) * 0.25f;
                        if (fa < kNFrames) row[0] = oa;
                        if (fa + 16 < kNFrames) row[16] = obv;
                    }
                }
            }
            // tiles of mine that hold no real sample: log-mel is exactly -10 everywhere
            const float silent = (fmaxf(-10.0f, floor_v) + 4.0f) * 0.25f;
            for (int tile = rank + n_my * kCluster; tile < kTilesPerClip; tile += kCluster) {
                const int fa = tile * kTile + pa;
                float* of = ob + fa;
                for (int q = 0; q < nf; ++q) {
                    float* row = of + kt.fref[warp][q].m * kNFrames;
                    if (fa < kNFrames) row[0] = silent;
                    if (fa + 16 < kNFrames) row[16] = silent;
                }
            }
        }
        cc = cn;
    }
    // all TMEM reads are complete (tcgen05.wait::ld inside tmem_ld_x16); release the allocation
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();   // also keeps every CTA's shared memory alive until its peers have read cta_max
    if (warp == 0) tmem_dealloc_512(tmem_base);
}

// ---- host side -----------------------------------------------------------------------------------
inline cudaError_t configure(int /*n_mels*/, int* max_clusters) {
    cudaError_t e = cudaFuncSetAttribute(logmel_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCluster * 148);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kCluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, logmel_cluster_kernel, &cfg);
    if (e != cudaSuccess) return e;
    if (n < 1) return cudaErrorLaunchOutOfResources;
    *max_clusters = n;
    return cudaSuccess;
}

inline cudaError_t launch(const ClipArgs& a, const Tables* d_tables, const Tables& h_tables, int /*sm_count*/,
                          int max_clusters, cudaStream_t st, int* n_launches) {
    const int n_clusters = a.B < max_clusters ? a.B : max_clusters;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kCluster * n_clusters);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kCluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    *n_launches = 1;
    return cudaLaunchKernelEx(&cfg, logmel_cluster_kernel, a, h_tables.mel, static_cast<const float*>(d_tables->win_lane));
}

}  // namespace fused
}  // namespace wlm
