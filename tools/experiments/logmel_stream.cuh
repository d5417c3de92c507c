// Streaming variant of the fused log-mel kernel (sm_100a): the same stages, index maps and arithmetic as
// logmel_fused.cuh (its stage1 / stage2 / mel_fixed / tile_* functions are used as they are, so the features are
// bit-identical), with the warps of a CTA DECOUPLED.
//
// Why.  tools/ubench_fe.cu runs stage 1 + stage 2 of the kernel with NO synchronisation between warps: 8 warps need
// 4,115 cycles per 64 frames per SM, 16 warps 3,982 -- the max of the FMA-pipe and shared-memory floors.  The dataflow
// kernel of logmel_fused.cuh needs 7,400: its 16 warps hand every half-tile over through single-buffered raw and P
// buffers, so the warps of a group move in lockstep through shared-memory-bound and FMA-bound phases and the two pipes
// are busy one after the other instead of at the same time.  Shared memory was full (Y alone is 6.8 KB per warp), so the
// buffers could not be doubled there.  Here
//
//   FRONT END  8 warps (not 16): stage 1 + stage 2 of a 32-frame half-tile back to back, warp-private Y.  Half the Y
//              space buys a DOUBLE-buffered raw tile (TMA two half-tiles ahead) and a TRIPLE-buffered P, so a front-end
//              warp only ever waits for something that happened two half-tiles ago.
//   BACK END   8 warps, one filter run (<= 16 filters) each, lane = frame: mel gather over P, mel power retained in tensor
//              memory, clip-end max exchange over distributed shared memory, read-back + log10 + clamp + scale + store.
//              The front end never sees a clip boundary.
//
// A thread-block cluster of 6 CTAs owns a clip: half-tile u of its 94 goes to CTA u mod 6, which retains up to 16 of them
// (16 slots x 16 columns per back-end warp, two warps per TMEM lane quarter = 512 columns).  Static clip assignment
// (cluster c of n takes clips first + c, + n, ...): this kernel serves the batches whose clips all cost the same; ragged
// batches and the SMs clusters cannot cover stay with logmel_fused.cuh (clip queue / flat twin).
//
// Shared memory (bytes):  raw 2 x 22,400 | Y 8 x 6,784 | P 3 x 26,752 | mbarriers + scratch 1024  = 180,352
#pragma once
#include "logmel_fused.cuh"   /* -I whisper_context_biasing_b200/csrc */

namespace wlm {
namespace stream {

using namespace fused;

constexpr int kFe = 8;                         // front-end warps (= fused::kGroupWarps: stage1 / stage2 index with it)
constexpr int kBe = 8;                         // back-end warps: one filter run each
constexpr int kSWarps = kFe + kBe;
constexpr int kSThreads = kSWarps * 32;
#ifndef WLM_RAW_BUFS
#define WLM_RAW_BUFS 2
#endif
#ifndef WLM_P_BUFS
#define WLM_P_BUFS 3
#endif
constexpr int kRawBufs = WLM_RAW_BUFS;        // half-tiles of PCM in flight (TMA prefetch distance)
constexpr int kPBufs = WLM_P_BUFS;
constexpr int kSlots = (kTilesPerClip + kCluster - 1) / kCluster;            // 16 half-tiles per CTA and clip
constexpr int kColsPerBe = kSlots * kTmemColsPerTile;                        // 256; two back-end warps per lane quarter
static_assert(kFe == kGroupWarps, "stage1 / stage2 address the warp inside a group of kGroupWarps");
static_assert(2 * kColsPerBe <= 512, "the retained mel power must fit the 512 TMEM columns");
constexpr int kSSmemY = kFe * kYWarpFloat2 * 8;
constexpr int kSSmemBytes = kRawBufs * kSmemRaw + kSSmemY + kPBufs * kSmemP + kSmemMisc;

// the sequence of half-tiles (and empty clips) of one CTA: clip b = first + worker, + n_workers, ...; inside a clip the
// half-tiles rank, rank + 6, ... that hold real samples.  A clip in which the CTA owns none still yields one (empty) step
// so that the back end takes part in that clip's max exchange.
struct Cursor {
    int b, j, n_my;
    bool valid;
    ClipCtx cc;
};
__device__ __forceinline__ void cursor_open(Cursor& s, const ClipArgs& a, int b, int rank) {
    s.b = b; s.j = 0; s.n_my = 0;
    s.valid = b < a.B;
    s.cc.b = b; s.cc.len = 0; s.cc.n_act = 0; s.cc.base = 0;
    if (s.valid) {
        s.cc = clip_ctx(a, b);
        s.n_my = s.cc.n_act > rank ? (s.cc.n_act - rank + kCluster - 1) / kCluster : 0;
    }
}
__device__ __forceinline__ void cursor_step(Cursor& s, const ClipArgs& a, int stride, int rank) {
    const int steps = s.n_my > 0 ? s.n_my : 1;
    if (s.j + 1 < steps) ++s.j;
    else cursor_open(s, a, s.b + stride, rank);
}
// front end: the next step that has a half-tile
__device__ __forceinline__ void cursor_next_tile(Cursor& s, const ClipArgs& a, int stride, int rank) {
    do { cursor_step(s, a, stride, rank); } while (s.valid && s.n_my == 0);
}

// Back-end waits sleep between polls: the back end has two half-tiles of slack, and a spinning warp issues an
// instruction every few cycles on the scheduler it shares with two front-end warps.
#ifndef WLM_BE_SLEEP_NS
#define WLM_BE_SLEEP_NS 200
#endif
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t bar, uint32_t parity) {
    while (true) {
        uint32_t done;
        asm volatile(
            "{\n.reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        __nanosleep(WLM_BE_SLEEP_NS);
    }
}

template <int NMELS, class OutT>
__global__ void __launch_bounds__(kSThreads, 1)
logmel_stream_kernel(const ClipArgs a, const __grid_constant__ KernelTables kt, const float* __restrict__ win_lane) {
    namespace cg = cooperative_groups;
    extern __shared__ __align__(128) unsigned char smem[];
    // this CTA is resident: once all of them are, the flat kernel (a programmatic dependent launch) may take the free SMs
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // warp-uniform for the compiler

    float* raw0 = reinterpret_cast<float*>(smem);
    float2* Y0 = reinterpret_cast<float2*>(smem + kRawBufs * kSmemRaw);
    float* P0 = reinterpret_cast<float*>(smem + kRawBufs * kSmemRaw + kSSmemY);
    unsigned char* misc = smem + kRawBufs * kSmemRaw + kSSmemY + kPBufs * kSmemP;
    // mbarriers (8 B each)
    const uint32_t bar_raw = smem_u32(misc);             // [kRawBufs] TMA landed a half-tile's PCM          (tx, 1 arrival)
    const uint32_t bar_pfull = smem_u32(misc + 32);      // [kPBufs] all front-end warps stored the power    (8)
    const uint32_t bar_pfree = smem_u32(misc + 64);      // [kPBufs] all back-end warps finished the mel stage (8)
    uint32_t* raw_readers = reinterpret_cast<uint32_t*>(misc + 96);   // [kRawBufs] front-end warps done with a raw buffer
    const uint32_t bar_max = smem_u32(misc + 256);                    // [2] clip-end max exchange, by clip parity
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc + 272);    // TMEM base address
    uint32_t* cta_cnt = reinterpret_cast<uint32_t*>(misc + 288);      // [2] back-end warps that have contributed
    int* cta_max = reinterpret_cast<int*>(misc + 320);                // [2] running max of the CTA (float bits, >= 0)
    float* clip_max = reinterpret_cast<float*>(misc + 352);           // [2][kCluster] written by the peers

    cg::cluster_group cluster = cg::this_cluster();
    const int rank = static_cast<int>(cluster.block_rank());
    const int worker = static_cast<int>(blockIdx.x) / kCluster;
    const int stride = static_cast<int>(gridDim.x) / kCluster;
    const int first_clip = a.clip_first + worker;

    if (tid == 0) {
        for (int i = 0; i < kRawBufs; ++i) { mbar_init(bar_raw + 8 * i, 1); raw_readers[i] = 0; }
        for (int i = 0; i < kPBufs; ++i) { mbar_init(bar_pfull + 8 * i, kFe); mbar_init(bar_pfree + 8 * i, kBe); }
        mbar_init(bar_max, 1);          // one arrival (this CTA, with the byte count) + 6 x 4 bytes from the peers
        mbar_init(bar_max + 8, 1);
        cta_cnt[0] = cta_cnt[1] = 0;
        cta_max[0] = cta_max[1] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc_512(smem_u32(tmem_slot));
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();     // (also a CTA barrier) every peer's mbarriers exist before anyone can arrive on them remotely
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kFe) {
        // ============================================================================================
        // FRONT END.  Half-tile n of this CTA lives in raw buffer n & 1 and P buffer n % 3.
        //   wait raw | load | FFT-25 | the last warp re-arms the buffer with the half-tile kRawBufs ahead | store (own Y) |
        //   __syncwarp | load own Y | FFT-16, |X|^2 | wait P free (mel of half-tile n - 3) | store P | arrive P full
        // ============================================================================================
        const int wg = warp;
        float2* Y = Y0 + wg * kYWarpFloat2;
        // the imaginary part of slot 0 in the warp's Y: zeros, written once (stage 1 never stores there)
        Y[(lane >> 1) * kYN1 + kYLanes + (lane & 1)] = make_float2(0.f, 0.f);
        __syncwarp();
        float wv[25];                                    // Hann window at this lane's 25 sample positions
#pragma unroll
        for (int t = 0; t < 25; ++t) wv[t] = win_lane[(lane & 15) * 25 + t];

        Cursor cur, nxt;                                 // the half-tile in hand, and the one to prefetch (kRawBufs ahead)
        cursor_open(cur, a, first_clip, rank);
        if (cur.valid && cur.n_my == 0) cursor_next_tile(cur, a, stride, rank);
        nxt = cur;
#pragma unroll 1
        for (int i = 0; i < kRawBufs; ++i) {             // the first kRawBufs half-tiles: one lane issues them all
            if (tid == 0 && nxt.valid) tile_issue_tma(a, nxt.cc, rank + nxt.j * kCluster, raw0 + i * kRawFloats, bar_raw + 8 * i);
            if (nxt.valid) cursor_next_tile(nxt, a, stride, rank);
        }
        int rb = 0, rround = 0;                          // raw buffer / P buffer of the half-tile in hand and how often
        int pb = 0, pround = 0;                          // each has been used before (barrier phase)
        while (cur.valid) {
            float* raw = raw0 + rb * kRawFloats;
            float* P = P0 + pb * kPFloats;
            const int tile = rank + cur.j * kCluster;
#if defined(WLM_KO_RAW2)      /* knock-out: real PCM in the buffers (first round), never re-armed, never waited for again */
            if (rround == 0) mbar_wait(bar_raw + 8 * rb, 0);
#ifdef WLM_X_FIXUP
            tile_fixup(a, cur.cc, tile, raw, 0, tid);
#endif
#elif !defined(WLM_KO_RAW)
            mbar_wait(bar_raw + 8 * rb, rround & 1);
            tile_fixup(a, cur.cc, tile, raw, 0, tid);
#endif
            // This warp is done with the raw buffer once the 25-point DFTs have consumed its samples (the loads have then
            // completed by data dependence: no fence needed, and the warp is not held up between its loads and its
            // arithmetic); the last of the 8 warps re-arms the buffer with the half-tile after next.
            auto raw_done = [&]() {
#if defined(WLM_KO_RAW) || (defined(WLM_KO_RAW2) && !defined(WLM_X_ATOM) && !defined(WLM_X_TMA))
                return;
#endif
                __syncwarp();
                if (lane == 0) {
                    const uint32_t old = atomicAdd(raw_readers + rb, 1u);
                    if (old == kFe - 1) {
                        raw_readers[rb] = 0;
#if !defined(WLM_KO_RAW2) || defined(WLM_X_TMA)
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        if (nxt.valid) tile_issue_tma(a, nxt.cc, rank + nxt.j * kCluster, raw, bar_raw + 8 * rb);
#endif
                    }
                }
            };
#ifdef WLM_RAW_EARLY
            stage1(raw, Y, wv, wg, lane, [&]() { __threadfence_block(); raw_done(); }, [&]() {});
#else
            stage1(raw, Y, wv, wg, lane, [&]() {}, raw_done);
#endif
            __syncwarp();      // Y is private to the warp: this is the whole stage 1 -> stage 2 hand-over
            stage2(Y, P, wg, lane, [&]() {
                // the back end must have read this buffer's previous content (half-tile n - 3)
#ifndef WLM_KO_P
                if (pround > 0) mbar_wait(bar_pfree + 8 * pb, (pround - 1) & 1);
#endif
            });
            __syncwarp();
#ifndef WLM_KO_P
            if (lane == 0) mbar_arrive(bar_pfull + 8 * pb);
#endif
            if (++rb == kRawBufs) { rb = 0; ++rround; }
            if (++pb == kPBufs) { pb = 0; ++pround; }
            cursor_next_tile(cur, a, stride, rank);
            if (nxt.valid) cursor_next_tile(nxt, a, stride, rank);
        }
    } else {
        // ============================================================================================
        // BACK END: warp bw owns filter run bw, lane = frame.  Per step of the CTA's stream:
        //   F  output of the clip that ended before this one: ONE retained half-tile (the slot B is about to overwrite)
        //   B  mel stage of the step's half-tile: wait P full | sparse gather | retain the mel power | arrive P free
        //   D  the clip ended: warp max -> CTA max; the CTA's last back-end warp delivers it to all CTAs of the clip
        // ============================================================================================
        const int bw = warp - kFe;
        const uint32_t twin = tmem_base + (static_cast<uint32_t>((warp & 3) * 32) << 16) +
                              static_cast<uint32_t>((bw >> 2) * kColsPerBe);       // lane i <-> frame i of the half-tile
        const int nf = kt.nf[bw], m0 = kt.m0[bw];
        Cursor s;
        cursor_open(s, a, first_clip, rank);
        int pb = 0, pround = 0;
        int fin_seq = 0;                         // clips whose max this warp has taken; parity = slot of the exchange
        float mx = 0.f;                          // running max of the mel power of the clip in flight (>= 0)
        bool pend = false;                       // an output pass is owed (max delivered, not yet waited for)
        int pend_b = 0, pend_n_my = 0;
        int out_j = 0;                           // next retained half-tile of the pending clip to write out
        bool have_max = false;
        float floor_v = 0.f;
        OutT* const outp = static_cast<OutT*>(a.out);

        auto output_slot = [&](int j) {
            constexpr float kLog10_2 = 0.30102999566398120f;
            const int tile = rank + j * kCluster;
            const bool va = tile * kTile + lane < kNFrames;
            OutT* of = outp + (static_cast<int64_t>(pend_b) * a.n_mels + m0) * kNFrames + tile * kTile + lane;
            float p[kMaxFiltersPerWarp];
            tmem_wait_st();
            tmem_ld_x16(twin + j * kTmemColsPerTile, p);
            // rows in blocks of four: straight-line code inside a block, so four MUFU.LG2 chains overlap
#pragma unroll
            for (int q0 = 0; q0 < kMaxFiltersPerWarp; q0 += 4) {
                if (q0 < nf) {
                    float lg[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        lg[i] = fmaf(fmaxf(lg2_approx(p[q0 + i]) * kLog10_2, floor_v), 0.25f, 1.0f);   // TF-FE:158,161
#pragma unroll
                    for (int i = 0; i < 4; ++i)
                        if (va && q0 + i < nf) of[(q0 + i) * kNFrames] = to_out<OutT>(lg[i]);
                }
            }
        };
        auto output_step = [&]() {
            if (!have_max) {    // first step after the clip ended: the maxima of the 6 CTAs of the clip
                const int fpar = fin_seq & 1;
                mbar_wait_cluster(bar_max + fpar * 8, (fin_seq >> 1) & 1);
                float pmax = lane < kCluster ? clip_max[fpar * kCluster + lane] : 0.f;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) pmax = fmaxf(pmax, __shfl_xor_sync(0xffffffffu, pmax, o));
                ++fin_seq;
                have_max = true;
                const float gmax = log10_floor(pmax);                 // TF-FE:157
                floor_v = fmaxf(gmax - 8.0f, -10.0f);                 // TF-FE:158 (log-mel is never below -10)
                if (rank == 0 && bw == 0 && lane == 0 && a.gmax) a.gmax[pend_b] = gmax;
                // half-tiles of mine that hold no real sample: log-mel is exactly -10 everywhere
                const float silent = (floor_v + 4.0f) * 0.25f;
                OutT* ob = outp + (static_cast<int64_t>(pend_b) * a.n_mels + m0) * kNFrames + lane;
                for (int tile = rank + pend_n_my * kCluster; tile < kTilesPerClip; tile += kCluster) {
                    OutT* of = ob + tile * kTile;
                    if (tile * kTile + lane < kNFrames)
                        for (int q = 0; q < nf; ++q) of[q * kNFrames] = to_out<OutT>(silent);
                }
            }
            if (out_j < pend_n_my) output_slot(out_j++);
            if (out_j >= pend_n_my) pend = false;
        };

#ifdef WLM_KO_P
        s.valid = false;        // (knock-out: no back end at all)
#endif
        while (s.valid || pend) {
            const bool do_tile = s.valid && s.j < s.n_my;
#ifndef WLM_KO_OUT
            if (pend) output_step();                                   // ---- F
#else
            pend = false;
#endif
            if (do_tile) {                                             // ---- B
                const float* P = P0 + pb * kPFloats;
                const int tile = rank + s.j * kCluster;
                mbar_wait_sleepy(bar_pfull + 8 * pb, pround & 1);
                const uint32_t tcol = twin + s.j * kTmemColsPerTile;
                auto sink = [&](const float (&o)[kMaxFiltersPerWarp]) { tmem_st_x16(tcol, o); };
#ifdef WLM_KO_MEL
                const float m1 = P[lane] + static_cast<float>(tcol & 1);
#else
                const float m1 = NMELS == 0 ? mel_stage(kt, P, bw, lane, sink) : mel_fixed<NMELS == 0 ? 80 : NMELS>(kt, P, bw, lane, sink);
#endif
                if (tile * kTile + lane < kNFrames) mx = fmaxf(mx, m1);     // frames past 3000 do not exist
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_pfree + 8 * pb);
                if (++pb == kPBufs) { pb = 0; ++pround; }
            }
            const bool clip_ends = s.valid && s.j + 1 >= (s.n_my > 0 ? s.n_my : 1);
            if (clip_ends) {                                           // ---- D
                while (pend) output_step();      // (only when this clip had fewer half-tiles than the one before it)
                float wmax = mx;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) wmax = fmaxf(wmax, __shfl_xor_sync(0xffffffffu, wmax, o));
                mx = 0.f;
                const int cpar = fin_seq & 1;            // F has run: this is the parity of the clip ending now
                if (lane == 0) {
                    if (s.n_my > 0) atomicMax(cta_max + cpar, __float_as_int(wmax));     // non-negative floats order like their bits
                    __threadfence_block();
                    if (atomicAdd(cta_cnt + cpar, 1u) == kBe - 1) {
                        __threadfence_block();
                        const float m = __int_as_float(atomicExch(cta_max + cpar, 0));   // (reset for the clip after next)
                        atomicExch(cta_cnt + cpar, 0u);
                        // Fire-and-forget: st.async writes the value into the peer's clip_max and completes 4 bytes on the
                        // peer's bar_max; every CTA posts the expectation of 6 x 4 bytes for itself.
                        const uint32_t slot_l = smem_u32(clip_max + cpar * kCluster + rank), bar_l = bar_max + cpar * 8;
                        mbar_expect_tx(bar_l, 4u * kCluster);
                        for (int r = 0; r < kCluster; ++r) {
                            uint32_t slot_r, bar_r;
                            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(slot_r) : "r"(slot_l), "r"(r));
                            asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(bar_r) : "r"(bar_l), "r"(r));
                            asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b32 [%0], %1, [%2];"
                                         ::"r"(slot_r), "r"(__float_as_uint(m)), "r"(bar_r) : "memory");
                        }
                    }
                }
                pend = true;
                have_max = false;
                out_j = 0;
                pend_b = s.b;
                pend_n_my = s.n_my;
            }
            if (s.valid) cursor_step(s, a, stride, rank);
        }
    }
    // all TMEM reads are complete (tcgen05.wait::ld inside tmem_ld_x16); release the allocation
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    cluster.sync();   // also keeps every CTA's shared memory alive until its peers have delivered their last max
    if (warp == 0) tmem_dealloc_512(tmem_base);
}

// ---- host side -----------------------------------------------------------------------------------
template <class OutT>
inline KernelFn stream_kernel_for_t(int variant) {
#ifdef WLM_DEVICE_ONLY
    return logmel_stream_kernel<80, OutT>;
#endif
    if (variant == 80) return logmel_stream_kernel<80, OutT>;
    if (variant == 128) return logmel_stream_kernel<128, OutT>;
    return logmel_stream_kernel<0, OutT>;
}
inline KernelFn stream_kernel_for(int variant, int out_format) {
#ifdef WLM_DEVICE_ONLY
    return stream_kernel_for_t<float>(variant);
#endif
    if (out_format == WLM_OUT_BF16) return stream_kernel_for_t<__nv_bfloat16>(variant);
    if (out_format == WLM_OUT_F16) return stream_kernel_for_t<__half>(variant);
    return stream_kernel_for_t<float>(variant);
}
inline void fill_stream_config(cudaLaunchConfig_t* cfg, cudaLaunchAttribute* at, int n_clusters, cudaStream_t st) {
    memset(cfg, 0, sizeof(*cfg));
    cfg->gridDim = dim3(kCluster * n_clusters);
    cfg->blockDim = dim3(kSThreads);
    cfg->dynamicSmemBytes = kSSmemBytes;
    cfg->stream = st;
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kCluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg->attrs = at;
    cfg->numAttrs = 1;
}
inline cudaError_t configure_stream(int variant, int* max_clusters) {
    cudaError_t e = cudaSuccess;
    for (int fmt : {WLM_OUT_F32, WLM_OUT_BF16, WLM_OUT_F16}) {
        e = cudaFuncSetAttribute(stream_kernel_for(variant, fmt), cudaFuncAttributeMaxDynamicSharedMemorySize, kSSmemBytes);
        if (e != cudaSuccess) return e;
    }
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    fill_stream_config(&cfg, at, 148, nullptr);
    int n = 0;
    e = cudaOccupancyMaxActiveClusters(&n, stream_kernel_for(variant, WLM_OUT_F32), &cfg);
    if (e != cudaSuccess) return e;
    if (n < 1) return cudaErrorLaunchOutOfResources;
    *max_clusters = n;
    return cudaSuccess;
}

#ifndef WLM_DEVICE_ONLY
// Dense batch, static split: clips [0, B - n_flat) round-robin over the clusters of the streaming kernel, the last n_flat
// over the CTAs of the flat kernel of logmel_fused.cuh (a programmatic dependent launch on the SMs the clusters leave idle;
// it executes griddepcontrol.wait before it exits, so the pair completes in stream order).
// rounds_per_clip: cluster rounds a flat CTA needs for one clip (see fused::flat_clip_count).
inline cudaError_t launch_static(const ClipArgs& a0, const Tables* d_tables, const Tables& h_tables, int variant,
                                 int max_clusters, cudaStream_t st, int* n_launches, int flat_ctas, double rounds_per_clip,
                                 bool* flat_broken) {
    ClipArgs a = a0;
    a.clip_first = 0;
    a.worker_base = 0;
    a.flat_reserve = 0;
    a.flat_cap = 0x7fffffff;
    int n_flat = 0;
    if (flat_ctas > 0 && a.B >= 2 * max_clusters) {
        int k = 0;
        while ((k + 1) * flat_ctas < a.B &&
               (k + 1) * rounds_per_clip <= (a.B - (k + 1) * flat_ctas + max_clusters - 1) / max_clusters) ++k;
        n_flat = k * flat_ctas;
    }
    ClipArgs af = a;
    af.clip_first = a.B - n_flat;
    a.B -= n_flat;
    const int nc = a.B < max_clusters ? a.B : max_clusters;
    const int nf = n_flat < flat_ctas ? n_flat : flat_ctas;
    a.n_workers = nc;
    af.n_workers = nf;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute at[1];
    fill_stream_config(&cfg, at, nc, st);
    *n_launches = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, stream_kernel_for(variant, a.out_format), a, h_tables.mel,
                                       static_cast<const float*>(d_tables->win_lane));
    if (e != cudaSuccess || nf == 0) return e;
    cudaLaunchConfig_t fcfg;
    memset(&fcfg, 0, sizeof(fcfg));
    fcfg.gridDim = dim3(nf);
    fcfg.blockDim = dim3(kThreads);
    fcfg.dynamicSmemBytes = kSmemBytes;
    fcfg.stream = st;
    cudaLaunchAttribute fat[1];
    fat[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    fat[0].val.programmaticStreamSerializationAllowed = 1;
    fcfg.attrs = fat;
    fcfg.numAttrs = 1;
    *n_launches = 2;
    e = cudaLaunchKernelEx(&fcfg, kernel_for(variant, true, af.out_format, false), af, h_tables.mel,
                           static_cast<const float*>(d_tables->win_lane));
    if (e == cudaSuccess) return e;
    (void)cudaGetLastError();            // dependent launch refused: ordinary launch (runs after the cluster kernel)
    if (flat_broken) *flat_broken = true;
    fcfg.attrs = nullptr;
    fcfg.numAttrs = 0;
    return cudaLaunchKernelEx(&fcfg, kernel_for(variant, true, af.out_format, false), af, h_tables.mel,
                              static_cast<const float*>(d_tables->win_lane));
}
#endif  // WLM_DEVICE_ONLY

}  // namespace stream
}  // namespace wlm
