// Shared declarations between the C-ABI translation unit and the kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "wlm.h"

namespace wlm {

constexpr int kNfft = WLM_N_FFT;        // 400
constexpr int kHop = WLM_HOP;           // 160
constexpr int kNSamples = WLM_N_SAMPLES;  // 480000
constexpr int kNFrames = WLM_N_FRAMES;  // 3000
constexpr int kNFreq = WLM_N_FREQ;      // 201
constexpr int kMaxMels = 128;

// Sparse form of the triangular mel bank (host-built from the dense table the caller passes,
// so the weights stay bit-identical to the reference's float32 table).
// FFT bin k feeds filter lo[k] with weight w_lo[k] and filter lo[k]+1 with weight w_hi[k];
// a weight of 0 means "no contribution" (lo[k] is then still a valid filter index or -1 with
// both weights 0).
struct MelSparse {
    int16_t lo[kNFreq + 3];
    float w_lo[kNFreq + 3];
    float w_hi[kNFreq + 3];
};

// Work queue shared by the cluster kernel and its cluster-less twin (one per plan, device memory): the first
// `n_workers` clips are taken statically, one per worker; every further clip is `n_workers + next++`.  The last worker to
// leave resets both counters, so a launch always starts from {0, 0}.
struct ClipQueue {
    unsigned int next;
    unsigned int done;
};

// Arguments common to every log-mel kernel launch.
struct ClipArgs {
    const void* pcm;          // f32 or i16
    const int64_t* offsets;   // nullptr -> b * row_stride
    const int32_t* lengths;   // nullptr -> dense_len
    int64_t row_stride;
    int32_t dense_len;        // min(row_stride, 480000) when lengths == nullptr
    int32_t pcm_format;       // WLM_PCM_*
    int32_t n_mels;
    int32_t B;                // clips of the batch; a STATIC launch processes [clip_first, B), worker w the clips
    int32_t clip_first;       //   clip_first + w, + n_workers of its kernel, ... ; a DYNAMIC launch pulls from the queue
    ClipQueue* queue;         // see above
    int32_t n_workers;        // clusters + flat CTAs of this launch
    int32_t worker_base;      // index of this kernel's first worker: 0 (cluster kernel), number of clusters (flat kernel)
    int32_t flat_reserve;     // a flat CTA takes another clip only while at least this many clips are still unassigned
    int32_t flat_cap;         // clips a flat CTA may take at most
    void* out;                // [B][n_mels][3000], element type out_format
    float* gmax;              // [B] (may be workspace)
    int32_t out_format;       // WLM_OUT_*
};

}  // namespace wlm
