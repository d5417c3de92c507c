#!/usr/bin/env python
"""Benchmark of the log-mel hot path (BASELINE.json metric: log-mel audio-seconds/second).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c2|c3]

One "step" = one pass of the hot path over one batch of synthetic PCM.  Default workload is
BASELINE.json configs[1] (whisper-small 80-mel, batch 256 x 30 s, device-resident PCM, one
B200).  With N > 1 (launched by torchrun, one rank per GPU) every rank processes its own batch
of the same size -- the path shards by clip with no data-path collective -- so scaling is WEAK
and `value` is the whole-job audio-seconds/second.  `--workload c3` is BASELINE configs[2]
(large-v3 128-mel, 1024 clips split across the ranks: strong scaling).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for how every field is derived.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_SAMPLES = 480000
N_FRAMES = 3000
CLIP_SECONDS = 30.0

WORKLOADS = {
    # name: (n_mels, clips per GPU at N=1, strong?)
    "c2": dict(n_mels=80, batch=256, strong=False,
               label="C2: whisper-small 80-mel, batch 256 x 30 s f32 PCM per GPU (BASELINE configs[1])"),
    "c3": dict(n_mels=128, batch=1024, strong=True,
               label="C3: whisper-large-v3 128-mel, 1024 x 30 s sharded by clip (BASELINE configs[2])"),
    "c4": dict(n_mels=80, batch=4096, strong=True, variable=True,
               label="C4: 4096 variable-length clips (1-30 s, medical-jsonl word-rate proxy) ragged, pad/trim in-kernel "
                     "(BASELINE configs[3])"),
}


def c4_lengths(n, seed=3):
    """d = clip(0.4 + words/2.6, 1, 30) s with words ~ the transcript word-count histogram of the
    reference's test.jsonl (min 1, median 11, mean 11.3, p95 15, max 25; the audio itself is absent), plus a 2 %
    tail uniform 10-30 s and 0.5 % at 35 s (trim).  The jsonl is not on the GPU box: a clipped normal stands in."""
    import numpy as np

    rng = np.random.default_rng(seed)
    words = np.clip(np.rint(rng.normal(11.26, 2.4, n)), 1, 25)
    d = np.clip(0.4 + words / 2.6, 1.0, 30.0)
    tail = rng.random(n)
    d = np.where(tail < 0.02, rng.uniform(10.0, 30.0, n), d)
    d = np.where(tail > 0.995, 35.0, d)
    return (d * 16000).astype(np.int64) // 8 * 8      # multiples of 8 samples: clips are back to back in the ragged buffer


def algorithmic_bytes_per_clip(n_mels: int) -> int:
    # read 480000*4 B PCM + write n_mels*3000*4 B features
    return N_SAMPLES * 4 + n_mels * N_FRAMES * 4


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            d = json.load(open(path))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU through NVML while the timed region runs."""

    REASONS = {0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
               0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x100: "display_clock_setting"}

    def __init__(self, index: int, period_s: float = 0.002):
        super().__init__(daemon=True)
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_ev = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._stop_ev.is_set():
            try:
                self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                mask = int(self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_ev.set()
        if self.ok:
            self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def visible_nvml_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except Exception:
            return local_rank
    return local_rank


# ------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (HF WhisperFeatureExtractor through a
# DataLoader shaped like REF/data_utils/data_loader.py:170-172 + data_collator.py:64-76)
# ------------------------------------------------------------------------------------------------
def synth_noise_clips(n, seed):
    import numpy as np

    rng = np.random.default_rng(seed)
    return [(0.1 * rng.standard_normal(N_SAMPLES)).astype(np.float32) for _ in range(n)]


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    from oracle.hf_reference import time_reference_dataloader

    cores = len(os.sched_getaffinity(0))
    # bounded sample: ~cores * 12 clips per step keeps a step to a few seconds of host time
    clips_per_step = max(16, min(wl["batch"], cores * 16))
    clips_per_step = (clips_per_step + 15) // 16 * 16
    clips = synth_noise_clips(min(clips_per_step, 64), seed=1)
    clips = [clips[i % len(clips)] for i in range(clips_per_step)]
    t_total, n_total = 0.0, 0
    for _ in range(args.steps):            # every step has its own untimed warm-up epoch inside
        r = time_reference_dataloader(clips, wl["n_mels"], 16, cores, "default")
        t_total += r["seconds"]
        n_total += r["clips"]
    value = CLIP_SECONDS * n_total / t_total
    sample = (f"{clips_per_step} x 30 s white-noise clips per step through a torch DataLoader "
              f"(batch 16, {cores} workers, 1 torch thread each) running the unmodified "
              f"transformers WhisperFeatureExtractor per clip + feature_extractor.pad stack")
    line = {
        "impl": "reference", "metric": "log-mel audio-seconds/second", "value": value, "unit": "audio-s/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * t_total / max(1, args.steps), "higher_is_better": True,
        "scaling": "strong" if wl["strong"] else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["label"], "n_mels": wl["n_mels"], "sample_clips_per_step": clips_per_step},
        "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args, wl, rank, world, local_rank):
    import torch

    import whisper_context_biasing_b200 as W

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    json_fd = None
    if world > 1:
        import torch.distributed as dist_mod

        dist = dist_mod
        # NCCL writes "NCCL version ..." straight to file descriptor 1 (at any NCCL_DEBUG level): point fd 1 at stderr for
        # the duration of the run and keep the real stdout for the ONE JSON line
        sys.stdout.flush()
        json_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group(backend="nccl", device_id=dev)

    from whisper_context_biasing_b200.sharding import clip_shard

    n_mels = wl["n_mels"]
    if wl["strong"]:                       # one global batch, contiguous clips per rank
        lo, hi = clip_shard(wl["batch"], rank, world)
        B = hi - lo
    else:                                  # the same batch size on every rank
        B = wl["batch"]
    fe = W.B200WhisperFeatureExtractor(feature_size=n_mels, device=dev)

    # synthetic PCM (family F1, white Gaussian sigma 0.1), generated on the host, pinned
    g = torch.Generator(device="cpu").manual_seed(1000 + rank)
    variable = bool(wl.get("variable"))
    if not variable:
        host_pcm = torch.empty((B, N_SAMPLES), dtype=torch.float32).pin_memory()
        torch.randn((B, N_SAMPLES), generator=g, out=host_pcm)
        host_pcm.mul_(0.1)
        pcm = host_pcm.to(dev, non_blocking=True)
        clips = [host_pcm[b].numpy() for b in range(B)]
        true_audio_s = CLIP_SECONDS * B
        run_device = lambda: fe.extract_device(pcm, out=out)
        in_bytes = B * N_SAMPLES * 4
    else:
        import numpy as np

        lens = c4_lengths(wl["batch"])[lo:hi]
        offs = np.zeros(B, dtype=np.int64)
        offs[1:] = np.cumsum((lens[:-1] + 7) // 8 * 8)
        total = int(offs[-1] + (lens[-1] + 7) // 8 * 8)
        host_pcm = torch.empty((total,), dtype=torch.float32).pin_memory()
        torch.randn((total,), generator=g, out=host_pcm)
        host_pcm.mul_(0.1)
        pcm = host_pcm.to(dev, non_blocking=True)
        d_offs = torch.from_numpy(offs).to(dev)
        d_lens = torch.from_numpy(lens.astype(np.int32)).to(dev)
        flat = host_pcm.numpy()
        clips = [flat[o:o + n] for o, n in zip(offs, lens)]
        true_audio_s = float(np.minimum(lens, N_SAMPLES).sum()) / 16000.0
        run_device = lambda: fe.extract_device(pcm, lengths=d_lens, offsets=d_offs, out=out)
        in_bytes = int(np.minimum(lens, N_SAMPLES).sum()) * 4
    out = torch.empty((B, n_mels, N_FRAMES), dtype=torch.float32, device=dev)
    torch.cuda.synchronize(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident: PCM already in HBM -------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        run_device()
    barrier()
    sampler = ClockSampler(visible_nvml_index(local_rank))
    sampler.start()
    launches0 = fe.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for s in range(args.steps):
        evs[s][0].record()
        run_device()
        evs[s][1].record()
    t1.record()
    barrier()
    clocks = sampler.stop()
    launches = fe.launch_count - launches0
    dev_ms = t0.elapsed_time(t1)
    launch_ms = sum(a.elapsed_time(b) for a, b in evs) / args.steps

    # ---- end to end: pinned host PCM -> H2D -> kernels -> D2H of the per-clip max --------------
    gmax_host = torch.empty((B,), dtype=torch.float32).pin_memory()
    e2e_steps = max(2, min(args.steps, 5))

    def e2e_step():
        fe.extract_host(clips, out=out)
        gmax_host.copy_(out[:, 0, 0], non_blocking=True)      # D2H read of the step's result
        torch.cuda.current_stream(dev).synchronize()

    if args.no_e2e:
        e2e_steps, e2e_ms, e2e_wall_ms = 1, float("nan"), float("nan")
    else:
        for _ in range(2):
            e2e_step()
        barrier()
        w0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            e2e_step()
        e1.record()
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), 0.0)
        e2e_wall_ms = 1e3 * (time.perf_counter() - w0)

    # ---- same end-to-end step with int16 PCM (half the H2D bytes; x/32768 on the GPU) ------
    e2e16_ms = float("nan")
    if not args.no_e2e:
        import numpy as np

        clips16 = [np.ascontiguousarray(np.round(c * 32767.0).astype(np.int16)) for c in clips]
        pin16 = [torch.from_numpy(c).pin_memory() for c in clips16] if len(clips16) <= 512 else None
        if pin16 is not None:
            c16 = [t.numpy() for t in pin16]

            def e2e16_step():
                fe.extract_host(c16, out=out)
                gmax_host.copy_(out[:, 0, 0], non_blocking=True)
                torch.cuda.current_stream(dev).synchronize()

            for _ in range(2):
                e2e16_step()
            barrier()
            f0 = torch.cuda.Event(enable_timing=True)
            f1 = torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(e2e_steps):
                e2e16_step()
            f1.record()
            barrier()
            e2e16_ms = f0.elapsed_time(f1)

    # ---- max over ranks ----------------------------------------------------------------------
    if dist is not None:
        t = torch.tensor([dev_ms, e2e_ms, launch_ms, e2e_wall_ms, e2e16_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, launch_ms, e2e_wall_ms, e2e16_ms = [float(v) for v in t.tolist()]

    total_clips = wl["batch"] if wl["strong"] else B * world
    value = CLIP_SECONDS * total_clips * args.steps / (dev_ms * 1e-3)
    e2e_value = CLIP_SECONDS * total_clips * e2e_steps / (e2e_ms * 1e-3)
    peak, peak_src = measured_peaks()
    alg_bytes = in_bytes + n_mels * N_FRAMES * 4 * B      # valid PCM read + full output written
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9

    if rank == 0:
        line = {
            "metric": "log-mel audio-seconds/second", "value": value, "unit": "audio-s/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if wl["strong"] else "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["label"], "n_mels": n_mels, "clips_per_gpu": B, "clip_seconds": 30,
                       "true_audio_seconds_per_gpu": true_audio_s,
                       "l2_policy": f"inputs larger than L2: {in_bytes / 1e6:.0f} MB PCM + "
                                    f"{B * n_mels * N_FRAMES * 4 / 1e6:.0f} MB features per step",
                       "timing": "CUDA events on the launch stream, barrier + synchronize both sides, max over ranks"},
            "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": in_bytes + B * 12,
                    "d2h_bytes_per_step": B * 4, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "wall_ms_per_step": e2e_wall_ms / e2e_steps,
                    "what": "wlm_logmel_host: pinned host f32 PCM -> chunked H2D overlapped with the kernels -> "
                            "features stay in HBM; D2H of one float per clip"},
            "e2e_int16": {"value": (CLIP_SECONDS * total_clips * e2e_steps / (e2e16_ms * 1e-3)) if e2e16_ms == e2e16_ms else None,
                          "unit": "audio-s/s", "h2d_bytes_per_step": in_bytes // 2 + B * 12, "d2h_bytes_per_step": B * 4,
                          "what": "same step with int16 PCM in pinned host memory (not the reference's input format: "
                                  "an optional ingest path, bit-identical to float(x)/32768)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": (int(KERNEL_DRAM_TRAFFIC_PER_CLIP[n_mels] * B) if (n_mels in KERNEL_DRAM_TRAFFIC_PER_CLIP and not variable) else None),
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg_bytes,
                         "launch_ms": launch_ms},
            "clocks": clocks,
        }
        if n_mels in ALGORITHMIC_MFLOP_PER_CLIP and not variable:
            sm_mhz = float((clocks or {}).get("sm_max_mhz") or 1965.0)
            fp32_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
            fp32_ach = ALGORITHMIC_MFLOP_PER_CLIP[n_mels] * 1e6 * B / (launch_ms * 1e-3) / 1e12
            line["roofline_fp32"] = {"bound": "fp32-cuda-core", "achieved": fp32_ach, "peak": fp32_peak, "unit": "TFLOP/s",
                                     "frac": fp32_ach / fp32_peak, "peak_source": "nominal: 148 SMs x 128 lanes x 2 x max SM clock",
                                     "ncu_pct_of_peak_while_active": NCU_PIPE_PCT[n_mels],
                                     "note": "the FFT is add-heavy (FADD2 = one issue, two lanes, no multiply): the FMA-pipe "
                                             "counter, not the flop fraction, says how busy the CUDA cores are"}
        if not args.no_cpu_baseline and world >= 1:
            line["cpu_baseline"] = cpu_baseline(n_mels)
        if json_fd is not None:
            os.write(json_fd, (json.dumps(line) + "\n").encode())
        else:
            print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


# dram__bytes_read.sum + dram__bytes_write.sum per clip, summed over the TWO kernels of a step (cluster kernel + the flat
# kernel that runs under it on the 16 SMs the clusters cannot cover), from the `ncu --set full` captures summarised under
# profiles/ (r01d_c2_80mel_summary.txt: 461.06 + 189.44 + 30.79 + 0.06 MB for 256 clips; r01d_c3_128mel_summary.txt:
# 1813.79 + 1399.15 + 153.70 + 82.48 MB for 1024 clips).  Algorithmic: 2.88 / 3.456 MB per clip; the measured traffic is
# slightly lower because the tail of the output is still dirty in L2 at kernel end (ncu runs the two kernels one after the
# other, so the flat kernel's clamp pass finds its clip in L2 there).
KERNEL_DRAM_TRAFFIC_PER_CLIP = {80: 681350000 / 256, 128: 3449120000 / 1024}

# FP32 side of the roofline (SURVEY 8d convention: 32.49 / 33.23 MFLOP per 30 s clip; nominal CUDA-core peak =
# 148 SMs x 128 lanes x 2 flop x SM clock).  The pipe utilisations are the ncu counters of the same captures
# (sm__pipe_fma_cycles_active, l1tex__data_pipe_lsu_wavefronts_mem_shared, smsp__issue_active: % of peak while active).
ALGORITHMIC_MFLOP_PER_CLIP = {80: 32.49, 128: 33.23}
NCU_PIPE_PCT = {80: {"fma_pipe": 51.6, "shared_memory_wavefronts": 48.1, "issue_slots": 61.2},
                128: {"fma_pipe": 50.2, "shared_memory_wavefronts": 47.2, "issue_slots": 61.0}}


def cpu_baseline(n_mels):
    """The reference's CPU implementation on this box's host cores, bounded sample (~10-30 s)."""
    from oracle.hf_reference import time_reference_dataloader

    cores = len(os.sched_getaffinity(0))
    n = max(64, min(cores * 64, 2048))           # ~10-30 s of single-core work
    n = (n + 15) // 16 * 16
    base = synth_noise_clips(min(n, 32), seed=1)
    clips = [base[i % len(base)] for i in range(n)]
    r = time_reference_dataloader(clips, n_mels, 16, cores, "default")
    return {"value": r["audio_s_per_s"], "unit": "audio-s/s", "cores": cores, "kind": "reference",
            "sample": f"{r['clips']} x 30 s white-noise clips, torch DataLoader batch 16, {cores} workers x 1 thread, "
                      f"unmodified transformers WhisperFeatureExtractor per clip + pad stack ({r['seconds']:.1f} s)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host end-to-end leg (profiling runs)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return
    if world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (one rank per GPU); running rank 0 only", file=sys.stderr)
    run_ours(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
