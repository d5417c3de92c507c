"""CPU restatement of the Whisper log-mel front-end (TEST INFRASTRUCTURE, not product).

This file is the *oracle* for the hot path `16 kHz PCM -> input_features [B, n_mels, 3000] f32`.
Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu-baseline / `--impl reference`
legs may import it.  The product path (`whisper_context_biasing_b200`) never does and fails
loudly when its CUDA library is missing.

Where the algorithm lives
-------------------------
The reference repo (thanh-nt25/Whisper-context-biasing) contains no arithmetic for this path.
It calls the un-vendored third-party dependency `transformers` (pinned `== 4.51.3` in the
reference's `requirements.txt:1`; 5.5.0 is what is installed in this image):

    REF/data_utils/data_loader.py:170-172     feature_extractor(audio, sampling_rate=16000).input_features
    REF/data_utils/data_collator.py:17-24     prepare_dataset(...) same call
    REF/data_utils/data_collator.py:64-76     feature_extractor.pad(..., padding="longest", return_tensors="pt")

so this module restates the *published* algorithm of
`transformers/models/whisper/feature_extraction_whisper.py` (TF-FE below),
`transformers/audio_utils.py` (TF-AU) and `transformers/feature_extraction_sequence_utils.py`
(TF-SU), line numbers for transformers 5.5.0.

Parity pinning
--------------
The reference has no tests, golden vectors or fixtures for this path, so
the restatement is pinned against outputs of the live third-party implementation generated in the
build container by `tests/golden/make_golden.py` (committed fixtures under `tests/golden/`), and
`tests/test_oracle.py` additionally compares it with the live `WhisperFeatureExtractor` whenever
`transformers` is importable.
"""
from __future__ import annotations

import numpy as np

SAMPLING_RATE = 16000          # TF-FE:72
N_FFT = 400                    # TF-FE:75
HOP_LENGTH = 160               # TF-FE:73
CHUNK_LENGTH = 30              # TF-FE:74
N_SAMPLES = CHUNK_LENGTH * SAMPLING_RATE      # 480000, TF-FE:91
NB_MAX_FRAMES = N_SAMPLES // HOP_LENGTH       # 3000,   TF-FE:92
N_FREQ = 1 + N_FFT // 2                       # 201,    TF-FE:96


# ----------------------------------------------------------------------------------------------
# mel filter bank  (TF-AU:263-296, 299-332, 356-375, 453-544 with the ctor args of TF-FE:95-103)
# ----------------------------------------------------------------------------------------------
def hertz_to_mel_slaney(freq):
    """Slaney mel scale: linear below 1 kHz (200/3 Hz per mel), log above (TF-AU:286-296)."""
    freq = np.asarray(freq, dtype=np.float64)
    min_log_hertz, min_log_mel = 1000.0, 15.0
    logstep = 27.0 / np.log(6.4)
    mels = 3.0 * freq / 200.0
    log_region = freq >= min_log_hertz
    with np.errstate(divide="ignore"):
        mels = np.where(log_region, min_log_mel + np.log(np.maximum(freq, 1e-300) / min_log_hertz) * logstep, mels)
    return mels


def mel_to_hertz_slaney(mels):
    """Inverse of the above (TF-AU:322-332)."""
    mels = np.asarray(mels, dtype=np.float64)
    min_log_hertz, min_log_mel = 1000.0, 15.0
    logstep = np.log(6.4) / 27.0
    freq = 200.0 * mels / 3.0
    log_region = mels >= min_log_mel
    return np.where(log_region, min_log_hertz * np.exp(logstep * (mels - min_log_mel)), freq)


def mel_filter_bank(n_mels: int, n_freq: int = N_FREQ, sampling_rate: int = SAMPLING_RATE,
                    min_frequency: float = 0.0, max_frequency: float = 8000.0) -> np.ndarray:
    """[n_freq, n_mels] float64 triangular filters, Slaney scale + Slaney area norm.

    TF-AU:516-519 edge frequencies, :528 FFT-bin frequencies, :356-375 triangles
    `max(0, min(down, up))`, :532-535 area normalisation `2 / (f[i+2] - f[i])`.
    """
    mel_min = hertz_to_mel_slaney(min_frequency)
    mel_max = hertz_to_mel_slaney(max_frequency)
    mel_freqs = np.linspace(mel_min, mel_max, n_mels + 2)
    filter_freqs = mel_to_hertz_slaney(mel_freqs)
    fft_freqs = np.linspace(0, sampling_rate // 2, n_freq)

    filter_diff = np.diff(filter_freqs)
    slopes = filter_freqs[None, :] - fft_freqs[:, None]
    down = -slopes[:, :-2] / filter_diff[:-1]
    up = slopes[:, 2:] / filter_diff[1:]
    fb = np.maximum(0.0, np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2:n_mels + 2] - filter_freqs[:n_mels])
    return fb * enorm[None, :]


def hann_window(n: int = N_FFT, dtype=np.float64) -> np.ndarray:
    """Periodic Hann `0.5 - 0.5 cos(2 pi i / n)`; `torch.hann_window(400)` at TF-FE:141,
    `np.hanning(401)[:-1]` at TF-AU:593-607."""
    i = np.arange(n, dtype=np.float64)
    return (0.5 - 0.5 * np.cos(2.0 * np.pi * i / n)).astype(dtype)


# ----------------------------------------------------------------------------------------------
# pad / trim to 30 s   (TF-SU:327-332 truncate, :268-278 right zero pad, TF-FE:296-303)
# ----------------------------------------------------------------------------------------------
def pad_or_trim(clips, n_samples: int = N_SAMPLES) -> np.ndarray:
    """list of 1-D arrays (any length >= 0) -> [B, n_samples] float32, `x[:n]` then right zeros."""
    out = np.zeros((len(clips), n_samples), dtype=np.float32)
    for b, x in enumerate(clips):
        x = np.asarray(x, dtype=np.float32).reshape(-1)[:n_samples]
        out[b, : x.shape[0]] = x
    return out


def frame_mask(lengths, n_samples: int = N_SAMPLES, hop: int = HOP_LENGTH) -> np.ndarray:
    """int32 [B, 3000] attention mask: sample mask `[0, min(L, n))` subsampled `[::hop]`
    (TF-FE:328-337; 480000 % 160 == 0 so no trailing trim)."""
    lengths = np.minimum(np.asarray(lengths, dtype=np.int64), n_samples)
    pos = np.arange(0, n_samples, hop, dtype=np.int64)
    return (pos[None, :] < lengths[:, None]).astype(np.int32)


# ----------------------------------------------------------------------------------------------
# the hot path
# ----------------------------------------------------------------------------------------------
def _frames(padded_clip: np.ndarray, dtype) -> np.ndarray:
    """reflect-pad 200 each side of the *already padded/trimmed* buffer (torch.stft center=True,
    TF-FE:149; np.pad(..., 'reflect') TF-AU:769-771) and cut frames t=0..2999 over
    padded[160 t, 160 t + 400).  Frame 3000 is computed by torch and dropped (TF-FE:150)."""
    x = np.pad(padded_clip.astype(dtype), (N_FFT // 2, N_FFT // 2), mode="reflect")
    idx = HOP_LENGTH * np.arange(NB_MAX_FRAMES)[:, None] + np.arange(N_FFT)[None, :]
    return x[idx]


def log_mel_spectrogram(padded: np.ndarray, n_mels: int = 80, precision: str = "f64",
                        mel_filters: np.ndarray | None = None, return_gmax: bool = False):
    """[B, 480000] float32 (already padded/trimmed) -> [B, n_mels, 3000] float32.

    precision "f64" follows the numpy fallback (`_np_extract_fbank_features`, TF-FE:105-133 +
    `spectrogram`, TF-AU:768-832): fp64 window/FFT/mel, float32 at the end.
    precision "f32" follows the torch default path (TF-FE:135-164): fp32 window, fp32 STFT,
    `abs()**2`, fp32 mel matmul, clamp 1e-10, log10, max-8, (x+4)/4.
    """
    padded = np.asarray(padded, dtype=np.float32)
    if padded.ndim == 1:
        padded = padded[None]
    assert padded.shape[1] == N_SAMPLES
    dt = np.float64 if precision == "f64" else np.float32
    fb = mel_filter_bank(n_mels) if mel_filters is None else np.asarray(mel_filters, dtype=np.float64)
    fbT = fb.T.astype(dt)                                   # [M, 201]   TF-FE:152
    win = hann_window(N_FFT, dt)
    out = np.empty((padded.shape[0], n_mels, NB_MAX_FRAMES), dtype=np.float32)
    gmax = np.empty((padded.shape[0],), dtype=np.float32)
    for b in range(padded.shape[0]):
        fr = _frames(padded[b], dt) * win[None, :]           # [3000, 400]
        spec = np.fft.rfft(fr, axis=1)                       # un-normalised one-sided DFT, k=0..200
        if precision == "f32":
            spec = spec.astype(np.complex64)
        power = (np.abs(spec) ** 2).astype(dt)               # TF-FE:150 / TF-AU:807-808
        mel = fbT @ power.T                                  # [M, 3000]  TF-FE:153 / TF-AU:812-813
        log_spec = np.log10(np.maximum(mel, dt(1e-10)))      # TF-FE:155 / TF-AU:813,818-819
        g = log_spec.max()
        log_spec = np.maximum(log_spec, g - dt(8.0))         # TF-FE:156-160 / :129
        log_spec = (log_spec + dt(4.0)) / dt(4.0)            # TF-FE:161 / :130
        out[b] = log_spec.astype(np.float32)
        gmax[b] = g
    return (out, gmax) if return_gmax else out


def extract(clips, n_mels: int = 80, precision: str = "f64") -> np.ndarray:
    """ragged list of PCM clips -> [B, n_mels, 3000] float32 (pad/trim + log-mel): what
    `WhisperFeatureExtractor.__call__` returns as `.input_features` (TF-FE:189-342)."""
    if isinstance(clips, np.ndarray) and clips.ndim == 1:
        clips = [clips]
    return log_mel_spectrogram(pad_or_trim(list(clips)), n_mels=n_mels, precision=precision)


# ----------------------------------------------------------------------------------------------
# seeded synthetic PCM families -- shared by tests, golden script and bench
# ----------------------------------------------------------------------------------------------
FAMILIES = ("noise", "sine", "chirp", "speech", "int16", "gap", "zeros", "tiny")


def synth_clip(family: str, n: int, seed: int) -> np.ndarray:
    """float32 PCM in [-1, 1], 16 kHz mono, `np.random.default_rng(seed)`."""
    if n == 0:
        return np.zeros(0, dtype=np.float32)
    rng = np.random.default_rng(seed)
    t = np.arange(n, dtype=np.float64) / SAMPLING_RATE
    if family == "noise":          # F1 white Gaussian sigma 0.1
        x = 0.1 * rng.standard_normal(n)
    elif family == "sine":         # F2 440 Hz A=0.5
        x = 0.5 * np.sin(2 * np.pi * 440.0 * t + rng.uniform(0, 2 * np.pi))
    elif family == "chirp":        # F3 linear chirp 100 -> 7100 Hz A=0.5
        dur = max(n / SAMPLING_RATE, 1e-9)
        x = 0.5 * np.sin(2 * np.pi * (100.0 * t + 0.5 * (7000.0 / dur) * t * t))
    elif family == "speech":       # F4 exp-decay-filtered noise x slow envelope, ~0.05 RMS
        from scipy.signal import lfilter  # scipy is in the image; test infrastructure only
        a = 0.95
        y = lfilter([1.0 - a], [1.0, -a], rng.standard_normal(n))      # one-pole lowpass
        env = np.abs(lfilter([1.0 - 0.9995], [1.0, -0.9995], rng.standard_normal(n)))
        env = env / (env.max() + 1e-12)
        x = y * env
        x = 0.05 * x / (np.sqrt(np.mean(x * x)) + 1e-12)
    elif family == "int16":        # F5 F1 rounded to the int16 grid
        x = np.round(0.1 * rng.standard_normal(n) * 32768.0) / 32768.0
    elif family == "gap":          # F6 F1 with the middle third exactly zero
        x = 0.1 * rng.standard_normal(n)
        x[n // 3: 2 * n // 3] = 0.0
    elif family == "zeros":        # F7
        x = np.zeros(n)
    elif family == "tiny":         # F8 sigma 1e-6
        x = 1e-6 * rng.standard_normal(n)
    else:
        raise ValueError(family)
    return np.clip(x, -1.0, 1.0).astype(np.float32)
