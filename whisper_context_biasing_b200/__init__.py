"""B200-native (sm_100a) log-mel front-end: a drop-in for the `WhisperFeatureExtractor` path of
thanh-nt25/Whisper-context-biasing.  Hand-written CUDA behind a C ABI (include/wlm.h, lib/libwlm.so);
Python is the host-side mirror of the reference's interface.  No CPU fallback."""
from .feature_extraction import B200WhisperFeatureExtractor, LogMelBatch, slaney_mel_filters  # noqa: F401

__all__ = ["B200WhisperFeatureExtractor", "LogMelBatch", "slaney_mel_filters"]
