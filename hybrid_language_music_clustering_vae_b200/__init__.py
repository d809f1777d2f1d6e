"""B200-native audio feature extractor with a librosa-compatible surface.

``import hybrid_language_music_clustering_vae_b200 as librosa`` covers the calls the
reference's preprocessing scripts make: ``stft``, ``power_to_db`` and
``feature.{melspectrogram, mfcc, spectral_centroid, spectral_bandwidth,
spectral_rolloff, zero_crossing_rate, rms}``; ``FeatureExtractor`` is the fused
batched entry that replaces their per-file loops.
"""
from .core import (FeatureExtractor, ParameterError, UnsupportedError, STAT_NAMES, get_extractor,
                   launch_count, measure_fp32_peak)
from .api import stft, power_to_db, feature
from . import pipeline, preprocessing, scaler, sharding, synth

__all__ = ["FeatureExtractor", "ParameterError", "UnsupportedError", "STAT_NAMES", "get_extractor",
           "launch_count", "measure_fp32_peak", "stft", "power_to_db", "feature", "preprocessing",
           "scaler", "sharding", "synth", "pipeline"]
