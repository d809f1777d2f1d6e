"""ctypes binding of ``libhlmc_b200.so`` (the C ABI declared in ``include/hlmc_b200.h``).

There is no CPU fallback: if the shared library is missing the import fails
loudly, and if no CUDA device is present ``hlmc_plan_create`` fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libhlmc_b200.so")

# every symbol include/hlmc_b200.h declares
EXPORTS = (
    "hlmc_abi_version", "hlmc_last_error", "hlmc_params_default", "hlmc_num_frames",
    "hlmc_launch_count", "hlmc_plan_create", "hlmc_plan_destroy", "hlmc_plan_set_path",
    "hlmc_plan_uses_fast_path", "hlmc_plan_mel_basis", "hlmc_plan_dct_basis",
    "hlmc_extract_device", "hlmc_melspectrogram_device", "hlmc_stft_device",
    "hlmc_power_to_db_device", "hlmc_pool_device", "hlmc_fix_frames_device",
    "hlmc_extract_host", "hlmc_last_transfer_bytes", "hlmc_measure_fp32_peak",
    "hlmc_plan_set_timing", "hlmc_plan_read_timing", "hlmc_extract_host_ex",
    "hlmc_chroma_workspace_bytes", "hlmc_extract_device_ex", "hlmc_pool_device_ex",
    "hlmc_extract_host_io", "hlmc_column_stats_device", "hlmc_standardize_device",
    "hlmc_extract_pooled_device", "hlmc_graph_create", "hlmc_graph_launch", "hlmc_graph_destroy",
    "hlmc_resampled_length", "hlmc_load_frontend_device", "hlmc_resample_taps",
    "hlmc_impute_stats_device", "hlmc_scaler_stats_f64_device", "hlmc_impute_scale_device",
)
ABI_VERSION = 2

HLMC_OK, HLMC_ERR_PARAM, HLMC_ERR_UNSUPPORTED, HLMC_ERR_CUDA, HLMC_ERR_NOMEM = 0, -1, -2, -3, -4
PAD_MODES = {"constant": 0, "reflect": 1, "edge": 2}
REF_VALUE, REF_MAX = 0, 1


class HlmcParams(C.Structure):
    """Mirror of ``struct hlmc_params`` (field order and types must match the header)."""

    _fields_ = [
        ("sr", C.c_int32), ("n_fft", C.c_int32), ("hop_length", C.c_int32), ("win_length", C.c_int32),
        ("center", C.c_int32), ("pad_mode", C.c_int32), ("n_mels", C.c_int32), ("fmin", C.c_float),
        ("fmax", C.c_float), ("htk", C.c_int32), ("mel_norm", C.c_int32), ("power", C.c_float),
        ("n_mfcc", C.c_int32), ("lifter", C.c_float), ("ref_mode", C.c_int32), ("ref_value", C.c_float),
        ("amin", C.c_float), ("top_db", C.c_float), ("roll_percent", C.c_float),
        ("zcr_threshold", C.c_float),
    ]


class HlmcHostIo(C.Structure):
    """Mirror of ``struct hlmc_host_io``."""

    _fields_ = [
        ("wave", C.c_void_p), ("sample_format", C.c_int32), ("B", C.c_int64), ("n_valid", C.c_int64),
        ("pitch", C.c_int64), ("n_total", C.c_int64), ("logmel", C.c_void_p), ("mfcc", C.c_void_p),
        ("stats", C.c_void_p), ("chroma", C.c_void_p), ("tuning", C.c_void_p), ("pooled", C.c_void_p),
        ("status", C.c_void_p), ("pooled_with_chroma", C.c_int32), ("chunk_clips", C.c_int64),
        ("n_streams", C.c_int32),
        # ABI 2
        ("channels", C.c_int32), ("sr_in", C.c_int32), ("reserved0", C.c_int32),
        ("fixed_logmel", C.c_void_p), ("fixed_frames", C.c_int64), ("wave_out", C.c_void_p),
        ("valid_frames", C.c_void_p),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' at the repo root). "
            "This package has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32, f32 = C.c_void_p, C.c_int64, C.c_int32, C.c_float
    PP = C.POINTER(HlmcParams)
    lib.hlmc_abi_version.restype = C.c_int
    lib.hlmc_last_error.restype = C.c_char_p
    lib.hlmc_params_default.argtypes = [PP]
    lib.hlmc_params_default.restype = None
    lib.hlmc_num_frames.argtypes = [PP, i64]
    lib.hlmc_num_frames.restype = i64
    lib.hlmc_launch_count.restype = i64
    lib.hlmc_plan_create.argtypes = [PP, vp, vp, C.c_int, C.POINTER(vp)]
    lib.hlmc_plan_destroy.argtypes = [vp]
    lib.hlmc_plan_destroy.restype = None
    lib.hlmc_plan_set_path.argtypes = [vp, C.c_int]
    lib.hlmc_plan_uses_fast_path.argtypes = [vp]
    lib.hlmc_plan_mel_basis.argtypes = [vp, vp]
    lib.hlmc_plan_dct_basis.argtypes = [vp, vp]
    lib.hlmc_extract_device.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, vp]
    lib.hlmc_melspectrogram_device.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp]
    lib.hlmc_stft_device.argtypes = [vp, vp, i64, i64, i64, vp, vp]
    lib.hlmc_power_to_db_device.argtypes = [vp, vp, i64, i64, i64, i32, f32, f32, f32, vp, C.c_int, vp]
    lib.hlmc_pool_device.argtypes = [vp, vp, vp, vp, i64, i64, vp, vp]
    lib.hlmc_fix_frames_device.argtypes = [vp, vp, i64, i64, i64, i64, C.c_int, vp]
    lib.hlmc_extract_host.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, i64, C.c_int]
    lib.hlmc_extract_host_ex.argtypes = [vp, vp, C.c_int, i64, i64, i64, i64, vp, vp, vp, vp, vp, i64, C.c_int]
    lib.hlmc_chroma_workspace_bytes.argtypes = [vp, i64, i64]
    lib.hlmc_chroma_workspace_bytes.restype = i64
    lib.hlmc_extract_device_ex.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.hlmc_graph_create.argtypes = [vp, vp, i64, i64, i64, vp, vp, vp, vp, vp, C.POINTER(vp)]
    lib.hlmc_graph_launch.argtypes = [vp, vp]
    lib.hlmc_graph_destroy.argtypes = [vp]
    lib.hlmc_graph_destroy.restype = None
    lib.hlmc_extract_pooled_device.argtypes = [vp, vp, i64, i64, i64, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp, i64, vp]
    lib.hlmc_pool_device_ex.argtypes = [vp, vp, vp, vp, vp, i64, i64, vp, vp]
    lib.hlmc_extract_host_io.argtypes = [vp, C.POINTER(HlmcHostIo)]
    lib.hlmc_column_stats_device.argtypes = [vp, i64, i64, vp, vp, C.c_int, vp]
    lib.hlmc_standardize_device.argtypes = [vp, vp, i64, i64, vp, vp, C.c_int, vp]
    lib.hlmc_resampled_length.argtypes = [i64, i32, i32]
    lib.hlmc_resampled_length.restype = i64
    lib.hlmc_load_frontend_device.argtypes = [vp, vp, C.c_int, C.c_int, i64, i64, i64, i32, vp, i64, i64, vp, vp]
    lib.hlmc_resample_taps.argtypes = [i32, i32, vp, i64]
    lib.hlmc_resample_taps.restype = i64
    lib.hlmc_impute_stats_device.argtypes = [vp, i64, i64, vp, vp, C.c_int, vp]
    lib.hlmc_scaler_stats_f64_device.argtypes = [vp, i64, i64, vp, vp, vp, C.c_int, vp]
    lib.hlmc_impute_scale_device.argtypes = [vp, i64, i64, vp, i64, vp, vp, vp, vp, vp, C.c_int, vp]
    lib.hlmc_last_transfer_bytes.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.hlmc_last_transfer_bytes.restype = None
    lib.hlmc_measure_fp32_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    lib.hlmc_plan_set_timing.argtypes = [vp, C.c_int]
    lib.hlmc_plan_read_timing.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(i64)]
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is C.c_int and name not in ("hlmc_abi_version",):
            fn.restype = C.c_int
    if lib.hlmc_abi_version() != ABI_VERSION:
        raise ImportError(f"{LIB_PATH} has ABI {lib.hlmc_abi_version()}, this package needs {ABI_VERSION}: rebuild it")
    return lib


lib = _load()


def last_error() -> str:
    msg = lib.hlmc_last_error()
    return msg.decode("utf-8", "replace") if msg else ""
