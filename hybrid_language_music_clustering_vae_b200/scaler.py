"""sklearn.preprocessing.StandardScaler with the column statistics and the transform on the GPU.

For the (N, 131072) flattened mel images of ``src/1_preprocessing_advanced.py:376-382``
(SURVEY.md 8f-4).  The result is a genuine, picklable ``sklearn`` ``StandardScaler`` whose
``mean_ / var_ / scale_`` came from the device, so ``mel_scaler.pkl`` stays loadable by the
reference.  Across GPUs the per-rank statistics are combined with Chan's parallel formula:
the only collective anywhere on this path (two all-reduces of D float64 values).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import lib
from .core import _check


def combine_stats(counts, means, m2s):
    """Chan et al. combination of per-shard (n, mean, sum of squared deviations) -> global."""
    counts = np.asarray(counts, dtype=np.float64)
    n = counts.sum()
    means = np.asarray(means, dtype=np.float64)
    m2s = np.asarray(m2s, dtype=np.float64)
    if n == 0:
        return 0.0, np.zeros_like(means[0]), np.zeros_like(m2s[0])
    mean = (counts[:, None] * means).sum(axis=0) / n
    m2 = (m2s + counts[:, None] * (means - mean[None, :]) ** 2).sum(axis=0)
    return n, mean, m2


def column_stats_device(x):
    """(N, D) float32 CUDA tensor -> (mean, m2) float64 CUDA tensors (m2 = sum of squared deviations)."""
    import torch

    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2 and x.is_contiguous()
    N, D = x.shape
    mean = torch.empty((D,), dtype=torch.float64, device=x.device)
    m2 = torch.empty((D,), dtype=torch.float64, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _check(lib.hlmc_column_stats_device(C.c_void_p(x.data_ptr()), N, D, C.c_void_p(mean.data_ptr()),
                                        C.c_void_p(m2.data_ptr()), x.device.index, C.c_void_p(stream)))
    return mean, m2


def make_sklearn_scaler(n, mean, var):
    """A fitted StandardScaler carrying the given statistics (scale_ via sklearn's own zero handling)."""
    from sklearn.preprocessing import StandardScaler
    from sklearn.preprocessing._data import _handle_zeros_in_scale, _is_constant_feature

    sc = StandardScaler()
    sc.mean_ = np.asarray(mean, dtype=np.float64)
    sc.var_ = np.asarray(var, dtype=np.float64)
    sc.n_samples_seen_ = np.int64(n)
    sc.n_features_in_ = int(sc.mean_.shape[0])
    constant = _is_constant_feature(sc.var_, sc.mean_, sc.n_samples_seen_)
    sc.scale_ = _handle_zeros_in_scale(np.sqrt(sc.var_), copy=False, constant_mask=constant)
    return sc


def fit_transform_device(x, group=None, inplace=False):
    """StandardScaler().fit_transform(x) for an (N, D) float32 CUDA tensor.

    With ``torch.distributed`` initialised and ``group`` given (or the default group), ``x`` is
    this rank's shard of the rows and the statistics are global.  Returns (y, sklearn scaler).
    """
    import torch

    x = x.contiguous()
    N, D = x.shape
    mean, m2 = column_stats_device(x)
    n_total = float(N)
    try:
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            cnt = torch.tensor([float(N)], dtype=torch.float64, device=x.device)
            dist.all_reduce(cnt, group=group)
            s1 = mean * float(N)
            dist.all_reduce(s1, group=group)
            gmean = s1 / cnt
            m2 = m2 + float(N) * (mean - gmean) ** 2
            dist.all_reduce(m2, group=group)
            mean, n_total = gmean, float(cnt.item())
    except ImportError:  # pragma: no cover
        pass
    var = (m2 / max(n_total, 1.0)).cpu().numpy()
    sc = make_sklearn_scaler(n_total, mean.cpu().numpy(), var)
    mean32 = torch.from_numpy(sc.mean_.astype(np.float32)).to(x.device)
    scale32 = torch.from_numpy(sc.scale_.astype(np.float32)).to(x.device)
    y = x if inplace else torch.empty_like(x)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _check(lib.hlmc_standardize_device(C.c_void_p(x.data_ptr()), C.c_void_p(y.data_ptr()), N, D,
                                       C.c_void_p(mean32.data_ptr()), C.c_void_p(scale32.data_ptr()),
                                       x.device.index, C.c_void_p(stream)))
    return y, sc
