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


# ---------------------------------------------------------------------------
# The scripts' tabular normalisation ([R] src/1_preprocessing.py:303-311,
# src/1_preprocessing_advanced.py:384-391) on device, float64:
#   np.where(np.isinf(X), np.nan, X) -> SimpleImputer(strategy="mean") -> StandardScaler()
# ---------------------------------------------------------------------------
def make_sklearn_imputer(statistics):
    """A fitted ``SimpleImputer(strategy="mean")`` carrying the given per-column means (NaN = column without
    any observed value, which sklearn drops on transform)."""
    from sklearn.impute import SimpleImputer

    im = SimpleImputer(strategy="mean")
    im.statistics_ = np.asarray(statistics, dtype=np.float64)
    im.n_features_in_ = int(im.statistics_.shape[0])
    im._fit_dtype = np.dtype(np.float64)
    im._fill_dtype = np.dtype(np.float64)
    im.indicator_ = None
    return im


def fit_transform_tabular_device(x, group=None):
    """The scripts' inf -> nan, mean-impute, StandardScaler chain for an (N, D) float64 CUDA tensor.

    With ``torch.distributed`` initialised, ``x`` is this rank's rows and the statistics are global
    (sums / counts add, (mean, m2) combine with Chan's formula).  Returns
    ``(imputed (N, D'), scaled (N, D'), sklearn SimpleImputer, sklearn StandardScaler)`` where D' drops the
    columns without a single finite value, as sklearn does."""
    import torch

    assert x.is_cuda and x.dtype == torch.float64 and x.dim() == 2
    x = x.contiguous()
    N, D = x.shape
    dev = x.device
    stream = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
    csum = torch.empty((D,), dtype=torch.float64, device=dev)
    ccnt = torch.empty((D,), dtype=torch.int64, device=dev)
    _check(lib.hlmc_impute_stats_device(ptr(x), N, D, ptr(csum), ptr(ccnt), dev.index, stream))
    dist = None
    try:
        import torch.distributed as tdist

        if tdist.is_available() and tdist.is_initialized() and tdist.get_world_size(group) > 1:
            dist = tdist
    except ImportError:  # pragma: no cover
        pass
    n_total = float(N)
    if dist is not None:
        dist.all_reduce(csum, group=group)
        dist.all_reduce(ccnt, group=group)
    fill = csum / ccnt.to(torch.float64)                 # NaN where a column has no finite entry
    keep = torch.nonzero(ccnt > 0).flatten().to(torch.int32)
    Do = int(keep.numel())
    imputer = make_sklearn_imputer(fill.cpu().numpy())
    fill_safe = torch.nan_to_num(fill, nan=0.0)
    mean = torch.empty((D,), dtype=torch.float64, device=dev)
    m2 = torch.empty((D,), dtype=torch.float64, device=dev)
    _check(lib.hlmc_scaler_stats_f64_device(ptr(x), N, D, ptr(fill_safe), ptr(mean), ptr(m2), dev.index, stream))
    if dist is not None:
        cnt = torch.tensor([float(N)], dtype=torch.float64, device=dev)
        dist.all_reduce(cnt, group=group)
        s1 = mean * float(N)
        dist.all_reduce(s1, group=group)
        gmean = s1 / cnt
        m2 = m2 + float(N) * (mean - gmean) ** 2
        dist.all_reduce(m2, group=group)
        mean, n_total = gmean, float(cnt.item())
    keep64 = keep.to(torch.int64)
    mean_k, var_k = mean[keep64], (m2 / max(n_total, 1.0))[keep64]
    scaler = make_sklearn_scaler(n_total, mean_k.cpu().numpy(), var_k.cpu().numpy())
    mean_d = torch.from_numpy(scaler.mean_).to(dev)
    scale_d = torch.from_numpy(scaler.scale_).to(dev)
    imputed = torch.empty((N, Do), dtype=torch.float64, device=dev)
    scaled = torch.empty((N, Do), dtype=torch.float64, device=dev)
    _check(lib.hlmc_impute_scale_device(ptr(x), N, D, ptr(keep), Do, ptr(fill_safe), ptr(mean_d), ptr(scale_d),
                                        ptr(imputed), ptr(scaled), dev.index, stream))
    return imputed, scaled, imputer, scaler
