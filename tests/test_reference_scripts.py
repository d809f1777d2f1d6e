"""The reference's own function bodies, executed here.

CPU: with ``librosa`` = the oracle they must reproduce the oracle's script-level helpers exactly (so the helpers
ARE the reference's code path, not a paraphrase).  GPU: with ``librosa`` = this package they must agree with the
oracle at north_star's tolerances - the drop-in claim of SURVEY.md 8(b), tested with the scripts' own code.
Also executes the scripts' normalise-and-save cells and compares every file with what ``preprocessing.save_*``
writes, then loads the directories exactly as the three VAE scripts do."""
import os
import pickle

import numpy as np
import pytest

import refscripts as rs
from oracle import librosa_oracle as orc

needs_ref = pytest.mark.skipif(not rs.available(), reason="/root/reference is not present on this box")
SR = 22050


def _clips(n=66150):
    rng = np.random.default_rng(77)
    t = np.arange(n) / SR
    white = (0.1 * rng.standard_normal(n)).astype(np.float32)
    harm = sum(np.sin(2 * np.pi * 220.0 * k * t + k) / k for k in range(1, 12)) * np.exp(-t)
    harm = (0.4 * harm / np.abs(harm).max() + 1e-4 * rng.standard_normal(n)).astype(np.float32)
    half = white.copy()
    half[n // 2:] = 0.0
    return [white, harm, half]


@needs_ref
def test_oracle_helpers_are_the_reference_function_bodies():
    lib = rs.oracle_librosa()
    basic = rs.load_functions(rs.BASIC, rs.BASIC_FUNCS, lib)
    adv = rs.load_functions(rs.ADVANCED, rs.ADV_FUNCS, lib)
    assert basic.CONFIG["n_fft"] == 2048 and basic.CONFIG["n_mfcc"] == 40 and adv.CONFIG["fixed_time_steps"] == 1024
    for y in _clips(30000):
        assert np.array_equal(basic.extract_mel_spectrogram(y, SR), orc.basic_extract_mel_spectrogram(y, SR))
        assert np.array_equal(basic.extract_mfcc(y, SR), orc.basic_extract_mfcc(y, SR))
        ref_stats, got_stats = basic.extract_spectral_features(y, SR), orc.extract_spectral_features(y, SR)
        assert list(ref_stats) == list(got_stats)
        for k in ref_stats:
            assert np.array_equal(ref_stats[k], got_stats[k]) and ref_stats[k].dtype == got_stats[k].dtype
        f = basic.extract_all_features(y, SR)
        assert f.shape == (370,) and f.dtype == np.float64
        assert np.array_equal(f, orc.extract_all_features(y, SR))
        assert np.array_equal(adv.extract_mel_spectrogram(y, SR), orc.adv_extract_mel_spectrogram(y, SR))
        g = adv.extract_flattened_features(y, SR)
        assert g.shape == (290,) and np.array_equal(g, orc.extract_flattened_features(y, SR))


def _cols_close(got, want, n_mfcc, T):
    """Pooled vector against the oracle's, group by group, at the tolerances of the underlying features."""
    nm = 128
    assert got.shape == want.shape
    o = 2 * nm
    assert np.abs(got[:o] - want[:o]).max() <= 0.01, "log-mel mean/std columns"
    if n_mfcc:
        scale = max(np.abs(want[o:o + 2 * n_mfcc]).max(), 1e-6)
        assert np.abs(got[o:o + 2 * n_mfcc] - want[o:o + 2 * n_mfcc]).max() <= 1e-4 * scale, "MFCC columns"
        o += 2 * n_mfcc
    for s in range(5):
        for j in (0, 1):
            g, w = got[o + 2 * s + j], want[o + 2 * s + j]
            tol = 1e-4 * max(abs(want[o + 2 * s]), abs(w), 1e-12)
            if s == 2:
                tol += 2 * (SR / 2048) / T          # a rolloff tie may move one frame by one bin
            assert abs(g - w) <= tol, f"stat {s} col {j}: {g} vs {w}"
    o += 10
    assert np.abs(got[o:] - want[o:]).max() <= 1e-4, "chroma columns"


@needs_ref
@pytest.mark.gpu
def test_reference_functions_run_on_this_package(built):
    """`import hybrid_language_music_clustering_vae_b200 as librosa` under the scripts' own functions."""
    hl = built
    basic = rs.load_functions(rs.BASIC, rs.BASIC_FUNCS, hl)
    adv = rs.load_functions(rs.ADVANCED, rs.ADV_FUNCS, hl)
    clips = _clips() + [np.concatenate(_clips(66150) * 4)[:250000]]          # T = 130, and T = 489
    clips.append(np.tile(_clips()[0], 11)[:661500])                          # 30 s: the crop branch, T = 1292
    for y in clips:
        T = 1 + len(y) // 512
        mel = basic.extract_mel_spectrogram(y, SR)
        want_mel = orc.basic_extract_mel_spectrogram(y, SR)
        assert mel.shape == want_mel.shape and mel.dtype == np.float32
        assert np.abs(mel - want_mel).max() <= 0.01
        mf, want_mf = basic.extract_mfcc(y, SR), orc.basic_extract_mfcc(y, SR)
        assert mf.shape == want_mf.shape == (40, T)
        assert np.abs(mf - want_mf).max() <= 1e-4 * np.abs(want_mf).max()
        st, want_st = basic.extract_spectral_features(y, SR), orc.extract_spectral_features(y, SR)
        assert list(st) == list(want_st)
        for k in st:
            assert st[k].shape == want_st[k].shape == (1, T) and st[k].dtype == want_st[k].dtype, k
            if k == "spectral_rolloff":
                assert (np.abs(st[k] - want_st[k]) > 1e-3).sum() <= 2
            else:
                assert np.abs(st[k] - want_st[k]).max() <= 1e-4 * max(np.abs(want_st[k]).max(), 1e-12), k
        ch, want_ch = basic.extract_chroma_features(y, SR), orc.extract_chroma_features(y, SR)
        assert ch.shape == want_ch.shape == (12, T) and np.abs(ch - want_ch).max() <= 1e-4
        f = basic.extract_all_features(y, SR)
        assert f.shape == (370,) and f.dtype == np.float64
        _cols_close(f, orc.extract_all_features(y, SR), 40, T)
        img, want_img = adv.extract_mel_spectrogram(y, SR), orc.adv_extract_mel_spectrogram(y, SR)
        assert img.shape == want_img.shape == (128, 1024) and img.dtype == np.float32
        assert np.abs(img - want_img).max() <= 0.01
        g = adv.extract_flattened_features(y, SR)
        assert g.shape == (290,) and g.dtype == np.float64
        _cols_close(g, orc.extract_flattened_features(y, SR), 0, T)
        # ... and the package's own batched helpers give the same rows as the scripts' functions on it
        assert np.allclose(hl.preprocessing.extract_all_features(y, SR), f, rtol=1e-5, atol=1e-5)
        assert np.allclose(hl.preprocessing.extract_flattened_features(y, SR), g, rtol=1e-5, atol=1e-5)
        assert np.abs(hl.preprocessing.extract_mel_spectrogram_fixed(y, SR) - img).max() <= 1e-5


# ---------------------------------------------------------------------------
# processed_data1 / processed_data2: files, dtypes and contents
# ---------------------------------------------------------------------------
def _fake_features(n, d, rng):
    x = rng.standard_normal((n, d)) * rng.uniform(0.5, 30.0, d) + rng.uniform(-50, 50, d)
    x[3, 7] = np.nan
    x[5, 11] = np.inf
    x[9, 11] = -np.inf
    return x


def _pickle_load(path):
    with open(path, "rb") as f:
        return pickle.load(f)


def _run_reference_save_cells(path, first, last, ns, out_dir):
    import pandas as pd

    ns = dict(ns, np=np, pd=pd, os=os, pickle=pickle, OUTPUT_PATH=str(out_dir))
    exec(rs.top_level_block(path, first, last), ns)
    return ns


def _same_dir(a, b, pickles, arrays, rtol=0.0):
    assert sorted(os.listdir(a)) == sorted(os.listdir(b))
    for name in arrays:
        x, y = np.load(os.path.join(a, name), allow_pickle=True), np.load(os.path.join(b, name), allow_pickle=True)
        assert x.shape == y.shape and x.dtype == y.dtype, name
        if x.dtype.kind == "f":
            if rtol:
                scale = np.abs(x[np.isfinite(x)]).max()
                assert np.allclose(x, y, rtol=rtol, atol=rtol * scale, equal_nan=True), name
            else:
                assert np.array_equal(x, y, equal_nan=True), name
        else:
            assert np.array_equal(x, y), name
    assert open(os.path.join(a, "metadata.csv")).read() == open(os.path.join(b, "metadata.csv")).read()
    for name in pickles:
        p, q = _pickle_load(os.path.join(a, name)), _pickle_load(os.path.join(b, name))
        assert type(p) is type(q), name
        if isinstance(p, dict):
            assert p == q
            continue
        assert repr(sorted(p.get_params().items())) == repr(sorted(q.get_params().items())), name
        for attr in ("statistics_", "mean_", "var_", "scale_", "n_samples_seen_", "n_features_in_"):
            if hasattr(p, attr):
                u, v = np.asarray(getattr(p, attr), dtype=np.float64), np.asarray(getattr(q, attr), dtype=np.float64)
                assert np.allclose(u, v, rtol=max(rtol, 1e-15), atol=1e-12 if rtol else 0.0, equal_nan=True), (name, attr)


def _check_processed_data1(hl, tmp_path, device):
    import pandas as pd

    rng = np.random.default_rng(5)
    feats = _fake_features(40, 370, rng)
    labels = [("rock", "pop", "folk")[i % 3] for i in range(40)]
    meta = [{"language": "bn" if i % 2 else "en", "genre": labels[i], "filename": f"{i}.wav"} for i in range(40)]
    ref_dir, our_dir = tmp_path / "ref1", tmp_path / "ours1"
    ref_dir.mkdir()
    basic = rs.load_functions(rs.BASIC, (), None)
    _run_reference_save_cells(rs.BASIC, "from sklearn.preprocessing import StandardScaler", "pickle.dump(CONFIG",
                              dict(features=feats.copy(), labels=list(labels), metadata=[dict(m) for m in meta],
                                   CONFIG=basic.CONFIG), ref_dir)
    df = pd.DataFrame(meta)
    df["label"] = labels
    hl.preprocessing.save_processed_data1(str(our_dir), feats.copy(), labels, df, config=basic.CONFIG, device=device)
    _same_dir(str(ref_dir), str(our_dir), ("scaler.pkl", "imputer.pkl", "config.pkl"),
              ("features_raw.npy", "features_normalized.npy", "labels.npy"), rtol=0.0 if device is None else 1e-12)
    # [R] Simple_VAE.py:31-33 reads it like this
    features = np.load(os.path.join(our_dir, "features_normalized.npy"))
    lab = np.load(os.path.join(our_dir, "labels.npy"), allow_pickle=True)
    metadata = pd.read_csv(os.path.join(our_dir, "metadata.csv"))
    assert features.shape == (40, 370) and features.dtype == np.float64 and not np.isnan(features).any()
    assert len(lab) == 40 and list(metadata.columns) == ["language", "genre", "filename", "label"]
    raw = np.load(os.path.join(our_dir, "features_raw.npy"))
    assert np.isnan(raw[3, 7]) and np.isinf(raw[5, 11])            # saved as extracted, not imputed


def _check_processed_data2(hl, tmp_path, device):
    import pandas as pd

    rng = np.random.default_rng(6)
    n = 12
    mel = (rng.standard_normal((n, 128, 1024)) * 12 - 40).astype(np.float32)
    flat = _fake_features(n, 290, rng)
    labels = np.array([("rock", "pop")[i % 2] for i in range(n)])
    lyr = rng.standard_normal((n, 768)).astype(np.float32)
    meta = [{"language": "bn", "genre": labels[i], "filename": f"{i}.wav", "file_id": str(i)} for i in range(n)]
    ref_dir, our_dir = tmp_path / "ref2", tmp_path / "ours2"
    ref_dir.mkdir()
    adv = rs.load_functions(rs.ADVANCED, (), None)
    from sklearn.preprocessing import StandardScaler

    _run_reference_save_cells(rs.ADVANCED, "mel_scaler = StandardScaler()", "pickle.dump(CONFIG",
                              dict(mel_spectrograms=mel.copy(), flattened_features=flat.copy(), labels=labels.copy(),
                                   lyrics_embeddings=lyr.copy(), metadata_list=[dict(m) for m in meta],
                                   CONFIG=adv.CONFIG, StandardScaler=StandardScaler), ref_dir)
    df = pd.DataFrame(meta)
    df["label"] = labels
    hl.preprocessing.save_processed_data2(str(our_dir), mel.copy(), flat.copy(), labels, lyr, df, config=adv.CONFIG,
                                          device_scaler=device)
    _same_dir(str(ref_dir), str(our_dir), ("mel_scaler.pkl", "flat_scaler.pkl", "imputer.pkl", "config.pkl"),
              ("mel_spectrograms_raw.npy", "mel_spectrograms_normalized.npy", "features_raw.npy",
               "features_normalized.npy", "lyrics_embeddings.npy", "labels.npy"),
              rtol=0.0 if device is None else 2e-6)
    # [R] Convolutional_VAE.py:39-46, Conditional_VAE.py:54-66 read it like this
    audio = np.load(os.path.join(our_dir, "mel_spectrograms_normalized.npy"))[:, np.newaxis, :, :]
    text = np.load(os.path.join(our_dir, "lyrics_embeddings.npy"))
    hand = np.load(os.path.join(our_dir, "features_normalized.npy"))
    metadata = pd.read_csv(os.path.join(our_dir, "metadata.csv"))
    assert audio.shape == (n, 1, 128, 1024) and audio.dtype == np.float32
    assert text.shape == (n, 768) and hand.shape == (n, 290) and hand.dtype == np.float64
    assert list(metadata.columns) == ["language", "genre", "filename", "file_id", "label"]
    assert np.load(os.path.join(our_dir, "mel_spectrograms_raw.npy")).dtype == np.float32


@needs_ref
def test_processed_data_dirs_equal_the_scripts_own_save_cells(built, tmp_path):
    _check_processed_data1(built, tmp_path, None)
    _check_processed_data2(built, tmp_path, None)


@needs_ref
@pytest.mark.gpu
def test_processed_data_dirs_with_device_normalisation(built, tmp_path):
    _check_processed_data1(built, tmp_path, 0)
    _check_processed_data2(built, tmp_path, 0)
