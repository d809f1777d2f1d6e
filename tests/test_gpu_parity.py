"""Parity of the CUDA path (through the C ABI) against the CPU oracle; needs a B200."""
import os

import numpy as np
import pytest

from oracle import librosa_oracle as orc
from parity import oracle_clip, compare_clip, assert_clip, LOGMEL_TOL_DB, REL_TOL

pytestmark = pytest.mark.gpu
SR = 22050


def _run(hl, y, generic=False, pitch=None, **kw):
    import torch

    ex = hl.FeatureExtractor(ref=np.max, **kw)
    if generic:
        ex.force_generic(True)
    if pitch is None:
        d = torch.from_numpy(y).cuda()
    else:
        store = torch.zeros((y.shape[0], pitch), dtype=torch.float32, device="cuda")
        store[:, : y.shape[1]] = torch.from_numpy(y).cuda()
        d = store[:, : y.shape[1]]
    out = ex.extract_device(d)
    torch.cuda.synchronize()
    res = {k: v.cpu().numpy() for k, v in out.items()}
    ex.close()
    return res


def _check_batch(hl, y, generic=False, pitch=None, **kw):
    res = _run(hl, y, generic=generic, pitch=pitch, **kw)
    okw = {k: v for k, v in kw.items()}
    okw.setdefault("n_mfcc", 20)
    flips = 0
    for i in range(y.shape[0]):
        want = oracle_clip(y[i], **okw)
        got = {k: res[k][i] for k in ("logmel", "mfcc", "stats") if k in res}
        m = compare_clip(got, want, n_fft=kw.get("n_fft", 2048), roll_percent=kw.get("roll_percent", 0.85), y=y[i])
        assert_clip(m, where=f"clip {i} {kw} generic={generic}")
        flips += m.get("rolloff_flips", 0)
    frames = y.shape[0] * res["logmel"].shape[-1]
    assert flips <= max(2, frames // 200), f"too many rolloff tie flips: {flips}/{frames}"
    assert not res["status"].any()
    return res


@pytest.mark.parametrize("generic", [False, True])
def test_mixture_3s_full_set(built, generic):
    """configs[0]-shaped: 3 s clips, the 1_preprocessing.py feature set."""
    y = built.synth.synth_batch(40, 66150, seed=20260)
    res = _check_batch(built, y, generic=generic, n_mfcc=40)
    assert res["logmel"].shape == (40, 128, 130) and res["mfcc"].shape == (40, 40, 130)
    assert res["stats"].shape == (40, 5, 130)


def test_30s_clips(built):
    """GTZAN-shaped (configs[2]): 30 s clips, T = 1292."""
    y = built.synth.synth_batch(3, 661500, seed=20262)
    res = _check_batch(built, y, n_mfcc=40)
    assert res["logmel"].shape == (3, 128, 1292)


@pytest.mark.parametrize("pad_mode", ["constant", "reflect", "edge"])
@pytest.mark.parametrize("generic", [False, True])
def test_pad_modes(built, pad_mode, generic):
    y = built.synth.synth_batch(6, 9000, seed=3)
    _check_batch(built, y, generic=generic, pad_mode=pad_mode, n_mfcc=13)


@pytest.mark.parametrize("generic", [False, True])
def test_center_false(built, generic):
    y = built.synth.synth_batch(4, 10000, seed=4)
    res = _check_batch(built, y, generic=generic, center=False, n_mfcc=13)
    assert res["logmel"].shape[-1] == 1 + (10000 - 2048) // 512


@pytest.mark.parametrize("n", [1, 2, 511, 1025, 2047, 2048, 2049, 4097])
@pytest.mark.parametrize("pad_mode", ["constant", "reflect"])
def test_ragged_lengths(built, n, pad_mode):
    """Frame counts and values for clips shorter than / around one frame."""
    y = (0.1 * np.random.default_rng(n).standard_normal((3, n))).astype(np.float32)
    res = _check_batch(built, y, pad_mode=pad_mode, n_mfcc=13)
    assert res["logmel"].shape[-1] == 1 + n // 512


@pytest.mark.parametrize("pitch_extra", [0, 1, 2, 3, 5])
def test_row_pitch_and_misalignment(built, pitch_extra):
    """Rows that are not 16-byte aligned (TMA needs the shifted-landing path) and odd pitches."""
    n = 8191
    y = built.synth.synth_batch(5, n, seed=9)
    _check_batch(built, y, pitch=n + pitch_extra, n_mfcc=13)


@pytest.mark.parametrize("n_fft,hop", [(512, 128), (1024, 256), (4096, 1024), (256, 64), (8192, 2048)])
def test_n_fft_sweep(built, n_fft, hop):
    """configs[4]'s FFT sizes; n_fft=512/256 with 128 mels has empty filters and must not crash."""
    y = built.synth.synth_batch(4, 30000, seed=n_fft)
    _check_batch(built, y, n_fft=n_fft, hop_length=hop, n_mfcc=20)


@pytest.mark.parametrize("kw", [
    dict(n_mels=64, n_mfcc=20), dict(n_mels=40, n_mfcc=13, htk=True), dict(fmin=100.0, fmax=8000.0),
    dict(win_length=1024), dict(window="hamming"), dict(window=("kaiser", 8.0)), dict(hop_length=300),
    dict(hop_length=1000, n_mfcc=12), dict(n_mels=200, n_mfcc=40), dict(norm=None), dict(roll_percent=0.5),
])
@pytest.mark.parametrize("generic", [False, True])
def test_parameter_coverage(built, kw, generic):
    y = built.synth.synth_batch(4, 12000, seed=17)
    _check_batch(built, y, generic=generic, **kw)


def test_power_one_and_db_options(built):
    import torch

    hl = built
    y = hl.synth.synth_batch(4, 12000, seed=21)
    for ref, top_db, amin in ((1.0, 80.0, 1e-10), (np.max, None, 1e-10), (0.5, 40.0, 1e-6)):
        ex = hl.FeatureExtractor(power=1.0, ref=ref, top_db=top_db, amin=amin, n_mfcc=13, lifter=22)
        out = ex.extract_device(torch.from_numpy(y).cuda())
        for i in range(len(y)):
            mel = orc.melspectrogram(y=y[i], sr=SR, power=1.0)
            want = orc.power_to_db(mel, ref=ref, top_db=top_db, amin=amin)
            assert np.abs(out["logmel"][i].cpu().numpy() - want).max() <= LOGMEL_TOL_DB
            mf = orc.mfcc(y=y[i], sr=SR, n_mfcc=13, lifter=22, power=1.0)
            assert np.abs(out["mfcc"][i].cpu().numpy() - mf).max() <= REL_TOL * np.abs(mf).max()


def test_plan_tables_match_oracle(built):
    import scipy.fftpack

    ex = built.FeatureExtractor(n_mfcc=40)
    assert np.abs(ex.mel_basis() - orc.mel(sr=SR, n_fft=2048)).max() < 1e-7
    D = scipy.fftpack.dct(np.eye(128, dtype=np.float64), axis=0, type=2, norm="ortho")[:40]
    assert np.abs(ex.dct_basis() - D).max() < 1e-7
    ex2 = built.FeatureExtractor(n_mels=40, htk=True, fmin=50.0, fmax=7000.0, n_fft=1024, n_mfcc=0)
    assert np.abs(ex2.mel_basis() - orc.mel(sr=SR, n_fft=1024, n_mels=40, htk=True, fmin=50.0, fmax=7000.0)).max() < 1e-7


def test_nonfinite_clips_are_flagged_not_fatal(built):
    """[R] 1_preprocessing.py:248-251: a bad file is skipped, the rest of the batch is unaffected."""
    import torch

    hl = built
    y = hl.synth.synth_batch(6, 20000, seed=5)
    clean = _run(hl, y, n_mfcc=13)
    bad = y.copy()
    bad[1, 777] = np.nan
    bad[4, 19999] = np.inf
    res = _run(hl, bad, n_mfcc=13)
    assert res["status"].tolist() == [0, 1, 0, 0, 1, 0]
    for i in (0, 2, 3, 5):
        for k in ("logmel", "mfcc", "stats"):
            assert np.array_equal(res[k][i], clean[k][i]), (k, i)
    with pytest.raises(hl.ParameterError):
        hl.feature.melspectrogram(y=bad[1])
    with pytest.raises(hl.ParameterError):
        hl.preprocessing.extract_all_features(bad[4], SR)


def test_host_pipeline_equals_device_path(built):
    import torch

    hl = built
    y = hl.synth.synth_batch(37, 22050, seed=8)
    ex = hl.FeatureExtractor(ref=np.max, n_mfcc=40)
    dev = ex.extract_device(torch.from_numpy(y).cuda(), pooled=True)
    torch.cuda.synchronize()
    for chunk, streams in ((0, 3), (5, 2), (37, 1), (1, 4)):
        host = ex.extract_host(y, pooled=True, chunk_clips=chunk, n_streams=streams)
        for k in ("logmel", "mfcc", "stats", "status", "pooled"):
            assert np.array_equal(host[k], dev[k].cpu().numpy()), (k, chunk, streams)
        h2d, d2h = ex.last_transfer_bytes()
        assert h2d == y.nbytes and d2h == sum(host[k].nbytes for k in ("logmel", "mfcc", "stats", "status", "pooled"))
    # strided host input (row pitch > n)
    wide = np.zeros((37, 22050 + 7), np.float32)
    wide[:, :22050] = y
    host = ex.extract_host(wide[:, :22050])
    assert np.array_equal(host["logmel"], dev["logmel"].cpu().numpy())
    # pooled columns = np.mean / np.std over frames, in the scripts' order
    lm, mf, st = host["logmel"], host["mfcc"], host["stats"]
    want = np.concatenate([lm.mean(-1), lm.std(-1), mf.mean(-1), mf.std(-1),
                           np.stack([st.mean(-1), st.std(-1)], axis=-1).reshape(len(y), 10)], axis=1)
    got = dev["pooled"].cpu().numpy()
    assert got.shape == (37, 346)
    assert np.abs(got - want).max() <= 2e-4 * np.abs(want).max()


def test_multi_gpu_thread_sharding_single_device(built):
    """extract_multi_gpu with the one visible device twice: shards meet in host memory, no collective."""
    hl = built
    y = hl.synth.synth_batch(11, 22050, seed=12)
    ref = hl.FeatureExtractor(ref=np.max, n_mfcc=40).extract_host(y)
    out = hl.sharding.extract_multi_gpu(y, [0, 0], dict(ref=np.max, n_mfcc=40))
    for k in ("logmel", "mfcc", "stats", "status"):
        assert np.array_equal(out[k], ref[k]), k


def test_librosa_compatible_functions(built):
    """Each librosa call the scripts make, by its own name and signature."""
    import torch

    hl = built
    y = hl.synth.synth_batch(3, 22050, seed=30)[1]
    D = hl.stft(y)
    Do = orc.stft(y)
    assert D.shape == Do.shape and D.dtype == np.complex64
    assert np.abs(D - Do).max() <= 2e-6 * np.abs(Do).max()
    mel = hl.feature.melspectrogram(y=y, sr=SR, n_mels=128, n_fft=2048, hop_length=512)
    melo = orc.melspectrogram(y=y, sr=SR)
    assert mel.dtype == np.float32 and np.abs(mel - melo).max() <= 1e-5 * melo.max()
    db = hl.power_to_db(mel, ref=np.max)
    assert np.abs(db - orc.power_to_db(melo, ref=np.max)).max() <= LOGMEL_TOL_DB
    assert db.max() == 0.0
    mf = hl.feature.mfcc(y=y, sr=SR, n_mfcc=40, n_fft=2048, hop_length=512)
    mfo = orc.mfcc(y=y, sr=SR, n_mfcc=40)
    assert mf.shape == (40, 44) and np.abs(mf - mfo).max() <= REL_TOL * np.abs(mfo).max()
    for name in ("spectral_centroid", "spectral_bandwidth"):
        got = getattr(hl.feature, name)(y=y, sr=SR, hop_length=512)
        want = getattr(orc, name)(y=y, sr=SR, hop_length=512)
        assert got.shape == want.shape == (1, 44) and got.dtype == np.float64
        assert np.abs(got - want).max() <= REL_TOL * np.abs(want).max()
    ro = hl.feature.spectral_rolloff(y=y, sr=SR, hop_length=512)
    assert np.mean(np.abs(ro - orc.spectral_rolloff(y=y, sr=SR)) < 1e-3) >= 0.95
    z = hl.feature.zero_crossing_rate(y, hop_length=512)
    assert np.array_equal(z, orc.zero_crossing_rate(y, hop_length=512)) and z.dtype == np.float64
    r = hl.feature.rms(y=y, hop_length=512)
    ro_ = orc.rms(y=y, hop_length=512)
    assert r.dtype == np.float32 and np.abs(r - ro_).max() <= REL_TOL * ro_.max()
    # batched leading dimension and CUDA tensors
    yb = hl.synth.synth_batch(3, 22050, seed=30)
    mb = hl.feature.melspectrogram(y=yb, sr=SR)
    assert mb.shape == (3, 128, 44) and np.array_equal(mb[1], mel)
    mc = hl.feature.melspectrogram(y=torch.from_numpy(yb).cuda(), sr=SR)
    assert mc.is_cuda and np.array_equal(mc.cpu().numpy(), mb)
    # per-clip ref=max for batched power_to_db
    dbb = hl.power_to_db(mb, ref=np.max)
    assert np.array_equal(dbb[1], db) and all(dbb[i].max() == 0.0 for i in range(3))


def test_script_level_functions_and_layout(built, tmp_path):
    hl = built
    pp = hl.preprocessing
    y = hl.synth.synth_batch(5, 66150, seed=40)
    # 1_preprocessing.py
    lm = pp.extract_mel_spectrogram(y[0], SR)
    assert np.abs(lm - orc.basic_extract_mel_spectrogram(y[0], SR)).max() <= LOGMEL_TOL_DB
    mf = pp.extract_mfcc(y[0], SR)
    assert np.abs(mf - orc.basic_extract_mfcc(y[0], SR)).max() <= REL_TOL * np.abs(mf).max()
    sp = pp.extract_spectral_features(y[0], SR)
    assert list(sp) == ["spectral_centroid", "spectral_bandwidth", "spectral_rolloff", "zcr", "rms"]
    assert sp["spectral_centroid"].shape == (1, 130) and sp["rms"].dtype == np.float32
    f = pp.extract_all_features(y[0], SR)
    want = orc.extract_all_features(y[0], SR, with_chroma=False)
    assert f.shape == (370,) and f.dtype == np.float64
    assert np.abs(f[346:] - orc.extract_all_features(y[0], SR)[346:]).max() <= 1e-4   # chroma on device
    # pooled means / stds of quantities that each meet the per-frame tolerance; the two
    # rolloff columns may move by a tie flip (one 10.77 Hz bin in one of 130 frames)
    d = np.abs(f[:346] - want)
    tol = 1e-3 + 1e-4 * np.abs(want)
    tol[340:342] = 3 * (SR / 2048)
    assert np.all(d <= tol), np.nonzero(d > tol)
    assert np.all(pp.extract_all_features(y[0], SR, chroma="zeros")[346:] == 0.0)   # explicit policy, logged
    fb = pp.extract_all_features_batch(y, SR, chroma="nan")
    # (f came through the kernel variant with the piptrack epilogue, which synthesises the Hann window in registers;
    #  fb through the default variant, which reads the window from its table: equal to rounding, not bitwise)
    assert fb.shape == (5, 370) and np.isnan(fb[:, 346:]).all()
    assert np.allclose(fb[0, :346], f[:346], rtol=1e-5, atol=1e-5)
    fb = pp.extract_all_features_batch(y, SR)
    # 1_preprocessing_advanced.py
    mel, flat, status = pp.process_batch_advanced(y, SR)
    assert mel.shape == (5, 128, 1024) and mel.dtype == np.float32 and flat.shape == (5, 290)
    want_mel = orc.adv_extract_mel_spectrogram(y[2], SR)
    assert np.abs(mel[2] - want_mel).max() <= LOGMEL_TOL_DB       # T=130 -> right-padded with the clip minimum
    long = np.zeros((2, 661500), np.float32)
    long[:, :66150] = y[:2]
    mel_l, flat_l, _ = pp.process_batch_advanced(long, SR)
    assert np.abs(mel_l[1] - orc.adv_extract_mel_spectrogram(long[1], SR)).max() <= LOGMEL_TOL_DB   # T=1292 -> crop
    wantf = orc.extract_flattened_features(long[1], SR, with_chroma=False)
    assert np.abs(flat_l[1, :266] - wantf).max() <= 2e-3 * max(1.0, np.abs(wantf).max())
    # on-disk layout
    labels = np.array(["a", "b", "a", "b", "a"])
    raw, norm = pp.save_processed_data1(str(tmp_path / "p1"), fb, labels)
    assert np.load(tmp_path / "p1" / "features_raw.npy").shape == (5, 370)
    assert np.load(tmp_path / "p1" / "features_normalized.npy").dtype == np.float64
    for name in ("labels.npy", "scaler.pkl", "imputer.pkl", "config.pkl"):
        assert (tmp_path / "p1" / name).exists()
    pp.save_processed_data2(str(tmp_path / "p2"), mel, flat, labels, lyrics_embeddings=np.zeros((5, 768), np.float32))
    assert np.load(tmp_path / "p2" / "mel_spectrograms_normalized.npy").shape == (5, 128, 1024)
    assert np.load(tmp_path / "p2" / "mel_spectrograms_raw.npy").dtype == np.float32
    assert np.load(tmp_path / "p2" / "features_raw.npy").shape == (5, 290)
    for name in ("mel_scaler.pkl", "flat_scaler.pkl", "imputer.pkl", "config.pkl", "lyrics_embeddings.npy"):
        assert (tmp_path / "p2" / name).exists()


def test_size_independent_properties_full_batch(built):
    """BASELINE configs[1] at full size (10,000 x 3 s): properties that need no oracle."""
    import torch

    hl = built
    B, n = 10000, 66150
    g = torch.Generator(device="cuda").manual_seed(1)
    y = torch.randn((B, n), device="cuda", generator=g) * 0.1
    y[7] = 0.0
    y[8, n // 2:] = 0.0
    ex = hl.FeatureExtractor(ref=np.max, n_mfcc=40)
    a = ex.extract_device(y)
    a = {k: v.clone() for k, v in a.items()}
    assert a["logmel"].shape == (B, 128, 130)
    mx = a["logmel"].amax(dim=(1, 2))
    assert torch.all(mx == 0.0), "ref=max: every clip's maximum is exactly 0 dB"
    assert float(a["logmel"].min()) >= -80.0 and torch.all(a["logmel"][7] == 0.0)
    assert torch.all(a["logmel"][8][:, 70:] == -80.0), "frames in the silent half sit exactly on the top_db floor"
    assert torch.all(a["stats"][7] == 0.0) and not a["status"].any()
    assert torch.isfinite(a["mfcc"]).all() and torch.isfinite(a["stats"]).all()
    # deterministic, and independent of batch composition / position
    b = ex.extract_device(y)
    for k in ("logmel", "mfcc", "stats"):
        assert torch.equal(a[k], b[k]), k
    idx = torch.tensor([9999, 5, 4321, 8, 7], device="cuda")
    c = ex.extract_device(y[idx].contiguous())
    for k in ("logmel", "mfcc", "stats"):
        assert torch.equal(c[k], a[k][idx]), k
    # scaling a clip by 2: log-mel(ref=max) unchanged, MFCC[0] shifts by 10*log10(4)*sqrt(128), rms doubles
    s = ex.extract_device((y[:64] * 2.0).contiguous())
    assert (s["logmel"] - a["logmel"][:64]).abs().max() <= 1e-3
    # generic kernel agrees with the register-FFT kernel on the same inputs
    ex.force_generic(True)
    gsub = ex.extract_device(y[:256].contiguous())
    assert (gsub["logmel"] - a["logmel"][:256]).abs().max() <= 2e-3
    assert (gsub["mfcc"] - a["mfcc"][:256]).abs().max() <= 1e-4 * float(a["mfcc"][:256].abs().max())
    # and a sample of the full batch against the oracle
    yc = y[[0, 8, 5000, 9999]].cpu().numpy()
    for j, i in enumerate([0, 8, 5000, 9999]):
        want = oracle_clip(yc[j], n_mfcc=40)
        got = {k: a[k][i].cpu().numpy() for k in ("logmel", "mfcc", "stats")}
        assert_clip(compare_clip(got, want), where=f"full batch clip {i}")


def test_pcm16_and_device_pad_front_end(built):
    """[R] load_audio_file: librosa.load of a PCM16 file (int16 / 32768) + np.pad to the target length,
    with the conversion and the padding done on the device."""
    hl = built
    rng = np.random.default_rng(77)
    y16 = rng.integers(-20000, 20000, size=(7, 22050)).astype(np.int16)
    yf = (y16.astype(np.float32) / np.float32(32768.0))
    padded = np.zeros((7, 66150), np.float32)
    padded[:, :22050] = yf
    ex = hl.FeatureExtractor(ref=np.max, n_mfcc=40)
    ref = ex.extract_host(padded)
    a = ex.extract_host(y16, pad_to=66150, chunk_clips=3)
    assert ex.last_transfer_bytes()[0] == y16.nbytes
    b = ex.extract_host(yf, pad_to=66150)
    assert ex.last_transfer_bytes()[0] == yf.nbytes
    for k in ("logmel", "mfcc", "stats", "status"):
        assert np.array_equal(a[k], ref[k]) and np.array_equal(b[k], ref[k]), k
    want = oracle_clip(padded[2], n_mfcc=40)
    assert_clip(compare_clip({k: a[k][2] for k in ("logmel", "mfcc", "stats")}, want), where="pcm16 clip 2")
    with pytest.raises(hl.ParameterError):
        ex.extract_host(y16, pad_to=100)


def test_chroma_stft_and_tuning_on_device(built):
    """librosa.feature.chroma_stft incl. estimate_tuning (piptrack -> median -> histogram).
    The tuning is an arg-max over a 100-bin histogram: where the oracle's own histogram has a near tie
    the device may pick the other bin; that is allowed only in that case."""
    import torch

    hl = built
    for n in (22050, 66150):
        y = hl.synth.synth_batch(20, n, seed=31)
        ex = hl.FeatureExtractor(n_mfcc=40, ref=np.max)
        out = ex.extract_device(torch.from_numpy(y).cuda(), chroma=True, pooled=True)
        ch, tu = out["chroma"].cpu().numpy(), out["tuning"].cpu().numpy()
        assert out["pooled"].shape == (20, 370)
        ties = 0
        for i in range(len(y)):
            S = np.abs(orc.stft(y[i])) ** 2
            t_or = orc.estimate_tuning(S=S, sr=SR, n_fft=2048, bins_per_octave=12)
            if abs(tu[i] - t_or) > 1e-6:
                p, m = orc.piptrack(S=S, sr=SR, n_fft=2048)
                sel = p[(m >= np.median(m[p > 0])) & (p > 0)]
                res = np.mod(12 * np.log2(sel / 27.5), 1.0)
                res[res >= 0.5] -= 1.0
                counts, edges = np.histogram(res, np.linspace(-0.5, 0.5, 101))
                got = int(round((tu[i] + 0.5) * 100))
                assert counts[got] == counts.max(), (f"tuning {tu[i]} vs {t_or}: the device's bin holds {counts[got]} "
                                                     f"candidates, the oracle's arg-max {counts.max()} (clip {i})")
                ties += 1
            want = orc.chroma_stft(y=y[i], sr=SR, tuning=float(tu[i]))
            assert ch[i].shape == want.shape == (12, 1 + n // 512)
            assert np.abs(ch[i] - want).max() <= 1e-4, f"chroma clip {i}"
            assert ch[i].max() <= 1.0 + 1e-6
        assert ties <= 2
        # pooled chroma columns = mean / std over frames
        po = out["pooled"].cpu().numpy()
        assert np.abs(po[:, 346:358] - ch.mean(-1)).max() <= 1e-5 and np.abs(po[:, 358:370] - ch.std(-1)).max() <= 1e-5
    # a workspace without room for the power-spectrum stash takes the recomputing kernel: same chroma
    import ctypes as C
    from hybrid_language_music_clustering_vae_b200._lib import lib
    yd = torch.from_numpy(y).cuda()
    B, T = len(y), 1 + n // 512
    big = int(lib.hlmc_chroma_workspace_bytes(ex._plan, B, n))
    small = big - B * T * (32 * 32 + 4) * 4
    bufs = {k: torch.empty(sh, dtype=dt, device="cuda") for k, sh, dt in (
        ("lm", (B, 128, T), torch.float32), ("mf", (B, 40, T), torch.float32), ("st", (B, 5, T), torch.float32),
        ("sta", (B,), torch.int32), ("cm", (B,), torch.float32), ("ch", (B, 12, T), torch.float32),
        ("tu", (B,), torch.float32), ("wk", (small,), torch.uint8))}
    pt = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.hlmc_extract_device_ex(ex._plan, pt(yd), B, n, n, pt(bufs["lm"]), pt(bufs["mf"]), pt(bufs["st"]),
                                    pt(bufs["sta"]), pt(bufs["cm"]), pt(bufs["ch"]), pt(bufs["tu"]), pt(bufs["wk"]),
                                    small, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0 and np.array_equal(bufs["tu"].cpu().numpy(), tu)
    assert np.abs(bufs["ch"].cpu().numpy() - ch).max() <= 2e-6
    # the host pipeline produces the same chroma (chunked, overlapped copies)
    hp = ex.extract_host(y, chroma=True, pooled=True, chunk_clips=7)
    assert np.array_equal(hp["chroma"], ch) and np.array_equal(hp["tuning"], tu) and np.array_equal(hp["pooled"], po)
    # pooled columns only: the fused dB + DCT + pooling kernel (one-pass shifted sums instead of two passes)
    hp2 = ex.extract_host(y, logmel=False, mfcc=False, stats=False, pooled=True, chroma="pooled")
    assert "chroma" not in hp2 and np.abs(hp2["pooled"] - po).max() <= 2e-5 * np.abs(po).max()
    assert np.array_equal(hp2["pooled"][:, 336:], po[:, 336:])          # statistics + chroma columns: same arithmetic
    c1 = hl.feature.chroma_stft(y=y[3], sr=SR, n_fft=2048, hop_length=512)
    assert c1.shape == (12, 130) and np.array_equal(c1, ch[3])


@pytest.mark.parametrize("n_fft,hop", [(1024, 256), (4096, 1024), (512, 128), (256, 100)])
def test_chroma_stft_for_other_n_fft(built, n_fft, hop):
    """chroma_stft + the tuning estimate for the n_fft values the scripts do not use (BASELINE configs[4] sweeps them):
    candidates and projection come from the shared-memory FFT kernel; same tolerances as the n_fft = 2048 path."""
    import torch

    hl = built
    y = hl.synth.synth_batch(6, 30000, seed=n_fft + 3)
    ex = hl.FeatureExtractor(n_fft=n_fft, hop_length=hop, n_mfcc=13)
    out = ex.extract_device(torch.from_numpy(y).cuda(), chroma=True, pooled=True)
    ch, tu = out["chroma"].cpu().numpy(), out["tuning"].cpu().numpy()
    ties = 0
    for b in range(len(y)):
        S = np.abs(orc.stft(y[b], n_fft=n_fft, hop_length=hop)) ** 2
        t_or = orc.estimate_tuning(S=S, sr=SR, n_fft=n_fft, bins_per_octave=12)
        if abs(tu[b] - t_or) > 1e-6:          # only an exact tie of the oracle's own histogram may move the arg-max
            p_, m_ = orc.piptrack(S=S, sr=SR, n_fft=n_fft)
            sel = p_[(m_ >= np.median(m_[p_ > 0])) & (p_ > 0)]
            res = np.mod(12 * np.log2(sel / 27.5), 1.0)
            res[res >= 0.5] -= 1.0
            counts, _edges = np.histogram(res, np.linspace(-0.5, 0.5, 101))
            assert counts[int(round((tu[b] + 0.5) * 100))] == counts.max(), (b, tu[b], t_or)
            ties += 1
        want = orc.chroma_stft(y=y[b], sr=SR, n_fft=n_fft, hop_length=hop, tuning=float(tu[b]))
        assert ch[b].shape == want.shape == (12, 1 + 30000 // hop)
        assert np.abs(ch[b] - want).max() <= 1e-4, (b, np.abs(ch[b] - want).max())
    assert ties <= 1
    # the pooled chroma columns are the mean / std of exactly these rows
    po = out["pooled"].cpu().numpy()
    assert np.allclose(po[:, -24:-12], ch.mean(axis=2), atol=1e-5) and np.allclose(po[:, -12:], ch.std(axis=2), atol=1e-5)
    c1 = hl.feature.chroma_stft(y=y[2], sr=SR, n_fft=n_fft, hop_length=hop)
    assert np.array_equal(c1, ch[2])


def test_empty_and_degenerate_batches(built):
    """Empty batch, empty clip, single sample, single clip: shapes and errors as librosa would give them."""
    import torch

    hl = built
    ex = hl.FeatureExtractor(ref=np.max, n_mfcc=40)
    out = ex.extract_device(torch.zeros((0, 22050), device="cuda"), pooled=True)
    assert out["logmel"].shape == (0, 128, 44) and out["mfcc"].shape == (0, 40, 44) and out["pooled"].shape == (0, 346)
    host = ex.extract_host(np.zeros((0, 22050), np.float32))
    assert host["logmel"].shape == (0, 128, 44) and host["status"].shape == (0,)
    with pytest.raises(hl.ParameterError):
        ex.extract_host(np.zeros((3, 0), np.float32))             # librosa: input too short
    with pytest.raises(hl.ParameterError):
        hl.FeatureExtractor(center=False).extract_host(np.zeros((2, 1000), np.float32))   # n < n_fft, uncentered
    one = ex.extract_host(np.full((1, 1), 0.5, np.float32))
    assert one["logmel"].shape == (1, 128, 1) and one["logmel"].max() == 0.0
    # a 1-D clip is a batch of one
    y = hl.synth.synth_batch(1, 5000, seed=2)[0]
    a = ex.extract_host(y)
    assert a["logmel"].shape == (1, 128, 10)


def test_device_standard_scaler_matches_sklearn(built, tmp_path):
    """StandardScaler over the (N, 131072) flattened mel images, statistics and transform on device."""
    import pickle
    import torch
    from sklearn.preprocessing import StandardScaler

    hl = built
    rng = np.random.default_rng(3)
    X = (rng.standard_normal((57, 4096)) * rng.uniform(0.1, 30, 4096) - 40.0).astype(np.float32)
    X[:, 7] = -80.0                      # a constant column (all frames on the top_db floor)
    ref = StandardScaler()
    Yr = ref.fit_transform(X)
    Yd, sc = hl.scaler.fit_transform_device(torch.from_numpy(X).cuda())
    assert np.allclose(sc.mean_, ref.mean_, rtol=1e-12, atol=1e-12) and np.allclose(sc.var_, ref.var_, rtol=1e-10, atol=1e-12)
    assert np.array_equal(sc.scale_ == 1.0, ref.scale_ == 1.0) and sc.n_samples_seen_ == 57
    assert np.abs(Yd.cpu().numpy() - Yr).max() <= 2e-6 * max(1.0, np.abs(Yr).max())
    assert np.abs(sc.transform(X) - Yr).max() <= 2e-6 * max(1.0, np.abs(Yr).max())     # a genuine sklearn object
    pickle.loads(pickle.dumps(sc))
    # shard combination (what the all-reduce computes) == global statistics
    parts = [X[:20], X[20:41], X[41:]]
    st = [hl.scaler.column_stats_device(torch.from_numpy(p).cuda()) for p in parts]
    n, mean, m2 = hl.scaler.combine_stats([len(p) for p in parts], [s[0].cpu().numpy() for s in st],
                                          [s[1].cpu().numpy() for s in st])
    assert n == 57 and np.allclose(mean, ref.mean_, rtol=1e-12, atol=1e-12) and np.allclose(m2 / n, ref.var_, rtol=1e-10, atol=1e-12)
    # through the processed_data2 writer
    mel = rng.uniform(-80, 0, size=(6, 128, 1024)).astype(np.float32)
    flat = rng.standard_normal((6, 290))
    a, _ = hl.preprocessing.save_processed_data2(str(tmp_path / "dev"), mel, flat, np.arange(6), device_scaler=0)
    b, _ = hl.preprocessing.save_processed_data2(str(tmp_path / "cpu"), mel, flat, np.arange(6))
    assert a.shape == (6, 128, 1024) and a.dtype == np.float32 and np.abs(a - b).max() <= 5e-6


@pytest.mark.parametrize("n_fft", [1024, 512, 4096])
def test_subwarp_register_fft_kernel(built, n_fft):
    """n_fft = 1024 / 512 run the frames_sub kernel (16 / 8 lanes per frame, 2 / 4 frames per warp) and
    n_fft = 4096 the two-pass frames_fast_4096 kernel (even / odd bins): the same cases the 2048 kernel
    is put through, plus agreement with the generic kernel."""
    import torch

    hl = built
    hop = n_fft // 4
    assert hl.FeatureExtractor(n_fft=n_fft, hop_length=hop).uses_fast_path()
    y = hl.synth.synth_batch(12, 20000, seed=n_fft + 1)
    for kw in (dict(), dict(pad_mode="reflect"), dict(pad_mode="edge"), dict(center=False), dict(hop_length=hop + 37),
               dict(n_mels=40, n_mfcc=13), dict(n_mels=64, htk=True), dict(win_length=n_fft // 2), dict(window="hamming")):
        kw = dict(dict(n_fft=n_fft, hop_length=hop, n_mfcc=20), **kw)
        _check_batch(hl, y, **kw)
    for n in (1, 2, n_fft // 2 - 1, n_fft - 1, n_fft, n_fft + 1, 3 * n_fft + 5):
        yy = (0.1 * np.random.default_rng(n).standard_normal((5, n))).astype(np.float32)
        res = _check_batch(hl, yy, n_fft=n_fft, hop_length=hop, n_mfcc=13, pad_mode="reflect")
        assert res["logmel"].shape[-1] == 1 + n // hop
    for extra in (1, 2, 3):
        _check_batch(hl, hl.synth.synth_batch(5, 6001, seed=4), pitch=6001 + extra, n_fft=n_fft, hop_length=hop, n_mfcc=13)
    # bitwise independence of batch composition, and agreement with the shared-memory FFT kernel
    ex = hl.FeatureExtractor(n_fft=n_fft, hop_length=hop, n_mfcc=20, ref=np.max)
    yd = torch.from_numpy(hl.synth.synth_batch(64, 30000, seed=9)).cuda()
    a = {k: v.clone() for k, v in ex.extract_device(yd).items()}
    idx = torch.tensor([63, 5, 17], device="cuda")
    c = ex.extract_device(yd[idx].contiguous())
    for k in ("logmel", "mfcc", "stats"):
        assert torch.equal(c[k], a[k][idx]), k
    ex.force_generic(True)
    gk = ex.extract_device(yd)
    assert (gk["logmel"] - a["logmel"]).abs().max() <= LOGMEL_TOL_DB     # two float32 FFTs, each within tolerance
    assert (gk["mfcc"] - a["mfcc"]).abs().max() <= 1e-4 * float(a["mfcc"].abs().max())
    bad = yd.clone()
    bad[3, 100] = float("nan")
    ex.force_generic(False)
    assert ex.extract_device(bad)["status"].cpu().tolist() == [0, 0, 0, 1] + [0] * 60


@pytest.mark.parametrize("kw", [dict(), dict(n_mels=40, n_mfcc=13), dict(n_mels=96, n_mfcc=0), dict(n_mels=33, n_mfcc=8),
                                dict(n_fft=1024, hop_length=256), dict(ref=1.0, top_db=None)])
def test_fused_pooled_only_path(built, kw):
    """hlmc_extract_pooled_device never writes log-mel / MFCC to HBM; its columns equal
    np.mean / np.std over frames of the full outputs (float32 accumulation, so 1e-5 of the column scale)."""
    import torch

    hl = built
    ex = hl.FeatureExtractor(**dict(dict(ref=np.max, n_mfcc=40), **kw))
    for n in (22050, 300, 200000):          # 44 frames, 1 frame, 391 frames (several 128-frame tiles)
        y = hl.synth.synth_batch(9, n, seed=n)
        yd = torch.from_numpy(y).cuda()
        full = {k: v.cpu().numpy().astype(np.float64) for k, v in ex.extract_device(yd).items()}
        got = ex.extract_pooled_device(yd)
        po = got["pooled"].cpu().numpy()
        parts = [full["logmel"].mean(-1), full["logmel"].std(-1)]
        if ex.n_mfcc > 0:
            parts += [full["mfcc"].mean(-1), full["mfcc"].std(-1)]
        st = full["stats"]
        parts.append(np.stack([st.mean(-1), st.std(-1)], axis=-1).reshape(len(y), 10))
        want = np.concatenate(parts, axis=1)
        assert po.shape == want.shape == (9, ex.pooled_width(ex.n_mfcc > 0))
        scale = np.maximum(np.abs(want).max(axis=0, keepdims=True), 1.0)
        assert np.abs(po - want).max() <= 2e-4 * scale.max() and (np.abs(po - want) / scale).max() <= 5e-5, kw
        assert np.array_equal(got["status"].cpu().numpy(), full["status"].astype(np.int32))
    # constant rows pool to exactly zero variance (all-zero clip: every band sits on the dB floor)
    z = ex.extract_pooled_device(torch.zeros((2, 22050), device="cuda"))["pooled"].cpu().numpy()
    nm = ex.n_mels
    assert np.all(z[:, nm:2 * nm] == 0.0)


def test_fused_pooling_at_the_largest_filterbank(built):
    """n_mels = 256 with 128 coefficients: the fused kernel's largest configuration (two bands per thread,
    64 KB DCT table) against db_dct + pool_kernel through the host pipeline."""
    hl = built
    ex = hl.FeatureExtractor(n_mels=256, n_mfcc=128, ref=np.max)
    y = hl.synth.synth_batch(5, 22050, seed=3)
    a = ex.extract_host(y, pooled=True)
    b = ex.extract_host(y, logmel=False, mfcc=False, stats=False, pooled=True)
    assert b["pooled"].shape == (5, 2 * 256 + 2 * 128 + 10)
    assert np.abs(a["pooled"] - b["pooled"]).max() <= 2e-5 * np.abs(a["pooled"]).max()


def test_cuda_graph_replay_equals_the_eager_call(built):
    """hlmc_graph_*: extract_device captured once, replayed on new audio written into the captured input buffer."""
    import torch

    hl = built
    ex = hl.FeatureExtractor(ref=np.max, n_mfcc=40)
    y0 = torch.from_numpy(hl.synth.synth_batch(3, 22050, seed=1)).cuda()
    g = ex.capture_device(y0)
    for seed in (2, 3):
        y = torch.from_numpy(hl.synth.synth_batch(3, 22050, seed=seed)).cuda()
        g.waves.copy_(y)
        got = {k: v.clone() for k, v in g.replay().items()}
        torch.cuda.synchronize()
        want = ex.extract_device(y)
        for k in ("logmel", "mfcc", "stats", "status"):
            assert torch.equal(got[k], want[k]), k
    g.close()
    with pytest.raises(hl.ParameterError):
        ex.capture_device(torch.zeros((0, 22050), device="cuda"))


def _random_cases(n_cases, seed):
    """Seeded random parameter combinations over every kernel family (the draw is fixed, so a failure reproduces)."""
    rng = np.random.default_rng(seed)
    cases = []
    for _ in range(n_cases):
        n_fft = int(rng.choice([512, 1024, 2048, 2048, 2048, 4096]))
        kw = dict(n_fft=n_fft,
                  hop_length=int(rng.choice([n_fft // 4, n_fft // 2, n_fft // 8, int(rng.integers(50, n_fft))])),
                  pad_mode=str(rng.choice(["constant", "reflect", "edge"])), center=bool(rng.random() < 0.8),
                  n_mels=int(rng.choice([128, 128, 40, 64, 96, 20])), htk=bool(rng.random() < 0.25),
                  power=float(rng.choice([2.0, 2.0, 1.0])), roll_percent=float(rng.choice([0.85, 0.5, 0.95])))
        kw["n_mfcc"] = int(min(kw["n_mels"], rng.choice([13, 20, 40])))
        if rng.random() < 0.3:
            kw["window"] = str(rng.choice(["hamming", "blackman"]))
        if rng.random() < 0.25:
            kw["fmin"], kw["fmax"] = float(rng.choice([0.0, 50.0, 300.0])), float(rng.choice([4000.0, 8000.0, 11025.0]))
        if rng.random() < 0.2:
            kw["win_length"] = n_fft // 2
        n = int(rng.choice([n_fft + int(rng.integers(0, 3 * n_fft)), int(rng.integers(2 * n_fft, 40000))]))
        cases.append((n, int(rng.integers(0, 4)), kw))
    return cases


@pytest.mark.parametrize("n,pitch_extra,kw", _random_cases(36, seed=20261018))
def test_random_parameter_combinations(built, n, pitch_extra, kw):
    """Every kernel family against the oracle on seeded random combinations of the librosa parameters, clip
    lengths and row pitches (mis-aligned rows take the hand-built staging path)."""
    y = built.synth.synth_batch(5, n, seed=n)
    _check_batch(built, y, pitch=n + pitch_extra, **kw)


@pytest.mark.parametrize("kw", [dict(), dict(window="hamming"), dict(n_mels=40, n_mfcc=13), dict(n_mels=64, htk=True),
                                dict(n_mels=256), dict(pad_mode="reflect", power=1.0), dict(win_length=1024),
                                dict(n_fft=1024, hop_length=256), dict(n_fft=512, hop_length=128),
                                dict(n_fft=1024, hop_length=300, window="hamming", n_mels=40, n_mfcc=13),
                                dict(n_fft=512, hop_length=128, n_mels=256, pad_mode="edge"),
                                dict(n_fft=4096, hop_length=1024), dict(n_fft=4096, hop_length=1000, n_mels=64, pad_mode="reflect")])
def test_tensor_memory_tables_match_shared_memory_tables(built, kw):
    """The n_fft = 2048 kernel reads its per-lane tables (window, twiddles, banded mel weights, gather offsets) from
    Tensor Memory (tcgen05.ld); HLMC_PATH_FAST_SMEM_TABLES runs the same pipeline with the tables in shared memory.
    The two differ only in how table values are rounded (the shared-memory variant synthesises the Hann window and
    rotates one split twiddle per lane in registers, the TMEM variant reads values rounded once from float64), i.e.
    they are two float32 FFTs that each meet the oracle tolerances; against each other they are held to the same."""
    import torch

    hl = built
    y = np.concatenate([hl.synth.synth_batch(24, 30011, seed=77), np.zeros((1, 30011), np.float32)])   # odd length: edge frames too
    yd = torch.from_numpy(y).cuda()
    ex = hl.FeatureExtractor(n_mfcc=kw.get("n_mfcc", 20), ref=np.max, **{k: v for k, v in kw.items() if k != "n_mfcc"})
    a = {k: v.clone() for k, v in ex.extract_device(yd).items()}
    ex.set_path(2)
    b = ex.extract_device(yd)
    assert (a["logmel"] - b["logmel"]).abs().max().item() <= LOGMEL_TOL_DB
    for k in ("mfcc", "stats"):
        for c in range(a[k].shape[0]):
            d = (a[k][c] - b[k][c]).abs().amax(dim=-1)
            if k == "mfcc":             # 1e-4 of the clip's largest coefficient, as against the oracle
                assert d.max().item() <= REL_TOL * max(b[k][c].abs().max().item(), 1e-6), (k, c, d.tolist())
            else:                       # per statistic; rolloff (row 2) may flip by one bin at a tie; zcr (row 3) is exact
                assert torch.equal(a[k][c][3], b[k][c][3])
                scale = b[k][c].abs().amax(dim=-1).clamp_min(1e-6)
                keep = torch.tensor([0, 1, 4], device=d.device)
                assert (d[keep] <= REL_TOL * scale[keep]).all(), (k, c, d.tolist(), scale.tolist())
    # the chroma variant (piptrack epilogue + power-spectrum stash) on both table paths
    assert ex.uses_fast_path()
    if not kw:
        ex.set_path(0)
        ca = {k: v.clone() for k, v in ex.extract_device(yd, chroma=True).items()}
        ex.set_path(2)
        cb = ex.extract_device(yd, chroma=True)
        assert torch.equal(ca["tuning"], cb["tuning"])
        assert (ca["chroma"] - cb["chroma"]).abs().max().item() <= 1e-5


def test_caller_supplied_filterbanks(built):
    """hlmc_plan_create(mel_basis=...): a caller's own filterbank replaces librosa.filters.mel.  A banded one (wider
    triangles than Slaney's) still fits the Tensor-Memory tables or the shared-memory ones; a fully dense one fits
    neither and runs on the shared-memory FFT kernel.  Both must equal basis @ |STFT|^2 of the oracle."""
    import torch

    hl = built
    rng = np.random.default_rng(12)
    y = hl.synth.synth_batch(4, 20000, seed=8)
    F = 1025
    tri = np.zeros((24, F), np.float32)
    for m in range(24):                                  # overlapping triangles 120 bins wide
        c, hw = 40 * m + 30, 60
        k = np.arange(max(0, c - hw), min(F, c + hw + 1))
        tri[m, k] = 1.0 - np.abs(k - c) / float(hw + 1)
    dense = rng.uniform(0.0, 1.0, (16, F)).astype(np.float32)
    for name, basis in (("banded", tri), ("dense", dense)):
        ex = hl.FeatureExtractor(n_mels=basis.shape[0], n_mfcc=8, ref=np.max, mel_basis=basis)
        assert np.array_equal(ex.mel_basis(), basis)
        out = ex.extract_device(torch.from_numpy(y).cuda())
        lm = out["logmel"].cpu().numpy()
        for b in range(len(y)):
            S = np.abs(orc.stft(y[b])) ** 2
            want = orc.power_to_db(basis.astype(np.float64) @ S, ref=np.max)
            assert np.abs(lm[b] - want).max() <= LOGMEL_TOL_DB, (name, b, np.abs(lm[b] - want).max())


def test_small_batch_db_kernel_is_bitwise_the_large_batch_one(built):
    """Batches up to 24,000 frames run db_dct_small (four lanes per frame), larger ones db_dct (a thread per two
    frames): same table, same summation order - log-mel and MFCC must not depend on the batch size, bit for bit."""
    import torch

    hl = built
    for kw in (dict(n_mfcc=40), dict(n_mfcc=13, n_mels=40), dict(n_mfcc=20, n_mels=127, ref=1.0, top_db=None)):
        ref = kw.pop("ref", np.max)
        top_db = kw.pop("top_db", 80.0)
        ex = hl.FeatureExtractor(ref=ref, top_db=top_db, **kw)
        y = torch.from_numpy(hl.synth.synth_batch(8, 22050, seed=90)).cuda()      # 44 frames per clip
        small = {k: v.clone() for k, v in ex.extract_device(y).items()}            # 352 frames
        big = ex.extract_device(torch.cat([y] * 80, dim=0))                         # 28,160 frames
        for k in ("logmel", "mfcc", "stats"):
            assert torch.equal(big[k][:8], small[k]) and torch.equal(big[k][-8:], small[k]), (kw, k)
