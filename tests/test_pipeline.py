"""The runnable drop-ins of the two preprocessing scripts (`pipeline.py`) on a small synthetic dataset:
folders of WAV files + a metadata CSV in, processed_data1/ and processed_data2/ out."""
import os
import pickle
import wave

import numpy as np
import pytest

from oracle import librosa_oracle as orc

SR = 22050


def _tone(n, sr, f0, seed):
    rng = np.random.default_rng(seed)
    t = np.arange(n) / sr
    x = 0.3 * np.sin(2 * np.pi * f0 * t) + 0.1 * np.sin(2 * np.pi * 3.1 * f0 * t + 1.0) + 0.02 * rng.standard_normal(n)
    return np.clip(np.rint(x * 32768.0), -32768, 32767).astype(np.int16)


def _write(path, frames, sr):
    with wave.open(str(path), "wb") as w:
        w.setnchannels(frames.shape[1])
        w.setsampwidth(2)
        w.setframerate(sr)
        w.writeframes(np.ascontiguousarray(frames, dtype="<i2").tobytes())


@pytest.fixture()
def dataset(tmp_path):
    import pandas as pd

    rows, raw = [], {}
    layout = {"Bangla_Datasets": {"folk": [("b1", 22050, 1), ("b2", 44100, 2)], "rock": [("b3", 22050, 1)]},
              "English_Datasets": {"pop": [("e1", 48000, 1), ("e2", 22050, 2)], "jazz": [("e3", 22050, 1)]}}
    for top, genres in layout.items():
        for genre, files in genres.items():
            d = tmp_path / top / genre
            d.mkdir(parents=True)
            for k, (fid, sr, ch) in enumerate(files):
                n = int((0.9 + 0.4 * k) * sr)
                fr = np.stack([_tone(n, sr, 180.0 * (1 + len(rows)) + 40 * c, len(rows) + c) for c in range(ch)], axis=1)
                _write(d / f"{fid}.wav", fr, sr)
                raw[fid] = (fr, sr)
                lyrics = {"b3": "short", "e2": float("nan")}.get(fid, "la la la a long enough line of lyrics")
                # the genre comes from the metadata, not from the folder name
                rows.append({"ID": fid, "genre": "jazz" if genre == "jazz" else genre.upper(), "lyrics": lyrics})
            (d / "notes.txt").write_text("not audio")
            (d / "zz_unlisted.wav").write_bytes(b"RIFF")          # not in the metadata: skipped before it is opened
    meta = tmp_path / "updated_metadata.csv"
    pd.DataFrame(rows).to_csv(meta, index=False)
    return tmp_path, meta, raw


def test_collect_audio_files_follows_the_scripts_filters(built, dataset):
    root, meta, _raw = dataset
    pl = built.pipeline
    dirs = [(str(root / "Bangla_Datasets"), "bn"), (str(root / "English_Datasets"), "en")]
    basic, skipped = pl.collect_audio_files(dirs, str(meta), 160)
    assert sorted(f["file_id"] for f in basic) == ["b1", "b2", "b3", "e1", "e2", "e3"] and skipped["not_in_metadata"] == 4
    assert {f["file_id"]: f["genre"] for f in basic}["b1"] == "FOLK" and all("lyrics" not in f for f in basic)
    adv, skipped = pl.collect_audio_files(dirs, str(meta), 200, advanced=True)
    assert sorted(f["file_id"] for f in adv) == ["b1", "b2", "e1"]
    assert skipped == {"not_in_metadata": 4, "jazz_excluded": 1, "empty_lyrics": 1, "short_lyrics": 1}
    one, _ = pl.collect_audio_files(dirs, str(meta), 1)
    assert len(one) <= 4                                            # max_samples_per_class caps each genre folder


@pytest.mark.gpu
def test_basic_and_advanced_drop_ins_end_to_end(built, dataset, tmp_path):
    import pandas as pd

    root, meta, raw = dataset
    pl, pp = built.pipeline, built.preprocessing
    out1, out2 = tmp_path / "processed_data1", tmp_path / "processed_data2"
    args = ["--bangla", str(root / "Bangla_Datasets"), "--english", str(root / "English_Datasets"),
            "--metadata", str(meta), "--duration", "2"]
    s1 = pl.main(["basic"] + args + ["--out", str(out1)])
    assert s1["processed"] == 6 and s1["failed"] == [] and s1["features_shape"] == (6, 370)
    assert sorted(os.listdir(out1)) == sorted(["features_raw.npy", "features_normalized.npy", "labels.npy",
                                               "metadata.csv", "scaler.pkl", "imputer.pkl", "config.pkl"])
    feats = np.load(out1 / "features_raw.npy")
    md = pd.read_csv(out1 / "metadata.csv")
    assert list(md.columns) == ["language", "genre", "filename", "label"] and len(md) == 6
    cfg = dict(pp.BASIC_CONFIG, duration=2)
    for row, name in zip(feats, md["filename"]):
        fr, sr = raw[os.path.splitext(name)[0]]
        audio, _ = orc.load_audio_file_pcm16(fr, sr, cfg)
        want = orc.extract_all_features(audio, SR, cfg)
        assert np.abs(row[:256] - want[:256]).max() <= 0.011
        assert np.abs(row[256:336] - want[256:336]).max() <= 2e-4 * np.abs(want[256:336]).max()
        assert np.allclose(row[336:346], want[336:346], rtol=2e-4, atol=SR / 2048 / 40)
        assert np.abs(row[346:] - want[346:]).max() <= 2e-4
    with open(out1 / "config.pkl", "rb") as f:
        assert pickle.load(f)["duration"] == 2
    norm = np.load(out1 / "features_normalized.npy")
    assert norm.shape == (6, 370) and np.abs(norm.mean(0)).max() < 1e-9

    s2 = pl.main(["advanced"] + args + ["--out", str(out2), "--normalise-on-device"])
    assert s2["processed"] == 3 and s2["mel_shape"] == (3, 128, 1024) and s2["flat_shape"] == (3, 290)
    assert not s2["lyrics_embeddings_written"] and "lyrics_embeddings.npy" not in os.listdir(out2)
    mel = np.load(out2 / "mel_spectrograms_raw.npy")
    md2 = pd.read_csv(out2 / "metadata.csv")
    assert list(md2.columns) == ["language", "genre", "filename", "file_id", "label"]
    cfg2 = dict(pp.ADV_CONFIG, duration=2)
    for img, fid in zip(mel, md2["file_id"]):
        fr, sr = raw[str(fid)]
        audio, _ = orc.load_audio_file_pcm16(fr, sr, cfg2)
        assert np.abs(img - orc.adv_extract_mel_spectrogram(audio, SR, cfg2)).max() <= 0.011
    # with an embedder the file appears, one row per kept clip
    s3 = pl.run_advanced(pl.collect_audio_files([(str(root / "Bangla_Datasets"), "bn")], str(meta), 200, advanced=True)[0],
                         str(tmp_path / "pd2b"), cfg2, lyrics_embedder=lambda texts: np.ones((len(texts), 768), np.float32))
    assert s3["lyrics_embeddings_written"] and np.load(tmp_path / "pd2b" / "lyrics_embeddings.npy").shape == (2, 768)
