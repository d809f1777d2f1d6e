"""Pins the oracle (and through it the CUDA path) to REAL librosa - for whoever has librosa.

librosa is not installable in the build image (no network, no wheel), so parity is "unpinned" there
(DESIGN.md section 2).  Run this once on any machine with librosa >= 0.10:

    pip install librosa
    python tests/golden/make_librosa_golden.py          # writes tests/golden/librosa_v1.npz

and commit the file.  tests/test_librosa_golden.py then checks the oracle (CPU suite) and the CUDA path
(GPU suite) against it at north_star's tolerances; without the file those tests skip.  The script itself needs
only numpy + librosa: it reads the seeded clips from golden_v1.npz / torchaudio_v1.npz (already committed), so
the inputs are bit-identical everywhere.  It exercises every row of SURVEY.md 8(a): stft, melspectrogram,
power_to_db, mfcc, spectral_centroid / bandwidth / rolloff, zero_crossing_rate, rms, chroma_stft (incl.
estimate_tuning), the scripts' crop / pad to 1024 frames and their pooled 370 / 290 vectors, and
librosa.resample(res_type="polyphase") plus librosa.load's default soxr_hq for the front end.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SR = 22050


def clips():
    g = np.load(os.path.join(HERE, "golden_v1.npz"))
    t = np.load(os.path.join(HERE, "torchaudio_v1.npz"))
    return {"a": g["a_y"], "b": g["b_y"], "t": t["y"]}


KW = {
    "a": dict(n_fft=2048, hop_length=512, n_mels=128, n_mfcc=40, pad_mode="constant"),
    "b": dict(n_fft=512, hop_length=128, n_mels=40, n_mfcc=13, pad_mode="reflect"),
    "t": dict(n_fft=2048, hop_length=512, n_mels=128, n_mfcc=40, pad_mode="constant"),
}


def main():
    import librosa

    out = {"librosa_version": np.array(librosa.__version__), "numpy_version": np.array(np.__version__)}
    for case, ys in clips().items():
        kw = KW[case]
        stft_kw = dict(n_fft=kw["n_fft"], hop_length=kw["hop_length"], pad_mode=kw["pad_mode"])
        acc = {k: [] for k in ("stft", "mel", "logmel", "logmel_ref1", "mfcc", "centroid", "bandwidth", "rolloff",
                               "zcr", "rms", "chroma", "tuning", "all370", "flat290", "fixed1024")}
        for y in ys:
            y = np.ascontiguousarray(y, dtype=np.float32)
            acc["stft"].append(librosa.stft(y, **stft_kw))
            mel = librosa.feature.melspectrogram(y=y, sr=SR, n_mels=kw["n_mels"], **stft_kw)
            acc["mel"].append(mel)
            acc["logmel"].append(librosa.power_to_db(mel, ref=np.max))
            acc["logmel_ref1"].append(librosa.power_to_db(mel))
            acc["mfcc"].append(librosa.feature.mfcc(y=y, sr=SR, n_mfcc=kw["n_mfcc"], n_mels=kw["n_mels"], **stft_kw))
            acc["centroid"].append(librosa.feature.spectral_centroid(y=y, sr=SR, **stft_kw))
            acc["bandwidth"].append(librosa.feature.spectral_bandwidth(y=y, sr=SR, **stft_kw))
            acc["rolloff"].append(librosa.feature.spectral_rolloff(y=y, sr=SR, **stft_kw))
            acc["zcr"].append(librosa.feature.zero_crossing_rate(y, frame_length=kw["n_fft"], hop_length=kw["hop_length"]))
            acc["rms"].append(librosa.feature.rms(y=y, frame_length=kw["n_fft"], hop_length=kw["hop_length"],
                                                  pad_mode=kw["pad_mode"]))
            if kw["n_fft"] == 2048 and kw["pad_mode"] == "constant":
                S = np.abs(librosa.stft(y, **stft_kw)) ** 2
                acc["tuning"].append(librosa.estimate_tuning(S=S, sr=SR, n_fft=2048))
                acc["chroma"].append(librosa.feature.chroma_stft(y=y, sr=SR, n_fft=2048, hop_length=512))
                # the scripts' own vectors ([R] src/1_preprocessing.py:105-129, _advanced.py:97-156)
                lm = librosa.power_to_db(librosa.feature.melspectrogram(y=y, sr=SR, n_mels=128, n_fft=2048,
                                                                        hop_length=512), ref=np.max)
                mf = librosa.feature.mfcc(y=y, sr=SR, n_mfcc=40, n_fft=2048, hop_length=512)
                five = [librosa.feature.spectral_centroid(y=y, sr=SR, hop_length=512),
                        librosa.feature.spectral_bandwidth(y=y, sr=SR, hop_length=512),
                        librosa.feature.spectral_rolloff(y=y, sr=SR, hop_length=512),
                        librosa.feature.zero_crossing_rate(y, hop_length=512),
                        librosa.feature.rms(y=y, hop_length=512)]
                ch = acc["chroma"][-1]
                f = list(np.mean(lm, axis=1)) + list(np.std(lm, axis=1)) + list(np.mean(mf, axis=1)) + list(np.std(mf, axis=1))
                g = list(np.mean(lm, axis=1)) + list(np.std(lm, axis=1))
                for x in five:
                    f += [np.mean(x), np.std(x)]
                    g += [np.mean(x), np.std(x)]
                tail = list(np.mean(ch, axis=1)) + list(np.std(ch, axis=1))
                acc["all370"].append(np.array(f + tail))
                acc["flat290"].append(np.array(g + tail))
                fixed = lm[:, :1024] if lm.shape[1] > 1024 else np.pad(
                    lm, ((0, 0), (0, 1024 - lm.shape[1])), mode="constant", constant_values=lm.min())
                acc["fixed1024"].append(fixed)
        for k, v in acc.items():
            if v:
                out[f"{case}_{k}"] = np.stack(v)
    # front end: resampling of a seeded 44.1 kHz / 48 kHz signal
    rng = np.random.default_rng(99)
    for sr_in in (44100, 48000, 16000):
        x = (0.2 * rng.standard_normal(sr_in // 2)).astype(np.float32)
        out[f"resample_in_{sr_in}"] = x
        out[f"resample_polyphase_{sr_in}"] = librosa.resample(x, orig_sr=sr_in, target_sr=SR, res_type="polyphase")
        try:
            out[f"resample_soxr_hq_{sr_in}"] = librosa.resample(x, orig_sr=sr_in, target_sr=SR, res_type="soxr_hq")
        except Exception as e:      # soxr missing: not fatal
            print("soxr_hq unavailable:", e, file=sys.stderr)
    path = os.path.join(HERE, "librosa_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with librosa", librosa.__version__)


if __name__ == "__main__":
    main()
