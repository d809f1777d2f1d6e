"""Runnable drop-ins for the two preprocessing scripts: dataset folders + metadata CSV in,
``processed_data1/`` or ``processed_data2/`` out, with every numerical step on the GPU.

    python -m hybrid_language_music_clustering_vae_b200.pipeline basic    --bangla DIR --english DIR --metadata CSV --out DIR
    python -m hybrid_language_music_clustering_vae_b200.pipeline advanced --bangla DIR --english DIR --metadata CSV --out DIR

What stays on the host is what is not arithmetic: walking the folders, the metadata look-ups
([R] src/1_preprocessing.py:157-219, src/1_preprocessing_advanced.py:190-266), parsing the WAV containers, and
writing the files.  The per-file loop of the basic script ([R] :223-258) and the joblib map of the advanced one
([R] :286-314) become batched device calls.  The lyrics embedding of the advanced script
(sentence-transformers, [R] :320-343) is not part of this path: pass ``lyrics_embedder`` (texts -> (N, 768) array)
or the file ``lyrics_embeddings.npy`` is not written, and the run says so.
"""
from __future__ import annotations

import argparse
import logging
import os

import numpy as np

from . import preprocessing as pp

log = logging.getLogger(__name__)
_NO_LYRICS = {"nan", "none", "null", "instrumental", "", " "}


def collect_audio_files(lang_dirs, metadata_csv, max_samples_per_class, advanced=False):
    """The scripts' ``collect_audio_files``: ``lang_dirs`` = [(folder, language code), ...]; genre (and lyrics) come
    from the metadata CSV by file id, never from the folder name.  ``advanced`` adds the advanced script's filters
    (no jazz, lyrics present and at least 15 characters).  Returns (files, skipped-by-reason)."""
    import pandas as pd

    meta = pd.read_csv(metadata_csv)
    ids = meta["ID"].astype(str)
    genre_of = dict(zip(ids, meta["genre"]))
    lyrics_of = dict(zip(ids, meta["lyrics"].fillna(""))) if "lyrics" in meta else {}
    files, skipped = [], {"not_in_metadata": 0, "jazz_excluded": 0, "empty_lyrics": 0, "short_lyrics": 0}
    for folder, lang in lang_dirs:
        if not folder or not os.path.exists(folder):
            continue
        for genre_folder in os.listdir(folder):
            gdir = os.path.join(folder, genre_folder)
            if not os.path.isdir(gdir):
                continue
            wavs = [f for f in os.listdir(gdir) if f.endswith(".wav")][:max_samples_per_class]
            for name in wavs:
                fid = os.path.splitext(name)[0]
                if fid not in genre_of:
                    skipped["not_in_metadata"] += 1
                    continue
                info = {"path": os.path.join(gdir, name), "language": lang, "genre": genre_of[fid],
                        "filename": name, "file_id": fid}
                if advanced:
                    if str(info["genre"]).strip().lower() == "jazz":
                        skipped["jazz_excluded"] += 1
                        continue
                    text = lyrics_of.get(fid, "")
                    if not isinstance(text, str) or text.strip().lower() in _NO_LYRICS:
                        skipped["empty_lyrics"] += 1
                        continue
                    if len(text.strip()) < 15:
                        skipped["short_lyrics"] += 1
                        continue
                    info["lyrics"] = text
                files.append(info)
    return files, skipped


def run_basic(audio_files, out_dir, cfg=pp.BASIC_CONFIG, device=0, batch_files=256, normalise_on_device=False,
              chroma="device"):
    """[R] src/1_preprocessing.py cells 6-9 -> ``processed_data1/``.  Returns a summary dict."""
    import pandas as pd

    rows, labels, meta, failed = [], [], [], []
    for lo in range(0, len(audio_files), batch_files):
        chunk = audio_files[lo:lo + batch_files]
        feats, ok, errors = pp.process_files_basic(chunk, cfg, chroma, device)
        for info, f, good, err in zip(chunk, feats, ok, errors):
            if good:
                rows.append(f)
                labels.append(info["genre"])
                meta.append({"language": info["language"], "genre": info["genre"], "filename": info["filename"]})
            else:
                failed.append((info["path"], err))
    features = np.array(rows) if rows else np.zeros((0, 2 * cfg["n_mels"] + 2 * cfg["n_mfcc"] + 34))
    df = pd.DataFrame(meta)
    df["label"] = labels
    pp.save_processed_data1(out_dir, features, labels, df, config=cfg, device=device if normalise_on_device else None)
    return {"processed": len(rows), "failed": failed, "features_shape": tuple(features.shape)}


def run_advanced(audio_files, out_dir, cfg=pp.ADV_CONFIG, device=0, batch_files=64, lyrics_embedder=None,
                 normalise_on_device=False, chroma="device"):
    """[R] src/1_preprocessing_advanced.py cells 5-9 -> ``processed_data2/``.  Returns a summary dict."""
    import pandas as pd

    mels, flats, labels, lyrics, meta, failed = [], [], [], [], [], 0
    for lo in range(0, len(audio_files), batch_files):
        for res in pp.process_files_advanced(audio_files[lo:lo + batch_files], cfg, chroma, device):
            if res["status"] != "success":
                failed += 1
                continue
            mels.append(res["mel_spec"])
            flats.append(res["flat_feat"])
            labels.append(res["genre"])
            lyrics.append(res["lyrics"])
            meta.append({"language": res["language"], "genre": res["genre"], "filename": res["filename"],
                         "file_id": res["file_id"]})
    mel = np.array(mels) if mels else np.zeros((0, cfg["n_mels"], cfg["fixed_time_steps"]), np.float32)
    flat = np.array(flats) if flats else np.zeros((0, 2 * cfg["n_mels"] + 34))
    emb = None
    if lyrics_embedder is not None:
        texts = [str(t) if t and len(str(t)) > 0 else " " for t in lyrics]
        emb = np.asarray(lyrics_embedder(texts))
        assert len(emb) == len(mel), "Mismatch between audio and lyrics samples!"
    else:
        log.warning("no lyrics_embedder given: lyrics_embeddings.npy is NOT written (the sentence-transformers "
                    "encoder of the reference is outside this path)")
    df = pd.DataFrame(meta)
    df["label"] = labels
    pp.save_processed_data2(out_dir, mel, flat, np.array(labels), emb, df, config=cfg,
                            device_scaler=device if normalise_on_device else None)
    return {"processed": len(mels), "failed": failed, "mel_shape": tuple(mel.shape), "flat_shape": tuple(flat.shape),
            "lyrics_embeddings_written": emb is not None}


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__.split("\n\n")[0])
    ap.add_argument("script", choices=["basic", "advanced"])
    ap.add_argument("--bangla", default=None)
    ap.add_argument("--english", default=None)
    ap.add_argument("--metadata", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--duration", type=int, default=None, help="override CONFIG['duration'] (seconds)")
    ap.add_argument("--normalise-on-device", action="store_true")
    args = ap.parse_args(argv)
    cfg = dict(pp.BASIC_CONFIG if args.script == "basic" else pp.ADV_CONFIG)
    if args.duration:
        cfg["duration"] = args.duration
    files, skipped = collect_audio_files([(args.bangla, "bn"), (args.english, "en")], args.metadata,
                                         cfg["max_samples_per_class"], advanced=args.script == "advanced")
    print(f"Total audio files collected: {len(files)}; skipped: {skipped}")
    if not files:
        raise ValueError("No audio files collected! Check paths and metadata.")
    run = run_basic if args.script == "basic" else run_advanced
    summary = run(files, args.out, cfg, device=args.device, normalise_on_device=args.normalise_on_device)
    print(summary)
    return summary


if __name__ == "__main__":
    main()
