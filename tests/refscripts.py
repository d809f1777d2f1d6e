"""Run the REFERENCE's own code (read from /root/reference at test time, never copied into the repo)
against a stand-in ``librosa``: function bodies are taken out of the two preprocessing scripts with ``ast``
and executed with the module-level name ``librosa`` bound to the oracle (CPU tests) or to the product package
(GPU tests).  Absent reference (the GPU box) -> the callers skip."""
from __future__ import annotations

import ast
import os
import types

import numpy as np

REF_SRC = "/root/reference/src"
BASIC = os.path.join(REF_SRC, "1_preprocessing.py")
ADVANCED = os.path.join(REF_SRC, "1_preprocessing_advanced.py")


def available() -> bool:
    return os.path.exists(BASIC) and os.path.exists(ADVANCED)


def load_functions(path, names, librosa):
    """-> namespace holding CONFIG and the named functions of the script at `path`, with `librosa` as given."""
    tree = ast.parse(open(path).read(), filename=path)
    keep = []
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            keep.append(node)
        elif (isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name)
              and node.targets[0].id == "CONFIG"):
            keep.append(node)
    found = {n.name for n in keep if isinstance(n, ast.FunctionDef)}
    missing = set(names) - found
    assert not missing, f"{path}: no function(s) {sorted(missing)}"
    ns = {"np": np, "librosa": librosa, "__name__": "reference_extract"}
    exec(compile(ast.Module(body=keep, type_ignores=[]), path, "exec"), ns)
    return types.SimpleNamespace(**{k: ns[k] for k in list(names) + ["CONFIG"]})


def top_level_block(path, first_marker, last_marker):
    """Compiled top-level statements of the script from the first one whose source contains `first_marker` to
    the last one containing `last_marker` (print calls dropped): the scripts' normalise-and-save cells."""
    src = open(path).read()
    tree = ast.parse(src, filename=path)
    segs = [(node, ast.get_source_segment(src, node) or "") for node in tree.body]
    i0 = next(i for i, (_n, s) in enumerate(segs) if first_marker in s)
    i1 = max(i for i, (_n, s) in enumerate(segs) if last_marker in s)
    body = []
    for node, _s in segs[i0:i1 + 1]:
        if (isinstance(node, ast.Expr) and isinstance(node.value, ast.Call)
                and getattr(node.value.func, "id", "") == "print"):
            continue
        body.append(node)
    return compile(ast.Module(body=body, type_ignores=[]), path, "exec")


def oracle_librosa():
    """The oracle dressed as the ``librosa`` module the scripts import."""
    from oracle import librosa_oracle as orc

    feature = types.SimpleNamespace(
        melspectrogram=orc.melspectrogram, mfcc=orc.mfcc, spectral_centroid=orc.spectral_centroid,
        spectral_bandwidth=orc.spectral_bandwidth, spectral_rolloff=orc.spectral_rolloff,
        zero_crossing_rate=orc.zero_crossing_rate, rms=orc.rms, chroma_stft=orc.chroma_stft)
    return types.SimpleNamespace(feature=feature, power_to_db=orc.power_to_db, stft=orc.stft)


BASIC_FUNCS = ("extract_mel_spectrogram", "extract_mfcc", "extract_spectral_features", "extract_chroma_features",
               "extract_all_features")
ADV_FUNCS = ("extract_mel_spectrogram", "extract_flattened_features")
