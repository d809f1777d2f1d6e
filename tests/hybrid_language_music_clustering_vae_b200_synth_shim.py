"""Loads the package's synth module by path so golden generation works without the CUDA library."""
import importlib.util
import os

_p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                  "hybrid_language_music_clustering_vae_b200", "synth.py")
_spec = importlib.util.spec_from_file_location("_hlmc_synth", _p)
_m = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_m)
synth_clip, synth_batch, mixture_kinds = _m.synth_clip, _m.synth_batch, _m.mixture_kinds
