"""Imports the UNMODIFIED reference modules staged under baseline/_ref/ (oracle/build_ref.py) — TEST / BASELINE
INFRASTRUCTURE ONLY: used by bench.py's baseline legs (--impl reference, cpu_baseline, gpu_eager_baseline) and by
tests; never by e2e_tts_b200/.  Returns None when the files are not there (the callers then fall back to the oracle
port and say `kind: "port"`)."""
from __future__ import annotations

import os
import sys
import types
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_hifigan_class():
    """e2e_tts/models/vocoder/generator.py:13 `HifiGan`, imported the way e2e_tts/src/api/inference.py:7 exposes it
    (package `vocoder` on sys.path)."""
    if not os.path.exists(os.path.join(REF_DIR, "vocoder", "generator.py")):
        return None
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from vocoder.generator import HifiGan  # noqa
    return HifiGan


def build_reference_hifigan(config: dict, state_dict: dict):
    """Constructs the reference generator exactly as e2e_tts/src/api/utils.py:53-56 does (weight-norm hooks left on)."""
    cls = reference_hifigan_class()
    if cls is None:
        return None
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = cls(config)
        m.load_state_dict(state_dict)
    return m.eval()


def reference_stft_class():
    """e2e_tts/src/tools/stft.py:11 `TorchSTFT` with stub modules for the third-party imports that are not installed in
    this image (librosa, parselmouth, pyworld) and the unrelated `models.g2p`; librosa.filters.mel is bound to the
    restated librosa-0.9.2 algorithm (oracle/mel_oracle.py; basis parity unpinned, see that header)."""
    if not os.path.exists(os.path.join(REF_DIR, "tools", "stft.py")):
        return None
    from . import mel_oracle as mo
    if "librosa" not in sys.modules:
        lib = types.ModuleType("librosa")
        lib.filters = types.ModuleType("librosa.filters")
        lib.filters.mel = lambda sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw: mo.slaney_mel_basis(
            sr, n_fft, n_mels, fmin, fmax)
        lib.util = types.ModuleType("librosa.util")
        lib.util.normalize = lambda x, **kw: x
        sys.modules["librosa"] = lib
        sys.modules["librosa.filters"] = lib.filters
        sys.modules["librosa.util"] = lib.util
    for name in ("parselmouth", "pyworld"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if "models" not in sys.modules:
        models = types.ModuleType("models")
        g2p = types.ModuleType("models.g2p")
        g2p._symbols_to_sequence = lambda s: []
        models.g2p = g2p
        sys.modules["models"] = models
        sys.modules["models.g2p"] = g2p
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from tools.stft import TorchSTFT  # noqa
    return TorchSTFT
