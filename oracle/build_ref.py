"""Stages the UNMODIFIED reference modules of the hot path under baseline/_ref/ (git-ignored, NOT gpurun-ignored:
the files travel to the GPU box with the snapshot, the history stays free of reference sources).

    python -m oracle.build_ref          # run in the authoring container, where /root/reference exists

Copied verbatim (byte for byte; checked by sha256 after the copy):
    e2e_tts/models/vocoder/{__init__,generator,layers,function,discriminator,loss}.py -> baseline/_ref/vocoder/
        (generator.py:13-62 HifiGan is the class bench.py --impl reference and the gpu_eager_baseline leg time;
         __init__.py imports discriminator.py and loss.py, which only import torch)
    e2e_tts/src/tools/{stft,utils}.py                                                 -> baseline/_ref/tools/
        (stft.py:11-89 TorchSTFT; its imports librosa / parselmouth / pyworld are not installed anywhere in this
         image: baseline/_ref users install the same stub modules oracle/make_golden.py uses)
Nothing under e2e_tts_b200/ ever imports baseline/_ref; only bench.py's baseline legs and tests do.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

REF = "/root/reference/e2e_tts"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DST = os.path.join(ROOT, "baseline", "_ref")

FILES = [
    ("models/vocoder/__init__.py", "vocoder/__init__.py"),
    ("models/vocoder/generator.py", "vocoder/generator.py"),
    ("models/vocoder/layers.py", "vocoder/layers.py"),
    ("models/vocoder/function.py", "vocoder/function.py"),
    ("models/vocoder/discriminator.py", "vocoder/discriminator.py"),
    ("models/vocoder/loss.py", "vocoder/loss.py"),
    ("src/tools/stft.py", "tools/stft.py"),
    ("src/tools/utils.py", "tools/utils.py"),
]


def _sha(path: str) -> str:
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def build_ref(verbose: bool = True) -> bool:
    """Returns True when baseline/_ref is complete (copied now or already there)."""
    if not os.path.isdir(REF):
        ok = all(os.path.exists(os.path.join(DST, d)) for _, d in FILES)
        if verbose:
            print("oracle.build_ref: %s absent; baseline/_ref %s" % (REF, "already staged" if ok else "NOT staged"))
        return ok
    for src, dst in FILES:
        s, d = os.path.join(REF, src), os.path.join(DST, dst)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        if _sha(s) != _sha(d):
            raise RuntimeError("copy of %s differs from the source" % src)
    with open(os.path.join(DST, "MANIFEST.txt"), "w") as f:
        for src, dst in FILES:
            f.write("%s  %s  <- %s\n" % (_sha(os.path.join(DST, dst)), dst, src))
    if verbose:
        print("oracle.build_ref: staged %d unmodified reference files under %s" % (len(FILES), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if build_ref() else 1)
