"""Builds the CUDA library IN-TREE for sm_100a with nvcc (cross-compiles without a GPU).

    python -m e2e_tts_b200.build [--force]

Output: e2e_tts_b200/lib/libe2e_tts_b200.so (git-ignored; it travels to the GPU box with the tree).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libe2e_tts_b200.so")
SOURCES = ["voc.cu", "mel.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-shared", "-Xcompiler", "-fPIC",
]  # no --use_fast_math: tanhf / logf / sqrtf stay accurate, parity comes first


def _nvcc() -> str | None:
    cand = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    return cand if os.path.exists(cand) else None


def _stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(_HERE, "..", "include", "e2e_tts_b200.h"))
    return any(os.path.getmtime(d) > t for d in deps if os.path.isfile(d))


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile the library if it is missing or older than its sources.  Returns the .so path."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = _nvcc()
    if nvcc is None:
        if os.path.exists(LIB_PATH):
            return LIB_PATH  # cannot rebuild here; use what travelled with the tree
        raise RuntimeError("nvcc not found and %s does not exist" % LIB_PATH)
    os.makedirs(LIB_DIR, exist_ok=True)
    tmp = LIB_PATH + ".tmp.%d" % os.getpid()
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", tmp]
    cmd += [os.path.join(CSRC, s) for s in SOURCES]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed:\n%s\n%s" % (" ".join(cmd), res.stderr[-4000:]))
    if verbose:
        sys.stderr.write(res.stderr)
    os.replace(tmp, LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
