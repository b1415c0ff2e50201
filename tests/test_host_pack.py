"""The weight packer's host-side fp32 -> fp16 (round-to-nearest-even, saturating at +-65504 like the kernels'
cvt.rn.satfinite) and fp32 -> bf16 conversions (e2e_tts_b200/csrc/conv_host.cuh), against numpy / torch."""
import os
import shutil
import subprocess

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="needs nvcc")
def test_host_f16_bf16_conversions(tmp_path):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = str(tmp_path / "host_f16")
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "cuda", "host_f16.cu")], check=True,
                   capture_output=True)
    rng = np.random.default_rng(0)
    vals = np.concatenate([
        rng.standard_normal(4000).astype(np.float32) * np.float32(10.0) ** rng.integers(-9, 6, 4000).astype(np.float32),
        np.array([0.0, -0.0, 1.0, -1.0, 65504.0, 65519.9, 65520.0, 1e6, -1e6, 6.1e-5, 6.0e-5, 5.96e-8, 2.9e-8, 3.1e-8,
                  1e-10, np.float32(2.0 ** -24), np.float32(2.0 ** -25), np.float32(1.5 * 2.0 ** -24), 0.1, 1 / 3],
                 dtype=np.float32),
        # halfway cases between neighbouring halves (ties to even)
        (np.arange(1024, 1100, dtype=np.float32) + 0.5) * np.float32(2.0 ** -10),
    ])
    bits = vals.view(np.uint32)
    out = subprocess.run([exe], input="\n".join("%08x" % b for b in bits), capture_output=True, text=True, check=True).stdout
    got = np.array([[int(x, 16) for x in line.split()] for line in out.strip().splitlines()], dtype=np.uint32)
    assert got.shape == (len(vals), 2)
    with np.errstate(over="ignore"):
        want16 = np.clip(vals, -65504.0, 65504.0).astype(np.float16).view(np.uint16)   # saturating RNE
    assert np.array_equal(got[:, 0].astype(np.uint16), want16)
    wantbf = torch.from_numpy(vals).to(torch.bfloat16).view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(got[:, 1].astype(np.uint16), wantbf)
