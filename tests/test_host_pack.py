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


def _tiled8_off(b, t, chunk16, t8, c16):
    """epilogue.cuh tiled8_off: element offset of 16 channels of row t in [B][ceil(T/8)][C/16][8][16]."""
    return ((b * t8 + (t >> 3)) * c16 + chunk16) * 128 + (t & 7) * 16


def _prefetch_range(tiled, b, t0, t1, T, C):
    """epilogue.cuh prefetch_sum_rows: the element range [first, last) the slab producer asks L2 to fetch."""
    t1 = min(t1, T)
    t0 = max(t0, 0)
    if t1 <= t0:
        return None
    if tiled:
        t8 = (T + 7) >> 3
        return (b * t8 + (t0 >> 3)) * C * 8, (b * t8 + ((t1 + 7) >> 3)) * C * 8
    return (b * T + t0) * C, (b * T + t1) * C


@pytest.mark.parametrize("C,T,r_out", [(128, 27584, 118), (64, 55168, 246), (32, 110336, 4 * 126), (32, 96, 40)])
def test_sum_prefetch_range_covers_exactly_the_rows_the_epilogue_reads(C, T, r_out):
    """The rows a unit's last epilogue reads from the running-sum tensor form ONE contiguous range in the tiled8 layout
    (and in the natural one): every 16-channel item of rows [t0, t1) lies inside the prefetched range, the range is
    16-byte aligned and holds no whole 8-row group that the unit does not touch."""
    t8, c16 = (T + 7) >> 3, C // 16
    for b in (0, 3):
        n_tiles = (T + r_out - 1) // r_out
        for tile in sorted({0, 1, min(7, n_tiles - 1), n_tiles - 1}):
            t0, t1 = tile * r_out, tile * r_out + r_out
            first, last = _prefetch_range(1, b, t0, t1, T, C)
            assert first % 8 == 0 and last % 8 == 0 and last > first     # 16-byte pieces of bf16
            offs = [_tiled8_off(b, t, ch, t8, c16) for t in range(t0, min(t1, T)) for ch in range(c16)]
            assert min(offs) >= first and max(offs) + 16 <= last
            assert min(offs) - first < C * 8 and last - (max(offs) + 16) < C * 8   # tight to the 8-row group
            nf, nl = _prefetch_range(0, b, t0, t1, T, C)
            assert nf == (b * T + t0) * C and nl == (b * T + min(t1, T)) * C and nf % 8 == 0
    assert _prefetch_range(1, 0, T, T + r_out, T, C) is None
