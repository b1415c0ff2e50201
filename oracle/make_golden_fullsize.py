"""Full-size goldens for BASELINE configs 2 and 3 (TEST INFRASTRUCTURE; authoring container only):

    python -m oracle.make_golden_fullsize

Runs the UNMODIFIED reference generator (baseline/_ref staged by oracle/build_ref.py from /root/reference;
e2e_tts/models/vocoder/generator.py:13-62) on the seeded inputs tests/test_gpu_full_size.py uses - every one of the
16 x 5 s and 8 x 30 s utterances - and commits a strided sample of the reference waveforms:
    first / last 2048 samples of every utterance (per-layer zero padding at the utterance ends) and every 97th
    (cfg 2) / 197th (cfg 3) sample in between (the strides are coprime to every tile size, so tile seams are hit).
The GPU test regenerates the same inputs from the seeds and compares ALL utterances at those positions."""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from e2e_tts_b200 import synthetic as sy  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
CASES = [("full_cfg2", 16, 431, 21, 31, 97), ("full_cfg3", 8, 2584, 22, 32, 197)]   # name, B, T, weight seed, mel seed, stride
EDGE = 2048


def sample_index(n: int, stride: int) -> np.ndarray:
    idx = np.concatenate([np.arange(EDGE), np.arange(EDGE, n - EDGE, stride), np.arange(n - EDGE, n)])
    return idx.astype(np.int64)


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    for name, B, T, wseed, mseed, stride in CASES:
        sd = sy.make_state_dict(sy.DEFAULT_CONFIG, wseed, "strong")
        m = ref_loader.build_reference_hifigan(sy.DEFAULT_CONFIG, sd)
        assert m is not None, "run `python -m oracle.build_ref` first"
        mel = sy.mel_like(B, T, mseed)
        idx = sample_index(256 * T, stride)
        rows = []
        t0 = time.time()
        with torch.no_grad():
            for b in range(B):
                rows.append(m(mel[b:b + 1])[0, 0].numpy()[idx])
        wav = np.stack(rows).astype(np.float32)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), wav=wav, index=idx, B=B, T=T, weight_seed=wseed,
                            mel_seed=mseed, regime="strong", absmax=np.float32(np.abs(wav).max()))
        print("%s: %s in %.0f s, |wav|max %.3f" % (name, wav.shape, time.time() - t0, np.abs(wav).max()))


if __name__ == "__main__":
    main()
