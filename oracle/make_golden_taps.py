"""Per-layer goldens - TEST INFRASTRUCTURE, authoring container only:

    python -m oracle.make_golden_taps

Runs the UNMODIFIED reference generator (baseline/_ref, staged from /root/reference by oracle/build_ref.py) with
forward hooks on conv_pre, every ups[i], every resblocks[n] and conv_post (generator.py:18-33) on a small seeded input
and commits the hooked outputs as tests/golden/voc_taps_strong.npz (fp16-packed where lossless enough is not needed:
stored as float32).  tests/test_oracle.py checks the oracle's named intermediates against them, which pins the
restatement layer by layer and not only end to end."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from e2e_tts_b200 import synthetic as sy  # noqa: E402
from oracle import ref_loader  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "voc_taps_strong.npz")
SEED, MEL_SEED, B, T = 41, 141, 1, 2


def main():
    sd = sy.make_state_dict(sy.DEFAULT_CONFIG, SEED, "strong")
    m = ref_loader.build_reference_hifigan(sy.DEFAULT_CONFIG, sd)
    assert m is not None, "run `python -m oracle.build_ref` first"
    taps = {}

    def hook(name):
        return lambda mod, inp, out: taps.__setitem__(name, out.detach().clone())

    m.conv_pre.register_forward_hook(hook("conv_pre"))
    for i, u in enumerate(m.ups):
        u.register_forward_hook(hook("ups.%d" % i))
    for n, rb in enumerate(m.resblocks):
        rb.register_forward_hook(hook("resblocks.%d" % n))
    m.conv_post.register_forward_hook(hook("conv_post"))
    mel = sy.mel_like(B, T, MEL_SEED)
    with torch.no_grad():
        wav = m(mel)
    arrays = {("tap:" + k): v.numpy().astype(np.float32) for k, v in taps.items()}
    np.savez_compressed(OUT, mel=mel.numpy(), wav=wav.numpy(), seed=SEED, mel_seed=MEL_SEED, **arrays)
    print("wrote %s: %d taps, %.0f KB" % (OUT, len(arrays), os.path.getsize(OUT) / 1e3))


if __name__ == "__main__":
    main()
