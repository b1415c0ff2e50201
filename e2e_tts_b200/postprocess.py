"""Waveform post-processing of the synthesis path (SURVEY.md §8 f, N3): the reference's `combine_audio`
(e2e_tts/src/api/utils.py:108-117), which trims every utterance to `mel_len * hop_length`, scales by
`max_wav_value`, joins the utterances with `distance` samples of silence and casts to int16.

`combine_audio` below is the drop-in (same arguments as the method, hop_length / max_wav_value as keywords since
the reference reads them from `self`).  It accepts the float waveforms the reference passes, or the int16 PCM that
`HifiGan.forward_pcm16` already produced on the GPU (trim + scale + cast fused into the conv_post kernel): then it
only slices and concatenates."""
from __future__ import annotations

from typing import Sequence

import numpy as np


def _to_numpy(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    return np.asarray(a)


def combine_audio(audios: Sequence, lengths: Sequence[int], distance: int, hop_length: int = 256,
                  max_wav_value: float = 32768.0) -> np.ndarray:
    """utils.py:108-117.  audios[i]: 1-D float waveform (scaled here) or int16 PCM (already scaled)."""
    out = []
    sil = np.zeros(int(distance), dtype=np.int16)
    for i, audio in enumerate(audios):
        audio = _to_numpy(audio)[: int(lengths[i]) * hop_length]
        if audio.dtype != np.int16:
            # float32 * python float stays float32 (as in the reference); astype truncates toward zero
            audio = (audio * max_wav_value).astype("int16")
        out.extend([audio, sil])
    if not out:
        return np.zeros(0, dtype=np.int16)
    return np.concatenate(out).astype("int16")
