"""combine_audio (reference e2e_tts/src/api/utils.py:108-117): the oracle restatement against hand-computed vectors,
the drop-in against the oracle (float input and already-int16 input), on CPU."""
import numpy as np
import torch

import e2e_tts_b200 as pkg
from oracle import postprocess_oracle as po


def test_oracle_known_answers():
    a = [np.array([0.5, -0.5, 0.25, 0.999], dtype=np.float32), np.array([-0.75, 0.1, 0.1, 0.1], dtype=np.float32)]
    got = po.combine_audio(a, [1, 2], 3, hop_length=2)
    # utterance 0 trimmed to 2 samples, utterance 1 to 4; 3 zeros after each; truncation toward zero
    want = np.array([16384, -16384, 0, 0, 0, -24576, 3276, 3276, 3276, 0, 0, 0], dtype=np.int16)
    assert got.dtype == np.int16 and np.array_equal(got, want)


def test_dropin_matches_oracle_on_float_and_pcm_inputs():
    rng = np.random.default_rng(0)
    audios = [np.tanh(rng.standard_normal(256 * 9)).astype(np.float32) for _ in range(5)]
    lengths = [9, 3, 0, 7, 9]
    want = po.combine_audio(audios, lengths, 11025)
    got = pkg.combine_audio(audios, lengths, 11025)
    assert got.dtype == np.int16 and np.array_equal(got, want)
    got_t = pkg.combine_audio([torch.from_numpy(a) for a in audios], lengths, 11025)
    assert np.array_equal(got_t, want)
    # the int16 PCM HifiGan.forward_pcm16 produces (trunc(wav * 32768), zero beyond the length) gives the same stream
    pcm = []
    for a, n in zip(audios, lengths):
        q = np.trunc(a * np.float32(32768.0)).astype(np.int16)
        q[n * 256:] = 0
        pcm.append(q)
    assert np.array_equal(pkg.combine_audio(pcm, lengths, 11025), want)
    assert pkg.combine_audio([], [], 5).shape == (0,)
