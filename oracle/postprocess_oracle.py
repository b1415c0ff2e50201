"""TEST INFRASTRUCTURE ONLY (never imported by e2e_tts_b200/): CPU restatement of the reference's waveform
post-processing, `combine_audio` (e2e_tts/src/api/utils.py:108-117), line for line: trim to `lengths[i] * hop_length`,
`* max_wav_value`, a float64 `np.zeros(distance)` of silence after every utterance, one concatenate, `.astype("int16")`.
Pinned by tests/test_postprocess.py against hand-computed vectors (the reference holds no test or fixture for it,
SURVEY.md §4, and its class needs checkpoints and Coqui TTS to construct: parity unpinned by the reference)."""
import numpy as np


def combine_audio(audios, lengths, distance, hop_length=256, max_wav_value=32768.0):
    output_audio = []
    for i, audio in enumerate(audios):                      # utils.py:110
        audio = audio[: lengths[i] * hop_length]            # utils.py:111
        audio = (audio * max_wav_value)                     # utils.py:112
        disOfsil = np.zeros(distance)                       # utils.py:113
        output_audio.extend([audio, disOfsil])              # utils.py:115
    return np.concatenate(output_audio).astype("int16")     # utils.py:117
