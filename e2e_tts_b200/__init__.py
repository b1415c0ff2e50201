"""e2e_tts_b200 — B200-native (sm_100a) implementation of the e2e-tts synthesis hot path.

Drop-in names (same signatures as the reference, see each module's docstring):
    HifiGan, ResBlock1, ResBlock2, get_padding, init_weights   <- e2e_tts/models/vocoder/{generator,layers,function}.py
    TorchSTFT, generate_melspecs, dynamic_range_compression     <- e2e_tts/src/tools/{stft,utils}.py
    combine_audio                                               <- e2e_tts/src/api/utils.py:108-117 (+ HifiGan.forward_pcm16)
    iSTFT, inverse_stft                                         <- generator.py:65-119, e2e_tts/src/tools/stft.py:138-148
    crop_segments_and_mel                                       <- e2e_tts/src/tools/dataloader.py:364-373 (MelAudioLoader crop + mel, batched on the device)
    Postnet                                                     <- e2e_tts/models/acoustic/unsupervised_fastspeech2/layers.py:507-563
"""
from .vocoder import HifiGan, iSTFT, ResBlock1, ResBlock2, get_padding, init_weights, apply_weight_norm, LRELU_SLOPE  # noqa: F401
from .stft import TorchSTFT, generate_melspecs, crop_segments_and_mel, inverse_stft, dynamic_range_compression, dynamic_range_decompression  # noqa: F401

from .postprocess import combine_audio  # noqa: F401
from .postnet import Postnet  # noqa: F401
from .serving import HostPipeline  # noqa: F401

__all__ = ["combine_audio", "Postnet", "HostPipeline", "HifiGan", "iSTFT", "inverse_stft", "ResBlock1", "ResBlock2", "get_padding", "init_weights", "apply_weight_norm", "TorchSTFT", "generate_melspecs", "crop_segments_and_mel",
           "dynamic_range_compression", "dynamic_range_decompression"]
