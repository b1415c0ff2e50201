"""Synthetic checkpoints and inputs for benchmarks, smoke runs and tests (there is no network for real checkpoints).

`make_state_dict` writes a seeded random-init checkpoint in the reference's state-dict layout (234 tensors
`{layer}.{bias,weight_g,weight_v}` for the default HiFi-GAN V1 config, e2e_tts/models/vocoder/generator.py:14-35),
so the same file loads into the reference class and into e2e_tts_b200.HifiGan.  The generator is bit-identical to the
one the CPU oracle uses for the committed goldens (tests/test_synthetic.py checks that).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np
import torch

DEFAULT_CONFIG = {  # e2e_tts/config/model_config.yaml:75-82 (`hifigan:` mapping)
    "resblock": 1,
    "num_freq": 1025,
    "upsample_rates": [8, 8, 2, 2],
    "upsample_kernel_sizes": [16, 16, 4, 4],
    "upsample_initial_channel": 512,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
}

ISTFT_CONFIG = {  # e2e_tts/config/model_config.yaml:83-92 (`istft:` mapping, class iSTFT)
    "resblock": 1,
    "gen_istft_n_fft": 16,
    "gen_istft_hop_size": 4,
    "gen_istft_win_size": 16,
    "upsample_rates": [8, 8],
    "upsample_kernel_sizes": [16, 16],
    "upsample_initial_channel": 512,
    "resblock_kernel_sizes": [3, 7, 11],
    "resblock_dilation_sizes": [[1, 3, 5], [1, 3, 5], [1, 3, 5]],
}


def resblock_type(config: dict) -> int:
    """HifiGan compares `config['resblock'] == 1` (generator.py:19); iSTFT compares with the STRING '1' (:71)."""
    if "gen_istft_n_fft" in config:
        return 1 if config["resblock"] == "1" else 2
    return 1 if config["resblock"] == 1 else 2


def layer_names(config: dict) -> List[tuple]:
    """(state-dict prefix, kind, c_in, c_out, k) for every conv of the generator, in construction order."""
    c0 = config["upsample_initial_channel"]
    out = [("conv_pre", "conv", 80, c0, 7)]
    for i, (u, k) in enumerate(zip(config["upsample_rates"], config["upsample_kernel_sizes"])):
        out.append(("ups.%d" % i, "convt", c0 // 2 ** i, c0 // 2 ** (i + 1), k))
    ch = c0
    n = 0
    for i in range(len(config["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for k, d in zip(config["resblock_kernel_sizes"], config["resblock_dilation_sizes"]):
            if resblock_type(config) == 1:
                for m in range(3):
                    out.append(("resblocks.%d.convs1.%d" % (n, m), "conv", ch, ch, k))
                for m in range(3):
                    out.append(("resblocks.%d.convs2.%d" % (n, m), "conv", ch, ch, k))
            else:
                for m in range(2):
                    out.append(("resblocks.%d.convs.%d" % (n, m), "conv", ch, ch, k))
            n += 1
    n_post = config["gen_istft_n_fft"] + 2 if "gen_istft_n_fft" in config else 1
    out.append(("conv_post", "conv", ch, n_post, 7))
    return out


def make_state_dict(config: dict, seed: int, regime: str = "strong") -> Dict[str, torch.Tensor]:
    """regime "default": what the reference's constructor effectively produces (PyTorch's Conv default:
        v ~ U(+-1/sqrt(fan_in)), g = ||v||, bias ~ U(+-1/sqrt(fan_in))); waveform abs-max ~ 0.07.
    regime "strong": v ~ N(0, 1/sqrt(fan_in_eff)), g = ||v|| * U(0.8, 1.25) (so the weight-norm fold matters),
        bias ~ U(+-0.1); keeps activations O(1) through all stages; waveform abs-max ~ 0.8."""
    gen = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}
    rates = dict(("ups.%d" % i, u) for i, u in enumerate(config["upsample_rates"]))
    for name, kind, cin, cout, k in layer_names(config):
        shape = (cin, cout, k) if kind == "convt" else (cout, cin, k)
        if regime == "default":
            fan_in = shape[1] * k
            bound = 1.0 / np.sqrt(fan_in)
            v = (torch.rand(shape, generator=gen) * 2 - 1) * bound
            g = v.reshape(shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
            b = (torch.rand(cout, generator=gen) * 2 - 1) * bound
        elif regime == "strong":
            fan = cin * k / rates[name] if kind == "convt" else cin * k
            gain = {"conv_pre": 0.2, "conv_post": 0.5}.get(name, 1.0)
            v = torch.randn(shape, generator=gen) * (gain / np.sqrt(fan))
            g = v.reshape(shape[0], -1).norm(dim=1).reshape(-1, 1, 1)
            g = g * (0.8 + 0.45 * torch.rand(g.shape, generator=gen))
            b = (torch.rand(cout, generator=gen) * 2 - 1) * 0.1
        else:
            raise ValueError(regime)
        sd[name + ".bias"] = b.float()
        sd[name + ".weight_g"] = g.float()
        sd[name + ".weight_v"] = v.float()
    return sd


def mel_like(B: int, T: int, seed: int) -> torch.Tensor:
    """Log-mel-like input [B, 80, T]: N(-5, 2^2) clipped to [-11.5, 2]."""
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).clamp(-11.5, 2.0)
