"""CPU oracle for the HiFi-GAN generator — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
module; the product package (e2e_tts_b200) never does.

It is a plain fp32 (or fp64) restatement, in eager torch-on-CPU functional ops, of
    e2e_tts/models/vocoder/generator.py:37-53   HifiGan.forward
    e2e_tts/models/vocoder/layers.py:33-40      ResBlock1.forward
    e2e_tts/models/vocoder/layers.py:60-65      ResBlock2.forward
    e2e_tts/models/vocoder/function.py:16-17    get_padding
    e2e_tts/models/vocoder/generator.py:91-109  iSTFT.forward (istft_forward; `resblock == '1'` quirk of :71 included)
    e2e_tts/src/tools/stft.py:138-148           inverse_stft (+ inverse_stft_def, a float64 numpy definition)
plus the weight-norm fold torch applies in its pre-forward hook (w = g * v / ||v||, norm over all dims but 0).
A second, definition-level numpy implementation of the two convolutions (`conv1d_def`,
`conv_transpose1d_def`, SURVEY.md §8 a'5-6) cross-checks the torch ops on tiny shapes.

Pinning: the reference has no tests or golden vectors.  This oracle is pinned by running
the UNMODIFIED reference modules in the authoring container (oracle/make_golden.py, which imports
/root/reference) and committing their outputs under tests/golden/; tests/test_oracle.py checks the oracle
against those files.
"""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn.functional as F

LRELU_SLOPE = 0.1  # generator.py:10, layers.py:7

DEFAULT_CONFIG = {  # e2e_tts/config/model_config.yaml:75-82
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
    """HifiGan compares `config['resblock'] == 1` (generator.py:19); iSTFT compares with the STRING '1'
    (generator.py:71), so the shipped `resblock: 1` selects ResBlock2 there."""
    if "gen_istft_n_fft" in config:
        return 1 if config["resblock"] == "1" else 2
    return 1 if config["resblock"] == 1 else 2


def get_padding(kernel_size: int, dilation: int = 1) -> int:
    return int((kernel_size * dilation - dilation) / 2)  # function.py:16-17


def fold_weight_norm(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """torch.nn.utils.weight_norm(dim=0): w = g * v / ||v||_2 with the norm over dims 1.."""
    n = v.reshape(v.shape[0], -1).norm(dim=1).reshape(-1, *([1] * (v.dim() - 1)))
    return v * (g / n)


def layer_names(config: dict) -> List[tuple]:
    """(state-dict prefix, kind, c_in, c_out, k) for every conv of HifiGan.__init__ (generator.py:14-35)."""
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
    n_post = config["gen_istft_n_fft"] + 2 if "gen_istft_n_fft" in config else 1   # generator.py:86 / :33
    out.append(("conv_post", "conv", ch, n_post, 7))
    return out


def make_state_dict(config: dict, seed: int, regime: str = "strong") -> Dict[str, torch.Tensor]:
    """Deterministic synthetic checkpoint in the reference's 3-tensors-per-layer naming.

    regime "default": what the reference's constructor effectively produces:
        v ~ U(+-1/sqrt(fan_in)), g = ||v||, bias ~ U(+-1/sqrt(fan_in)); waveform abs-max ~ 0.07.
    regime "strong": v ~ N(0, 1/sqrt(fan_in_eff)), g = ||v|| * U(0.8, 1.25) (so the fold matters),
        bias ~ U(+-0.1); fan_in_eff = C_in*k for Conv1d and C_in*k/stride for ConvTranspose1d, which keeps
        activations O(1) through all four stages; waveform abs-max ~ 0.8.
    """
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
            gain = {"conv_pre": 0.2, "conv_post": 0.5}.get(name, 1.0)  # log-mel inputs are O(5); keep tanh unsaturated
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


def _weight(sd: dict, name: str, dtype) -> torch.Tensor:
    if name + ".weight" in sd:
        return sd[name + ".weight"].to(dtype)
    return fold_weight_norm(sd[name + ".weight_g"].to(dtype), sd[name + ".weight_v"].to(dtype))


def hifigan_forward(sd: dict, config: dict, mel: torch.Tensor, dtype=torch.float32,
                    taps: Optional[dict] = None) -> torch.Tensor:
    """mel [B, 80, T] -> wav [B, 1, prod(upsample_rates)*T].  `taps`, if given, receives named intermediates."""
    x = _trunk(sd, config, mel, dtype, taps)
    x = F.leaky_relu(x)                                                                 # :49 (slope 0.01)
    x = F.conv1d(x, _weight(sd, "conv_post", dtype), sd["conv_post.bias"].to(dtype), padding=3)   # :50
    if taps is not None:
        taps["conv_post"] = x
    return torch.tanh(x)                                                                # :51


def istft_forward(sd: dict, config: dict, mel: torch.Tensor, dtype=torch.float32):
    """class iSTFT, generator.py:91-109: mel [B, 80, T] -> (spec, phase), each [B, n_fft/2+1, prod(rates)*T + 1]."""
    n_fft = config["gen_istft_n_fft"]
    x = _trunk(sd, config, mel, dtype, None)                                            # :92-101 (same as HifiGan)
    x = F.leaky_relu(x)                                                                 # :102 (slope 0.01)
    x = F.pad(x, (1, 0), mode="reflect")                                                # :103 ReflectionPad1d((1, 0))
    x = F.conv1d(x, _weight(sd, "conv_post", dtype), sd["conv_post.bias"].to(dtype), padding=3)   # :104
    spec = torch.exp(x[:, :n_fft // 2 + 1, :])                                          # :105
    phase = torch.sin(x[:, n_fft // 2 + 1:, :])                                         # :106
    return spec, phase


def inverse_stft(magnitude: torch.Tensor, phase: torch.Tensor, n_fft=1024, hop_size=256, win_size=1024) -> torch.Tensor:
    """e2e_tts/src/tools/stft.py:138-148, line for line."""
    hann_window = torch.hann_window(win_size).to(magnitude.device)
    inverse_transform = torch.istft(magnitude * torch.exp(phase * 1j), n_fft=n_fft, hop_length=hop_size,
                                    win_length=win_size, window=hann_window)
    return inverse_transform.unsqueeze(-2)


def inverse_stft_def(mag: np.ndarray, phase: np.ndarray, n_fft: int, hop: int) -> np.ndarray:
    """Definition of the same transform in float64 numpy (win == n_fft, periodic Hann, center=True):
    overlap-add of w * irfft(frame) divided by the overlap-added w^2, n_fft/2 samples trimmed at both ends."""
    B, nb, F_ = mag.shape
    w = 0.5 - 0.5 * np.cos(2 * np.pi * np.arange(n_fft) / n_fft)
    X = mag.astype(np.float64) * np.exp(1j * phase.astype(np.float64))
    total = n_fft + hop * (F_ - 1)
    num, den = np.zeros((B, total)), np.zeros(total)
    for f in range(F_):
        num[:, hop * f: hop * f + n_fft] += np.fft.irfft(X[:, :, f], n=n_fft, axis=1) * w
        den[hop * f: hop * f + n_fft] += w * w
    lo, hi = n_fft // 2, total - n_fft // 2            # the trimmed edges are where the envelope vanishes
    return num[:, lo:hi] / den[lo:hi]


def _trunk(sd: dict, config: dict, mel: torch.Tensor, dtype, taps: Optional[dict]) -> torch.Tensor:
    """conv_pre and the upsampling stages (generator.py:38-48; identical in iSTFT.forward, :92-101)."""
    x = mel.to(dtype)
    b = lambda n: sd[n + ".bias"].to(dtype)
    nk = len(config["resblock_kernel_sizes"])
    x = F.conv1d(x, _weight(sd, "conv_pre", dtype), b("conv_pre"), padding=3)          # generator.py:38
    if taps is not None:
        taps["conv_pre"] = x
    for i, (u, k) in enumerate(zip(config["upsample_rates"], config["upsample_kernel_sizes"])):
        x = F.leaky_relu(x, LRELU_SLOPE)                                                # :40
        x = F.conv_transpose1d(x, _weight(sd, "ups.%d" % i, dtype), b("ups.%d" % i), stride=u,
                               padding=(k - u) // 2)                                    # :41
        if taps is not None:
            taps["ups.%d" % i] = x
        xs = None
        for j in range(nk):                                                             # :43-47
            n = i * nk + j
            ks = config["resblock_kernel_sizes"][j]
            dil = config["resblock_dilation_sizes"][j]
            y = x
            if resblock_type(config) == 1:                                              # layers.py:33-40
                for m in range(3):
                    p1, p2 = "resblocks.%d.convs1.%d" % (n, m), "resblocks.%d.convs2.%d" % (n, m)
                    xt = F.leaky_relu(y, LRELU_SLOPE)
                    xt = F.conv1d(xt, _weight(sd, p1, dtype), b(p1), dilation=dil[m], padding=get_padding(ks, dil[m]))
                    xt = F.leaky_relu(xt, LRELU_SLOPE)
                    xt = F.conv1d(xt, _weight(sd, p2, dtype), b(p2), dilation=1, padding=get_padding(ks, 1))
                    y = xt + y
            else:                                                                       # layers.py:60-65
                for m in range(2):
                    p1 = "resblocks.%d.convs.%d" % (n, m)
                    xt = F.leaky_relu(y, LRELU_SLOPE)
                    xt = F.conv1d(xt, _weight(sd, p1, dtype), b(p1), dilation=dil[m], padding=get_padding(ks, dil[m]))
                    y = xt + y
            if taps is not None:
                taps["resblocks.%d" % n] = y
            xs = y if xs is None else xs + y
        x = xs / nk                                                                     # :48
        if taps is not None:
            taps["stage.%d" % i] = x
    return x


# ----------------------------------------------------------------------------------------------------
# Definition-level numpy convolutions (loops; tiny shapes only)
# ----------------------------------------------------------------------------------------------------
def conv1d_def(x: np.ndarray, w: np.ndarray, bias: np.ndarray, dilation: int) -> np.ndarray:
    """y[co,t] = b[co] + sum_ci sum_j W[co,ci,j] * x[ci, t + (j-(k-1)/2)*d], zero outside [0,T)."""
    cout, cin, k = w.shape
    T = x.shape[1]
    y = np.zeros((cout, T), dtype=np.float64)
    for t in range(T):
        for j in range(k):
            tt = t + (j - (k - 1) // 2) * dilation
            if 0 <= tt < T:
                y[:, t] += w[:, :, j].astype(np.float64) @ x[:, tt].astype(np.float64)
    return y + bias[:, None]


def conv_transpose1d_def(x: np.ndarray, w: np.ndarray, bias: np.ndarray, stride: int) -> np.ndarray:
    """ConvTranspose1d(k = 2u, stride = u, padding = u/2): y[co,n] = b + sum_t sum_ci x[ci,t]*W[ci,co,n+u/2-u*t]."""
    cin, cout, k = w.shape
    u = stride
    pad = (k - u) // 2
    T = x.shape[1]
    y = np.zeros((cout, u * T), dtype=np.float64)
    for n in range(u * T):
        for t in range(T):
            j = n + pad - u * t
            if 0 <= j < k:
                y[:, n] += w[:, :, j].astype(np.float64).T @ x[:, t].astype(np.float64)
    return y + bias[:, None]
