"""BASELINE.json's full-size configurations on the GPU, checked through size-independent properties (the CPU oracle
would take minutes at these sizes; a few rows are still compared with it):
  cfg 2  16 x 5 s vocoder batch : determinism, batch independence, time-shift equivariance in the interior
                                  (the generator is a stack of convolutions: shifting the mel by one frame shifts the
                                  waveform by 256 samples wherever the 13-frame receptive halo does not see an edge),
                                  one utterance against the oracle
  cfg 3  8 x 30 s               : every utterance equals its single-utterance run (tiling never mixes utterances)
  cfg 5  1024 x 10 s mel        : rows equal their single-clip runs bit for bit, three rows against the oracle,
                                  energy^2 against the direct sum over the bins of the float64 definition"""
import numpy as np
import pytest
import torch

import e2e_tts_b200 as pkg
from oracle import hifigan_oracle as ho
from oracle import mel_oracle as mo

pytestmark = pytest.mark.gpu


def mel_like(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).clamp(-11.5, 2.0)


def build(seed):
    sd = ho.make_state_dict(ho.DEFAULT_CONFIG, seed, "strong")
    voc = pkg.HifiGan(ho.DEFAULT_CONFIG)
    voc.load_state_dict(sd)
    return voc.eval().to("cuda"), sd


def test_cfg2_batch16_x_5s_properties():
    voc, sd = build(21)
    B, T = 16, 431
    mel = mel_like(B, T, 31)
    with torch.no_grad():
        a = voc(mel.cuda())
        b = voc(mel.cuda())
        one = voc(mel[7:8].cuda())
        shifted = voc(torch.roll(mel, 1, dims=2).cuda())       # frame t -> t + 1 (frame 0 receives the last one)
    assert a.shape == (B, 1, 256 * T) and torch.isfinite(a).all()
    assert torch.equal(a, b)                                    # deterministic
    assert torch.equal(a[7:8], one)                             # batch items independent
    halo = 16                                                   # > 13 frames of receptive field per side
    lhs = shifted[:, :, 256 * (halo + 1): 256 * (T - halo)]
    rhs = a[:, :, 256 * halo: 256 * (T - halo - 1)]
    assert torch.equal(lhs, rhs)                                # bit-exact: the same tiles see the same operands...
    with torch.no_grad():
        ref = ho.hifigan_forward(sd, ho.DEFAULT_CONFIG, mel[3:4])
    scale = ref.abs().max().item()
    d = (a[3:4].cpu() - ref).abs()
    assert d.max().item() <= 2e-2 * scale and d.mean().item() <= 3e-3 * scale


def test_cfg3_batch8_x_30s_rows_match_single_runs():
    voc, _ = build(22)
    B, T = 8, 2584
    mel = mel_like(B, T, 32).cuda()
    with torch.no_grad():
        a = voc(mel)
        assert a.shape == (B, 1, 256 * T) and torch.isfinite(a).all()
        for i in (0, 5, 7):
            assert torch.equal(a[i:i + 1], voc(mel[i:i + 1]))


def test_cfg5_mel_1024_clips_x_10s():
    B, L = 1024, 220500
    g = torch.Generator(device="cuda").manual_seed(0)
    wav = torch.rand(B, L, device="cuda", generator=g) * 2 - 1
    stft = pkg.TorchSTFT()
    mel, energy = stft.mel_spectrogram(wav, return_energy=True)
    assert mel.shape == (B, 80, 861) and energy.shape == (B, 861) and torch.isfinite(mel).all()
    for i in (0, 511, 1023):
        m1, e1 = stft.mel_spectrogram(wav[i:i + 1], return_energy=True)
        assert torch.equal(mel[i:i + 1], m1) and torch.equal(energy[i:i + 1], e1)
    rows = wav[[0, 511, 1023]].cpu()
    ref_mel, ref_energy = mo.mel_spectrogram(rows, return_energy=True)
    got = mel[[0, 511, 1023]].cpu().double()
    lin, ref_lin = got.exp(), ref_mel.double().exp()
    bound = 1e-5 * ref_lin.max(dim=1, keepdim=True).values + 1e-7
    assert ((lin - ref_lin).abs() / bound).max().item() <= 1.0
    assert ((energy[[0, 511, 1023]].cpu() - ref_energy).abs() / ref_energy).max().item() <= 2e-5
