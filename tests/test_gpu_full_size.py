"""BASELINE.json's full-size configurations on the GPU, checked through size-independent properties (the CPU oracle
would take minutes at these sizes; a few rows are still compared with it):
  cfg 2  16 x 5 s vocoder batch : determinism, batch independence, time-shift equivariance in the interior
                                  (the generator is a stack of convolutions: shifting the mel by one frame shifts the
                                  waveform by 256 samples wherever the 13-frame receptive halo does not see an edge),
                                  EVERY utterance against the reference's waveform at the sampled positions of
                                  tests/golden/full_cfg2.npz (oracle/make_golden_fullsize.py: the unmodified reference
                                  generator run on the same seeded inputs; utterance ends + a stride coprime to the tiles)
  cfg 3  8 x 30 s               : every utterance equals its single-utterance run (tiling never mixes utterances), and
                                  every utterance against tests/golden/full_cfg3.npz
  cfg 5  1024 x 10 s mel        : rows equal their single-clip runs bit for bit, three rows against the oracle,
                                  energy^2 against the direct sum over the bins of the float64 definition"""
import numpy as np
import pytest
import torch

import os

import e2e_tts_b200 as pkg
from oracle import hifigan_oracle as ho
from oracle import mel_oracle as mo
import margins

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
# bf16 operands: max |err| <= 2 % of the utterance's abs-max.  The MEAN bound at full size is 3.5e-3
# instead of c6's provisional 3e-3 ("to be confirmed on hardware"): scripts/emulate_numerics.py shows that ANY
# implementation with bf16 tensor-core operands lands at 0.21-0.26 % mean / 1.3 % max on these seeds (operand rounding
# alone, everything else fp32; profiles/r02_numerics_emulation.txt), the kernels' extra bf16 storage points add 0.04-0.07 %.
# operand_dtype="fp16" (same speed) is held to bounds 5x tighter, see test_gpu_vocoder.py.
MAX_TOL, MEAN_TOL = 2e-2, 3.5e-3


def check_against_full_golden(name, wav, max_tol=MAX_TOL, mean_tol=MEAN_TOL, tag=""):
    """Every utterance of the batch vs the reference's strided waveform sample."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    idx = torch.from_numpy(g["index"])
    ref = torch.from_numpy(g["wav"])
    got = wav[:, 0, :].cpu()[:, idx]
    assert got.shape == ref.shape, (got.shape, ref.shape)
    worst = 0.0
    for b in range(ref.shape[0]):
        scale = ref[b].abs().max().item()
        d = (got[b] - ref[b]).abs()
        margins.record("%s%s utterance %d" % (tag, name, b), max_rel=d.max().item() / scale,
                       mean_rel=d.mean().item() / scale, bound_max=max_tol, bound_mean=mean_tol, scale=scale)
        assert d.max().item() <= max_tol * scale, "%s utterance %d: max err %.3g vs scale %.3g" % (name, b, d.max().item(), scale)
        assert d.mean().item() <= mean_tol * scale, "%s utterance %d: mean err %.3g" % (name, b, d.mean().item())
        edge = 2048   # the first / last 2048 entries are the utterance ends (per-layer zero padding)
        for sl in (slice(0, edge), slice(-edge, None)):
            assert d[sl].max().item() <= max_tol * scale
        worst = max(worst, d.max().item() / scale)
    return worst


def mel_like(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).clamp(-11.5, 2.0)


def build(seed, operand_dtype=None):
    sd = ho.make_state_dict(ho.DEFAULT_CONFIG, seed, "strong")
    voc = pkg.HifiGan(ho.DEFAULT_CONFIG, operand_dtype=operand_dtype)
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
    g = np.load(os.path.join(GOLD, "full_cfg2.npz"))
    assert int(g["weight_seed"]) == 21 and int(g["mel_seed"]) == 31 and int(g["B"]) == B and int(g["T"]) == T
    check_against_full_golden("full_cfg2", a)


@pytest.mark.parametrize("name,B,T,wseed,mseed", [("full_cfg2", 16, 431, 21, 31), ("full_cfg3", 8, 2584, 22, 32)])
def test_full_size_fp16_operands_every_utterance(name, B, T, wseed, mseed):
    """operand_dtype="fp16" at BASELINE's full sizes: every utterance within max 4e-3 / mean 6e-4 of the reference."""
    voc, _ = build(wseed, operand_dtype="fp16")
    with torch.no_grad():
        a = voc(mel_like(B, T, mseed).cuda())
    check_against_full_golden(name, a, 4e-3, 6e-4, tag="fp16 ")


def test_cfg3_batch8_x_30s_rows_match_single_runs():
    voc, _ = build(22)
    B, T = 8, 2584
    mel = mel_like(B, T, 32).cuda()
    with torch.no_grad():
        a = voc(mel)
        assert a.shape == (B, 1, 256 * T) and torch.isfinite(a).all()
        for i in (0, 5, 7):
            assert torch.equal(a[i:i + 1], voc(mel[i:i + 1]))
    g = np.load(os.path.join(GOLD, "full_cfg3.npz"))
    assert int(g["weight_seed"]) == 22 and int(g["mel_seed"]) == 32 and int(g["B"]) == B and int(g["T"]) == T
    check_against_full_golden("full_cfg3", a)


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
