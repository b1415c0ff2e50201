"""GPU parity for the iSTFTNet head (SURVEY.md §8 f, N1), through the C ABI (e2e_voc_forward_spec, e2e_istft_forward).

Tolerances: conv_post's output y feeds exp / sin, so the generator is compared on y itself: |log spec - log ref| and
|phase - ref| <= 2e-2 * max|y_ref| (bf16 operands / fp32 accumulate, the vocoder bound of test_gpu_vocoder.py);
inverse_stft is fp32 arithmetic on fp32 inputs: |wav - ref| <= 1e-5 * max|ref|."""
import os

import numpy as np
import pytest
import torch

import e2e_tts_b200 as pkg
from oracle import hifigan_oracle as ho
import margins

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def build(cfg, seed):
    sd = ho.make_state_dict(cfg, seed, "strong")
    g = pkg.iSTFT(cfg)
    g.load_state_dict(sd)
    return g.eval().to("cuda"), sd


def check_head(spec, phase, ref_spec, ref_phase, what):
    spec, phase = spec.double().cpu(), phase.double().cpu()
    assert spec.shape == ref_spec.shape and phase.shape == ref_phase.shape, (spec.shape, ref_spec.shape)
    assert torch.isfinite(spec).all() and torch.isfinite(phase).all()
    ylog = ref_spec.double().log()
    scale = max(ylog.abs().max().item(), 1.0)
    e1 = (spec.log() - ylog).abs().max().item()
    e2 = (phase - ref_phase.double()).abs().max().item()
    margins.record(what, log_spec_rel=e1 / scale, phase_rel=e2 / scale, bound=2e-2)
    assert e1 <= 2e-2 * scale, "%s: log-spec err %.3g vs scale %.3g" % (what, e1, scale)
    assert e2 <= 2e-2 * scale, "%s: phase err %.3g vs scale %.3g" % (what, e2, scale)


def test_generator_against_reference_golden():
    g = np.load(os.path.join(GOLD, "istft_strong.npz"))
    gen, _ = build(ho.ISTFT_CONFIG, int(g["seed"]))
    with torch.no_grad():
        spec, phase = gen(torch.from_numpy(g["mel"]).cuda())
    check_head(spec, phase, torch.from_numpy(g["spec"]), torch.from_numpy(g["phase"]), "golden")


@pytest.mark.parametrize("B,T,resblock", [(1, 1, 1), (3, 9, 1), (2, 40, "1"), (1, 431, 1)])
def test_generator_against_oracle(B, T, resblock):
    cfg = dict(ho.ISTFT_CONFIG, resblock=resblock)     # "1" selects ResBlock1 (generator.py:71), the int ResBlock2
    gen, sd = build(cfg, 40 + T)
    gm = torch.Generator().manual_seed(T)
    mel = (torch.randn(B, 80, T, generator=gm) * 2.0 - 5.0).clamp(-11.5, 2.0)
    with torch.no_grad():
        spec, phase = gen(mel.cuda())
        rs, rp = ho.istft_forward(sd, cfg, mel)
    assert spec.shape == (B, 9, 64 * T + 1)
    check_head(spec, phase, rs, rp, "B%d T%d" % (B, T))


def test_inverse_stft_against_reference_golden_and_definition():
    g = np.load(os.path.join(GOLD, "istft_strong.npz"))
    wav = pkg.inverse_stft(torch.from_numpy(g["spec"]).cuda(), torch.from_numpy(g["phase"]).cuda(), 16, 4, 16)
    ref = torch.from_numpy(g["wav"])
    assert wav.shape == ref.shape == (2, 1, 1536)
    assert (wav.cpu() - ref).abs().max().item() <= 1e-5 * ref.abs().max().item()
    # other small transforms, ragged frame counts, against the float64 definition
    rng = np.random.default_rng(0)
    for n_fft, hop, frames, B in ((16, 4, 2, 1), (16, 4, 1000, 3), (32, 8, 77, 2), (8, 2, 301, 2), (64, 16, 50, 1)):
        mag = np.exp(rng.standard_normal((B, n_fft // 2 + 1, frames))).astype(np.float32)
        ph = np.sin(rng.standard_normal((B, n_fft // 2 + 1, frames)) * 3).astype(np.float32)
        got = pkg.inverse_stft(torch.from_numpy(mag).cuda(), torch.from_numpy(ph).cuda(), n_fft, hop, n_fft).cpu().numpy()
        want = ho.inverse_stft_def(mag, ph, n_fft, hop)
        assert got.shape == (B, 1, hop * (frames - 1))
        assert np.abs(got[:, 0] - want).max() <= 1e-5 * np.abs(want).max(), (n_fft, hop, frames)
    with pytest.raises(Exception):
        pkg.inverse_stft(torch.zeros(1, 513, 4).cuda(), torch.zeros(1, 513, 4).cuda(), 1024, 256, 1024)


def test_end_to_end_waveform_matches_the_reference_pipeline():
    """spec, phase -> inverse_stft, as load_vocoder(use_complex=True) callers do (tools_for_model.py:45-50)."""
    gen, sd = build(ho.ISTFT_CONFIG, 77)
    gm = torch.Generator().manual_seed(5)
    mel = (torch.randn(2, 80, 30, generator=gm) * 2.0 - 5.0).clamp(-11.5, 2.0)
    with torch.no_grad():
        spec, phase = gen(mel.cuda())
        wav = pkg.inverse_stft(spec, phase, 16, 4, 16).cpu()
        rs, rp = ho.istft_forward(sd, ho.ISTFT_CONFIG, mel)
        ref = ho.inverse_stft(rs, rp, 16, 4, 16)
    assert wav.shape == ref.shape == (2, 1, 256 * 30)
    scale = ref.abs().max().item()
    assert (wav - ref).abs().max().item() <= 5e-2 * scale and (wav - ref).abs().mean().item() <= 1e-2 * scale
