"""CPU tests: the oracle against the golden vectors produced by the UNMODIFIED reference
(oracle/make_golden.py), plus definition-level cross-checks.  No GPU, no /root/reference."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import hifigan_oracle as ho
from oracle import mel_oracle as mo

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _cfg(resblock):
    cfg = dict(ho.DEFAULT_CONFIG)
    cfg["resblock"] = int(resblock)
    return cfg


@pytest.mark.parametrize("name", ["voc_default_init", "voc_strong_init", "voc_strong_resblock2"])
def test_vocoder_oracle_matches_reference_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cfg = _cfg(g["resblock"])
    sd = ho.make_state_dict(cfg, int(g["seed"]), str(g["regime"]))
    with torch.no_grad():
        got = ho.hifigan_forward(sd, cfg, torch.from_numpy(g["mel"]))
    want = torch.from_numpy(g["wav"])
    assert got.shape == want.shape
    assert (got - want).abs().max().item() <= 1e-5          # fp32 round-off of an identical op sequence


def test_state_dict_layout_is_the_reference_one():
    sd = ho.make_state_dict(ho.DEFAULT_CONFIG, 0, "default")
    assert len(sd) == 234
    assert sd["ups.0.weight_g"].shape == (512, 1, 1)          # ConvTranspose1d: weight-norm dim 0 is C_in
    assert sd["resblocks.11.convs2.2.weight_v"].shape == (32, 32, 11)
    n_eff = sum(v.numel() for k, v in sd.items() if not k.endswith("weight_g"))
    assert n_eff == 13926017


@pytest.mark.parametrize("k,d", [(3, 1), (7, 3), (11, 5)])
def test_conv1d_definition(k, d):
    g = torch.Generator().manual_seed(k * 10 + d)
    x = torch.randn(6, 70, generator=g)
    w = torch.randn(5, 6, k, generator=g)
    b = torch.randn(5, generator=g)
    want = torch.nn.functional.conv1d(x[None], w, b, dilation=d, padding=ho.get_padding(k, d))[0].numpy()
    got = ho.conv1d_def(x.numpy(), w.numpy(), b.numpy(), d)
    assert got.shape == want.shape
    np.testing.assert_allclose(got, want, atol=2e-5)


@pytest.mark.parametrize("u", [2, 8])
def test_conv_transpose1d_definition_and_polyphase(u):
    g = torch.Generator().manual_seed(u)
    cin, cout, T = 6, 4, 9
    x = torch.randn(cin, T, generator=g)
    w = torch.randn(cin, cout, 2 * u, generator=g)
    b = torch.randn(cout, generator=g)
    want = torch.nn.functional.conv_transpose1d(x[None], w, b, stride=u, padding=u // 2)[0].numpy()
    got = ho.conv_transpose1d_def(x.numpy(), w.numpy(), b.numpy(), u)
    assert want.shape == (cout, u * T)
    np.testing.assert_allclose(got, want, atol=2e-5)
    # polyphase form used by the CUDA path (SURVEY.md §8 a'5)
    xn, wn = x.numpy().astype(np.float64), w.numpy().astype(np.float64)
    poly = np.zeros_like(got)
    for q in range(T):
        for p in range(u):
            j0 = p + u // 2
            acc = wn[:, :, j0].T @ xn[:, q]
            if j0 < u:
                if q - 1 >= 0:
                    acc = acc + wn[:, :, j0 + u].T @ xn[:, q - 1]
            elif q + 1 < T:
                acc = acc + wn[:, :, j0 - u].T @ xn[:, q + 1]
            poly[:, q * u + p] = acc + b.numpy()
    np.testing.assert_allclose(poly, want, atol=2e-5)


def test_weight_norm_fold_dim0():
    v = torch.randn(7, 5, 3)
    g = torch.rand(7, 1, 1) + 0.5
    w = ho.fold_weight_norm(g, v)
    np.testing.assert_allclose(w.reshape(7, -1).norm(dim=1).numpy(), g.flatten().numpy(), rtol=1e-5)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(GOLD, "mel_*.npz"))))
def test_mel_oracle_matches_reference_golden(path):
    g = np.load(path)
    mel, energy = mo.mel_spectrogram(torch.from_numpy(g["wav"]), return_energy=True)
    assert mel.shape == g["mel"].shape and energy.shape == g["energy"].shape
    assert mel.shape[-1] == mo.num_frames(g["wav"].shape[1])
    np.testing.assert_allclose(mel.numpy(), g["mel"], atol=1e-6)
    np.testing.assert_allclose(energy.numpy(), g["energy"], rtol=1e-6)
    # and the float64 definition agrees with the fp32 reference op sequence to fp32 accuracy
    # (error of an fp32 FFT is relative to the frame's largest magnitude, so the bound is stated on the linear mel)
    lm, lin, en = mo.mel_spectrogram_f64(g["wav"])
    lin32 = np.exp(g["mel"].astype(np.float64))
    frame_max = lin.max(axis=1, keepdims=True)
    assert (np.abs(lin32 - np.maximum(lin, 1e-5)) / frame_max).max() < 1e-5
    assert np.abs(lm - g["mel"]).mean() < 1e-4
    np.testing.assert_allclose(en, g["energy"], rtol=2e-5)


def test_mel_basis_properties_and_torchaudio_crosscheck():
    fb = mo.slaney_mel_basis()
    assert fb.shape == (80, 513) and fb.dtype == np.float32
    assert int((fb != 0).sum()) == 727 and int(np.nonzero(fb.any(0))[0].max()) == 371
    torchaudio = pytest.importorskip("torchaudio")
    ta = torchaudio.functional.melscale_fbanks(513, 0.0, 8000.0, 80, 22050, norm="slaney", mel_scale="slaney").T
    assert np.abs(ta.numpy() - fb).max() < 2e-7


def test_mel_range_assert_and_silence_floor():
    with pytest.raises(AssertionError):
        mo.mel_spectrogram(torch.full((1, 2048), 1.01))
    mel, energy = mo.mel_spectrogram(torch.zeros(1, 2048), return_energy=True)
    assert torch.allclose(mel, torch.full_like(mel, float(np.log(1e-5))))          # clamp floor, -11.51
    assert torch.allclose(energy, torch.full_like(energy, float(np.sqrt(513e-9))), rtol=1e-5)


# ---- iSTFTNet head (SURVEY.md §8 f, N1): class iSTFT (generator.py:65-119) + inverse_stft (stft.py:138-148) ----
def test_istft_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "istft_strong.npz"))
    sd = ho.make_state_dict(ho.ISTFT_CONFIG, int(g["seed"]), "strong")
    with torch.no_grad():
        spec, phase = ho.istft_forward(sd, ho.ISTFT_CONFIG, torch.from_numpy(g["mel"]))
        wav = ho.inverse_stft(torch.from_numpy(g["spec"]), torch.from_numpy(g["phase"]), 16, 4, 16)
    assert spec.shape == g["spec"].shape == (2, 9, 64 * 6 + 1)
    assert (spec.log() - torch.from_numpy(g["spec"]).log()).abs().max().item() < 1e-4
    assert (phase - torch.from_numpy(g["phase"])).abs().max().item() < 1e-4
    assert torch.equal(wav, torch.from_numpy(g["wav"])) and wav.shape == (2, 1, 256 * 6)
    d = ho.inverse_stft_def(g["spec"], g["phase"], 16, 4)
    assert np.abs(d - g["wav"][:, 0]).max() < 1e-5 * np.abs(g["wav"]).max()


def test_istft_selects_resblock2_for_the_shipped_int_config():
    """generator.py:71 compares config['resblock'] with the string '1'."""
    names = [n for n, *_ in ho.layer_names(ho.ISTFT_CONFIG)]
    assert "resblocks.0.convs.1" in names and "resblocks.0.convs1.0" not in names and len(names) == 16
    cfg = dict(ho.ISTFT_CONFIG, resblock="1")
    assert "resblocks.0.convs1.2" in [n for n, *_ in ho.layer_names(cfg)]


def test_oracle_intermediates_against_reference_hooks():
    """per-layer goldens.  tests/golden/voc_taps_strong.npz holds the outputs of forward hooks on the
    UNMODIFIED reference generator's conv_pre, ups[i], resblocks[n] and conv_post (oracle/make_golden_taps.py); the
    oracle's named intermediates must match them layer by layer."""
    import os
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "voc_taps_strong.npz"))
    sd = ho.make_state_dict(ho.DEFAULT_CONFIG, int(g["seed"]), "strong")
    taps = {}
    with torch.no_grad():
        wav = ho.hifigan_forward(sd, ho.DEFAULT_CONFIG, torch.from_numpy(g["mel"]), taps=taps)
    assert (wav - torch.from_numpy(g["wav"])).abs().max().item() <= 1e-5
    names = [k[4:] for k in g.files if k.startswith("tap:")]
    assert len(names) == 1 + 4 + 12 + 1
    for n in names:
        ref = torch.from_numpy(g["tap:" + n])
        assert taps[n].shape == ref.shape, n
        assert (taps[n] - ref).abs().max().item() <= 1e-5 * max(1.0, ref.abs().max().item()), n
