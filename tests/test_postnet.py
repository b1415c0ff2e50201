"""Postnet (SURVEY.md §8 f, N2; reference e2e_tts/models/acoustic/unsupervised_fastspeech2/layers.py:507-563).
CPU: oracle vs the reference golden, BatchNorm folding, state-dict contract.  GPU (marked): parity through the C ABI
(e2e_postnet_forward); tolerance max|y - ref| <= 2e-2 * max|ref|, mean <= 3e-3 * max|ref| (bf16 operands, fp32
accumulate, bf16 tanh activations between the five convolutions)."""
import os

import numpy as np
import pytest
import torch

import e2e_tts_b200 as pkg
from oracle import postnet_oracle as pno

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLD, "postnet_default.npz"))
    sd = pno.make_state_dict(80, pno.DEFAULT_CONFIG, int(g["seed"]))
    with torch.no_grad():
        y = pno.postnet_forward(sd, pno.DEFAULT_CONFIG, torch.from_numpy(g["x"]))
    assert y.shape == g["y"].shape == (2, 37, 80)
    assert (y - torch.from_numpy(g["y"])).abs().max().item() < 1e-4


def test_state_dict_contract_and_batchnorm_fold():
    sd = pno.make_state_dict(80, pno.DEFAULT_CONFIG, 3)
    m = pkg.Postnet(80, pno.DEFAULT_CONFIG)
    assert set(m.state_dict().keys()) == set(sd.keys())          # 5 x (conv w, b, bn w, b, mean, var, count) = 35
    m.load_state_dict(sd)
    m.eval()
    # folded conv == conv followed by eval-mode BatchNorm1d
    x = torch.randn(2, 512, 9)
    w, b = m.folded(2)
    want = m.convolutions[2][1](m.convolutions[2][0].conv(x))
    got = torch.nn.functional.conv1d(x, w, b, padding=2)
    assert (got - want).abs().max().item() < 1e-4
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 4, 80))                                  # CPU tensor: no fallback
    with pytest.raises(RuntimeError):
        m.train()(torch.zeros(1, 4, 80))


@pytest.mark.gpu
@pytest.mark.parametrize("B,T,seed", [(2, 37, 41), (1, 1, 5), (16, 431, 6), (3, 130, 7)])
def test_gpu_parity(B, T, seed):
    sd = pno.make_state_dict(80, pno.DEFAULT_CONFIG, seed)
    m = pkg.Postnet(80, pno.DEFAULT_CONFIG)
    m.load_state_dict(sd)
    m = m.eval().to("cuda")
    if (B, T, seed) == (2, 37, 41):
        g = np.load(os.path.join(GOLD, "postnet_default.npz"))
        x, ref = torch.from_numpy(g["x"]), torch.from_numpy(g["y"])
    else:
        gen = torch.Generator().manual_seed(seed)
        x = (torch.randn(B, T, 80, generator=gen) * 2.0 - 5.0).clamp(-11.5, 2.0)
        with torch.no_grad():
            ref = pno.postnet_forward(sd, pno.DEFAULT_CONFIG, x)
    with torch.no_grad():
        y = m(x.cuda()).cpu()
        y2 = m(x.cuda(), add_input=True).cpu()
    assert y.shape == ref.shape and torch.isfinite(y).all()
    scale = ref.abs().max().item()
    d = (y - ref).abs()
    assert d.max().item() <= 2e-2 * scale and d.mean().item() <= 3e-3 * scale, (d.max().item(), d.mean().item(), scale)
    assert torch.allclose(y2, y + x, atol=1e-6)                  # the caller's `postnet(output) + output`, model.py:188
