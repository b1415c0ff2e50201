"""The synthetic checkpoint generator of the package (bench / smoke) is bit-identical to the one the oracle used for
the committed goldens, and the staged reference (baseline/_ref, when present) agrees with the oracle restatement."""
import os

import pytest
import torch

from e2e_tts_b200 import synthetic as sy
from oracle import hifigan_oracle as ho
from oracle import ref_loader


@pytest.mark.parametrize("regime", ["default", "strong"])
def test_state_dicts_identical_to_the_oracles(regime):
    for cfg_a, cfg_b in ((sy.DEFAULT_CONFIG, ho.DEFAULT_CONFIG), (sy.ISTFT_CONFIG, ho.ISTFT_CONFIG)):
        assert cfg_a == cfg_b
        a, b = sy.make_state_dict(cfg_a, 5, regime), ho.make_state_dict(cfg_b, 5, regime)
        assert list(a) == list(b)
        assert all(torch.equal(a[k], b[k]) for k in a)


def test_mel_like_is_seeded():
    assert torch.equal(sy.mel_like(2, 9, 3), sy.mel_like(2, 9, 3))
    assert sy.mel_like(2, 9, 3).shape == (2, 80, 9)


def test_staged_reference_matches_the_oracle():
    m_cls = ref_loader.reference_hifigan_class()
    if m_cls is None:
        pytest.skip("baseline/_ref not staged (python -m oracle.build_ref needs /root/reference)")
    sd = sy.make_state_dict(sy.DEFAULT_CONFIG, 9, "strong")
    m = ref_loader.build_reference_hifigan(sy.DEFAULT_CONFIG, sd)
    mel = sy.mel_like(1, 12, 4)
    with torch.no_grad():
        want = m(mel)
        got = ho.hifigan_forward(sd, sy.DEFAULT_CONFIG, mel)
    assert (got - want).abs().max().item() <= 1e-5
