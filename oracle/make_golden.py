"""Pins the oracle to the UNMODIFIED reference and writes the golden fixtures under tests/golden/.

Runs only in the authoring container (it imports /root/reference, which does not exist on the GPU box):

    python -m oracle.make_golden

1. Vocoder: imports the reference generator exactly as e2e_tts/src/api/inference.py:7 exposes it
   (`sys.path` -> e2e_tts/models; `from vocoder.generator import HifiGan`), loads the synthetic checkpoints
   of oracle.hifigan_oracle.make_state_dict, runs it on seeded inputs and
     (a) asserts the oracle restatement matches it to fp32 round-off,
     (b) stores input + reference output as tests/golden/voc_*.npz.
2. Mel: imports e2e_tts/src/tools/stft.py unmodified.  Its third-party imports that are not installed here
   (librosa, parselmouth, pyworld) and the unrelated `models.g2p` import of tools/utils.py are satisfied by stub
   modules; `librosa.filters.mel` is bound to oracle.mel_oracle.slaney_mel_basis (the restated librosa-0.9.2
   algorithm; see that module's header for why the basis is "parity unpinned").  Stores input + reference
   outputs as tests/golden/mel_*.npz and asserts the oracle matches.
"""
from __future__ import annotations

import os
import sys
import types
import warnings

import numpy as np
import torch

from . import hifigan_oracle as ho
from . import mel_oracle as mo

REF = "/root/reference/e2e_tts"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _import_reference_vocoder():
    sys.path.insert(0, os.path.join(REF, "models"))
    from vocoder.generator import HifiGan  # noqa
    return HifiGan


def _import_reference_stft():
    lib = types.ModuleType("librosa")
    lib.filters = types.ModuleType("librosa.filters")
    lib.filters.mel = lambda sr, n_fft, n_mels=128, fmin=0.0, fmax=None, **kw: mo.slaney_mel_basis(
        sr, n_fft, n_mels, fmin, fmax)
    lib.util = types.ModuleType("librosa.util")
    lib.util.normalize = lambda x, **kw: x
    sys.modules["librosa"] = lib
    sys.modules["librosa.filters"] = lib.filters
    sys.modules["librosa.util"] = lib.util
    for name in ("parselmouth", "pyworld"):
        sys.modules[name] = types.ModuleType(name)
    models = types.ModuleType("models")
    g2p = types.ModuleType("models.g2p")
    g2p._symbols_to_sequence = lambda s: []
    models.g2p = g2p
    sys.modules["models"] = models
    sys.modules["models.g2p"] = g2p
    sys.path.insert(0, os.path.join(REF, "src"))
    from tools.stft import TorchSTFT, generate_melspecs  # noqa
    return TorchSTFT, generate_melspecs


def mel_like(B, T, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(B, 80, T, generator=g) * 2.0 - 5.0).clamp(-11.5, 2.0)


def test_signals(seed: int, B: int, L: int) -> dict:
    g = torch.Generator().manual_seed(seed)
    t = torch.arange(L).float() / 22050.0
    noise = torch.rand(B, L, generator=g) * 2 - 1
    f = 80.0 + 3000.0 * torch.rand(B, 5, 1, generator=g)
    amp = torch.rand(B, 5, 1, generator=g) / 5.0
    tones = (amp * torch.sin(2 * np.pi * f * t[None, None, :])).sum(1)
    tones[:, L // 3: L // 3 + L // 5] = 0.0          # exact-zero span: exercises +1e-9 and the 1e-5 clamp
    full = torch.sign(torch.sin(2 * np.pi * 440.0 * t))[None, :].repeat(B, 1)  # +-1.0 full scale (range edge)
    full[:, :7] = 1.0
    return {"noise": noise, "tones_zero": tones, "fullscale": full}


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    HifiGan = _import_reference_vocoder()

    cfg2 = dict(ho.DEFAULT_CONFIG)
    cfg2["resblock"] = 2
    cases = [("voc_default_init", ho.DEFAULT_CONFIG, "default", 11, 2, 16),
             ("voc_strong_init", ho.DEFAULT_CONFIG, "strong", 12, 2, 16),
             ("voc_strong_resblock2", cfg2, "strong", 13, 1, 24)]
    for name, cfg, regime, seed, B, T in cases:
        sd = ho.make_state_dict(cfg, seed, regime)
        ref = HifiGan(cfg)
        ref.load_state_dict(sd)                      # as utils.py:54-55
        ref.eval()
        mel = mel_like(B, T, seed + 100)
        with torch.no_grad():
            want = ref(mel.transpose(1, 2).contiguous().transpose(1, 2))   # non-contiguous view, utils.py:144
            got = ho.hifigan_forward(sd, cfg, mel)
            got64 = ho.hifigan_forward(sd, cfg, mel, dtype=torch.float64)
        err = (got - want).abs().max().item()
        err64 = (got64.float() - want).abs().max().item()
        print("%s: |ref|max=%.4f  oracle-vs-reference max abs err fp32=%.3g fp64=%.3g" %
              (name, want.abs().max().item(), err, err64))
        assert err < 1e-5 and err64 < 1e-5, "oracle restatement does not match the reference"
        np.savez_compressed(os.path.join(OUT, name + ".npz"), mel=mel.numpy(), wav=want.numpy(), seed=seed,
                            regime=regime, resblock=cfg["resblock"])

    # iSTFTNet head: the reference's class iSTFT (generator.py:65-119) and inverse_stft (stft.py:138-148)
    from vocoder.generator import iSTFT  # noqa  (same sys.path entry as HifiGan)
    cfg = ho.ISTFT_CONFIG
    sd = ho.make_state_dict(cfg, 31, "strong")
    ref = iSTFT(cfg)
    ref.load_state_dict(sd)
    ref.eval()
    mel = mel_like(2, 6, 131)
    with torch.no_grad():
        spec, phase = ref(mel)
        ospec, ophase = ho.istft_forward(sd, cfg, mel)
    e1 = ((ospec.log() - spec.log()).abs().max().item(), (ophase - phase).abs().max().item())
    print("istft_strong: frames=%d oracle-vs-reference log-spec err=%.3g phase err=%.3g" % (spec.shape[-1], e1[0], e1[1]))
    assert max(e1) < 1e-4

    # Postnet (N2): the reference class in eval mode vs the oracle restatement
    from . import postnet_oracle as pno
    sys.path.insert(0, REF)
    from models.acoustic.unsupervised_fastspeech2.layers import Postnet  # noqa
    psd = pno.make_state_dict(80, pno.DEFAULT_CONFIG, 41)
    pref = Postnet(n_channels=80, config=pno.DEFAULT_CONFIG)
    pref.load_state_dict(psd)
    pref.eval()
    px = mel_like(2, 37, 141).transpose(1, 2).contiguous()          # [B, T, 80], log-mel-like values
    with torch.no_grad():
        pwant = pref(px)
        pgot = pno.postnet_forward(psd, pno.DEFAULT_CONFIG, px)
    perr = (pgot - pwant).abs().max().item()
    print("postnet_default: |ref|max=%.3f oracle-vs-reference max abs err=%.3g" % (pwant.abs().max().item(), perr))
    assert perr < 1e-4
    np.savez_compressed(os.path.join(OUT, "postnet_default.npz"), x=px.numpy(), y=pwant.numpy(), seed=41)

    TorchSTFT, generate_melspecs = _import_reference_stft()
    from tools.stft import inverse_stft  # noqa
    with torch.no_grad():
        wav = inverse_stft(spec, phase, 16, 4, 16)
        owav = ho.inverse_stft(spec, phase, 16, 4, 16)
    dwav = ho.inverse_stft_def(spec.numpy(), phase.numpy(), 16, 4)
    print("inverse_stft: len=%d oracle-vs-reference err=%.3g, float64 definition vs reference err=%.3g" %
          (wav.shape[-1], (owav - wav).abs().max().item(), np.abs(dwav - wav[:, 0].numpy()).max()))
    assert torch.equal(owav, wav) and np.abs(dwav - wav[:, 0].numpy()).max() < 1e-5 * float(wav.abs().max())
    np.savez_compressed(os.path.join(OUT, "istft_strong.npz"), mel=mel.numpy(), spec=spec.numpy(), phase=phase.numpy(),
                        wav=wav.numpy(), seed=31)

    stft = TorchSTFT()  # defaults == preprocessing_config.yaml:5-14
    import torchaudio
    fb = torchaudio.functional.melscale_fbanks(513, 0.0, 8000.0, 80, 22050, norm="slaney", mel_scale="slaney").T
    print("mel basis vs torchaudio: max abs diff %.3g" % (fb - torch.from_numpy(mo.slaney_mel_basis())).abs().max())
    for L, seed in ((6000, 21), (8192, 22)):
        for kind, wav in test_signals(seed, 2, L).items():
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mel, energy = stft.mel_spectrogram(wav, return_energy=True)
                mel2 = generate_melspecs(wav)
            assert torch.equal(mel, mel2)
            omel, oen = mo.mel_spectrogram(wav, return_energy=True)
            e1, e2 = (omel - mel).abs().max().item(), (oen - energy).abs().max().item()
            print("mel_%s_L%d: T=%d oracle-vs-reference mel err=%.3g energy err=%.3g" % (kind, L, mel.shape[-1], e1, e2))
            assert e1 == 0.0 and e2 == 0.0
            np.savez_compressed(os.path.join(OUT, "mel_%s_L%d.npz" % (kind, L)), wav=wav.numpy(), mel=mel.numpy(),
                                energy=energy.numpy())
    # the range assert of stft.py:56-57
    try:
        stft.mel_spectrogram(torch.full((1, 2048), 1.5))
        raise SystemExit("reference did not assert on out-of-range input")
    except AssertionError:
        print("reference asserts on |x| > 1: ok")


if __name__ == "__main__":
    main()
