"""CPU oracle for the STFT -> mel front-end — TEST INFRASTRUCTURE ONLY (see oracle/hifigan_oracle.py header).

Restates, line for line, e2e_tts/src/tools/stft.py:33,44,56-89 (TorchSTFT.mel_spectrogram),
stft.py:107-135 (generate_melspecs) and e2e_tts/src/tools/utils.py:22-28 (dynamic_range_compression) with the
same eager torch CPU ops the reference calls (F.pad reflect, torch.stft, matmul, clamp/log, norm).

The one piece that is NOT in /root/reference is the mel filterbank: stft.py:34-40 calls
`librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` from the third-party dependency librosa==0.9.2
(requirements.txt:6), which is not installed here and cannot be fetched.  `slaney_mel_basis` restates that
release's published algorithm (htk=False Slaney mel scale, norm="slaney" area normalisation, float32
result).  PARITY OF THE BASIS IS UNPINNED by the reference (it stores no basis and has no test); it is anchored
on the algorithm definition and cross-checked against torchaudio.functional.melscale_fbanks in
tests/test_oracle.py.

`mel_spectrogram_f64` is a definition-level float64 numpy version (explicit frames, periodic Hann, rfft) used
to state tolerances.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------------------------------
# librosa 0.9.2  filters.mel  (Slaney scale, Slaney norm)
# ----------------------------------------------------------------------------------------------------
def _hz_to_mel(f):
    f = np.asanyarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-300) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asanyarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    freqs = f_sp * m
    min_log_hz = 1000.0
    min_log_mel = min_log_hz / f_sp
    logstep = np.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), freqs)


def slaney_mel_basis(sr=22050, n_fft=1024, n_mels=80, fmin=0.0, fmax=8000.0) -> np.ndarray:
    if fmax is None:
        fmax = sr / 2.0
    weights = np.zeros((n_mels, 1 + n_fft // 2), dtype=np.float32)
    fftfreqs = np.linspace(0, sr / 2.0, 1 + n_fft // 2, endpoint=True)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


# ----------------------------------------------------------------------------------------------------
# stft.py:46-89
# ----------------------------------------------------------------------------------------------------
def dynamic_range_compression(x, C=1, clip_val=1e-5):
    return torch.log(torch.clamp(x, min=clip_val) * C)  # utils.py:22-28


def mel_spectrogram(wav: torch.Tensor, n_fft=1024, hop=256, win=1024, n_mels=80, sr=22050, fmin=0.0, fmax=8000.0,
                    return_energy=False, check_range=True, basis: np.ndarray | None = None):
    """wav [B, L] float32 in [-1, 1] -> log-mel [B, n_mels, T] (and energy [B, T])."""
    if check_range:
        assert torch.min(wav) >= -1 and torch.max(wav) <= 1          # stft.py:56-57
    if basis is None:
        basis = slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)
    mel_basis = torch.from_numpy(basis).float()
    pad = int((n_fft - hop) / 2)                                      # stft.py:33
    x = F.pad(wav.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)  # :60-64
    window = torch.hann_window(win)                                   # :44 (periodic)
    spec = torch.stft(x, n_fft=n_fft, hop_length=hop, win_length=win, window=window, center=False,
                      pad_mode="reflect", normalized=False, onesided=True, return_complex=True)  # :65-76
    spec = torch.view_as_real(spec)
    mag = torch.sqrt(spec.pow(2).sum(-1) + 1e-9)                      # :77
    mel = dynamic_range_compression(torch.matmul(mel_basis, mag))     # :80-81
    if return_energy:
        return mel, torch.norm(mag, dim=1)                            # :84
    return mel


def mel_spectrogram_f64(wav: np.ndarray, n_fft=1024, hop=256, n_mels=80, sr=22050, fmin=0.0, fmax=8000.0,
                        basis: np.ndarray | None = None):
    """Definition-level float64 version (SURVEY.md §8 a'7-9).  Returns (log-mel, linear mel, energy)."""
    wav = np.asarray(wav, dtype=np.float64)
    if basis is None:
        basis = slaney_mel_basis(sr, n_fft, n_mels, fmin, fmax)
    basis = basis.astype(np.float64)
    pad = (n_fft - hop) // 2
    B, L = wav.shape
    xp = np.pad(wav, ((0, 0), (pad, pad)), mode="reflect")
    T = 1 + (L + 2 * pad - n_fft) // hop
    n = np.arange(n_fft)
    w = 0.5 - 0.5 * np.cos(2 * np.pi * n / n_fft)
    idx = hop * np.arange(T)[:, None] + n[None, :]
    frames = xp[:, idx] * w                                           # [B, T, n_fft]
    X = np.fft.rfft(frames, axis=-1)                                  # [B, T, n_fft/2+1]
    mag = np.sqrt(X.real ** 2 + X.imag ** 2 + 1e-9)
    mel = np.einsum("mk,btk->bmt", basis, mag)
    energy = np.sqrt((mag ** 2).sum(-1))
    return np.log(np.maximum(mel, 1e-5)), mel, energy


def num_frames(L: int, n_fft=1024, hop=256) -> int:
    return 1 + (L + 2 * ((n_fft - hop) // 2) - n_fft) // hop
