"""Drop-in STFT -> log-mel front-end of e2e-tts on B200.

Mirrors e2e_tts/src/tools/stft.py:11-89 (`TorchSTFT`), :107-135 (`generate_melspecs`) and
e2e_tts/src/tools/utils.py:22-37 (`dynamic_range_compression/decompression`) as their callers use them
(tools_for_data.py:108-115,178-179; dataloader.py:82-86,155,344-346,372-373; textgrid2durations.py:101-103,137):

    stft = TorchSTFT(1024, 256, 1024, 80, 22050, 0.0, 8000.0)
    mel, energy = stft.mel_spectrogram(audio[B, L], return_energy=True)   # [B, 80, T], [B, T]

One fused CUDA kernel does reflect padding, Hann windowing, the 1024-point real FFT, magnitude, the sparse mel
filterbank, log-compression and the frame energy in a single pass over the audio.  CPU tensors are accepted, as
in the reference call sites: they are staged to the current CUDA device and the results come back as CPU tensors
(so the callers' `.numpy()` keeps working); CUDA tensors stay on the device.  There is no CPU implementation.
"""
from __future__ import annotations

import ctypes
import warnings
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _native


def dynamic_range_compression(x, C=1, clip_val=1e-5):
    """utils.py:22-28."""
    return torch.log(torch.clamp(x, min=clip_val) * C)


def dynamic_range_decompression(x, C=1):
    """utils.py:31-37."""
    return torch.exp(x) / C


def slaney_mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: Optional[float]) -> np.ndarray:
    """The filterbank the reference obtains from `librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)`
    (stft.py:34-40; librosa==0.9.2 defaults: Slaney mel scale, Slaney area normalisation, float32)."""
    if fmax is None:
        fmax = sr / 2.0
    f_sp = 200.0 / 3.0
    brk_hz, logstep = 1000.0, np.log(6.4) / 27.0
    brk_mel = brk_hz / f_sp

    def hz2mel(f):
        return brk_mel + np.log(f / brk_hz) / logstep if f >= brk_hz else f / f_sp

    def mel2hz(m):
        return np.where(m >= brk_mel, brk_hz * np.exp(logstep * (m - brk_mel)), f_sp * m)

    edges = mel2hz(np.linspace(hz2mel(float(fmin)), hz2mel(float(fmax)), n_mels + 2))   # band edges in Hz
    bins = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    width = np.diff(edges)
    rising = (bins[None, :] - edges[:-2, None]) / width[:-1, None]
    falling = (edges[2:, None] - bins[None, :]) / width[1:, None]
    fb = np.maximum(0.0, np.minimum(rising, falling)).astype(np.float32)
    fb *= (2.0 / (edges[2:] - edges[:-2]))[:, None]
    return fb


class _MelHandle:
    """Owns one e2e_mel* (device-side window, twiddles and sparse filterbank) on one CUDA device."""

    def __init__(self, n_fft: int, hop: int, win: int, n_mels: int, basis: np.ndarray, device: torch.device):
        self.device = device
        basis = np.ascontiguousarray(basis, dtype=np.float32)
        h = ctypes.c_void_p()
        with torch.cuda.device(device):
            rc = _native.lib().e2e_mel_create(n_fft, hop, win, n_mels,
                                              basis.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), ctypes.byref(h))
        _native.check(rc, "e2e_mel_create")
        self.h = h
        self.n_mels = n_mels
        self.n_fft, self.hop = int(n_fft), int(hop)

    def num_frames(self, L: int) -> int:
        return int(_native.lib().e2e_mel_num_frames(self.h, L))

    def run(self, wav: torch.Tensor, want_energy: bool, check_range: bool):
        B, L = wav.shape
        T = self.num_frames(L)
        if T < 1:
            raise ValueError("input too short: need more than %d samples (the reflect padding, (n_fft - hop) / 2)"
                             % ((self.n_fft - self.hop) // 2))
        dev = wav.device
        mel = torch.empty((B, self.n_mels, T), dtype=torch.float32, device=dev)
        energy = torch.empty((B, T), dtype=torch.float32, device=dev) if want_energy else None
        flag = torch.zeros(1, dtype=torch.int32, device=dev) if check_range else None
        with torch.cuda.device(dev):
            rc = _native.lib().e2e_mel_forward(self.h, wav.data_ptr(), B, L, wav.stride(0), mel.data_ptr(),
                                               energy.data_ptr() if want_energy else None,
                                               flag.data_ptr() if check_range else None,
                                               torch.cuda.current_stream(dev).cuda_stream)
        _native.check(rc, "e2e_mel_forward")
        return mel, energy, flag

    def __del__(self):
        try:
            if self.h:
                _native.lib().e2e_mel_destroy(self.h)
                self.h = None
        except Exception:
            pass


def _prepare(x: torch.Tensor) -> Tuple[torch.Tensor, bool]:
    """[B, L] float32 on a CUDA device, unit stride along L.  Returns (tensor, came_from_cpu)."""
    if not isinstance(x, torch.Tensor) or x.dim() != 2:
        raise ValueError("expected a [B, L] tensor")
    was_cpu = not x.is_cuda
    if was_cpu:
        if not torch.cuda.is_available():
            raise RuntimeError("e2e_tts_b200 mel front-end needs a CUDA (sm_100a) device; there is no CPU path")
        x = x.detach().to(torch.float32)
        x = (x if x.is_pinned() else x.contiguous()).to("cuda", non_blocking=True)
    else:
        x = x.detach()
        if x.dtype != torch.float32:
            x = x.float()
    if x.shape[1] > 1 and x.stride(1) != 1:
        x = x.contiguous()
    return x, was_cpu


class TorchSTFT(nn.Module):
    """stft.py:11-89.  Same constructor (positional use at dataloader.py:82-86) and public attributes."""

    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050,
                 mel_fmin=0.0, mel_fmax=8000.0, device=None):
        super().__init__()
        self.device = "cpu" if device is None else device
        self.sampling_rate = sampling_rate
        self.n_mel_channels = n_mel_channels
        self.filter_length = filter_length
        self.hop_length = hop_length
        self.win_length = win_length
        self.fmin = mel_fmin
        self.fmax = mel_fmax
        self.stft_pad = (int((filter_length - hop_length) / 2), int((filter_length - hop_length) / 2))
        basis = slaney_mel_filterbank(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax)
        self.register_buffer("mel_basis", torch.from_numpy(basis).float())
        self.window = torch.hann_window(win_length).to(self.device)
        self._handles: Dict[str, _MelHandle] = {}

    def __getstate__(self):
        d = self.__dict__.copy()   # copies / pickles start without native handles
        d["_handles"] = {}
        return d

    def _handle(self, device: torch.device) -> _MelHandle:
        key = str(device)
        h = self._handles.get(key)
        if h is None:
            h = _MelHandle(self.filter_length, self.hop_length, self.win_length, self.n_mel_channels,
                           self.mel_basis.detach().cpu().numpy(), device)
            self._handles[key] = h
        return h

    def inverse_tranform(self, magnitude: torch.Tensor, phase: torch.Tensor) -> torch.Tensor:
        """stft.py:90-100 (name as spelt there): torch.istft(magnitude * exp(i * phase)) with this module's sizes, via
        inverse_stft (supported for the small transforms of the iSTFTNet head, see there)."""
        return inverse_stft(magnitude, phase, self.filter_length, self.hop_length, self.win_length)

    def mel_spectrogram(self, input_data, center=False, return_energy=False, check_range=True):
        """stft.py:46-89.  input_data: [B, L] in [-1, 1] -> log-mel [B, n_mel_channels, T] (and energy [B, T]).
        Raises AssertionError for out-of-range samples like the reference (:56-57); `check_range=False` skips
        the (synchronising) flag read."""
        if center:
            raise NotImplementedError("center=True is not used by any e2e-tts caller and is not implemented")
        x, was_cpu = _prepare(input_data)
        mel, energy, flag = self._handle(x.device).run(x, return_energy, check_range)
        if check_range:
            assert int(flag.item()) == 0, "input samples must lie in [-1, 1]"
        if was_cpu:
            mel = mel.cpu()
            energy = energy.cpu() if energy is not None else None
        if return_energy is True:
            return mel, energy
        return mel


def crop_segments_and_mel(stft: "TorchSTFT", audio: torch.Tensor, lengths=None, segment_size: int = 8192, starts=None,
                          generator: Optional[torch.Generator] = None, check_range: bool = False):
    """Batched, on-device form of the vocoder-training crop + mel of MelAudioLoader.__getitem__
    (e2e_tts/src/tools/dataloader.py:364-373, the `load_mel_from_disk=False` branch), which the reference runs per item
    inside forked DataLoader workers (where a CUDA kernel cannot run - SURVEY.md §7): move it into the training step.

        audio    [B, Lmax] float32 CUDA tensor, already divided by max_wav_value (row b valid up to lengths[b])
        lengths  [B] ints (None: every row is Lmax long)
        starts   [B] ints (None: drawn like the reference, uniform in [0, len - segment_size] per row, from `generator`)
    returns (mel [B, n_mels, segment_size // hop], audio_seg [B, segment_size], mel_loss) - per row exactly
    `audio[start:start+segment_size]` (zero-padded at the end when the clip is shorter, dataloader.py:370) and
    `stft.mel_spectrogram` of it; the reference computes the same mel twice (mel, mel_loss), so one tensor is returned
    for both.  Crop = torch indexing on the device, mel = the CUDA kernel (e2e_mel_forward)."""
    if not isinstance(audio, torch.Tensor) or audio.dim() != 2 or not audio.is_cuda:
        raise ValueError("expected a [B, Lmax] CUDA tensor (there is no CPU path)")
    B, Lmax = audio.shape
    dev = audio.device
    seg = int(segment_size)
    if lengths is None:
        lens = torch.full((B,), Lmax, dtype=torch.int64, device=dev)
    else:
        lens = torch.as_tensor(lengths).to(device=dev, dtype=torch.int64)
        if lens.shape != (B,) or int(lens.max()) > Lmax or int(lens.min()) < 0:
            raise ValueError("lengths must be [B] values in [0, Lmax]")
    if starts is None:
        span = (lens - seg).clamp(min=0) + 1                      # random.randint(0, max_audio_start) is inclusive
        u = torch.rand(B, generator=generator, device=dev if generator is None or generator.device.type == "cuda" else "cpu")
        st = (u.to(dev).double() * span.double()).floor().to(torch.int64).clamp(max=span - 1)
    else:
        st = torch.as_tensor(starts).to(device=dev, dtype=torch.int64)
        if st.shape != (B,) or bool((st < 0).any()) or bool((st > (lens - seg).clamp(min=0)).any()):
            raise ValueError("starts must be [B] values in [0, max(len - segment_size, 0)]")
    idx = st[:, None] + torch.arange(seg, device=dev)[None, :]
    valid = idx < lens[:, None]
    audio_seg = torch.where(valid, audio.float().gather(1, idx.clamp(max=Lmax - 1)), audio.new_zeros((), dtype=torch.float32))
    mel = stft.mel_spectrogram(audio_seg, check_range=check_range)
    return mel, audio_seg, mel


_GM_HANDLES: Dict[tuple, _MelHandle] = {}


def generate_melspecs(y, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0.0,
                      fmax=8000.0, center=False) -> torch.Tensor:
    """stft.py:107-135: functional twin of TorchSTFT.mel_spectrogram that only WARNS on out-of-range input."""
    if center:
        raise NotImplementedError("center=True is not used by any e2e-tts caller and is not implemented")
    x, was_cpu = _prepare(y)
    key = (n_fft, num_mels, sampling_rate, hop_size, win_size, float(fmin), None if fmax is None else float(fmax),
           str(x.device))
    h = _GM_HANDLES.get(key)
    if h is None:
        h = _MelHandle(n_fft, hop_size, win_size, num_mels,
                       slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax), x.device)
        _GM_HANDLES[key] = h
    mel, _, flag = h.run(x, False, True)
    if int(flag.item()) != 0:
        warnings.warn("input has samples outside [-1, 1] (min %g, max %g)" % (float(x.min()), float(x.max())))
    return mel.cpu() if was_cpu else mel


def inverse_stft(magnitude: torch.Tensor, phase: torch.Tensor, n_fft: int = 1024, hop_size: int = 256,
                 win_size: int = 1024) -> torch.Tensor:
    """stft.py:138-148: torch.istft(magnitude * exp(i * phase), n_fft, hop_size, win_size, hann_window(win_size))
    .unsqueeze(-2).  magnitude, phase: [B, n_fft/2 + 1, frames] fp32 CUDA tensors -> [B, 1, hop_size * (frames - 1)].
    Runs the small-transform CUDA kernel (win_size == n_fft = 2^m <= 64, the iSTFTNet head's 16 / 4 / 16); other sizes
    are outside the synthesis path and raise (no CPU fallback)."""
    if magnitude.shape != phase.shape or magnitude.dim() != 3 or magnitude.shape[1] != n_fft // 2 + 1:
        raise ValueError("expected magnitude and phase of shape [B, %d, frames]" % (n_fft // 2 + 1))
    if not magnitude.is_cuda or not phase.is_cuda:
        raise RuntimeError("e2e_tts_b200.inverse_stft runs on CUDA (sm_100a) only; there is no CPU path")
    mag = magnitude.detach().float().contiguous()
    ph = phase.detach().float().contiguous()
    B, _, frames = mag.shape
    if B == 0:
        return mag.new_zeros((0, 1, hop_size * max(frames - 1, 0)))
    out = torch.empty((B, 1, hop_size * (frames - 1)), dtype=torch.float32, device=mag.device)
    with torch.cuda.device(mag.device):
        stream = torch.cuda.current_stream(mag.device).cuda_stream
        rc = _native.lib().e2e_istft_forward(mag.data_ptr(), ph.data_ptr(), B, frames, int(n_fft), int(hop_size),
                                             int(win_size), out.data_ptr(), stream)
        _native.check(rc, "e2e_istft_forward")
    return out

