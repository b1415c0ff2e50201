"""Quick device timing of the mel front-end (config 5 shape): python scripts/time_mel.py [B] [L] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import e2e_tts_b200 as pkg
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
L = int(sys.argv[2]) if len(sys.argv) > 2 else 220500
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
wav = torch.rand(B, L, device="cuda") * 2 - 1
stft = pkg.TorchSTFT()
for _ in range(3): stft.mel_spectrogram(wav, return_energy=True, check_range=False)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters): mel, en = stft.mel_spectrogram(wav, return_energy=True, check_range=False)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
T = mel.shape[-1]
bytes_ = B * (4 * L + 4 * 81 * T)
print("B=%d L=%d T=%d: %.3f ms -> %.0f audio-s/s, %.1f GB/s algorithmic (%.1f%% of 6544.3)" %
      (B, L, T, ms, B * L / 22050 / ms * 1e3, bytes_ / ms * 1e-6, bytes_ / ms * 1e-6 / 65.443))
