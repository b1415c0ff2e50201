"""Quick device timing of the iSTFTNet generator + inverse_stft: python scripts/time_istft.py [B] [T] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import e2e_tts_b200 as pkg
from oracle import hifigan_oracle as ho
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 431
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 20
g = pkg.iSTFT(ho.ISTFT_CONFIG)
g.load_state_dict(ho.make_state_dict(ho.ISTFT_CONFIG, 3, "strong"))
g = g.eval().to("cuda")
mel = (torch.randn(B, 80, T, device="cuda") * 2 - 5).clamp(-11.5, 2)
with torch.no_grad():
    for _ in range(3):
        wav = pkg.inverse_stft(*g(mel), 16, 4, 16)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        wav = pkg.inverse_stft(*g(mel), 16, 4, 16)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
print("iSTFTNet B=%d T=%d: %.3f ms/step -> %.0f audio-s/s (%d launches)" % (B, T, ms, B * 256 * T / 22050 / ms * 1e3, g.launches_per_forward() + 1))
