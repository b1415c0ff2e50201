"""Quick device timing of the vocoder forward (not the bench): python scripts/time_voc.py [B] [T] [iters]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import e2e_tts_b200 as pkg
from oracle import hifigan_oracle as ho

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 431
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
cfg = ho.DEFAULT_CONFIG
voc = pkg.HifiGan(cfg)
voc.load_state_dict(ho.make_state_dict(cfg, 1, "strong"))
voc = voc.eval().cuda()
g = torch.Generator().manual_seed(0)
mel = (torch.randn(B, 80, T, generator=g) * 2 - 5).clamp(-11.5, 2).cuda()
with torch.no_grad():
    t0 = time.time(); w = voc(mel); torch.cuda.synchronize(); print("first call %.1f ms" % ((time.time() - t0) * 1e3))
    for _ in range(3): voc(mel)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): voc(mel)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
audio_s = B * T * 256 / 22050.0
print("B=%d T=%d: %.3f ms/forward -> %.1f audio-s/s, %.1f TFLOP/s (%.1f%% of 1360.2)" %
      (B, T, ms, audio_s / ms * 1e3, B * T * 614105088 / ms * 1e-9, B * T * 614105088 / ms * 1e-9 / 13.602))
print("wav absmax %.3f finite %s launches %d" % (w.abs().max().item(), bool(torch.isfinite(w).all()), voc.launches_per_forward()))
