"""Quick device timing of the Postnet (N2): python scripts/time_postnet.py [B] [T] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import e2e_tts_b200 as pkg
from oracle import postnet_oracle as pno
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 431
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
m = pkg.Postnet(80, pno.DEFAULT_CONFIG)
m.load_state_dict(pno.make_state_dict(80, pno.DEFAULT_CONFIG, 1))
m = m.eval().to("cuda")
x = (torch.randn(B, T, 80, device="cuda") * 2 - 5).clamp(-11.5, 2)
with torch.no_grad():
    for _ in range(5): y = m(x, add_input=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): y = m(x, add_input=True)
    e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / iters
flops = 2.0 * B * T * 5 * (80 * 512 + 3 * 512 * 512 + 512 * 80)
print("Postnet B=%d T=%d: %.4f ms -> %.1f TFLOP/s algorithmic (7 launches; %d mel frames/s)" % (B, T, ms, flops / ms * 1e-9, B * T / ms * 1e3))
