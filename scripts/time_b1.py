"""B = 1 forward loop for the ncu launch list: python scripts/time_b1.py [B] [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import e2e_tts_b200 as pkg
from e2e_tts_b200 import synthetic as sy
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
voc = pkg.HifiGan(sy.DEFAULT_CONFIG); voc.load_state_dict(sy.make_state_dict(sy.DEFAULT_CONFIG, 1, "strong")); voc = voc.eval().cuda()
mels = [sy.mel_like(B, 431, 500 + i).cuda() for i in range(4)]
with torch.no_grad():
    for i in range(iters):
        voc(mels[i % 4])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        voc(mels[i % 4])
    e1.record(); torch.cuda.synchronize()
print("B=%d: %.4f ms per forward (fresh output buffers: eager launches)" % (B, e0.elapsed_time(e1) / iters))
