"""Per-utterance device time over the batch size (L2 residency of the stage tensors vs wave quantisation):
python scripts/time_batch_sweep.py [T]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import e2e_tts_b200 as pkg
from e2e_tts_b200 import synthetic as sy
import bench
T = int(sys.argv[1]) if len(sys.argv) > 1 else 431
voc = pkg.HifiGan(sy.DEFAULT_CONFIG); voc.load_state_dict(sy.make_state_dict(sy.DEFAULT_CONFIG, 1, "strong")); voc = voc.eval().cuda()
for rep in range(2):
    for B in (4, 6, 8, 10, 12, 16, 24, 32, 64):
        mels = [sy.mel_like(B, T, 500 + i).cuda() for i in range(4)]
        iters = max(20, int(2000 / (0.32 * B)))     # ~2 s per point: the sustained (power-capped) regime
        ms = bench.time_forward(voc, mels, iters)
        print("B=%3d: %.4f ms per forward, %.4f ms per utterance, %.0f audio-s/s" % (B, ms, ms / B, B * T * 256 / 22050 / ms * 1e3), flush=True)
        voc._workspaces.clear(); torch.cuda.empty_cache()
