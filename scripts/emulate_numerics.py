"""CPU emulation of where the bf16 roundings of the GPU path sit, to see what each one costs against the fp64 reference
(TEST / DESIGN TOOL; uses the oracle).  Convolutions run in fp64 on bf16-rounded operands (= exact products, wide
accumulation, like the tensor core's fp32 accumulate to first order).

    python scripts/emulate_numerics.py [T] [seeds...]

Switches (each True = that tensor is ROUNDED to bf16 when stored):
  res   the residual stream between the pairs of a resblock
  x0    the resblock input used as residual (the stage tensor the upsampler wrote)
  sum   the running sum over the resblocks of a stage
  y     the last stage's output as conv_post's input
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from oracle import hifigan_oracle as ho
from e2e_tts_b200 import synthetic as sy

def bf(x):
    return x.to(torch.bfloat16).to(torch.float64)

def lrelu(x, s=0.1):
    return torch.where(x >= 0, x, x * s)

def run(sd, cfg, mel, res, x0r, sm, yr):
    W = lambda n: bf(ho._weight(sd, n, torch.float64))
    b = lambda n: sd[n + ".bias"].double()
    nk = len(cfg["resblock_kernel_sizes"])
    x = F.conv1d(bf(mel.double()), W("conv_pre"), b("conv_pre"), padding=3)
    for i, (u, k) in enumerate(zip(cfg["upsample_rates"], cfg["upsample_kernel_sizes"])):
        a = bf(lrelu(x))
        x = F.conv_transpose1d(a, W("ups.%d" % i), b("ups.%d" % i), stride=u, padding=(k - u) // 2)
        a0 = bf(lrelu(x))                       # the stage tensor in HBM: bf16 leaky_relu(x)
        xin = torch.where(a0 >= 0, a0, a0 * 10) if x0r else x
        xs = None
        for j in range(nk):
            n = i * nk + j
            ks = cfg["resblock_kernel_sizes"][j]
            dil = cfg["resblock_dilation_sizes"][j]
            y = xin
            a = a0
            for m in range(3):
                p1, p2 = "resblocks.%d.convs1.%d" % (n, m), "resblocks.%d.convs2.%d" % (n, m)
                xt = F.conv1d(a, W(p1), b(p1), dilation=dil[m], padding=ho.get_padding(ks, dil[m]))
                xt = bf(lrelu(xt))
                xt = F.conv1d(xt, W(p2), b(p2), padding=ho.get_padding(ks, 1))
                y = xt + y
                a = bf(lrelu(y))
                if res:
                    y = torch.where(a >= 0, a, a * 10)    # residual recovered from the stored bf16 activation
            if xs is None:
                xs = y
            else:
                xs = xs + y
            if sm and j + 1 < nk:
                xs = bf(xs)
        x = xs / nk
    last = lrelu(x, 0.01)
    if yr:
        last = bf(last)
    out = torch.tanh(F.conv1d(last, ho._weight(sd, "conv_post", torch.float64), b("conv_post"), padding=3))
    return out

def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    seeds = [int(s) for s in sys.argv[2:]] or [1, 7, 21]
    cfg = sy.DEFAULT_CONFIG
    variants = [("V0 current: res,x0,sum,y bf16", True, True, True, True),
                ("V1 fp32 residual inside resblock", False, True, True, True),
                ("V2 + fp32 resblock sum", False, True, False, True),
                ("V3 + fp32 conv_post input", False, True, False, False),
                ("V4 + fp32 stage residual (operand roundings only)", False, False, False, False),
                ("V5 current but fp32 sum", True, True, False, True),
                ("V6 current but fp32 sum + fp32 y", True, True, False, False)]
    for seed in seeds:
        sd = sy.make_state_dict(cfg, seed, "strong")
        mel = sy.mel_like(2, T, seed + 100)
        ref = ho.hifigan_forward(sd, cfg, mel, dtype=torch.float64)
        scale = ref.abs().max().item()
        print("seed %d T %d scale %.3f" % (seed, T, scale))
        for name, r, x0, s, y in variants:
            out = run(sd, cfg, mel, r, x0, s, y)
            d = (out - ref).abs()
            print("   %-52s max %.3f %%  mean %.4f %%" % (name, 100 * d.max().item() / scale, 100 * d.mean().item() / scale))

main()
