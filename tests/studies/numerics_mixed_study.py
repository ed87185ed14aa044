#!/usr/bin/env python
"""CPU emulation of the mixed numeric mode (which layers keep fp16x2, every other GEMM layer = fp16 hi product + e4m3 corrections):
value-map / obstacle-map error against the fp32 oracle.  Planning aid for NBP.full_precision_layers; results in profiles/r02_numerics_mixed.txt.
Usage: python tests/studies/numerics_mixed_study.py S seed [seed ...]"""
import os, sys, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import nbp_torch as NT
LO = 2048.0
def q8(x): return x.clamp(-448, 448).to(torch.float8_e4m3fn).float()
def q8s(w):
    m = float(w.abs().max())
    if m == 0: return w
    s = 2.0 ** (8 - int(torch.frexp(torch.tensor(m))[1]))
    return q8(w * s) / s
FULL = set()
MODE = ["mix"]
def conv(x, sd, prefix, pad):
    w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
    gemm = w.shape[0] % 32 == 0 and w.shape[1] % 64 == 0
    xh = x.to(torch.float16).float(); xl = ((x - xh) * LO).to(torch.float16).float() / LO
    wh = w.to(torch.float16).float(); wl = ((w - wh) * LO).to(torch.float16).float() / LO
    if not gemm:
        return F.conv2d(xh + xl, w, b, padding=pad)
    main = F.conv2d(xh, wh, None, padding=pad)
    if prefix in FULL or MODE[0] == "x2":
        out = main + F.conv2d(xh, wl, None, padding=pad) + F.conv2d(xl, wh, None, padding=pad)
    else:
        xl8 = q8((x - xh) * LO) / LO
        xh8 = q8(x)
        out = main + F.conv2d(xh8, q8s(wl * LO) / LO, None, padding=pad) + F.conv2d(xl8, q8s(wh), None, padding=pad)
    return out + b.view(1, -1, 1, 1)
def err(o, r):
    return float((o[0]-r[0]).abs().max()/r[0].abs().max()), float((o[0]-r[0]).abs().mean()), float((o[1]-r[1]).abs().max())
if __name__ == "__main__":
    torch.set_num_threads(8)
    S = int(sys.argv[1]); seeds = [int(a) for a in sys.argv[2:]]
    sets = {"none": [], "enc123": ["Conv1.conv.3","Conv2.conv.0","Conv2.conv.3","Conv3.conv.0","Conv3.conv.3"],
            "enc12": ["Conv1.conv.3","Conv2.conv.0","Conv2.conv.3"],
            "enc1234": ["Conv1.conv.3","Conv2.conv.0","Conv2.conv.3","Conv3.conv.0","Conv3.conv.3","Conv4.conv.0","Conv4.conv.3"]}
    orig = NT._conv
    for seed in seeds:
        sd = NT.golden_state_dict(seed=seed)
        x = NT.count_like_input(1, S, seed=seed + 100)
        with torch.no_grad(): r = NT.forward(sd, x)
        NT._conv = conv
        for name, L in sets.items():
            FULL.clear(); FULL.update(L)
            with torch.no_grad(): o = NT.forward(sd, x)
            print(seed, name, "%.2e %.2e %.2e" % err(o, r), flush=True)
        NT._conv = orig
