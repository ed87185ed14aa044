#!/usr/bin/env python
"""CPU emulation of operand formats for the tcgen05 conv stack (planning aid for the pass count; no GPU needed).

The parity path stores every activation / weight as hi + lo/2048 (two fp16) and spends 3 tensor-core passes per flop
(A_hi*W_hi, A_hi*W_lo, A_lo*W_hi).  This script measures, on the BN-calibrated golden weights and count-like inputs, the
value-map / obstacle-map error against the fp32 oracle of cheaper schemes:

  fp16        : 1 pass, single fp16 plane (the documented fast mode)
  fp16x2      : 3 passes (the shipped parity format)
  fp16+fp8    : the fp16 main product + both correction products evaluated on the FP8 pipe (2x rate): activations rounded to
                e5m2 (fp16's exponent range, no scaling needed), weights to e4m3 with a per-layer power-of-two scale
                -> 1 + 0.5 + 0.5 = 2 pass-equivalents
  fp16+fp8(a) : only the A_lo*W_hi correction in fp8, A_hi*W_lo in fp16 -> 2.5 pass-equivalents
  fp16+mxfp8  : both corrections on the block-scaled FP8 pipe (kind::mxf8f6f4: e4m3 elements, one power-of-two scale per 32
                channels, for activations and weights) -> 2 pass-equivalents

Products are formed in fp32 on the CPU (F.conv2d), i.e. accumulation error is NOT modelled (the kernel's chunked accumulation
keeps it below the operand error).  Usage: python tests/studies/numerics_study.py [S] [seeds...]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import nbp_torch as NT

LO = 2048.0


def split16(x):
    hi = x.to(torch.float16).float()
    lo = ((x - hi) * LO).to(torch.float16).float() / LO
    return hi, lo


def q_e5m2(x):
    return x.to(torch.float8_e5m2).float()


def q_e4m3_scaled(w):
    m = float(w.abs().max())
    if m == 0.0:
        return w
    s = 2.0 ** (8 - int(torch.frexp(torch.tensor(m))[1]))          # |w|*s in [128, 256) at the maximum
    return (w * s).clamp(-448, 448).to(torch.float8_e4m3fn).float() / s


def q_mx_e4m3(x, block=32):
    """OCP MX block scaling as tcgen05 kind::mxf8f6f4 applies it: e4m3 elements, one power-of-two (ue8m0) scale per `block`
    consecutive elements along the channel (K) dimension."""
    n, c = x.shape[0], x.shape[1]
    if c % block:
        return q_e4m3_scaled(x)
    xb = x.reshape(n, c // block, block, *x.shape[2:])
    m = xb.abs().amax(dim=2, keepdim=True).clamp_min(1e-30)
    e = torch.ceil(torch.log2(m / 448.0))                          # smallest power of two that brings the block max under 448
    sc = torch.exp2(e)
    q = (xb / sc).clamp(-448, 448).to(torch.float8_e4m3fn).float() * sc
    return q.reshape(x.shape)


def make_conv(mode):
    def conv(x, sd, prefix, pad):
        w, b = sd[prefix + ".weight"], sd[prefix + ".bias"]
        gemm = w.shape[0] % 32 == 0 and w.shape[1] % 64 == 0       # the layers that run on the tensor cores
        xh, xl = split16(x)
        wh, wl = split16(w)
        if mode == "fp32":
            return F.conv2d(x, w, b, padding=pad)
        if not gemm:                                               # stem / psi / heads: CUDA cores, fp32 math on the stored operands
            xin = xh if mode == "fp16" else xh + xl
            return F.conv2d(xin, w, b, padding=pad)
        main = F.conv2d(xh, wh, None, padding=pad)
        if mode == "fp16":
            out = main
        elif mode == "fp16x2":
            out = main + F.conv2d(xh, wl, None, padding=pad) + F.conv2d(xl, wh, None, padding=pad)
        elif mode == "fp16+fp8":
            c1 = F.conv2d(q_e5m2(xh), q_e4m3_scaled(wl * LO), None, padding=pad) / LO
            c2 = F.conv2d(q_e5m2(xl * LO), q_e4m3_scaled(wh), None, padding=pad) / LO
            out = main + c1 + c2
        elif mode == "fp16+mxfp8":
            c1 = F.conv2d(q_mx_e4m3(xh), q_mx_e4m3(wl * LO), None, padding=pad) / LO
            c2 = F.conv2d(q_mx_e4m3(xl * LO), q_mx_e4m3(wh), None, padding=pad) / LO
            out = main + c1 + c2
        elif mode == "fp16+fp8(a)":
            c1 = F.conv2d(xh, wl, None, padding=pad)
            c2 = F.conv2d(q_e5m2(xl * LO), q_e4m3_scaled(wh), None, padding=pad) / LO
            out = main + c1 + c2
        else:
            raise ValueError(mode)
        return out + b.view(1, -1, 1, 1)
    return conv


def run(S, seed):
    sd = NT.golden_state_dict(seed=seed)
    x = NT.count_like_input(1, S, seed=seed + 100)
    res = {}
    orig = NT._conv
    try:
        outs = {}
        for mode in ("fp32", "fp16", "fp16x2", "fp16+fp8(a)", "fp16+fp8", "fp16+mxfp8"):
            NT._conv = make_conv(mode)
            with torch.no_grad():
                outs[mode] = NT.forward(sd, x)
        r1, r2 = outs["fp32"]
        for mode, (o1, o2) in outs.items():
            if mode == "fp32":
                continue
            res[mode] = (float((o1 - r1).abs().max() / r1.abs().max()), float((o1 - r1).abs().mean()), float((o2 - r2).abs().max()))
    finally:
        NT._conv = orig
    return res


if __name__ == "__main__":
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    seeds = [int(a) for a in sys.argv[2:]] or [9, 10, 11]
    torch.set_num_threads(os.cpu_count() or 1)
    print(f"S = {S}; columns: value-map max-rel error | value-map MAE | obstacle-map max abs error   (bar: 1e-3)")
    for seed in seeds:
        for mode, (e1, mae, e2) in run(S, seed).items():
            print(f"seed {seed:3d}  {mode:12s}  {e1:9.2e}  {mae:9.2e}  {e2:9.2e}")
