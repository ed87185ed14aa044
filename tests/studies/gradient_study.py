#!/usr/bin/env python
"""CPU emulation of where the whole-network gradient error of the train path comes from (planning aid, no GPU).

The train path keeps conv operands (activations, weights, dz) as fp16x2 (hi + lo/2048: ~22 significant bits), stores the raw conv
output z in the same format, and moves gradients between layers in fp32.  Measured on B200: every backward kernel matches fp64
autograd to <= 4e-6 on its own inputs, yet the whole-network gradient is 4e-3..5e-3 from the fp64 oracle (fp32 torch: 7e-4..4e-3).
Here the network runs in float64 with explicit rounding at exactly those storage points, so each can be switched off separately:

    A  operands of the forward convs (activations / weights)        Z  storage of the raw conv output z (read by BatchNorm fwd + bwd)
    G  dz as the operand of dgrad / wgrad                            F  fp32 results of dgrad / wgrad and fp32 gradients between layers

Tensor-core accumulation error is not modelled.  Usage: python tests/studies/gradient_study.py [B] [S]"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
from oracle import nbp_torch as NT


def r_planes(t, planes):
    """sum of `planes` fp16 numbers, each holding the residual of the previous ones scaled by 2048 (the fp16x2 / x3 formats)"""
    out = torch.zeros_like(t)
    res = t
    scale = 1.0
    for _ in range(planes):
        q = (res * scale).to(torch.float16).to(t.dtype) / scale
        out = out + q
        res = t - out
        scale *= 2048.0
    return out


def r_scaled(t, planes):
    """the same with the per-tensor power-of-two scale the gradient operands get (amax -> [128, 256))"""
    m = float(t.abs().max())
    if m == 0.0 or planes is None:
        return t
    s = 2.0 ** (8 - int(torch.frexp(torch.tensor(m))[1]))
    return r_planes(t * s, planes) / s


RND = {None: lambda t: t, 24: lambda t: t.float().double(), 22: lambda t: r_planes(t, 2), 33: lambda t: r_planes(t, 3)}


class QConv(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b, pad, cfg):
        xq, wq = RND[cfg["A"]](x), RND[cfg["A"]](w)
        z = F.conv2d(xq, wq, b, padding=pad)
        ctx.save_for_backward(xq, wq)
        ctx.pad, ctx.cfg = pad, cfg
        return RND[cfg["Z"]](z)

    @staticmethod
    def backward(ctx, dz):
        xq, wq = ctx.saved_tensors
        cfg = ctx.cfg
        dzq = dz if cfg["G"] is None else (dz.float().double() if cfg["G"] == 24 else r_scaled(dz, 2 if cfg["G"] == 22 else 3))
        with torch.enable_grad():
            x_, w_ = xq.detach().requires_grad_(True), wq.detach().requires_grad_(True)
            gx, gw = torch.autograd.grad(F.conv2d(x_, w_, None, padding=ctx.pad), (x_, w_), dzq)
        rf = RND[cfg["F"]]
        return rf(gx), rf(gw), dz.sum(dim=(0, 2, 3)), None, None


def step(B, S, cfg, seed=4):
    from test_train_gpu import _loss, _targets
    xb = NT.count_like_input(B, S, seed=seed).double()
    tgt_idx, tgt_val, layout = _targets(B, S)
    sd = {k: (v.double() if v.is_floating_point() else v.clone()) for k, v in NT.golden_state_dict(seed=9).items()}
    spec = {k: kind for k, _, kind in NT.state_dict_spec()}
    params = [k for k in sd if spec[k] in ("param", "conv_w", "conv_b", "bn_w", "bn_b")]
    for k in params:
        sd[k].requires_grad_(True)
    orig = NT._conv
    if cfg is not None:
        NT._conv = lambda x, sd_, prefix, pad: QConv.apply(x, sd_[prefix + ".weight"], sd_[prefix + ".bias"], pad, cfg)
    try:
        p1, p2 = NT.forward(sd, xb, training=True)
        loss = _loss(lambda a, b, c, d: NT.loss(sd, a, b, c, d), p1, p2, tgt_idx, tgt_val.double(), layout.double())
        loss.backward()
    finally:
        NT._conv = orig
    return {k: sd[k].grad.detach().clone() for k in params}, p1.detach()


def seeds_table(B, S, seeds):
    """shipped format vs fp32-like arithmetic over several input seeds: the spread IS the result (single decision flips)."""
    print(f"B={B} S={S}: global gradient error vs float64 per input seed   shipped (A22 Z22 G22 F24) | z in fp32 (A22 Z24) | fp32-like everywhere")
    for seed in seeds:
        g64, _ = step(B, S, None, seed)
        scale = max(float(g.norm()) for g in g64.values())
        keys = [k for k in g64 if float(g64[k].norm()) >= 1e-6 * scale]
        cat = lambda d: torch.cat([d[k].reshape(-1) for k in keys])
        ref = cat(g64)
        errs = [float((cat(step(B, S, cfg, seed)[0]) - ref).norm() / ref.norm())
                for cfg in (dict(A=22, Z=22, G=22, F=24), dict(A=22, Z=24, G=22, F=24), dict(A=24, Z=24, G=24, F=24))]
        print(f"  seed {seed:3d}   {errs[0]:9.2e}   {errs[1]:9.2e}   {errs[2]:9.2e}")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--seeds":
        torch.set_num_threads(os.cpu_count() or 1)
        seeds_table(int(sys.argv[2]), int(sys.argv[3]), [int(a) for a in sys.argv[4:]] or [4, 5, 6, 7, 8, 9])
        sys.exit(0)
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    torch.set_num_threads(os.cpu_count() or 1)
    g64, o64 = step(B, S, None)
    scale = max(float(g.norm()) for g in g64.values())
    keys = [k for k in g64 if float(g64[k].norm()) >= 1e-6 * scale]
    cat = lambda d: torch.cat([d[k].reshape(-1) for k in keys])
    ref = cat(g64)
    rows = [("shipped: A22 Z22 G22 F24", dict(A=22, Z=22, G=22, F=24)),
            ("z stored in fp32: A22 Z24 G22 F24", dict(A=22, Z=24, G=22, F=24)),
            ("z exact: A22 Z-- G22 F24", dict(A=22, Z=None, G=22, F=24)),
            ("operands 33 bit, z 22: A33 Z22 G22 F24", dict(A=33, Z=22, G=22, F=24)),
            ("operands 33 bit, z fp32: A33 Z24 G22 F24", dict(A=33, Z=24, G=22, F=24)),
            ("operands exact, z 22: A-- Z22 G22 F24", dict(A=None, Z=22, G=22, F=24)),
            ("only forward operand rounding: A22", dict(A=22, Z=None, G=None, F=None)),
            ("only z storage rounding: Z22", dict(A=None, Z=22, G=None, F=None)),
            ("only backward rounding: G22 F24", dict(A=None, Z=None, G=22, F=24)),
            ("everything fp32-like: A24 Z24 G24 F24", dict(A=24, Z=24, G=24, F=24))]
    print(f"B={B} S={S}: global relative error of all {len(keys)} gradients vs float64 | value-map error")
    for name, cfg in rows:
        g, o = step(B, S, cfg)
        e = float((cat(g) - ref).norm() / ref.norm())
        eo = float((o - o64).abs().max() / o64.abs().max())
        print(f"  {name:45s} {e:9.2e}   {eo:9.2e}")
