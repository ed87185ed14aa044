"""Train-mode forward and backward of the NBP network on the sm_100a kernels.

Reference semantics: ``NBP.forward`` in ``.train()`` mode (BatchNorm2d batch statistics + running-stat update,
next_best_path/networks/nbp_model.py:8-160) followed by autograd, as ``train_experience_data`` uses it
(next_best_path/utility/nbp_utils.py:378-390).  Exposed as one ``torch.autograd.Function`` so that the reference's own
``loss.backward()`` / ``AdamW.step()`` drive it unchanged; every tensor op inside is a C-ABI call into libnbp_b200.so:

  forward  : conv (tcgen05, raw output + bias) -> batch statistics (fp64 accumulation) -> affine + ReLU; attention gates
             with three train-mode BatchNorms each; CUDA-core stem / heads.
  backward : BatchNorm(+ReLU) backward -> dgrad = the SAME tcgen05 conv kernel on flipped/transposed weights with an fp32
             epilogue; wgrad = tcgen05 GEMM over channel-major operands (wgrad_tc.cu); pooling / upsampling / attention /
             head backward kernels.  Gradients travel as fp32 NHWC; GEMM operands are re-split to fp16x2 with a per-tensor
             power-of-two scale so that small gradients keep 22 significant bits.
"""
from __future__ import annotations

import ctypes

import torch

from .. import _lib
from . import nbp_model as M

_DEC_LEVELS = M._DEC_LEVELS
BN_MOMENTUM, BN_EPS = 0.1, 1e-5
# training keeps the forward / dgrad accumulation chains short: gradients are ill-conditioned (NBP_TRAIN_K_CHUNK: A/B switch)
TRAIN_K_CHUNK = int(__import__("os").environ.get("NBP_TRAIN_K_CHUNK", "3"))
WGRAD_MAX_K_TILES = int(__import__("os").environ.get("NBP_WGRAD_MAX_K_TILES", "0"))     # 0: split K only as far as needed to fill the GPU (bounding the in-TMEM chain showed no accuracy benefit, measured)


def _st():
    return torch.cuda.current_stream().cuda_stream


class _Tape:
    """Everything one forward pass saves for its backward pass."""

    def __init__(self, dev, B, cache=None, momentum=None):
        self.dev, self.B = dev, B
        self.momentum = BN_MOMENTUM if momentum is None else float(momentum)
        self.cache = cache if cache is not None else {}     # packed GEMM weights, valid until the parameters change
        self.L = _lib.lib()
        self.ws = torch.zeros(2 * 2048 + 16, dtype=torch.float64, device=dev)        # fp64 reduction scratch
        self.ws_big = torch.zeros(9 * 16 * 64 + 8 * 256 + 64, dtype=torch.float64, device=dev)
        self.ops = []            # backward closures, run in reverse
        self.capture = None      # dict: the tensors every discrete decision is taken from (module.capture_decisions, tests)
        self.grads = {}          # id(_Act) -> fp32 NHWC dense gradient [npix, c]
        self.pgrads = {}         # parameter name -> fp32 gradient tensor

    def new(self, h, w, c):
        return M._Act(torch.empty((self.B, h, w, 2 * c), dtype=torch.float16, device=self.dev), c, 2 * c, c, h, w)

    def f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.dev)

    def zeros(self, *shape):
        return torch.zeros(shape, dtype=torch.float32, device=self.dev)

    def pgrad(self, name, like):
        g = self.pgrads.get(name)
        if g is None:
            g = torch.zeros_like(like, dtype=torch.float32)
            self.pgrads[name] = g
        return g

    def grad_of(self, act, create=True):
        g = self.grads.get(id(act))
        if g is None and create:
            g = self.zeros(self.B * act.h * act.w, act.c)
            self.grads[id(act)] = g
        return g


def _chk(rc, what):
    _lib.check(rc, what)


def _packed(t, key, make):
    """Packed fp16x2 GEMM weights are a pure function of the parameter: computed once per optimizer step, reused by every
    micro-batch (the reference accumulates 8 micro-batches per step, nbp_utils.py:387-390)."""
    w = t.cache.get(key)
    if w is None:
        w = make()
        t.cache[key] = w
    return w


class _ZF32:
    """The raw pre-BatchNorm tensor z as plain fp32 NHWC (same 4 bytes per element as the split fp16x2 format).  BatchNorm subtracts
    the batch mean from z: with the 22 significant bits of fp16x2 that cancellation flips ReLU / max-pool decisions the fp32
    reference takes the other way, and those flips -- not the GEMM precision -- dominated the whole-network gradient error
    (tests/studies/gradient_study.py: 4.2e-3 -> 5.5e-5 at B=2, S=128).  The kernels take it through the `lo < 0` convention."""

    __slots__ = ("t", "c", "ld", "lo", "h", "w")

    def __init__(self, dev, B, h, w, c):
        self.t = torch.empty((B, h, w, c), dtype=torch.float32, device=dev)
        self.c, self.ld, self.lo, self.h, self.w = c, c, -1, h, w

    @property
    def ptr(self):
        return self.t.data_ptr()


def _raw_conv(t, name, w, bias, src, taps, dst):
    """dst (_ZF32) = conv(src) + bias (no normalisation): the raw pre-BatchNorm tensor z, fp32 epilogue of the tcgen05 kernel."""
    cout = w.shape[0]
    wp = _packed(t, ("fwd", name), lambda: M._pack_gemm_weight(_w2d(w), True))
    ones = _packed(t, ("ones", cout), lambda: torch.ones(cout, dtype=torch.float32, device=t.dev))
    d = _lib.ConvDesc(1, src.ptr, src.c, src.ld, src.lo, None, 0, 0, 0, t.B, src.h, src.w, taps, 0, wp.data_ptr(), cout,
                      ones.data_ptr(), bias.contiguous().data_ptr(), 0, dst.ptr, dst.ld, 0, 0, 1, TRAIN_K_CHUNK)
    _chk(t.L.nbp_conv_fwd(ctypes.byref(d), _st()), "nbp_conv_fwd(raw z)")


def _bn_stats(t, z, npix, bn):
    C = z.c
    stats = t.f32(4, C)
    _chk(t.L.nbp_bn_train_stats(z.ptr, z.ld, z.lo, npix, C, bn["weight"].data_ptr(), bn["bias"].data_ptr(),
                                bn["running_mean"].data_ptr(), bn["running_var"].data_ptr(), t.momentum, BN_EPS,
                                stats[0].data_ptr(), stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(),
                                t.ws.data_ptr(), _st()), "nbp_bn_train_stats")
    bn["num_batches_tracked"].add_(1)
    return stats            # rows: mean, invstd, scale, shift


def _w2d(w):
    """(Cout, Cin, k, k) -> (Cout, k*k*Cin) with K ordered [tap][ci]."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)


def _dgrad_w2d(w):
    """Weights of the data-gradient conv: dX[p] = sum_t' dz[p + off(t')] W[:, 8-t', :]^T -> (Cin, k*k*Cout)."""
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(w.shape[1], -1)


def _new_dz_operand(t, h, wd, cout, cin):
    """fp16x2 NHWC buffer for dz as a GEMM operand (channels padded to a multiple of 64: Att2_2 has 32) + the inverse-scale vector."""
    kpad = (cout + 63) // 64 * 64
    if kpad != cout:
        dzs = M._Act(torch.zeros((t.B, h, wd, 2 * kpad), dtype=torch.float16, device=t.dev), kpad, 2 * kpad, kpad, h, wd)
    else:
        dzs = t.new(h, wd, cout)
    return dzs, t.f32(max(cin, 1))


def _bn_bwd_to_operand(t, dy, z, npix, cout, cin, stats, bnp, bn_name, relu, h, wd):
    """BatchNorm(+ReLU) backward whose output IS the scaled fp16x2 operand of the following dgrad / wgrad GEMMs."""
    dzs, inv_vec = _new_dz_operand(t, h, wd, cout, cin)
    amax2 = t.f32(2)
    _chk(t.L.nbp_bn_bwd_split(dy.data_ptr(), dy.stride(0), z.ptr, z.ld, z.lo, npix, cout, stats[2].data_ptr(), stats[3].data_ptr(),
                              stats[0].data_ptr(), stats[1].data_ptr(), bnp["weight"].data_ptr(), 1 if relu else 0, None, 0, None,
                              t.pgrad(bn_name + ".weight", bnp["weight"]).data_ptr(), t.pgrad(bn_name + ".bias", bnp["bias"]).data_ptr(),
                              t.ws.data_ptr(), dzs.ptr, dzs.ld, dzs.lo, amax2.data_ptr(), inv_vec.data_ptr(), inv_vec.numel(), _st()), "nbp_bn_bwd_split")
    return dzs, inv_vec


def _conv_backward(t, name, w, src, dzs, inv_vec, taps, need_dsrc=True):
    """Given dz of z = conv(src, w) as a scaled fp16x2 operand (+ its inverse scale): accumulate dW into the parameter gradient,
    return d(src) (fp32 NHWC)."""
    B, h, wd = t.B, src.h, src.w
    npix = B * h * wd
    cout, cin = w.shape[0], src.c
    kpad = dzs.c
    # ---- weight gradient: tcgen05 GEMM over the pixel dimension, operands read in place (MN-major)
    dW = t.zeros(kpad, taps, cin)
    _chk(t.L.nbp_conv_wgrad(dzs.ptr, kpad, dzs.ld, dzs.lo, src.ptr, cin, src.ld, src.lo, B, h, wd, taps, inv_vec.data_ptr(),
                            dW.data_ptr(), WGRAD_MAX_K_TILES, _st()), "nbp_conv_wgrad")
    k = 3 if taps == 9 else 1
    t.pgrad(name + ".weight", w).add_(dW[:cout].view(cout, k, k, cin).permute(0, 3, 1, 2))
    t.pgrad(name + ".bias", w[:, 0, 0, 0])          # exactly zero in front of a train-mode BatchNorm; heads handle theirs
    if not need_dsrc:
        return None
    # ---- data gradient: the forward conv kernel on flipped / transposed weights, fp32 epilogue
    def make_dgrad():
        wd2 = _dgrad_w2d(w)
        if kpad != cout:
            assert taps == 1
            wd2 = torch.cat((wd2, torch.zeros(cin, kpad - cout, device=t.dev)), dim=1)
        return M._pack_gemm_weight(wd2, True)

    dsrc = t.f32(npix, cin)
    layer = {"w": _packed(t, ("dgrad", name), make_dgrad), "scale": inv_vec, "shift": t.zeros(cin), "c_out": cin}
    d = _lib.ConvDesc(1, dzs.ptr, kpad, dzs.ld, dzs.lo, None, 0, 0, 0, B, h, wd, taps, 0, layer["w"].data_ptr(), cin,
                      layer["scale"].data_ptr(), layer["shift"].data_ptr(), 0, dsrc.data_ptr(), cin, 0, 0, 1, TRAIN_K_CHUNK)
    _chk(t.L.nbp_conv_fwd(ctypes.byref(d), _st()), "nbp_conv_fwd(dgrad)")
    return dsrc


def _cbr(t, sd, conv, bn, src, taps=9, dst=None, relu=True):
    """conv + train-mode BatchNorm + ReLU.  Returns the output activation; pushes its backward on the tape."""
    w, b = sd[conv + ".weight"], sd[conv + ".bias"]
    bnp = {k: sd[f"{bn}.{k}"] for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")}
    cout = w.shape[0]
    npix = t.B * src.h * src.w
    z = _ZF32(t.dev, t.B, src.h, src.w, cout)
    _raw_conv(t, conv, w, b, src, taps, z)
    stats = _bn_stats(t, z, npix, bnp)
    y = dst if dst is not None else t.new(src.h, src.w, cout)
    _chk(t.L.nbp_affine_act(z.ptr, z.ld, z.lo, npix, cout, stats[2].data_ptr(), stats[3].data_ptr(), 1 if relu else 0,
                            y.t.data_ptr() + 2 * y.off, y.ld, y.lo, _st()), "nbp_affine_act")
    if t.capture is not None and relu:
        t.capture[bn] = ("affine", z.t, stats)

    def backward(dy, need_dsrc=True):
        """dy: fp32 [npix, ld_dy] view (first `cout` columns used).  Returns d(src) fp32 [npix, cin]."""
        dzs, inv_vec = _bn_bwd_to_operand(t, dy, z, npix, cout, src.c, stats, bnp, bn, relu, src.h, src.w)
        return _conv_backward(t, conv, w, src, dzs, inv_vec, taps, need_dsrc)

    return y, backward


def forward_train(sd, x, cache=None, capture=False, momentum=None):
    """sd: name -> CUDA fp32 tensor (parameters and BatchNorm buffers; buffers are updated in place).
    ``cache``: dict that outlives the call and holds the packed GEMM weights (the caller drops it when parameters change).
    Returns (out1, out2, tape)."""
    dev = x.device
    B, cin0, S, S2 = x.shape
    t = _Tape(dev, B, cache, momentum)
    if capture:
        t.capture = {}
    L = t.L
    st = _st()

    # ---- stem: Conv1.conv.0 (CUDA cores) raw + bias, then BN + ReLU
    w0, b0 = sd["Conv1.conv.0.weight"], sd["Conv1.conv.0.bias"]
    bn0 = {k: sd[f"Conv1.conv.1.{k}"] for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")}
    z0 = _ZF32(dev, B, S, S2, 64)
    w0p = w0.permute(2, 3, 1, 0).reshape(-1, 64).contiguous()
    ones64 = torch.ones(64, device=dev)
    _chk(L.nbp_conv_first(x.data_ptr(), B, cin0, S, S2, w0p.data_ptr(), ones64.data_ptr(), b0.contiguous().data_ptr(), 64, 0,
                          z0.ptr, z0.ld, z0.lo, 1, st), "nbp_conv_first(raw)")
    npix1 = B * S * S2
    stats0 = _bn_stats(t, z0, npix1, bn0)
    y0 = t.new(S, S2, 64)
    _chk(L.nbp_affine_act(z0.ptr, z0.ld, z0.lo, npix1, 64, stats0[2].data_ptr(), stats0[3].data_ptr(), 1, y0.ptr, y0.ld, y0.lo, st), "nbp_affine_act")
    if t.capture is not None:
        t.capture["Conv1.conv.1"] = ("affine", z0.t, stats0)

    def stem_backward(dy):
        dz, amax = t.f32(npix1, 64), t.f32(1)
        _chk(L.nbp_bn_bwd(dy.data_ptr(), dy.stride(0), z0.ptr, z0.ld, z0.lo, npix1, 64, stats0[2].data_ptr(), stats0[3].data_ptr(),
                          stats0[0].data_ptr(), stats0[1].data_ptr(), bn0["weight"].data_ptr(), 1, dz.data_ptr(), 64, amax.data_ptr(),
                          t.pgrad("Conv1.conv.1.weight", bn0["weight"]).data_ptr(), t.pgrad("Conv1.conv.1.bias", bn0["bias"]).data_ptr(),
                          t.ws.data_ptr(), st), "nbp_bn_bwd(stem)")
        dW = t.zeros(9 * cin0, 64)
        _chk(L.nbp_stem_wgrad(x.data_ptr(), B, cin0, S, S2, dz.data_ptr(), dW.data_ptr(), t.ws_big.data_ptr(), st), "nbp_stem_wgrad")
        t.pgrad("Conv1.conv.0.weight", w0).add_(dW.view(3, 3, cin0, 64).permute(3, 2, 0, 1))
        t.pgrad("Conv1.conv.0.bias", b0)

    x1, bw_c1b = _cbr(t, sd, "Conv1.conv.3", "Conv1.conv.4", y0)
    skips = {1: x1}
    enc_bw = {1: (bw_c1b, None, None, y0)}
    cur = x1
    for lvl in range(2, 6):
        p = t.new(cur.h // 2, cur.w // 2, cur.c)
        _chk(L.nbp_maxpool2x2(cur.ptr, B, cur.h, cur.w, cur.c, cur.ld, cur.lo, p.ptr, p.ld, p.lo, st), "nbp_maxpool2x2")
        if t.capture is not None:
            t.capture[f"pool{lvl}"] = ("pool", cur.t, cur.c)
        ya, bw_a = _cbr(t, sd, f"Conv{lvl}.conv.0", f"Conv{lvl}.conv.1", p)
        yb, bw_b = _cbr(t, sd, f"Conv{lvl}.conv.3", f"Conv{lvl}.conv.4", ya)
        enc_bw[lvl] = (bw_b, bw_a, cur, p)
        skips[lvl] = yb
        cur = yb

    def decoder_stage(d, lvl, dec):
        tg = f"{lvl}_{dec}"
        skip = skips[lvl - 1]
        f_l, h2, w2 = skip.c, skip.h, skip.w
        npix = B * h2 * w2
        up = t.new(h2, w2, d.c)
        _chk(L.nbp_upsample2x(d.ptr, B, d.h, d.w, d.c, d.ld, d.lo, up.ptr, up.ld, up.lo, st), "nbp_upsample2x")
        cat = t.new(h2, w2, 2 * f_l)
        g = cat.channels(f_l, f_l)
        _, bw_up = _cbr(t, sd, f"Up{tg}.up.1", f"Up{tg}.up.2", up, dst=g)
        # attention: two raw 1x1 convs, their train-mode BatchNorms, relu(sum), psi
        wg, bg = sd[f"Att{tg}.W_g.0.weight"], sd[f"Att{tg}.W_g.0.bias"]
        wx, bx = sd[f"Att{tg}.W_x.0.weight"], sd[f"Att{tg}.W_x.0.bias"]
        f_int = wg.shape[0]
        zg, zx = _ZF32(dev, B, h2, w2, f_int), _ZF32(dev, B, h2, w2, f_int)
        _raw_conv(t, f"Att{tg}.W_g.0", wg, bg, g, 1, zg)
        _raw_conv(t, f"Att{tg}.W_x.0", wx, bx, skip, 1, zx)
        bng = {k: sd[f"Att{tg}.W_g.1.{k}"] for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")}
        bnx = {k: sd[f"Att{tg}.W_x.1.{k}"] for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")}
        bn1 = {k: sd[f"Att{tg}.psi.1.{k}"] for k in ("weight", "bias", "running_mean", "running_var", "num_batches_tracked")}
        sg, sx = _bn_stats(t, zg, npix, bng), _bn_stats(t, zx, npix, bnx)
        a = t.new(h2, w2, f_int)
        _chk(L.nbp_att_pre(zg.ptr, zx.ptr, zg.ld, zg.lo, npix, f_int, sg[2].data_ptr(), sg[3].data_ptr(), sx[2].data_ptr(), sx[3].data_ptr(),
                           a.ptr, a.ld, a.lo, st), "nbp_att_pre")
        if t.capture is not None:
            t.capture[f"Att{tg}.relu"] = ("act", a.t, f_int)
        w_psi = sd[f"Att{tg}.psi.0.weight"].reshape(-1).contiguous()
        b_psi = sd[f"Att{tg}.psi.0.bias"]
        zpsi, stat4, psi = t.f32(npix), t.f32(4), t.f32(npix)
        _chk(L.nbp_psi_train(a.ptr, a.ld, a.lo, npix, f_int, w_psi.data_ptr(), b_psi.data_ptr(), bn1["weight"].data_ptr(), bn1["bias"].data_ptr(),
                             bn1["running_mean"].data_ptr(), bn1["running_var"].data_ptr(), t.momentum, BN_EPS, zpsi.data_ptr(), stat4.data_ptr(),
                             t.ws.data_ptr(), st), "nbp_psi_train")
        bn1["num_batches_tracked"].add_(1)
        _chk(L.nbp_att_apply(zpsi.data_ptr(), stat4[2:3].data_ptr(), stat4[3:4].data_ptr(), skip.ptr, skip.ld, skip.lo, npix, f_l,
                             cat.t.data_ptr(), cat.ld, 0, cat.lo, psi.data_ptr(), st), "nbp_att_apply")
        ya, bw_a = _cbr(t, sd, f"Up_conv{tg}.conv.0", f"Up_conv{tg}.conv.1", cat)
        yb, bw_b = _cbr(t, sd, f"Up_conv{tg}.conv.3", f"Up_conv{tg}.conv.4", ya)

        def backward():
            dyb = t.grads.pop(id(yb))
            dya = bw_b(dyb)
            dcat = bw_a(dya)                                             # [npix, 2*f_l]
            dskip = t.grad_of(skip)
            dpre, dt = t.f32(npix, f_int), t.f32(npix)
            _chk(L.nbp_att_bwd(dcat.data_ptr(), 2 * f_l, skip.ptr, skip.ld, skip.lo, psi.data_ptr(), zpsi.data_ptr(), npix, f_l,
                               stat4.data_ptr(), bn1["weight"].data_ptr(), a.ptr, a.ld, a.lo, f_int, w_psi.data_ptr(),
                               dskip.data_ptr(), 1, dpre.data_ptr(), dt.data_ptr(),
                               t.pgrad(f"Att{tg}.psi.0.weight", sd[f"Att{tg}.psi.0.weight"]).data_ptr(),
                               t.pgrad(f"Att{tg}.psi.1.weight", bn1["weight"]).data_ptr(), t.pgrad(f"Att{tg}.psi.1.bias", bn1["bias"]).data_ptr(),
                               t.ws_big.data_ptr(), st), "nbp_att_bwd")
            t.pgrad(f"Att{tg}.psi.0.bias", b_psi)
            # the two branch BatchNorms (no ReLU of their own: the ReLU mask is already in dpre)
            dzg, ivg = _bn_bwd_to_operand(t, dpre, zg, npix, f_int, g.c, sg, bng, f"Att{tg}.W_g.1", False, h2, w2)
            dzx, ivx = _bn_bwd_to_operand(t, dpre, zx, npix, f_int, skip.c, sx, bnx, f"Att{tg}.W_x.1", False, h2, w2)
            dg = _conv_backward(t, f"Att{tg}.W_g.0", wg, g, dzg, ivg, 1)
            dxs = _conv_backward(t, f"Att{tg}.W_x.0", wx, skip, dzx, ivx, 1)
            _chk(L.nbp_add_f32(dskip.data_ptr(), dxs.data_ptr(), f_l, npix, f_l, st), "nbp_add_f32")
            # d(up-conv output) = concat half + gate branch
            _chk(L.nbp_add_f32(dg.data_ptr(), dcat[:, f_l:].data_ptr(), 2 * f_l, npix, f_l, st), "nbp_add_f32")
            dup = bw_up(dg)                                              # [npix, d.c] at the upsampled resolution
            dd = t.grad_of(d)
            _chk(L.nbp_upsample2x_bwd(dup.data_ptr(), B, d.h, d.w, d.c, dd.data_ptr(), 1, st), "nbp_upsample2x_bwd")

        t.ops.append(backward)
        return yb

    def head(name, d, sigmoid):
        w, b = sd[name + ".weight"], sd[name + ".bias"]
        w2 = w[:, :, 0, 0].contiguous()
        out = torch.empty((B, w.shape[0], d.h, d.w), dtype=torch.float32, device=dev)
        _chk(L.nbp_conv1x1_head(d.ptr, d.c, d.ld, d.lo, w2.data_ptr(), b.data_ptr(), w.shape[0], 1 if sigmoid else 0, out.data_ptr(), None, B, d.h * d.w, 1, st),
             "nbp_conv1x1_head")

        def backward(dout):
            dd = t.f32(B * d.h * d.w, d.c)
            dw = t.zeros(w.shape[0], d.c)
            _chk(L.nbp_head_bwd(dout.contiguous().data_ptr(), out.data_ptr() if sigmoid else None, d.ptr, d.c, d.ld, d.lo, w2.data_ptr(), w.shape[0],
                                B, d.h * d.w, dd.data_ptr(), dw.data_ptr(), t.pgrad(name + ".bias", b).data_ptr(), t.ws_big.data_ptr(), st), "nbp_head_bwd")
            t.pgrad(name + ".weight", w).add_(dw.view_as(w))
            g0 = t.grads.get(id(d))
            if g0 is None:
                t.grads[id(d)] = dd
            else:
                g0.add_(dd)

        return out, backward

    x5 = skips[5]
    d = decoder_stage(x5, 5, 1)
    d41 = decoder_stage(d, 4, 1)
    out1, bw_h1 = head("Final1", d41, False)
    d = decoder_stage(x5, 5, 2)
    for lvl in (4, 3, 2):
        d = decoder_stage(d, lvl, 2)
    out2, bw_h2 = head("Final2.0", d, True)

    def run_backward(dout1, dout2):
        if dout1 is not None:
            bw_h1(dout1)
        else:
            t.grad_of(d41)
        if dout2 is not None:
            bw_h2(dout2)
        else:
            t.grad_of(d)
        for op in reversed(t.ops):                      # decoder stages, last built first
            op()
        # encoder, deepest level first
        for lvl in range(5, 1, -1):
            bw_b, bw_a, below, pooled = enc_bw[lvl]
            dy = t.grads.pop(id(skips[lvl]))
            dp = bw_a(bw_b(dy))                          # gradient of the pooled input
            dbelow = t.grad_of(below)
            _chk(L.nbp_maxpool2x2_bwd(dp.data_ptr(), below.ptr, below.ld, below.lo, B, below.h, below.w, below.c, dbelow.data_ptr(), 1, st),
                 "nbp_maxpool2x2_bwd")
        dy0 = enc_bw[1][0](t.grads.pop(id(x1)))
        stem_backward(dy0)
        pg = t.pgrads
        # the backward closures reference the tape and each other: drop them now so that the activations are returned to the
        # allocator before the next micro-batch starts (otherwise they wait for Python's cycle collector and the pool grows)
        t.ops.clear(); t.grads.clear(); enc_bw.clear(); skips.clear(); t.run_backward = None
        return pg

    t.run_backward = run_backward
    return out1, out2, t


def decisions_from_capture(capture):
    """The discrete decisions the backward kernels take, as NCHW tensors keyed like ``oracle.nbp_torch.Decisions``: ReLU masks
    (bool) and 2x2 max-pool window indices (long, first maximum in window order).  They are recomputed here from the very tensors
    the kernels read, with the kernels' expressions: ``fmaf(z, scale, shift) > 0`` (bn_bwd_*; evaluated in float64, where the
    product is exact, so the sign is the fused operation's), ``a > 0`` on the stored fp16x2 value (psi_bwd), and
    ``v[k] > v[best]`` over the stored fp16x2 values (maxpool_bwd).  Filled when ``module.capture_decisions`` is true; test
    support for the flip-robust gradient parity test, not used by training itself."""
    out = {}
    for name, (kind, t, aux) in capture.items():
        if kind == "affine":
            y = t.double() * aux[2].double() + aux[3].double()                     # (B,h,w,C)
            out[name] = (y > 0).permute(0, 3, 1, 2).contiguous()
        else:
            c = aux
            v = t[..., :c].float() + t[..., c:2 * c].float() / M.LO_SCALE
            v = v.permute(0, 3, 1, 2)
            if kind == "act":
                out[name] = (v > 0).contiguous()
            else:
                B, C, H, W = v.shape
                win = v.reshape(B, C, H // 2, 2, W // 2, 2).permute(0, 1, 2, 4, 3, 5).reshape(B, C, H // 2, W // 2, 4)
                best = torch.zeros(win.shape[:-1], dtype=torch.long, device=win.device)
                bv = win[..., 0]
                for k in range(1, 4):
                    upd = win[..., k] > bv
                    best = torch.where(upd, torch.full_like(best, k), best)
                    bv = torch.where(upd, win[..., k], bv)
                out[name] = best
    return out


class NBPTrainFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, x, names, *params):
        sd = dict(zip(names, params))
        for k, v in module.named_buffers():
            sd[k] = v
        # packed weights survive across micro-batches until an in-place update (optimizer.step, load_state_dict) bumps a version
        key = (str(x.device),) + tuple((p.data_ptr(), p._version) for p in params)
        cache = getattr(module, "_train_pack", None)
        if cache is None or cache.get("key") != key:
            cache = {"key": key}
            module._train_pack = cache
        out1, out2, tape = forward_train(sd, x.contiguous().float(), cache, capture=bool(getattr(module, "capture_decisions", False)),
                                          momentum=getattr(module, "bn_momentum", None))
        if tape.capture is not None:
            module.last_decisions = decisions_from_capture(tape.capture)
            tape.capture = None
        ctx.tape, ctx.names, ctx.params = tape, names, params
        return out1, out2

    @staticmethod
    def backward(ctx, dout1, dout2):
        pg = ctx.tape.run_backward(dout1, dout2)
        grads = []
        for n, p in zip(ctx.names, ctx.params):
            g = pg.get(n)
            grads.append(g.view_as(p) if g is not None else torch.zeros_like(p))
        ctx.tape = None
        return (None, None, None) + tuple(grads)
