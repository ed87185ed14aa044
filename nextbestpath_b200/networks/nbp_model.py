"""``NBP`` -- the coverage-gain / obstacle-map attention U-Net, B200-native.

Drop-in for ``next_best_path/networks/nbp_model.py`` (class NBP :64-173): same constructor, same 327
``state_dict`` keys in the same order (so ``AiMDoom_*_best_val.pth`` loads), real ``nn.Parameter``s,
``forward(x) -> (out1 (B,8,S/4,S/4), out2 (B,1,S,S))`` in NCHW fp32, ``loss(...)``, ``log_vars``.

The module tree below only HOLDS parameters.  ``forward`` never calls torch convolution / cuDNN: on a
CUDA tensor in eval mode it runs the hand-written sm_100a pipeline in ``libnbp_b200.so`` (tcgen05 implicit
GEMM for the 3x3 / 1x1 contractions with folded BatchNorm + ReLU epilogues, CUDA-core kernels for the
5-channel stem, pooling, up-sampling, attention gates and the 8- / 1-channel heads).  There is no CPU
fallback: CPU tensors raise.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _lib

_ENC = (("Conv1", None, 64), ("Conv2", 64, 128), ("Conv3", 128, 256), ("Conv4", 256, 512), ("Conv5", 512, 1024))
_DEC_LEVELS = {1: (5, 4), 2: (5, 4, 3, 2)}
_LEVEL_CH = {5: 1024, 4: 512, 3: 256, 2: 128}


def _conv_bn_relu_pair(cin, cout):
    """Two (3x3 conv, BN, ReLU) groups held in a Sequential named ``conv`` (indices 0,1,2 / 3,4,5)."""
    holder = nn.Module()
    holder.conv = nn.Sequential(nn.Conv2d(cin, cout, 3, 1, 1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True),
                                nn.Conv2d(cout, cout, 3, 1, 1), nn.BatchNorm2d(cout), nn.ReLU(inplace=True))
    return holder


def _up_stage(cin, cout):
    holder = nn.Module()
    holder.up = nn.Sequential(nn.Upsample(scale_factor=2), nn.Conv2d(cin, cout, 3, 1, 1), nn.BatchNorm2d(cout),
                              nn.ReLU(inplace=True))
    return holder


def _gate(f_g, f_l, f_int):
    holder = nn.Module()
    holder.W_g = nn.Sequential(nn.Conv2d(f_g, f_int, 1), nn.BatchNorm2d(f_int))
    holder.W_x = nn.Sequential(nn.Conv2d(f_l, f_int, 1), nn.BatchNorm2d(f_int))
    holder.psi = nn.Sequential(nn.Conv2d(f_int, 1, 1), nn.BatchNorm2d(1), nn.Sigmoid())
    holder.relu = nn.ReLU(inplace=True)
    return holder


class NBP(nn.Module):
    def __init__(self, img_ch=5, output_ch1=8, output_ch2=1):
        super().__init__()
        self.Maxpool = nn.MaxPool2d(kernel_size=2, stride=2)
        for name, cin, cout in _ENC:
            setattr(self, name, _conv_bn_relu_pair(img_ch if cin is None else cin, cout))
        for dec in (1, 2):
            for lvl in _DEC_LEVELS[dec]:
                c = _LEVEL_CH[lvl]
                setattr(self, f"Up{lvl}_{dec}", _up_stage(c, c // 2))
                setattr(self, f"Att{lvl}_{dec}", _gate(c // 2, c // 2, c // 4))
                setattr(self, f"Up_conv{lvl}_{dec}", _conv_bn_relu_pair(c, c // 2))
            if dec == 1:
                self.Final1 = nn.Conv2d(256, output_ch1, kernel_size=1)
        self.Final2 = nn.Sequential(nn.Conv2d(64, output_ch2, kernel_size=1), nn.Sigmoid())
        self.log_vars = nn.Parameter(torch.zeros(2))
        self.img_ch, self.output_ch1, self.output_ch2 = img_ch, output_ch1, output_ch2
        self._packed = None            # (key, dict) cache of packed fp16 weights / folded affines
        self.max_chunk = 32            # scenes per pass through the pipeline (bounds activation memory)

    # ------------------------------------------------------------------ reference API
    def forward(self, x):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError("nextbestpath_b200.NBP runs on CUDA tensors only (no CPU fallback); "
                               "the CPU oracle lives in oracle/nbp_torch.py and is test infrastructure")
        if self.training or torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()) and x.requires_grad:
            if self.training:
                raise NotImplementedError("train-mode forward/backward of the B200 NBP is not built yet (round 2); "
                                          "call .eval() for the inference rollout path")
        if x.dim() != 4 or x.shape[1] != self.img_ch:
            raise RuntimeError(f"expected (B,{self.img_ch},S,S) input, got {tuple(x.shape)}")
        if x.shape[2] % 16 or x.shape[3] % 16:
            raise RuntimeError("spatial size must be a multiple of 16 (four 2x2 poolings)")
        x = x.contiguous().float()
        outs1, outs2 = [], []
        pk = self._pack(x.device)
        for b0 in range(0, x.shape[0], self.max_chunk):
            o1, o2 = _forward_eval(pk, x[b0:b0 + self.max_chunk])
            outs1.append(o1); outs2.append(o2)
        if len(outs1) == 1:
            return outs1[0], outs2[0]
        return torch.cat(outs1), torch.cat(outs2)

    def loss(self, pred1, target1, pred2, target2):
        """nbp_model.py:162-173: MSE/(2 s1^2) + log s1 + BCE/s2^2 + log s2 with s = exp(log_vars)."""
        s1, s2 = torch.exp(2 * self.log_vars[0]), torch.exp(2 * self.log_vars[1])
        l1 = (1.0 / (2.0 * s1)) * F.mse_loss(pred1, target1) + self.log_vars[0]
        l2 = (1.0 / s2) * F.binary_cross_entropy(pred2, target2) + self.log_vars[1]
        return l1 + l2

    # ------------------------------------------------------------------ weight packing (host-side prep)
    def _pack(self, device):
        key = (str(device),) + tuple((p.data_ptr(), p._version) for p in self.state_dict(keep_vars=True).values())
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        sd = {k: v.detach().to(device=device, dtype=torch.float32) if v.is_floating_point() else v for k, v in self.state_dict().items()}
        pk = pack_state_dict(sd)
        self._packed = (key, pk)
        return pk


def _affine(sd, conv, bn, eps=1e-5):
    scale = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + eps)
    shift = sd[bn + ".bias"] + (sd[conv + ".bias"] - sd[bn + ".running_mean"]) * scale
    return scale.contiguous(), shift.contiguous()


def _pack3x3(w):
    """(Cout, Cin, 3, 3) -> fp16 [Cout][tap = ky*3+kx][Cin]."""
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).to(torch.float16).contiguous()


def pack_state_dict(sd):
    """Folded / packed parameters for the eval pipeline, from a float32 state_dict on the target device."""
    pk = {}

    def conv3(name, conv, bn):
        s, b = _affine(sd, conv, bn)
        pk[name] = {"w": _pack3x3(sd[conv + ".weight"]), "scale": s, "shift": b, "c_out": sd[conv + ".weight"].shape[0]}

    w0 = sd["Conv1.conv.0.weight"]
    s, b = _affine(sd, "Conv1.conv.0", "Conv1.conv.1")
    pk["stem"] = {"w": w0.permute(2, 3, 1, 0).reshape(-1, w0.shape[0]).contiguous(), "scale": s, "shift": b, "c_in": w0.shape[1]}
    conv3("Conv1.b", "Conv1.conv.3", "Conv1.conv.4")
    for lvl in range(2, 6):
        conv3(f"Conv{lvl}.a", f"Conv{lvl}.conv.0", f"Conv{lvl}.conv.1")
        conv3(f"Conv{lvl}.b", f"Conv{lvl}.conv.3", f"Conv{lvl}.conv.4")
    for dec in (1, 2):
        for lvl in _DEC_LEVELS[dec]:
            t = f"{lvl}_{dec}"
            conv3(f"Up{t}", f"Up{t}.up.1", f"Up{t}.up.2")
            conv3(f"Up_conv{t}.a", f"Up_conv{t}.conv.0", f"Up_conv{t}.conv.1")
            conv3(f"Up_conv{t}.b", f"Up_conv{t}.conv.3", f"Up_conv{t}.conv.4")
            # attention: one 1x1 GEMM over concat(g, x) with the two BN scales folded into the weights
            sg, bg = _affine(sd, f"Att{t}.W_g.0", f"Att{t}.W_g.1")
            sx, bx = _affine(sd, f"Att{t}.W_x.0", f"Att{t}.W_x.1")
            wg = sd[f"Att{t}.W_g.0.weight"][:, :, 0, 0] * sg[:, None]
            wx = sd[f"Att{t}.W_x.0.weight"][:, :, 0, 0] * sx[:, None]
            f_int = wg.shape[0]
            sp, bp = _affine(sd, f"Att{t}.psi.0", f"Att{t}.psi.1")
            pk[f"Att{t}"] = {"w": torch.cat((wg, wx), dim=1).to(torch.float16).contiguous(),
                             "scale": torch.ones(f_int, device=wg.device), "shift": (bg + bx).contiguous(), "c_out": f_int,
                             "w_psi": sd[f"Att{t}.psi.0.weight"].reshape(-1).contiguous(),
                             "psi_scale": float(sp.item()), "psi_shift": float(bp.item())}
    pk["Final1"] = {"w": sd["Final1.weight"][:, :, 0, 0].contiguous(), "b": sd["Final1.bias"].contiguous()}
    pk["Final2"] = {"w": sd["Final2.0.weight"][:, :, 0, 0].contiguous(), "b": sd["Final2.0.bias"].contiguous()}
    return pk


# ---------------------------------------------------------------------- the eval pipeline (C ABI calls)
def _stream():
    return torch.cuda.current_stream().cuda_stream


def _conv(layer, src0, c0, ld0, n, h, w, taps, dst, dst_ld, dst_c_off, relu=True, src1=None, c1=0, ld1=0):
    d = _lib.ConvDesc(src0, c0, ld0, src1, c1, ld1, n, h, w, taps, layer["w"].data_ptr(), layer["c_out"],
                      layer["scale"].data_ptr(), layer["shift"].data_ptr(), 1 if relu else 0, dst, dst_ld, dst_c_off)
    _lib.check(_lib.lib().nbp_conv_fwd(ctypes.byref(d), _stream()), "nbp_conv_fwd")


def _forward_eval(pk, x):
    L = _lib.lib()
    dev = x.device
    B, _, S, S2 = x.shape
    st = _stream()
    new = lambda h, w, c: torch.empty((B, h, w, c), dtype=torch.float16, device=dev)

    def double_conv(name, src, c_in, h, w):
        c_out = pk[name + ".a"]["c_out"]
        t = new(h, w, c_out)
        _conv(pk[name + ".a"], src.data_ptr(), c_in, c_in, B, h, w, 9, t.data_ptr(), c_out, 0)
        y = new(h, w, c_out)
        _conv(pk[name + ".b"], t.data_ptr(), c_out, c_out, B, h, w, 9, y.data_ptr(), c_out, 0)
        return y

    # ---- encoder
    a = new(S, S2, 64)
    stem = pk["stem"]
    _lib.check(L.nbp_conv_first(x.data_ptr(), B, stem["c_in"], S, S2, stem["w"].data_ptr(), stem["scale"].data_ptr(),
                                stem["shift"].data_ptr(), 64, a.data_ptr(), 64, st), "nbp_conv_first")
    x1 = new(S, S2, 64)
    _conv(pk["Conv1.b"], a.data_ptr(), 64, 64, B, S, S2, 9, x1.data_ptr(), 64, 0)
    del a
    skips = {1: (x1, 64, S, S2)}
    cur, c, h, w = x1, 64, S, S2
    for lvl in range(2, 6):
        p = new(h // 2, w // 2, c)
        _lib.check(L.nbp_maxpool2x2(cur.data_ptr(), B, h, w, c, c, p.data_ptr(), c, st), "nbp_maxpool2x2")
        h, w = h // 2, w // 2
        cur = double_conv(f"Conv{lvl}", p, c, h, w)
        c = pk[f"Conv{lvl}.a"]["c_out"]
        skips[lvl] = (cur, c, h, w)

    def decoder_stage(d, c_d, h_d, w_d, lvl, dec):
        """Up{lvl}_{dec} -> Att{lvl}_{dec} -> cat -> Up_conv{lvl}_{dec} (nbp_model.py:124-129)."""
        t = f"{lvl}_{dec}"
        skip, f_l, h2, w2 = skips[lvl - 1]
        up = new(h2, w2, c_d)
        _lib.check(L.nbp_upsample2x(d.data_ptr(), B, h_d, w_d, c_d, c_d, up.data_ptr(), c_d, st), "nbp_upsample2x")
        cat = new(h2, w2, 2 * f_l)                       # [skip*psi | up-conv output]
        g_ptr = cat.data_ptr() + 2 * f_l                 # channel offset f_l in fp16 bytes
        _conv(pk[f"Up{t}"], up.data_ptr(), c_d, c_d, B, h2, w2, 9, cat.data_ptr(), 2 * f_l, f_l)
        del up
        att = pk[f"Att{t}"]
        f_int = att["c_out"]
        arelu = new(h2, w2, f_int)
        _conv(att, g_ptr, f_l, 2 * f_l, B, h2, w2, 1, arelu.data_ptr(), f_int, 0, relu=True,
              src1=skip.data_ptr(), c1=f_l, ld1=f_l)
        _lib.check(L.nbp_att_gate(arelu.data_ptr(), f_int, skip.data_ptr(), f_l, f_l, att["w_psi"].data_ptr(),
                                  att["psi_scale"], att["psi_shift"], cat.data_ptr(), 2 * f_l, 0, B * h2 * w2, st), "nbp_att_gate")
        del arelu
        y = double_conv(f"Up_conv{t}", cat, 2 * f_l, h2, w2)
        return y, f_l, h2, w2

    x5, c5, h5, w5 = skips[5]
    # ---- decoder 1 -> value map at S/4
    d, cd, hd, wd = decoder_stage(x5, c5, h5, w5, 5, 1)
    d, cd, hd, wd = decoder_stage(d, cd, hd, wd, 4, 1)
    out1 = torch.empty((B, pk["Final1"]["w"].shape[0], hd, wd), dtype=torch.float32, device=dev)
    _lib.check(L.nbp_conv1x1_head(d.data_ptr(), cd, cd, pk["Final1"]["w"].data_ptr(), pk["Final1"]["b"].data_ptr(),
                                  out1.shape[1], 0, out1.data_ptr(), B, hd * wd, st), "nbp_conv1x1_head")
    # ---- decoder 2 -> obstacle map at S
    d, cd, hd, wd = decoder_stage(x5, c5, h5, w5, 5, 2)
    for lvl in (4, 3, 2):
        d, cd, hd, wd = decoder_stage(d, cd, hd, wd, lvl, 2)
    out2 = torch.empty((B, pk["Final2"]["w"].shape[0], hd, wd), dtype=torch.float32, device=dev)
    _lib.check(L.nbp_conv1x1_head(d.data_ptr(), cd, cd, pk["Final2"]["w"].data_ptr(), pk["Final2"]["b"].data_ptr(),
                                  out2.shape[1], 1, out2.data_ptr(), B, hd * wd, st), "nbp_conv1x1_head")
    return out1, out2
