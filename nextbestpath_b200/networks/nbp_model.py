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
import os

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
        # "mixed" : the parity path.  The five encoder layers the network is most sensitive to (``full_precision_layers``) run
        #           "fp16x2"; every other GEMM layer computes its hi product in fp16 and both correction products on the e4m3 pipe
        #           (nbp_conv_desc mode 2: 2 tensor pass-equivalents instead of 3).  ~1.2e-4 / 4e-4 of the fp32 reference
        #           (value / obstacle map; bar 1e-3).
        # "fp16x2": split-fp16 operands everywhere, 3 passes, fp32-grade results (1e-5 of the fp32 reference).
        # "fp16"  : single fp16 plane, 1 pass, ~7e-3 of the fp32 reference (not a parity mode).
        self.precision = "mixed"
        self.full_precision_layers = ("Conv1.b", "Conv2.a", "Conv2.b", "Conv3.a", "Conv3.b")
        # eval forward = replay of a captured CUDA graph (see _run_eval).  static_outputs: return the graph's own output buffers
        # (overwritten by the next forward) instead of copies -- for callers that consume the maps before the next call.
        self.bn_momentum = 0.1         # nn.BatchNorm2d default, as the reference instantiates it (nbp_model.py:12); train mode only
        self.use_cuda_graph = True
        self.static_outputs = False
        self._graphs = {}
        self.last_value_max = None

    # ------------------------------------------------------------------ reference API
    def forward(self, x):
        if not isinstance(x, torch.Tensor) or not x.is_cuda:
            raise RuntimeError("nextbestpath_b200.NBP runs on CUDA tensors only (no CPU fallback); "
                               "the CPU oracle lives in oracle/nbp_torch.py and is test infrastructure")
        if x.dim() != 4 or x.shape[1] != self.img_ch:
            raise RuntimeError(f"expected (B,{self.img_ch},S,S) input, got {tuple(x.shape)}")
        if x.shape[2] % 16 or x.shape[3] % 16:
            raise RuntimeError("spatial size must be a multiple of 16 (four 2x2 poolings)")
        x = x.contiguous().float()
        if self.training:
            # train mode: batch-statistics BatchNorm + autograd through the CUDA backward kernels (nbp_train.py)
            from .nbp_train import NBPTrainFunction
            named = [(n, p) for n, p in self.named_parameters() if n != "log_vars"]
            for n, p in named:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous():
                    raise RuntimeError(f"parameter {n} must be a contiguous fp32 CUDA tensor")
            return NBPTrainFunction.apply(self, x, tuple(n for n, _ in named), *[p for _, p in named])
        pk = self._pack(x.device)
        out1, out2, vmax = self._run_eval(pk, x)
        self.last_value_max = vmax                     # (B, S/4, S/4): max over the 8 headings (nbp_planning.py:193), fused
        return out1, out2

    def _run_eval(self, pk, x):
        """Eval forward of the whole batch in chunks of ``max_chunk`` scenes.  With ``use_cuda_graph`` the chunk loop is
        captured once per (weights, shape) into a CUDA graph reading a module-owned input buffer and replayed afterwards
        (~40 kernels per chunk; tensor maps, launch parameters and the activation arena are frozen in the graph)."""
        B, _, S, S2 = x.shape
        dev = x.device
        if not self.use_cuda_graph or torch.cuda.is_current_stream_capturing():
            out1 = torch.empty((B, self.output_ch1, S // 4, S2 // 4), dtype=torch.float32, device=dev)
            out2 = torch.empty((B, self.output_ch2, S, S2), dtype=torch.float32, device=dev)
            vmax = torch.empty((B, S // 4, S2 // 4), dtype=torch.float32, device=dev)
            self._eval_chunks(pk, x, out1, out2, vmax)
            return out1, out2, vmax
        key = (id(pk), B, S, S2, dev.index, self.max_chunk)
        g = self._graphs.get(key)
        if g is None:
            if len(self._graphs) >= 4:                  # bound the memory held by graph pools (shapes seen long ago)
                self._graphs.pop(next(iter(self._graphs)))
            g = _EvalGraph(self, pk, x)
            self._graphs[key] = g
        return g.replay(x, clone=not self.static_outputs)

    def _eval_chunks(self, pk, x, out1, out2, vmax):
        for b0 in range(0, x.shape[0], self.max_chunk):
            b1 = min(x.shape[0], b0 + self.max_chunk)
            _forward_eval(pk, x[b0:b1], out1[b0:b1], out2[b0:b1], vmax[b0:b1])

    def e4m3_saturation_count(self) -> int:
        """Number of (pixel, 32-channel group) stores, since the weights were last packed, in which an activation left the e4m3
        window of the "mixed" mode (|x| > 3584) and fell back to fp16 precision.  0 in the calibrated operating range."""
        pk = self._packed[1] if self._packed is not None else None
        return int(pk["sat_count"].item()) if pk is not None and "sat_count" in pk else 0

    def loss(self, pred1, target1, pred2, target2):
        """nbp_model.py:162-173: MSE/(2 s1^2) + log s1 + BCE/s2^2 + log s2 with s = exp(log_vars)."""
        s1, s2 = torch.exp(2 * self.log_vars[0]), torch.exp(2 * self.log_vars[1])
        l1 = (1.0 / (2.0 * s1)) * F.mse_loss(pred1, target1) + self.log_vars[0]
        l2 = (1.0 / s2) * F.binary_cross_entropy(pred2, target2) + self.log_vars[1]
        return l1 + l2

    # ------------------------------------------------------------------ weight packing (host-side prep)
    def _pack(self, device):
        if self.precision not in ("mixed", "fp16x2", "fp16"):
            raise ValueError(f"unknown precision {self.precision!r}")
        key = (str(device), self.precision, tuple(self.full_precision_layers)) + tuple((p.data_ptr(), p._version) for p in self.state_dict(keep_vars=True).values())
        if self._packed is not None and self._packed[0] == key:
            return self._packed[1]
        sd = {k: v.detach().to(device=device, dtype=torch.float32) if v.is_floating_point() else v for k, v in self.state_dict().items()}
        pk = pack_state_dict(sd, precise=self.precision != "fp16",
                             e4m3_layers=None if self.precision != "mixed" else (lambda name: name not in self.full_precision_layers))
        self._packed = (key, pk)
        self._graphs.clear()                            # graphs bake in the packed-weight pointers
        return pk


def _affine(sd, conv, bn, eps=1e-5):
    scale = sd[bn + ".weight"] / torch.sqrt(sd[bn + ".running_var"] + eps)
    shift = sd[bn + ".bias"] + (sd[conv + ".bias"] - sd[bn + ".running_mean"]) * scale
    return scale.contiguous(), shift.contiguous()


LO_SCALE = 2048.0
# A/B switch (NBP_FUSE_DOT=0): psi of the attention gates and the Final2 head as separate kernels instead of the conv kernel's dot epilogue
FUSE_DOT = os.environ.get("NBP_FUSE_DOT", "1") != "0"
# x * psi written by the attention GEMM's epilogue too (nbp_conv_desc.gate_src; needs FUSE_DOT).  Built, tested, and OFF: measured on B200
# it is a wash (8.71-8.81 ms per 32-scene forward without, 8.68-8.90 with) -- the thread-per-pixel epilogue reads and writes the skip
# tensor in 32-byte pieces, which costs what the coalesced streaming kernel it replaces cost
FUSE_GATE = os.environ.get("NBP_FUSE_GATE", "0") != "0"
# Final1 (8 headings + their max) in the dot epilogue of Up_conv4_1.b (nbp_conv_desc.dot_n = 8).  Built, tested, and OFF: measured a wash
# (9.05-9.08 ms per 32-scene forward with the separate 75 us head kernel, 9.10-9.12 ms fused: 8 x 128 FMAs and 256 weight loads per pixel
# in a thread-per-pixel epilogue cost what the streaming head kernel costs)
FUSE_HEAD8 = os.environ.get("NBP_FUSE_HEAD8", "0") != "0"


def _pack_gemm_weight(w2d, precise):
    """(Cout, K) fp32 -> the B operand of nbp_conv_fwd.  fast: fp16 [Cout][K].  precise (fp16x2): per tile of
    BN = 128|64|32 output channels, BN rows of hi = fp16(w) followed by BN rows of lo = fp16((w - hi) * 2048)."""
    hi = w2d.to(torch.float16)
    if not precise:
        return hi.contiguous()
    lo = ((w2d - hi.float()) * LO_SCALE).to(torch.float16)
    cout, k = w2d.shape
    bn = 128 if cout % 128 == 0 else 64 if cout % 64 == 0 else 32
    return torch.stack((hi.view(cout // bn, bn, k), lo.view(cout // bn, bn, k)), dim=1).reshape(2 * cout, k).contiguous()


def _e4m3_bytes(x):
    return x.clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)


def _e4m3_weight_scale(w2d):
    """Power of two s with max|w| * s in (224, 448]: the e4m3 weight rows use the top of the format's range."""
    m = float(w2d.abs().max())
    if m == 0.0 or m != m:
        return 1.0
    import math
    return 2.0 ** math.floor(math.log2(448.0 / m))


def _pack_gemm_weight_e4m3(w2d, s):
    """(Cout, K) fp32 -> the B operand of nbp_conv_fwd mode 2: per tile of BN output channels, BN rows of hi = fp16(w), then BN
    rows holding, per 64-element K slice, 64 bytes e4m3((w - hi) * 2048 * s) followed by 64 bytes e4m3(w * s) -- the slice's
    correction products A_hi8 . W_lo8 + A_lo8 . W_hi8 are then one K = 128 e4m3 reduction.  Stored as fp16 [2*Cout][K]."""
    hi = w2d.to(torch.float16)
    cout, k = w2d.shape
    assert k % 64 == 0
    lo8 = _e4m3_bytes((w2d - hi.float()) * (LO_SCALE * s)).view(cout, k // 64, 1, 64)
    hi8 = _e4m3_bytes(w2d * s).view(cout, k // 64, 1, 64)
    p8 = torch.cat((lo8, hi8), dim=2).reshape(cout, 2 * k).contiguous().view(torch.float16)          # (cout, k) fp16 container
    bn = 128 if cout % 128 == 0 else 64 if cout % 64 == 0 else 32
    return torch.stack((hi.view(cout // bn, bn, k), p8.view(cout // bn, bn, k)), dim=1).reshape(2 * cout, k).contiguous()


def pack_state_dict(sd, precise=True, e4m3_layers=None):
    """Folded / packed parameters for the eval pipeline, from a float32 state_dict on the target device.
    ``e4m3_layers``: predicate(layer name) -> True for the GEMM layers that run nbp_conv_desc mode 2 (fp16 + e4m3 corrections)."""
    pk = {"precise": bool(precise)}
    if e4m3_layers is not None:
        # device counter of e4m3 saturation events (nbp_conv_desc.sat_count): stays 0 unless inputs leave the calibrated range
        pk["sat_count"] = torch.zeros(1, dtype=torch.int64, device=sd["Final1.weight"].device)

    def gemm_entry(name, blocks, s_aff, b_aff, c_out, extra=None):
        """blocks: list of (Cout, K) fp32 weight matrices packed one after the other (1, or 4 parity blocks of an up-conv)."""
        mode = 0 if not precise else 2 if (e4m3_layers is not None and e4m3_layers(name)) else 1
        ent = {"scale": s_aff, "shift": b_aff, "c_out": c_out, "mode": mode, "lo_scale": 1.0 / LO_SCALE}
        if mode == 2:
            sw = _e4m3_weight_scale(torch.cat([b.reshape(-1) for b in blocks]))
            ent["w"] = torch.cat([_pack_gemm_weight_e4m3(b, sw) for b in blocks], dim=0).contiguous()
            ent["lo_scale"] = 1.0 / (LO_SCALE * sw)
        else:
            ent["w"] = torch.cat([_pack_gemm_weight(b, precise) for b in blocks], dim=0).contiguous()
        if extra:
            ent.update(extra)
        pk[name] = ent

    def conv3(name, conv, bn):
        s, b = _affine(sd, conv, bn)
        w = sd[conv + ".weight"]                                  # (Cout, Cin, 3, 3) -> [Cout][tap = ky*3+kx][Cin]
        gemm_entry(name, [w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)], s, b, w.shape[0])

    _ROWS = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}       # parity -> kernel rows/cols summed into 2x2 tap 0 / tap 1

    def up3(name, conv, bn):
        """nn.Upsample(2, nearest) + 3x3 conv == four parity-specific 2x2 convs on the source (up2x mode of nbp_conv_fwd):
        output (2y+py, 2x+px) reads source rows y-1+py+ty with the kernel rows that land on them pre-summed."""
        s, b = _affine(sd, conv, bn)
        w = sd[conv + ".weight"]                                  # (Cout, Cin, 3, 3)
        blocks = []
        for py in (0, 1):
            for px in (0, 1):
                taps = []
                for ty in (0, 1):
                    for tx in (0, 1):
                        acc = 0
                        for ky in _ROWS[py][ty]:
                            for kx in _ROWS[px][tx]:
                                acc = acc + w[:, :, ky, kx]
                        taps.append(acc)                          # (Cout, Cin)
                blocks.append(torch.stack(taps, dim=1).reshape(w.shape[0], -1))
        gemm_entry(name, blocks, s, b, w.shape[0])

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
            up3(f"Up{t}", f"Up{t}.up.1", f"Up{t}.up.2")
            conv3(f"Up_conv{t}.a", f"Up_conv{t}.conv.0", f"Up_conv{t}.conv.1")
            conv3(f"Up_conv{t}.b", f"Up_conv{t}.conv.3", f"Up_conv{t}.conv.4")
            # attention: one 1x1 GEMM over concat(g, x) with the two BN scales folded into the weights
            sg, bg = _affine(sd, f"Att{t}.W_g.0", f"Att{t}.W_g.1")
            sx, bx = _affine(sd, f"Att{t}.W_x.0", f"Att{t}.W_x.1")
            wg = sd[f"Att{t}.W_g.0.weight"][:, :, 0, 0] * sg[:, None]
            wx = sd[f"Att{t}.W_x.0.weight"][:, :, 0, 0] * sx[:, None]
            f_int = wg.shape[0]
            sp, bp = _affine(sd, f"Att{t}.psi.0", f"Att{t}.psi.1")
            gemm_entry(f"Att{t}", [torch.cat((wg, wx), dim=1)], torch.ones(f_int, device=wg.device), (bg + bx).contiguous(), f_int,
                       extra={"w_psi": sd[f"Att{t}.psi.0.weight"].reshape(-1).contiguous(),
                              "psi_scale": float(sp.item()), "psi_shift": float(bp.item())})
    pk["Final1"] = {"w": sd["Final1.weight"][:, :, 0, 0].contiguous(), "b": sd["Final1.bias"].contiguous()}
    pk["Final2"] = {"w": sd["Final2.0.weight"][:, :, 0, 0].contiguous(), "b": sd["Final2.0.bias"].contiguous(),
                    "b_host": float(sd["Final2.0.bias"].reshape(-1)[0].item())}
    return pk


# ---------------------------------------------------------------------- the eval pipeline (C ABI calls)
def _stream():
    return torch.cuda.current_stream().cuda_stream


class _Act:
    """An NHWC fp16 activation: `c` channels per pixel in `planes` planes (hi [, lo*2048]); a view may start at a
    channel offset of a wider buffer (concat fusion)."""

    __slots__ = ("t", "c", "ld", "lo", "off", "h", "w", "fmt")

    def __init__(self, t, c, ld, lo, h, w, off=0, fmt=1):
        self.t, self.c, self.ld, self.lo, self.h, self.w, self.off = t, c, ld, lo, h, w, off
        self.fmt = fmt              # second-plane format: 1 = fp16 lo * 2048, 2 = e4m3 pair (nbp_conv_desc mode 2)

    @property
    def ptr(self):
        return self.t.data_ptr() + 2 * self.off

    def channels(self, off, c):
        return _Act(self.t, c, self.ld, self.lo, self.h, self.w, self.off + off, self.fmt)


def _conv(pk, layer, B, src0, taps, dst, relu=True, src1=None, up2x=False, k_chunk=0, pool=None, dot=None, gate=None):
    """``dot`` = (w [c_out] fp32, scale, shift, sigmoid, out fp32 [B,h,w] or None[, n, bias [n] or None, max [B,h,w] or None]): the dot
    epilogue of nbp_conv_desc -- the layer's output is contracted with ``w`` per pixel instead of being stored (``dst`` is None then);
    with ``n`` > 1, ``w`` is [n][c_out] and ``out`` [B,n,h,w] (a fused 1x1 convolution to n channels, plus its max over the channels).  ``gate`` (an _Act in the sources' format, with
    ``dot``): ``dst`` receives gate * f(dot) instead (Attention_block's x * psi)."""
    mode = layer.get("mode", 1 if pk["precise"] else 0)
    if mode and (src0.fmt != mode or (src1 is not None and src1.fmt != mode)):
        raise RuntimeError(f"conv in mode {mode} got sources in format {src0.fmt}" + (f"/{src1.fmt}" if src1 is not None else ""))
    d = _lib.ConvDesc(mode, src0.ptr, src0.c, src0.ld, src0.lo,
                      src1.ptr if src1 is not None else None, src1.c if src1 is not None else 0,
                      src1.ld if src1 is not None else 0, src1.lo if src1 is not None else 0,
                      B, src0.h, src0.w, taps, 1 if up2x else 0, layer["w"].data_ptr(), layer["c_out"],
                      layer["scale"].data_ptr(), layer["shift"].data_ptr(), 1 if relu else 0,
                      dst.t.data_ptr() if dst is not None else None, dst.ld if dst is not None else 0, dst.off if dst is not None else 0,
                      dst.lo if dst is not None else 0, 0, k_chunk,
                      (dst.fmt if mode else 0) if dst is not None else 0, (pool.fmt if pool is not None else 0) if mode else 0, layer.get("lo_scale", 1.0 / LO_SCALE),
                      pk["sat_count"].data_ptr() if "sat_count" in pk else None, pool.ptr if pool is not None else None, pool.ld if pool is not None else 0, pool.lo if pool is not None else 0,
                      dot[0].data_ptr() if dot is not None else None, dot[4].data_ptr() if dot is not None and dot[4] is not None else None,
                      dot[1] if dot is not None else 0.0, dot[2] if dot is not None else 0.0, (1 if dot[3] else 0) if dot is not None else 0,
                      gate.ptr if gate is not None else None, gate.c if gate is not None else 0, gate.ld if gate is not None else 0,
                      gate.lo if gate is not None else 0,
                      dot[5] if dot is not None and len(dot) > 5 else 0,
                      dot[6].data_ptr() if dot is not None and len(dot) > 6 and dot[6] is not None else None,
                      dot[7].data_ptr() if dot is not None and len(dot) > 7 and dot[7] is not None else None)
    _lib.check(_lib.lib().nbp_conv_fwd(ctypes.byref(d), _stream()), "nbp_conv_fwd")


class _EvalGraph:
    """One captured eval forward: static input / output buffers + the CUDA graph of every kernel of every chunk."""

    def __init__(self, module, pk, x):
        B, C, S, S2 = x.shape
        dev = x.device
        self.pk = pk                                    # keeps the packed weights (whose pointers the graph holds) alive
        self.x = torch.empty_like(x)
        self.out1 = torch.empty((B, module.output_ch1, S // 4, S2 // 4), dtype=torch.float32, device=dev)
        self.out2 = torch.empty((B, module.output_ch2, S, S2), dtype=torch.float32, device=dev)
        self.vmax = torch.empty((B, S // 4, S2 // 4), dtype=torch.float32, device=dev)
        self.x.copy_(x)
        L = _lib.lib()
        # eager warm-up on a side stream (one-time kernel attribute set-up must not happen inside the capture)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            _forward_eval(pk, self.x[: min(B, module.max_chunk)], self.out1[: min(B, module.max_chunk)],
                          self.out2[: min(B, module.max_chunk)], self.vmax[: min(B, module.max_chunk)])
        torch.cuda.current_stream(dev).wait_stream(side)
        n0 = L.nbp_launch_count()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            module._eval_chunks(pk, self.x, self.out1, self.out2, self.vmax)
        self.n_kernels = int(L.nbp_launch_count() - n0)

    def replay(self, x, clone):
        if x.data_ptr() != self.x.data_ptr():
            self.x.copy_(x)
        self.graph.replay()
        _lib.lib().nbp_count_launches(self.n_kernels)
        if clone:
            return self.out1.clone(), self.out2.clone(), self.vmax.clone()
        return self.out1, self.out2, self.vmax


def _forward_eval(pk, x, out1, out2, vmax):
    """One chunk: x (B,5,S,S) fp32 -> out1 (B,8,S/4,S/4), out2 (B,1,S,S), vmax (B,S/4,S/4), written in place.
    Every activation is written in the second-plane format of the layers that read it (``mode`` of the consuming GEMM layers)."""
    L = _lib.lib()
    dev = x.device
    B, _, S, S2 = x.shape
    st = _stream()
    planes = 2 if pk["precise"] else 1
    mode = lambda name: pk[name].get("mode", 1) or 1          # single-plane fp16 tensors carry fmt 1 (no second plane is touched)

    def new(h, w, c, fmt):
        return _Act(torch.empty((B, h, w, planes * c), dtype=torch.float16, device=dev), c, planes * c,
                    c if planes == 2 else 0, h, w, 0, fmt)

    def same(*names):
        m = {mode(n) for n in names}
        if len(m) != 1:
            raise RuntimeError(f"layers {names} read the same tensor and must run the same numeric mode")
        return m.pop()

    def double_conv(name, src, out_fmt, pool_fmt=None, dot=None):
        """conv_block (nbp_model.py:8-21); ``pool_fmt``: the second conv also writes MaxPool2d(2,2) of its output (:113-121);
        ``dot``: the second conv feeds a fused 1-channel 1x1 head instead of storing its output (returns None, None)."""
        c_out = pk[name + ".a"]["c_out"]
        t = new(src.h, src.w, c_out, mode(name + ".b"))
        _conv(pk, pk[name + ".a"], B, src, 9, t)
        if dot is not None:
            _conv(pk, pk[name + ".b"], B, t, 9, None, dot=dot)
            return None, None
        y = new(src.h, src.w, c_out, out_fmt)
        p = new(src.h // 2, src.w // 2, c_out, pool_fmt) if pool_fmt is not None else None
        _conv(pk, pk[name + ".b"], B, t, 9, y, pool=p)
        return y, p

    # who reads the encoder outputs: x_l (l = 1..4) is the skip of decoder stage l+1 (attention conv, gate, concat) in decoder 2 and,
    # for l >= 3, decoder 1; its pooled copy feeds Conv{l+1}.a; x5 feeds the two Up5 convs
    stage_fmt = {lvl: same(*[f"{k}{lvl}_{dec}" + sfx for dec in (1, 2) if lvl in _DEC_LEVELS[dec]
                             for k, sfx in (("Att", ""), ("Up_conv", ".a"))]) for lvl in (5, 4, 3, 2)}

    # ---- encoder
    a = new(S, S2, 64, mode("Conv1.b"))
    stem = pk["stem"]
    _lib.check(L.nbp_conv_first(x.data_ptr(), B, stem["c_in"], S, S2, stem["w"].data_ptr(), stem["scale"].data_ptr(),
                                stem["shift"].data_ptr(), 64, 1, a.ptr, a.ld, a.lo, a.fmt, st), "nbp_conv_first")
    x1 = new(S, S2, 64, stage_fmt[2])
    p = new(S // 2, S2 // 2, 64, mode("Conv2.a"))
    _conv(pk, pk["Conv1.b"], B, a, 9, x1, pool=p)
    del a
    skips = {1: x1}
    for lvl in range(2, 5):
        skips[lvl], p = double_conv(f"Conv{lvl}", p, stage_fmt[lvl + 1], pool_fmt=mode(f"Conv{lvl + 1}.a"))
    skips[5], _ = double_conv("Conv5", p, same("Up5_1", "Up5_2"))

    def decoder_stage(d, lvl, dec, out_fmt, dot=None):
        """Up{lvl}_{dec} -> Att{lvl}_{dec} -> cat -> Up_conv{lvl}_{dec} (nbp_model.py:124-129)."""
        t = f"{lvl}_{dec}"
        skip = skips[lvl - 1]
        f_l = skip.c
        fmt = stage_fmt[lvl]
        cat = new(skip.h, skip.w, 2 * f_l, fmt)          # channels [skip*psi | up-conv output]
        g = cat.channels(f_l, f_l)
        _conv(pk, pk[f"Up{t}"], B, d, 4, g, up2x=True)   # upsample fused: d is read at its own (half) resolution
        att = pk[f"Att{t}"]
        gated = cat.channels(0, f_l)
        if FUSE_GATE and FUSE_DOT and att["c_out"] <= 128:
            # the whole gate in the attention GEMM's epilogue: psi = sigmoid(BN(w_psi . a)) (dot epilogue) and x * psi written into the
            # concat buffer by the same threads; neither `a` nor psi go to memory and the skip tensor is read once (TMA) + once from L2
            _conv(pk, att, B, g, 1, gated, relu=True, src1=skip, dot=(att["w_psi"], att["psi_scale"], att["psi_shift"], True, None), gate=skip)
        elif FUSE_DOT and att["c_out"] <= 128:
            # psi = sigmoid(BN(w_psi . a)) leaves the attention GEMM's epilogue directly (dot epilogue): `a` never goes to memory
            psi = torch.empty((B, skip.h, skip.w), dtype=torch.float32, device=dev)
            _conv(pk, att, B, g, 1, None, relu=True, src1=skip, dot=(att["w_psi"], att["psi_scale"], att["psi_shift"], True, psi))
            _lib.check(L.nbp_att_scale(psi.data_ptr(), skip.ptr, f_l, skip.ld, skip.lo,
                                       gated.t.data_ptr(), gated.ld, gated.off, gated.lo, B * skip.h * skip.w, fmt, st), "nbp_att_scale")
            del psi
        else:
            arelu = new(skip.h, skip.w, att["c_out"], 1)     # read by the gate kernel only (f_int can be 32 < one e4m3 group): 22-bit format
            _conv(pk, att, B, g, 1, arelu, relu=True, src1=skip)
            _lib.check(L.nbp_att_gate(arelu.ptr, arelu.c, arelu.ld, arelu.lo, skip.ptr, f_l, skip.ld, skip.lo,
                                      att["w_psi"].data_ptr(), att["psi_scale"], att["psi_shift"],
                                      gated.t.data_ptr(), gated.ld, gated.off, gated.lo, B * skip.h * skip.w, fmt, st), "nbp_att_gate")
            del arelu
        return double_conv(f"Up_conv{t}", cat, out_fmt, dot=dot)[0]

    def head(name, d, sigmoid, out, out_max=None):
        _lib.check(L.nbp_conv1x1_head(d.ptr, d.c, d.ld, d.lo, pk[name]["w"].data_ptr(), pk[name]["b"].data_ptr(),
                                      out.shape[1], 1 if sigmoid else 0, out.data_ptr(),
                                      out_max.data_ptr() if out_max is not None else None, B, d.h * d.w, d.fmt, st), "nbp_conv1x1_head")

    assert out1.is_contiguous() and out2.is_contiguous() and vmax.is_contiguous()
    # ---- decoder 1 -> value map at S/4 (+ its max over the 8 headings)
    d = decoder_stage(skips[5], 5, 1, mode("Up4_1"))
    f1 = pk["Final1"]
    if FUSE_HEAD8 and FUSE_DOT and f1["w"].shape[0] <= 8 and f1["w"].shape[1] in (32, 64, 128):
        # Final1 (1x1 conv to the 8 headings, nbp_model.py:89,133) and its max over the headings inside the epilogue of Up_conv4_1's second conv
        decoder_stage(d, 4, 1, 1, dot=(f1["w"].reshape(-1), 1.0, 0.0, False, out1, f1["w"].shape[0], f1["b"], vmax))
    else:
        d = decoder_stage(d, 4, 1, 1)                     # read by the CUDA-core head only: keep the 22-bit format
        head("Final1", d, False, out1, vmax)
    # ---- decoder 2 -> obstacle map at S
    d = decoder_stage(skips[5], 5, 2, mode("Up4_2"))
    for lvl in (4, 3):
        d = decoder_stage(d, lvl, 2, mode(f"Up{lvl - 1}_2"))
    f2 = pk["Final2"]
    if FUSE_DOT and f2["w"].shape[0] == 1 and f2["w"].shape[1] in (32, 64, 128):
        # Final2 (1x1 conv to one channel + sigmoid, nbp_model.py:106-108) inside the epilogue of Up_conv2_2's second conv
        decoder_stage(d, 2, 2, 1, dot=(f2["w"].reshape(-1), 1.0, f2["b_host"], True, out2))
    else:
        head("Final2", decoder_stage(d, 2, 2, 1), True, out2)
