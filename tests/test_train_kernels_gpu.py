"""GPU unit parity of the backward tensor-core kernels: nbp_conv_wgrad (tcgen05 GEMM over channel-major operands)
and the data-gradient use of nbp_conv_fwd (flipped / transposed weights, fp32 epilogue), against fp64 torch autograd."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

from nextbestpath_b200 import _lib
from nextbestpath_b200.networks import nbp_model as M
from nextbestpath_b200.networks import nbp_train as T

pytestmark = pytest.mark.gpu
DEV = "cuda:0"

CASES = [
    # n, h, w, cin, cout, taps
    (2, 16, 16, 64, 64, 9),
    (2, 8, 8, 128, 64, 9),        # x is the row operand (cin > cout)
    (1, 32, 32, 64, 128, 9),
    (3, 4, 4, 256, 128, 9),       # rows padded from 4 to 8 pixels
    (2, 16, 16, 64, 32, 1),       # attention 1x1, BLOCK_N = 32, dgrad K padded 32 -> 64
    (1, 64, 64, 64, 64, 9),       # one K slice = one image row
    (2, 8, 16, 128, 256, 1),
    (2, 128, 128, 64, 64, 9),     # long pixel reduction (32768 pixels)
    (1, 128, 128, 128, 64, 9),
]


def _act_split(x):
    """NCHW fp32 -> _Act (NHWC fp16x2) on the device."""
    a = x.permute(0, 2, 3, 1).contiguous()
    hi = a.to(torch.float16)
    lo = ((a - hi.float()) * 2048.0).to(torch.float16)
    t = torch.cat((hi, lo), dim=-1).contiguous().to(DEV)
    c = a.shape[-1]
    return M._Act(t, c, 2 * c, c, a.shape[1], a.shape[2])


@pytest.mark.parametrize("n,h,w,cin,cout,taps", CASES)
def test_conv_backward_kernels(n, h, w, cin, cout, taps):
    g = torch.Generator().manual_seed(n * 100 + h + cin + cout + taps)
    k = 3 if taps == 9 else 1
    x = torch.randn(n, cin, h, w, generator=g)
    wt = torch.randn(cout, cin, k, k, generator=g) / (cin * taps) ** 0.5
    dz = torch.randn(n, cout, h, w, generator=g) * 1e-5            # small gradients: exercises the power-of-two scaling
    xd, wd_ = x.double().requires_grad_(True), wt.double().requires_grad_(True)
    F.conv2d(xd, wd_, padding=k // 2).backward(dz.double())
    tape = T._Tape(torch.device(DEV), n)
    src = _act_split(x)
    dz_nhwc = dz.permute(0, 2, 3, 1).reshape(-1, cout).contiguous().to(DEV)
    amax = dz_nhwc.abs().max().reshape(1).contiguous()
    # dz as the scaled fp16x2 GEMM operand (in the network nbp_bn_bwd_split writes it directly; here the stand-alone converter)
    dzs, inv_vec = T._new_dz_operand(tape, h, w, cout, cin)
    T._chk(tape.L.nbp_to_split_nhwc(dz_nhwc.data_ptr(), cout, n * h * w, cout, amax.data_ptr(), dzs.ptr, dzs.ld, dzs.lo, inv_vec.data_ptr(), inv_vec.numel(),
                                    torch.cuda.current_stream().cuda_stream), "nbp_to_split_nhwc")
    dsrc = T._conv_backward(tape, "layer", wt.to(DEV), src, dzs, inv_vec, taps)
    torch.cuda.synchronize()
    dW = tape.pgrads["layer.weight"].cpu()
    rel = lambda a, b: float((a.double() - b).norm() / b.norm())
    e_w = rel(dW, wd_.grad)
    e_x = rel(dsrc.view(n, h, w, cin).permute(0, 3, 1, 2).cpu(), xd.grad)
    print(f"n={n} h={h} w={w} cin={cin} cout={cout} taps={taps}: wgrad rel {e_w:.2e}, dgrad rel {e_x:.2e}")
    assert e_w <= 2e-5 and e_x <= 2e-5


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 16, 16, 64, 64), (2, 128, 128, 64, 64), (1, 32, 32, 128, 256)])
def test_conv_bn_relu_block_forward_backward(n, h, w, cin, cout):
    """One conv_block half (3x3 conv + train-mode BatchNorm + ReLU, nbp_model.py:11-13) forward and backward against
    fp64 torch autograd: output, running statistics, d(input), d(weight), d(gamma), d(beta)."""
    g = torch.Generator().manual_seed(n + h + cin + cout)
    x = torch.relu(torch.randn(n, cin, h, w, generator=g)) * 2.0
    wt = torch.randn(cout, cin, 3, 3, generator=g) / (cin * 9) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    gamma, beta = torch.rand(cout, generator=g) + 0.5, torch.randn(cout, generator=g) * 0.1
    rm, rv = torch.zeros(cout), torch.ones(cout)
    dy = torch.randn(n, cout, h, w, generator=g) * 1e-4
    # fp64 reference
    xd = x.double().requires_grad_(True)
    P = [t.double().requires_grad_(True) for t in (wt, b, gamma, beta)]
    rmd, rvd = rm.double().clone(), rv.double().clone()
    yref = F.relu(F.batch_norm(F.conv2d(xd, P[0], P[1], padding=1), rmd, rvd, P[2], P[3], True, 0.1, 1e-5))
    yref.backward(dy.double())
    # ours
    tape = T._Tape(torch.device(DEV), n)
    sd = {"c.weight": wt.to(DEV), "c.bias": b.to(DEV), "b.weight": gamma.to(DEV), "b.bias": beta.to(DEV),
          "b.running_mean": rm.to(DEV), "b.running_var": rv.to(DEV), "b.num_batches_tracked": torch.zeros((), dtype=torch.long, device=DEV)}
    src = _act_split(x)
    y, bw = T._cbr(tape, sd, "c", "b", src)
    dsrc = bw(dy.permute(0, 2, 3, 1).reshape(-1, cout).contiguous().to(DEV))
    torch.cuda.synchronize()
    rel = lambda a, bb: float((a.double() - bb).norm() / bb.norm())
    yv = (y.t[..., :cout].float() + y.t[..., cout:].float() / 2048.0).cpu().permute(0, 3, 1, 2)
    errs = {"y": rel(yv, yref.detach()), "running_mean": rel(sd["b.running_mean"].cpu(), rmd), "running_var": rel(sd["b.running_var"].cpu(), rvd),
            "dx": rel(dsrc.view(n, h, w, cin).permute(0, 3, 1, 2).cpu(), xd.grad), "dW": rel(tape.pgrads["c.weight"].cpu(), P[0].grad),
            "dgamma": rel(tape.pgrads["b.weight"].cpu(), P[2].grad), "dbeta": rel(tape.pgrads["b.bias"].cpu(), P[3].grad)}
    print(f"n={n} h={h} w={w} cin={cin} cout={cout}: " + ", ".join(f"{k} {v:.1e}" for k, v in errs.items()))
    assert all(v <= 1e-4 for v in errs.values()), errs
