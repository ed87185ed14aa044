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
    dsrc = T._conv_backward(tape, "layer", wt.to(DEV), src, dz_nhwc, amax, taps)
    torch.cuda.synchronize()
    dW = tape.pgrads["layer.weight"].cpu()
    rel = lambda a, b: float((a.double() - b).norm() / b.norm())
    e_w = rel(dW, wd_.grad)
    e_x = rel(dsrc.view(n, h, w, cin).permute(0, 3, 1, 2).cpu(), xd.grad)
    print(f"n={n} h={h} w={w} cin={cin} cout={cout} taps={taps}: wgrad rel {e_w:.2e}, dgrad rel {e_x:.2e}")
    assert e_w <= 2e-5 and e_x <= 2e-5
