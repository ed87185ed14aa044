import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch, torch.nn.functional as F
from nextbestpath_b200.networks import NBP, nbp_train as T
from oracle import nbp_torch as NT
from test_train_gpu import _targets, _loss
DEV='cuda:0'; B,S=2,128
rec = {}
orig = T._conv_backward
def spy(t, name, w, src, dz, amax, taps, need_dsrc=True):
    out = orig(t, name, w, src, dz, amax, taps, need_dsrc)
    if name in ("Up_conv2_2.conv.3", "Up_conv4_1.conv.3", "Up_conv2_2.conv.0"):
        c = src.c
        base = src.t[..., src.off:src.off + c].double() + src.t[..., src.off + src.lo: src.off + src.lo + c].double() / 2048.0
        x = base.permute(0, 3, 1, 2).contiguous()
        dzn = dz.double().view(t.B, src.h, src.w, -1).permute(0, 3, 1, 2).contiguous()
        k = 3 if taps == 9 else 1
        with torch.enable_grad():
            wd = w.detach().double().clone().requires_grad_(True)
            xd = x.detach().clone().requires_grad_(True)
        with torch.enable_grad():
            F.conv2d(xd, wd, padding=k // 2).backward(dzn)
        rec[name] = (wd.grad.clone(), xd.grad.clone(), out.clone() if out is not None else None, float(amax.item()), float(dz.abs().max().item()), float(dz.abs().median().item()))
    return out
T._conv_backward = spy
net = NBP(); net.load_state_dict(NT.golden_state_dict(seed=9)); net.to(DEV).train()
xb = NT.count_like_input(B,S,seed=4); ti,tv,lay = _targets(B,S)
p1,p2 = net(xb.to(DEV)); loss=_loss(net.loss,p1,p2,ti,tv,lay); loss.backward(); torch.cuda.synchronize()
rel = lambda a,b: float((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30))
params = dict(net.named_parameters())
for name,(gw,gx,dx,amax,mx,med) in rec.items():
    print(name, 'wgrad kernel vs torch-from-same-inputs', rel(params[name+'.weight'].grad, gw), 'dgrad', rel(dx.view(B, gx.shape[2], gx.shape[3], -1).permute(0,3,1,2), gx) if dx is not None else None, 'amax', amax, 'true max', mx, 'median', med)
