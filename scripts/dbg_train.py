import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import torch
from nextbestpath_b200.networks import NBP
from oracle import nbp_torch as NT
from test_train_gpu import _targets, _loss, _oracle_step
DEV='cuda:0'; B,S=2,128
net = NBP(); net.load_state_dict(NT.golden_state_dict(seed=9)); net.to(DEV).train()
xb = NT.count_like_input(B,S,seed=4); ti,tv,lay = _targets(B,S)
p1,p2 = net(xb.to(DEV)); loss=_loss(net.loss,p1,p2,ti,tv,lay); loss.backward(); torch.cuda.synchronize()
r1,r2,rl,g64,sd64 = _oracle_step(xb,ti,tv,lay,torch.float64)
rel = lambda a,b: float((a.double()-b.double()).norm()/b.double().norm().clamp_min(1e-30))
print('fwd', rel(p1.detach().cpu(), r1), rel(p2.detach().cpu(), r2))
for n,p in net.named_parameters():
    g=g64[n]
    print(f"{n:32s} ref_norm {float(g.norm()):.3e}  err {rel(p.grad.cpu(), g):.2e}")
