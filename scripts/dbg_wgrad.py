import sys, ctypes
sys.path.insert(0, '.')
import torch
from nextbestpath_b200 import _lib
from nextbestpath_b200.networks import nbp_model as M, nbp_train as T
DEV='cuda:0'
L=_lib.lib()
dbg = torch.zeros(4, dtype=torch.int32).pin_memory()
cudart = torch.cuda.cudart()
# pinned torch memory is mapped & device-accessible under UVA: its host pointer is usable on the device
_lib.check(L.nbp_debug_attach_wgrad(dbg.data_ptr()), 'attach')
def run(n,h,w,cin,cout,taps):
    g = torch.Generator().manual_seed(1)
    k = 3 if taps==9 else 1
    x = torch.randn(n,cin,h,w,generator=g); wt = torch.randn(cout,cin,k,k,generator=g)/(cin*taps)**0.5
    dz = torch.randn(n,cout,h,w,generator=g)
    a = x.permute(0,2,3,1).contiguous(); hi=a.half(); lo=((a-hi.float())*2048).half()
    src = M._Act(torch.cat((hi,lo),-1).contiguous().to(DEV), cin, 2*cin, cin, h, w)
    tape = T._Tape(torch.device(DEV), n)
    dzn = dz.permute(0,2,3,1).reshape(-1,cout).contiguous().to(DEV)
    amax = dzn.abs().max().reshape(1).contiguous()
    try:
        T._conv_backward(tape,'l',wt.to(DEV),src,dzn,amax,taps,need_dsrc=False)
        torch.cuda.synchronize()
        xd, wd = x.double(), wt.double().requires_grad_(True)
        torch.nn.functional.conv2d(xd, wd, padding=k//2).backward(dz.double())
        mine = tape.pgrads['l.weight'].cpu().double()
        print('nan count', int(torch.isnan(mine).sum()), 'mine norm', float(mine[~torch.isnan(mine)].norm()), 'ref norm', float(wd.grad.norm()), 'sample', mine.flatten()[:4].tolist(), wd.grad.flatten()[:4].tolist())
        e = float((mine-wd.grad).norm()/wd.grad.norm())
        print((n,h,w,cin,cout,taps), 'OK rel', e, flush=True)
    except Exception as ex:
        print((n,h,w,cin,cout,taps), 'FAIL', str(ex)[:100], 'dbg', dbg.tolist(), flush=True)
        raise SystemExit(1)
import os
case = eval(os.environ.get('CASE','(1,64,64,64,64,9)'))
run(*case)
