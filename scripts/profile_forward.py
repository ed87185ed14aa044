#!/usr/bin/env python
"""One eval forward of the NBP network (kernel by kernel, no graph) for profiling under ncu.

    ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,... --clock-control none \
        -k regex:conv_gemm python scripts/profile_forward.py mixed 32 256

Prints the layer order so that the i-th conv_gemm launch of the LAST forward can be matched to its layer."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nextbestpath_b200 import synthetic as syn

precision = sys.argv[1] if len(sys.argv) > 1 else "mixed"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
S = int(sys.argv[3]) if len(sys.argv) > 3 else 256
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = "cuda:0"
net = syn.calibrated_nbp(dev, seed=9)
net.precision = precision
net.use_cuda_graph = False
x = syn.count_like_input(B, S, seed=3).to(dev)
with torch.no_grad():
    net(x)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()          # ncu --profile-from-start off: only the forwards below are captured
    e0.record()
    for _ in range(reps):
        net(x)
    e1.record()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print(f"{precision} B={B} S={S}: {e0.elapsed_time(e1) / reps:.3f} ms per forward (kernel by kernel, {reps} reps), env "
      + " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("NBP_CONV")))
order = ["Conv1.b"] + [f"Conv{l}.{ab}" for l in range(2, 6) for ab in "ab"]
for dec, lvls in ((1, (5, 4)), (2, (5, 4, 3, 2))):
    for l in lvls:
        order += [f"Up{l}_{dec}", f"Att{l}_{dec}", f"Up_conv{l}_{dec}.a", f"Up_conv{l}_{dec}.b"]
print("conv launch order:", " ".join(order))
