"""Per-micro-batch timing and allocator statistics of the NBP training step (the probe that found the tape reference cycles,
profiles/r01_train_bench.txt): python scripts/train_step_probe.py [micro_batch]."""
import sys, os, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nextbestpath_b200.networks import NBP
from nextbestpath_b200.train import sparse_value_loss
from nextbestpath_b200 import synthetic as syn    # seeded weights / synthetic count images come from the product package
dev = torch.device("cuda", 0)
net = NBP(); net.load_state_dict(syn.seeded_nbp_state_dict(net, 9)); net.to(dev).train()
opt = torch.optim.AdamW(net.parameters(), lr=1e-3)
S, K, b = 256, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator().manual_seed(0)
def mb(i):
    x = syn.count_like_input(b, S, seed=i).to(dev)
    tp = torch.stack((torch.randint(0, 8, (b, K), generator=g), torch.randint(0, S // 4, (b, K), generator=g), torch.randint(0, S // 4, (b, K), generator=g)), -1).to(dev)
    return (x, tp, (torch.rand(b, K, generator=g) * 10).to(dev), (torch.rand(b, 1, S, S, generator=g) < 0.2).float().to(dev))
mbs = [mb(i) for i in range(4)]
for step in range(3):
    opt.zero_grad(set_to_none=True)
    for i, (x, tp, tg, lay) in enumerate(mbs):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        o1, o2 = net(x)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        loss = sparse_value_loss(net, o1, o2, tp, tg, lay)
        loss.backward()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        st = torch.cuda.memory_stats()
        print(f"step {step} mb {i}: fwd {1e3*(t1-t0):.1f} ms bwd {1e3*(t2-t1):.1f} ms  alloc_retries {st['num_alloc_retries']} segs {st['segment.all.current']} reserved {st['reserved_bytes.all.current']/1e9:.1f} GB cudaMallocs {st['segment.all.allocated']}")
    t0 = time.perf_counter(); opt.step(); torch.cuda.synchronize(); print(f"  opt.step {1e3*(time.perf_counter()-t0):.1f} ms")
