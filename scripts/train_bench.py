#!/usr/bin/env python
"""Training-step benchmark (BASELINE.json configs[2] shape: NBP fwd + loss + bwd + AdamW on 256x256 map tiles, data parallel).

    python scripts/train_bench.py --tiles 8 --micro 4 --steps 3                          # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... scripts/train_bench.py ...      # N GPUs, NCCL gradient all-reduce

Per rank and per optimizer step: `tiles` tiles in micro-batches of `micro` (gradients accumulated, one all-reduce, one AdamW
step).  Reports tiles/s over all ranks and the algorithmic TFLOP/s (547 GFLOP per tile = 3 x forward)."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch.distributed as dist
from nextbestpath_b200.networks import NBP
from nextbestpath_b200.train import FlatGradAllReduce, train_step
from nextbestpath_b200 import ops
from nextbestpath_b200 import synthetic as syn    # seeded weights / synthetic count images come from the product package


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tiles", type=int, default=8); ap.add_argument("--micro", type=int, default=4)
    ap.add_argument("--grid", type=int, default=256); ap.add_argument("--steps", type=int, default=3); ap.add_argument("--warmup", type=int, default=1)
    a = ap.parse_args()
    world, rank, local = int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local); dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    net = NBP(); net.load_state_dict(syn.seeded_nbp_state_dict(net, 9)); net.to(dev)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)    # nbp_utils.py:228
    red = FlatGradAllReduce(net.parameters())
    S, K = a.grid, 64
    g = torch.Generator().manual_seed(rank)
    mbs = []
    for i in range(0, a.tiles, a.micro):
        b = min(a.micro, a.tiles - i)
        x = syn.count_like_input(b, S, seed=100 * rank + i).to(dev)
        tp = torch.stack((torch.randint(0, 8, (b, K), generator=g), torch.randint(0, S // 4, (b, K), generator=g),
                          torch.randint(0, S // 4, (b, K), generator=g)), -1).to(dev)
        mbs.append((x, tp, (torch.rand(b, K, generator=g) * 10).to(dev), (torch.rand(b, 1, S, S, generator=g) < 0.2).float().to(dev)))
    for _ in range(a.warmup):
        loss = train_step(net, opt, mbs, red)
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        loss = train_step(net, opt, mbs, red)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
    if rank == 0:
        tiles_s = world * a.tiles / (ms / 1e3)
        print(json.dumps({"metric": "nbp_train_tiles_per_sec", "value": tiles_s, "unit": "tiles/s", "n_gpus": world, "ms_per_optimizer_step": ms,
                          "tiles_per_gpu_per_step": a.tiles, "micro_batch": a.micro, "grid": S, "loss": loss,
                          "algorithmic_tflops": tiles_s * 3 * {128: 45.603, 256: 182.411, 512: 729.645}[S] / 1e3,
                          "grad_allreduce_bytes": red.flat.numel() * 4 if world > 1 else 0, "gpu_launches_per_step": (ops.launch_count() - l0) // a.steps,
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9, "precision": "fp16x2 operands, fp32 accumulate"}))
    if world > 1: dist.destroy_process_group()


if __name__ == "__main__":
    main()
