"""Data-parallel training support for the NBP network (BASELINE.json configs[2]).

The reference has no multi-GPU hook for NBP (``train_nbp.py:27-30`` is ``if params.ddp: pass``); its single-GPU loop is
``train_experience_data`` (next_best_path/utility/nbp_utils.py:340-395): micro-batches of <= 56 tiles, gradients accumulated
over 8 micro-batches without loss scaling, then one AdamW step.  ``train_step`` below is that loop with ONE collective per
optimizer step: an NCCL all-reduce (SUM, then / world) of the flat fp32 gradient buffer (49.96 M parameters = 199.9 MB) over
NVLink / NVSwitch.  BatchNorm statistics stay per rank, as in the reference's single-GPU run.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """One persistent flat fp32 buffer for all gradients; ``sync()`` = copy in, all-reduce, scale, copy out."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()

    def sync(self):
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if world == 1:
            return 0
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            else:
                v.copy_(p.grad)
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.div_(world)
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                p.grad = v.clone()
            else:
                p.grad.copy_(v)
        return self.flat.numel() * 4


def reduce_scalar(x: torch.Tensor) -> torch.Tensor:
    """Mean of a scalar over ranks for logging (reduce_tensor, macarons/utility/macarons_utils.py:235-240)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return x
    y = x.detach().clone()
    dist.all_reduce(y, op=dist.ReduceOp.SUM)
    return y / dist.get_world_size()


def sparse_value_loss(model, out1, out2, target_pixels, target_gains, layout):
    """The sparse gather of nbp_utils.py:373-381 + NBP.loss: target_pixels (B,K,3) long = (channel, gx, gy)."""
    B = out1.shape[0]
    b_idx = torch.arange(B, device=out1.device).view(B, 1).expand(B, target_pixels.shape[1])
    pred = out1[b_idx, target_pixels[..., 0], target_pixels[..., 1], target_pixels[..., 2]]
    return model.loss(pred, target_gains, out2, layout)


def train_step(model, optimizer, micro_batches, reducer: FlatGradAllReduce | None = None):
    """One optimizer step over a list of micro-batches (inputs, target_pixels, target_gains, layout), gradients accumulated
    un-scaled as the reference does (nbp_utils.py:383-390).  Returns the mean micro-batch loss (rank-local)."""
    model.train()
    optimizer.zero_grad(set_to_none=True)
    total = 0.0
    for (x, tp, tg, layout) in micro_batches:
        out1, out2 = model(x)
        loss = sparse_value_loss(model, out1, out2, tp, tg, layout)
        loss.backward()
        total += float(loss.detach())
    if reducer is not None:
        reducer.sync()
    optimizer.step()
    return total / max(len(micro_batches), 1)
