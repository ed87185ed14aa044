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
    """One persistent flat fp32 buffer that IS the gradient storage: every ``p.grad`` is a view into it, so backward passes
    accumulate straight into the buffer (autograd adds in place into an existing ``.grad``), ``zero()`` is one memset and
    ``sync()`` all-reduces the buffer where it lies -- no copy in or out.  If a caller detached the views
    (``optimizer.zero_grad()`` sets ``.grad`` to None, as the reference's loop does, nbp_utils.py:389), ``sync()`` copies those
    gradients in once and re-attaches.

    Every parameter always has a gradient tensor (zeros if nothing flowed into it) on every world size, so a 1-GPU and an
    N-GPU run apply the same AdamW update (weight decay included) to a parameter without gradient."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=torch.float32, device=dev)
        self.views, o = [], 0
        for p in self.params:
            self.views.append(self.flat[o:o + p.numel()].view_as(p))
            o += p.numel()
        self._ev = None
        self._attach(copy=True)

    def _attach(self, copy):
        for p, v in zip(self.params, self.views):
            if p.grad is not v:
                if copy and p.grad is not None:
                    v.copy_(p.grad)
                p.grad = v

    def zero(self):
        """optimizer.zero_grad() for attached gradients: one memset, the views stay attached."""
        self._attach(copy=False)
        self.flat.zero_()

    def sync(self):
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        for p, v in zip(self.params, self.views):           # gradients a caller detached (or never attached) come in once
            if p.grad is not v:
                if p.grad is None:
                    v.zero_()
                else:
                    v.copy_(p.grad)
                p.grad = v
        if world == 1:
            return 0
        timed = self.flat.is_cuda
        if timed:
            self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self._ev[0].record()
        dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.flat.div_(world)
        if timed:
            self._ev[1].record()
        return self.flat.numel() * 4

    def nonfinite(self) -> bool:
        """True if any gradient is inf / nan (one reduction + one host read, like GradScaler's found_inf check)."""
        return not bool(torch.isfinite(self.flat).all())

    def last_ms(self) -> float:
        """Device time of the last all-reduce (+ the division) on the calling stream; 0 for a single process."""
        if self._ev is None:
            return 0.0
        self._ev[1].synchronize()
        return self._ev[0].elapsed_time(self._ev[1])


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


def train_step(model, optimizer, micro_batches, reducer: FlatGradAllReduce | None = None, skip_nonfinite: bool = False):
    """One optimizer step over a list of micro-batches (inputs, target_pixels, target_gains, layout), gradients accumulated
    as the reference does (nbp_utils.py:383-390; its GradScaler multiplies by a constant that unscale_ removes again).
    ``skip_nonfinite``: drop the step if any (all-reduced) gradient is inf / nan -- the reference's GradScaler.step() does the same
    through its found_inf check; off by default because it costs one host synchronisation per step.
    Returns the mean micro-batch loss (rank-local)."""
    model.train()
    if reducer is not None:
        reducer.zero()
    else:
        optimizer.zero_grad(set_to_none=True)
    losses = []
    for (x, tp, tg, layout) in micro_batches:
        out1, out2 = model(x)
        loss = sparse_value_loss(model, out1, out2, tp, tg, layout)
        loss.backward()
        losses.append(loss.detach())
    if reducer is not None:
        reducer.sync()
        if skip_nonfinite and reducer.nonfinite():
            return float("nan")
    optimizer.step()
    return float(torch.stack(losses).mean()) if losses else 0.0
