"""CPU (gloo, world_size 2): the gradient all-reduce of the data-parallel training step averages per-rank gradients,
and the optimizer then takes identical steps on every rank."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from nextbestpath_b200.train import FlatGradAllReduce, reduce_scalar


class _Tiny(nn.Module):
    def __init__(self):
        super().__init__()
        torch.manual_seed(0)
        self.a = nn.Linear(4, 3)
        self.b = nn.Linear(3, 2)
        self.unused = nn.Parameter(torch.zeros(2))      # like log_vars when a loss term is absent

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _Tiny()
        red = FlatGradAllReduce(net.parameters())
        opt = torch.optim.AdamW(net.parameters(), lr=1e-2)
        g = torch.Generator().manual_seed(100 + rank)
        x = torch.randn(8, 4, generator=g)
        loss = net(x).pow(2).mean()
        loss.backward()
        local = [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in net.parameters()]
        nbytes = red.sync()
        opt.step()
        q.put((rank, [t.numpy() for t in local], [p.grad.numpy().copy() for p in net.parameters()],
               [p.detach().numpy().copy() for p in net.parameters()], nbytes, float(reduce_scalar(loss.detach()))))
    finally:
        dist.destroy_process_group()


def test_flat_gradient_allreduce_two_ranks():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, l0, g0, w0, n0, m0), (_, l1, g1, w1, n1, m1) = res
    import numpy as np
    for a, b, ga, gb, wa, wb in zip(l0, l1, g0, g1, w0, w1):
        assert np.allclose(ga, (a + b) / 2, atol=1e-7) and np.array_equal(ga, gb)      # averaged, identical on both ranks
        assert np.array_equal(wa, wb)                                                   # so the replicas stay in sync
    assert n0 == n1 == 4 * sum(p.numel() for p in _Tiny().parameters())
    assert m0 == m1


def test_single_process_keeps_gradients_and_attaches_the_flat_views():
    net = _Tiny()
    net(torch.randn(2, 4)).sum().backward()
    before = {n: (p.grad.clone() if p.grad is not None else torch.zeros_like(p)) for n, p in net.named_parameters()}
    red = FlatGradAllReduce(net.parameters())                      # adopts the existing gradients
    assert red.sync() == 0                                         # single process: no collective
    for (n, p), v in zip(net.named_parameters(), red.views):
        assert p.grad is v and torch.equal(p.grad, before[n])      # same values, now living in the flat buffer
    # backward accumulates straight into the flat buffer; zero() is one memset and keeps the views attached
    net(torch.randn(2, 4)).sum().backward()
    assert all(p.grad is v for p, v in zip(net.parameters(), red.views)) and float(red.flat.abs().sum()) > 0
    red.zero()
    assert float(red.flat.abs().sum()) == 0 and all(p.grad is v for p, v in zip(net.parameters(), red.views))
    # a caller that detaches (optimizer.zero_grad() sets .grad to None, as the reference loop does) is picked up by sync()
    torch.optim.SGD(net.parameters(), lr=0.1).zero_grad()
    net(torch.ones(2, 4)).sum().backward()
    g = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
    red.sync()
    for (n, p), v in zip(net.named_parameters(), red.views):
        assert p.grad is v and torch.equal(v, g.get(n, torch.zeros_like(p)))
    assert not red.nonfinite()
    red.flat[0] = float("inf")
    assert red.nonfinite()
