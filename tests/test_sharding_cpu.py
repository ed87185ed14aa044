"""CPU (gloo, world_size 2): the N>1 host logic of the rollout -- scene sharding and max-over-ranks timing."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from nextbestpath_b200.rollout import interpolated_poses, max_over_ranks, shard_scenes


def test_shard_scenes_partitions_exactly():
    for n in (1, 7, 128, 256, 257):
        for world in (1, 2, 3, 4, 8):
            blocks = [shard_scenes(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == n
            for (f0, c0), (f1, _) in zip(blocks, blocks[1:]):
                assert f1 == f0 + c0
            counts = [c for _, c in blocks]
            assert max(counts) - min(counts) <= 1
    assert shard_scenes(256, 8, 3) == (96, 32)          # 256 scenes -> 32 per GPU at 8
    with pytest.raises(ValueError):
        shard_scenes(4, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = shard_scenes(9, world, rank)
        # every rank reports a different elapsed time; all must agree on the max
        t = max_over_ranks(10.0 + 5.0 * rank)
        owned = torch.zeros(9, dtype=torch.int32)
        owned[first:first + count] = 1
        dist.all_reduce(owned)                      # test-only collective: every scene is owned exactly once
        q.put((rank, first, count, t, owned.tolist()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_sharding_and_timing():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 4)]
    assert all(r[3] == 15.0 for r in res)
    assert all(r[4] == [1] * 9 for r in res)


def test_interpolated_poses_wraps_azimuth():
    old = torch.tensor([[0.0, 1.0, 0.0, 0.0, 0.0], [3.0, 1.0, 0.0, 0.0, 315.0]])
    new = torch.tensor([[0.0, 1.0, 0.0, 0.0, 315.0], [3.0, 1.0, 3.0, 0.0, 0.0]])
    p = interpolated_poses(old, new, [0, 7], [7, 0])
    assert p.shape == (4, 2, 5)
    assert torch.allclose(p[:, 0, 4], torch.tensor([-11.25, -22.5, -33.75, 315.0]))      # 0 -> 315 goes through -45, not +315
    assert torch.allclose(p[:, 1, 4], torch.tensor([326.25, 337.5, 348.75, 0.0]))
    assert torch.equal(p[3], new)
