"""GPU, BASELINE configs[1] sizes (256x456 frames, 256x256 grid, ~4k-triangle scenes, gathering factor 0.05): properties that do
not need the (slow) CPU oracle -- conservation of points through back-projection and binning, depth range, render idempotence,
batch-composition independence of the network."""
import numpy as np
import pytest
import torch

from nextbestpath_b200 import synthetic as syn
from nextbestpath_b200.networks import NBP
from nextbestpath_b200.rollout import RolloutEngine
from oracle import nbp_torch as NT

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_full_size_rollout_properties():
    B, n_steps, S, H, W = 6, 3, 256, 256, 456
    scenes = [syn.make_scene(300 + i, "simple") for i in range(B)]
    walks = [syn.random_walk(sc, n_steps + 1, seed=300 + i) for i, sc in enumerate(scenes)]
    poses = np.stack([w[0] for w in walks]); az = np.stack([w[1] for w in walks])
    net = NBP(); net.load_state_dict(NT.golden_state_dict(seed=9)); net.to(DEV).eval()
    eng = RolloutEngine(scenes, net, DEV, S=S, H=H, W=W, max_steps=n_steps + 1, gathering_factor=0.05, seed=5)
    eng.reset(poses[:, 0])
    key0 = eng.frames[0].clone()
    # ---- render: idempotent, depth of every hit at or beyond the clip plane, misses are exactly -1
    eng._render(eng.frame_R[0], eng.frame_T[0], eng.scene_ids, list(range(B)), eng.frames[0])
    assert torch.equal(key0, eng.frames[0])
    hit = key0 > -1
    assert hit.float().mean() > 0.5 and (key0[hit] >= 0.5).all() and (key0[~hit] == -1).all()
    expected_len = torch.zeros(B, dtype=torch.int64)
    for t in range(n_steps):
        frames_before = eng.frames[0].clone()
        mv = eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1])
        out = eng.step(mv)
        # ---- back-projection: each of the 5 frames of a step contributes exactly int(0.05 * n_valid) points
        used = torch.cat((frames_before[None], eng.frames[1:4], frames_before[None]))          # key (stage A), interp 1-3, key again (stage E)
        n_valid = ((used > -1) & (used < 70.0)).flatten(2).sum(2)                               # (5, B)
        expected_len += (n_valid.double() * 0.05).floor().long().sum(0).cpu()
        assert torch.equal(eng.cloud_len.cpu().long(), expected_len)
        # ---- binning: the 4 slab images hold exactly the points inside the grid window and inside a slab
        grid = out.model_input
        assert (grid == grid.round()).all() and (grid >= 0).all()
        for b in range(B):
            n = int(eng.cloud_len[b]) - int((n_valid[1:, b].double() * 0.05).floor().sum())      # cloud as it was when the grid was built
            p = eng.cloud[b, :n]
            pose = torch.tensor(poses[b, t], device=DEV)
            r = torch.round((-(p[:, 2] - pose[2]) + 40.0) * (S / 80.0)); c = torch.round((-(p[:, 0] - pose[0]) + 40.0) * (S / 80.0))
            inside = (r >= 0) & (r < S) & (c >= 0) & (c < S)
            nb = int(eng.n_bounds[b]); bounds = eng.slab_bounds[b, :nb]
            slab = torch.bucketize(p[:, 1].contiguous(), bounds) - 1
            assert int(grid[b, :4].sum()) == int((inside & (slab >= 0) & (slab < 4)).sum())
            assert int(grid[b, 4].sum()) <= 1 + 4 * t and grid[b, 4].sum() >= 1
        # ---- world points lie inside the scene's bounding box; slack: the reference's un-projection tables are offset from the
        # rasteriser's pixel centres by up to a pixel, i.e. up to 2*tan(30 deg)*70/256 = 0.32 units at the sensor range
        for b in range(B):
            p = eng.cloud[b, : int(eng.cloud_len[b])]
            lo = torch.tensor(scenes[b].verts.min(0), device=DEV) - 0.5; hi = torch.tensor(scenes[b].verts.max(0), device=DEV) + 0.5
            assert ((p >= lo) & (p <= hi)).all()
    assert eng.overflow.item() == 0
    # ---- network: a scene's maps do not depend on which other scenes share the batch / chunk
    with torch.no_grad():
        o1a, o2a = net(out.model_input)
        o1b, o2b = net(out.model_input[2:5].contiguous())
    assert torch.equal(o1a[2:5], o1b) and torch.equal(o2a[2:5], o2b)
    assert torch.isfinite(o1a).all() and (o2a >= 0).all() and (o2a <= 1).all()
