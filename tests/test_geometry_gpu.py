"""GPU parity: raster / back-project / grid-scatter CUDA kernels (through the C ABI) against the CPU oracle
and the reference-generated fixtures.  Integer / index results must be bit-exact; the geometry kernels are
pinned op-for-op to the oracle, so fp32 results are compared bit-for-bit too."""
import os

import numpy as np
import pytest
import torch

from nextbestpath_b200 import ops, synthetic as syn
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _pack_scenes(scenes):
    verts = torch.from_numpy(np.concatenate([s.verts for s in scenes])).to(DEV)
    faces = torch.from_numpy(np.concatenate([s.faces for s in scenes]).astype(np.int32)).to(DEV)
    vo = torch.tensor(np.concatenate([[0], np.cumsum([len(s.verts) for s in scenes])]), dtype=torch.int64, device=DEV)
    fo = torch.tensor(np.concatenate([[0], np.cumsum([len(s.faces) for s in scenes])]), dtype=torch.int64, device=DEV)
    return verts, faces, vo, fo, [len(s.faces) for s in scenes]


def _cams(poses):
    p = torch.as_tensor(poses, dtype=torch.float32)
    return O.camera_rt(p[:, :3], p[:, 3:])


def _render_gpu(scenes, view_scene, R, T, H, W, want_faces=True):
    verts, faces, vo, fo, fc = _pack_scenes(scenes)
    vs = torch.tensor(view_scene, dtype=torch.int32, device=DEV)
    return ops.raster_depth(verts, faces, vo, fo, vs, R.reshape(-1, 9).contiguous().to(DEV), T.contiguous().to(DEV), H, W,
                            fc, view_scene, want_faces=want_faces)


def _dense(idx, val, shape):
    a = np.zeros(int(np.prod(shape)), dtype=val.dtype)
    a[idx] = val
    return a.reshape(shape)


# ------------------------------------------------------------------------------------------- raster
@pytest.mark.parametrize("H,W", [(256, 456), (64, 114), (40, 40)])
def test_raster_bit_exact_vs_oracle(H, W):
    scenes = [syn.make_scene(1, "simple"), syn.make_scene(2, tri_budget=900), syn.make_scene(3, tri_budget=7000)]
    view_scene, poses = [], []
    for si, sc in enumerate(scenes):
        p, _ = syn.random_walk(sc, 3, seed=10 + si)
        p[1, 3] = 18.0           # a tilted view: floor triangles cross the clip plane under the camera
        p[2, 4] += 22.5          # an interpolated heading
        for q in p:
            view_scene.append(si); poses.append(q)
    R, T = _cams(np.stack(poses))
    z, p2f = _render_gpu(scenes, view_scene, R, T, H, W)
    torch.cuda.synchronize()
    z, p2f = z.cpu().numpy(), p2f.cpu().numpy()
    hit = 0
    for v, si in enumerate(view_scene):
        zo, fo_ = O.render_depth(scenes[si].verts, scenes[si].faces, R[v].numpy(), T[v].numpy(), H, W, nthreads=8)
        assert np.array_equal(z[v], zo), f"view {v}: {(z[v] != zo).sum()} zbuf pixels differ"
        assert np.array_equal(p2f[v], fo_), f"view {v}: pix_to_face differs"
        hit += (zo > -1).sum()
    assert hit > 0.5 * len(view_scene) * H * W          # the views actually see geometry


def test_raster_view_groups_bound_the_workspace(monkeypatch):
    """ops.raster_depth sends large batches through in groups of views (RASTER_MAX_VIEW_FACES): same frames, bounded scratch."""
    scenes = [syn.make_scene(4, tri_budget=900), syn.make_scene(5, tri_budget=2500)]
    view_scene, poses = [], []
    for si, sc in enumerate(scenes):
        p, _ = syn.random_walk(sc, 4, seed=20 + si)
        for q in p:
            view_scene.append(si); poses.append(q)
    R, T = _cams(np.stack(poses))
    z0, f0 = _render_gpu(scenes, view_scene, R, T, 64, 114)
    monkeypatch.setattr(ops, "RASTER_MAX_VIEW_FACES", 3000)            # 1-3 views per call
    ops._ws.bufs.pop(("raster", z0.device), None)
    z1, f1 = _render_gpu(scenes, view_scene, R, T, 64, 114)
    assert torch.equal(z0, z1) and torch.equal(f0, f1)
    assert ops._ws.bufs[("raster", z0.device)].numel() <= ops._lib.lib().nbp_raster_workspace_bytes(3, 3000 + 2600)


def test_raster_edge_cases():
    R, T = torch.eye(3)[None], torch.zeros(1, 3)
    # empty mesh: every pixel is a miss
    sc = syn.Scene(np.zeros((3, 3), np.float32), np.zeros((0, 3), np.int32), np.zeros((1, 1), bool), 0, 1, np.zeros(2))
    z = _render_gpu([sc], [0], R, T, 32, 57, want_faces=False)
    assert (z == -1).all()
    # everything behind the camera / degenerate / crossing the clip plane: identical to the oracle
    v = np.array([[-50, -50, -5], [50, -50, -5], [50, 50, -5], [0, 0, 3], [0, 0, 3], [0, 0, 3],
                  [-1, -0.2, -2], [1, -0.2, -2], [0, -0.2, 30], [0.3, 0.1, 0.0], [2, 1, 4], [-2, 1.5, 6]], dtype=np.float32)
    f = np.array([[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10, 11]], dtype=np.int32)
    sc = syn.Scene(v, f, np.zeros((1, 1), bool), 0, 1, np.zeros(2))
    z, p2f = _render_gpu([sc], [0], R, T, 64, 114)
    zo, fo_ = O.render_depth(v, f, R[0].numpy(), T[0].numpy(), 64, 114)
    assert np.array_equal(z[0].cpu().numpy(), zo) and np.array_equal(p2f[0].cpu().numpy(), fo_)
    assert (zo > -1).any()


# ------------------------------------------------------------------------------------------- back-projection
def _frames(n_views=4, H=64, W=114, seed=5):
    sc = syn.make_scene(seed, "simple")
    p, _ = syn.random_walk(sc, n_views, seed=seed)
    R, T = _cams(p)
    z = _render_gpu([sc], [0] * n_views, R, T, H, W, want_faces=False)
    return sc, R, T, z


def test_backproject_all_points_bit_exact():
    sc, R, T, z = _frames()
    n_frames = z.shape[0]
    frame_scene = torch.tensor([0, 1, 0, 1], dtype=torch.int32, device=DEV)
    cap = 4 * z.shape[1] * z.shape[2]
    cloud = torch.zeros((2, cap, 3), device=DEV)
    cloud_len = torch.zeros(2, dtype=torch.int32, device=DEV)
    fv = torch.zeros(n_frames, dtype=torch.int32, device=DEV); fk = torch.zeros_like(fv)
    ops.backproject_append(z, R.reshape(-1, 9).contiguous().to(DEV), T.to(DEV), frame_scene, cloud, cloud_len,
                           fov_range=30.0, gathering_factor=1.0, frame_valid=fv, frame_kept=fk)
    torch.cuda.synchronize()
    zc = z.cpu().numpy()
    ref = [O.partial_point_cloud(zc[f], R[f].numpy(), T[f].numpy(), 30.0, 1.0) for f in range(n_frames)]
    assert fv.tolist() == [len(r) for r in ref] and fk.tolist() == fv.tolist()
    for s in range(2):
        want = np.concatenate([ref[f] for f in range(n_frames) if frame_scene[f].item() == s])     # frame order
        assert cloud_len[s].item() == len(want)
        assert np.array_equal(cloud[s, : len(want)].cpu().numpy(), want)
    assert min(len(r) for r in ref) > 100 and any(len(r) < zc[0].size for r in ref)      # the range filter bites


def test_backproject_subsample_is_a_k_subset():
    sc, R, T, z = _frames(n_views=6, H=96, W=171, seed=7)
    n_frames = z.shape[0]
    frame_scene = torch.zeros(n_frames, dtype=torch.int32, device=DEV)
    uid = torch.arange(n_frames, dtype=torch.int32, device=DEV)
    cap = 20000
    cloud = torch.zeros((1, cap, 3), device=DEV); cloud_len = torch.zeros(1, dtype=torch.int32, device=DEV)
    fv = torch.zeros(n_frames, dtype=torch.int32, device=DEV); fk = torch.zeros_like(fv)
    ops.backproject_append(z, R.reshape(-1, 9).contiguous().to(DEV), T.to(DEV), frame_scene, cloud, cloud_len,
                           frame_uid=uid, fov_range=70.0, gathering_factor=0.05, seed=1234, frame_valid=fv, frame_kept=fk)
    torch.cuda.synchronize()
    zc = z.cpu().numpy()
    got = cloud[0, : cloud_len[0].item()].cpu().numpy()
    off = 0
    rows = []
    for f in range(n_frames):
        full = O.partial_point_cloud(zc[f], R[f].numpy(), T[f].numpy(), 70.0, 1.0)
        k = int(len(full) * 0.05)                                    # macarons_utils.py:2836
        assert fv[f].item() == len(full) and fk[f].item() == k
        part = got[off: off + k]; off += k
        # every kept point is one of the frame's valid points (bit-exact), none twice, in row-major order
        key = {tuple(r): i for i, r in enumerate(map(tuple, full.view(np.uint32).tolist()))}
        idx = [key[tuple(r)] for r in part.view(np.uint32).tolist()]
        assert len(set(idx)) == k and idx == sorted(idx)
        rows.append(np.array(idx) / max(len(full), 1))
    assert off == len(got)
    # the subset is spread over the frame, not a prefix: mean rank ~ 0.5, and frames differ
    allr = np.concatenate(rows)
    assert 0.45 < allr.mean() < 0.55 and allr.max() > 0.95 and allr.min() < 0.05
    # same seed/uid -> same subset (counter-based), different seed -> different subset
    cloud2 = torch.zeros_like(cloud); len2 = torch.zeros_like(cloud_len)
    ops.backproject_append(z, R.reshape(-1, 9).contiguous().to(DEV), T.to(DEV), frame_scene, cloud2, len2,
                           frame_uid=uid, fov_range=70.0, gathering_factor=0.05, seed=1234)
    assert torch.equal(cloud, cloud2)
    cloud3 = torch.zeros_like(cloud); len3 = torch.zeros_like(cloud_len)
    ops.backproject_append(z, R.reshape(-1, 9).contiguous().to(DEV), T.to(DEV), frame_scene, cloud3, len3,
                           frame_uid=uid, fov_range=70.0, gathering_factor=0.05, seed=99)
    assert len3.item() == cloud_len.item() and not torch.equal(cloud, cloud3)


def test_backproject_overflow_and_empty_frame():
    sc, R, T, z = _frames(n_views=2)
    z[1].fill_(-1.0)                                                 # a frame that saw nothing
    frame_scene = torch.zeros(2, dtype=torch.int32, device=DEV)
    cloud = torch.zeros((1, 1000, 3), device=DEV); cloud_len = torch.zeros(1, dtype=torch.int32, device=DEV)
    ovf = torch.zeros(1, dtype=torch.int32, device=DEV)
    fv = torch.zeros(2, dtype=torch.int32, device=DEV)
    ops.backproject_append(z, R.reshape(-1, 9).contiguous().to(DEV), T.to(DEV), frame_scene, cloud, cloud_len,
                           fov_range=None, gathering_factor=1.0, frame_valid=fv, overflow=ovf)
    assert cloud_len.item() == 1000 and fv[1].item() == 0 and ovf.item() == fv[0].item() - 1000


# ------------------------------------------------------------------------------------------- grid scatter
def _scatter_inputs(n_scenes, n_pts, seed, ragged=True):
    g = np.random.default_rng(seed)
    cap = ((max(n_pts, 4) + 3) // 4) * 4
    cloud = np.zeros((n_scenes, cap, 3), np.float32)
    lens = np.zeros(n_scenes, np.int32)
    pose = np.zeros((n_scenes, 5), np.float32)
    bounds = np.zeros((n_scenes, 6), np.float32); nb = np.zeros(n_scenes, np.int32)
    traj = np.zeros((n_scenes, 32, 3), np.float32); tl = np.zeros(n_scenes, np.int32)
    ybs = []
    for s in range(n_scenes):
        n = n_pts if not ragged else int(g.integers(0, n_pts + 1)) if s else n_pts
        lens[s] = n
        lo, hi = g.uniform(-3, 1), g.uniform(7, 12)
        # wall-like clustering: half the points on a few vertical lines
        pts = np.stack([g.uniform(-60, 60, n), g.uniform(lo - 0.5, hi + 0.5, n), g.uniform(-60, 60, n)], 1).astype(np.float32)
        pts[: n // 2, 0] = np.round(pts[: n // 2, 0] / 12) * 12
        cloud[s, :n] = pts
        pose[s] = (g.uniform(-20, 20), 1.0, g.uniform(-20, 20), 0.0, 45.0 * g.integers(8))
        yb = O.y_bins_from_verts(torch.tensor([[0, lo, 0], [0, hi, 0]], dtype=torch.float32)).numpy()
        ybs.append(yb)
        b = yb[:-1]
        bounds[s, : len(b)] = b; nb[s] = len(b)
        tl[s] = g.integers(1, 33)
        traj[s, : tl[s]] = pose[s, :3] + g.uniform(-45, 45, (tl[s], 3)).astype(np.float32)
    return cloud, lens, pose, bounds, nb, traj, tl, ybs


@pytest.mark.parametrize("S", [128, 256, 512])
def test_grid_scatter_bit_exact_vs_oracle(S):
    cloud, lens, pose, bounds, nb, traj, tl, ybs = _scatter_inputs(5, 30000, seed=S)
    t = lambda a: torch.from_numpy(a).to(DEV)
    out = ops.grid_scatter(t(cloud), t(lens), t(pose), t(bounds), t(nb), S, traj=t(traj), traj_len=t(tl))
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    six = 0
    for s in range(cloud.shape[0]):
        want = O.build_model_input(cloud[s, : lens[s]], pose[s], ybs[s][:-1], traj[s, : tl[s]], S)
        assert np.array_equal(out[s], want), f"scene {s}: {(out[s] != want).sum()} cells differ"
        six += len(ybs[s]) == 6
    assert out[:, :4].sum() > 10000 and out[:, 4].sum() > 5


@pytest.mark.parametrize("S,n_pts", [(256, 150000), (512, 150000), (128, 40000)])
def test_grid_scatter_large_clouds_take_the_hash_path_bit_exact(S, n_pts):
    """Clouds of >= 32768 points per scene are counted per 16384-point chunk in a shared-memory hash table before they reach the grid
    (grid_scatter_hash).  Ragged lengths (scene 0 full, the others anywhere from empty), clustered and uniform points: at S = 512 a
    chunk holds more distinct cells than the table has slots, so the crowded-table fallback runs too.  Same bits as the numpy oracle."""
    cloud, lens, pose, bounds, nb, traj, tl, ybs = _scatter_inputs(4, n_pts, seed=S + 1)
    t = lambda a: torch.from_numpy(a).to(DEV)
    out = ops.grid_scatter(t(cloud), t(lens), t(pose), t(bounds), t(nb), S, traj=t(traj), traj_len=t(tl))
    torch.cuda.synchronize()
    out = out.cpu().numpy()
    for s in range(cloud.shape[0]):
        want = O.build_model_input(cloud[s, : lens[s]], pose[s], ybs[s][:-1], traj[s, : tl[s]], S)
        assert np.array_equal(out[s], want), f"scene {s}: {(out[s] != want).sum()} cells differ"
    assert out[:, :4].sum() > n_pts // 4


def test_grid_scatter_matches_reference_fixture(golden_dir):
    """The fixture was produced by the reference's own bucketize + transform_points_to_n_pieces +
    map_points_to_n_imgs (tests/golden/make_golden.py)."""
    g = np.load(os.path.join(golden_dir, "mapbuilder.npz"))
    pts = g["points"]; n = len(pts); cap = (n + 3) // 4 * 4
    cloud = torch.zeros((1, cap, 3), device=DEV); cloud[0, :n] = torch.from_numpy(pts).to(DEV)
    lens = torch.tensor([n], dtype=torch.int32, device=DEV)
    b = g["y_bins"][:-1]
    bounds = torch.zeros((1, 6), device=DEV); bounds[0, : len(b)] = torch.from_numpy(b).to(DEV)
    nb = torch.tensor([len(b)], dtype=torch.int32, device=DEV)
    pose = torch.from_numpy(g["pose"]).view(1, 5).to(DEV)
    for S in (128, 256, 512):
        out = ops.grid_scatter(cloud, lens, pose, bounds, nb, S)
        ref = _dense(g[f"grid{S}_idx"], g[f"grid{S}_val"], (4, S, S))
        assert np.array_equal(out[0, :4].cpu().numpy(), ref)
        assert (out[0, 4] == 0).all()
        p2d = torch.from_numpy(g["p2d"]).to(DEV)
        cells = ops.point_cells(p2d[0].contiguous(), (S, S), (-40, 40))
        assert np.array_equal(cells.cpu().numpy(), g[f"cells{S}"])
        # plain map_points_to_n_imgs on the whole cloud == sum over slabs + dropped points
        img = ops.map_points(p2d.contiguous(), (S, S), (-40, 40))
        want = O.map_points(g["p2d"][0], S)
        assert np.array_equal(img[0].cpu().numpy(), want)


def test_grid_scatter_empty_and_six_boundaries():
    # empty clouds give zero grids; a scene whose arange has 6 elements drops the points above b4
    cloud = torch.zeros((2, 8, 3), device=DEV)
    cloud[1, :3] = torch.tensor([[0.0, 0.6, 0.0], [0.0, 9.9, 0.0], [0.0, 10.6, 0.0]], device=DEV)
    lens = torch.tensor([0, 3], dtype=torch.int32, device=DEV)
    # a 6-element y_bins as torch.arange produces for ~10 % of scenes; built explicitly here because
    # which (min_y, max_y) pairs trigger it depends on the host's float rounding
    yb = np.array([0.5, 2.875, 5.25, 7.625, 10.0, 12.375], dtype=np.float32)
    cloud[1, 1, 1] = float(yb[4]) - 0.01        # in slab 3
    cloud[1, 2, 1] = float(yb[4]) + 0.01        # bin 4 -> dropped by the reference
    bounds = torch.zeros((2, 6), device=DEV); bounds[:, :5] = torch.from_numpy(yb[:-1]).to(DEV)
    nb = torch.tensor([5, 5], dtype=torch.int32, device=DEV)
    pose = torch.zeros((2, 5), device=DEV)
    out = ops.grid_scatter(cloud, lens, pose, bounds, nb, 128).cpu().numpy()
    assert out[0].sum() == 0
    want = O.build_model_input(cloud[1, :3].cpu().numpy(), np.zeros(5, np.float32), yb[:-1], np.zeros((0, 3), np.float32), 128)
    assert np.array_equal(out[1], want) and out[1].sum() == 2
