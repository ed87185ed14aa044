"""GPU: the reference-signature shims -- next_best_path/utility/utils.py map builder functions and the
Camera methods of macarons/utility/macarons_utils.py:2743-2847 -- called exactly as nbp_planning.py calls them, against the
reference-generated fixture and the oracle."""
import os

import numpy as np
import pytest
import torch

from nextbestpath_b200 import synthetic as syn
from nextbestpath_b200.utility import camera as C
from nextbestpath_b200.utility import utils as U
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "mapbuilder.npz"))


def _dense(idx, val, shape):
    a = np.zeros(int(np.prod(shape)), dtype=val.dtype)
    a[idx] = val
    return a.reshape(shape)


def test_map_builder_functions_as_the_driver_calls_them():
    """nbp_planning.py:114-127: bucketize, then per slab transform_points_to_n_pieces + map_points_to_n_imgs."""
    pts = torch.from_numpy(GOLD["points"]).to(DEV)
    pose = torch.from_numpy(GOLD["pose"]).to(DEV)
    p2d = U.transform_points_to_n_pieces(pts, pose, DEV)
    assert p2d.shape == (1, len(pts), 2) and np.array_equal(p2d.cpu().numpy(), GOLD["p2d"])
    y_bins = torch.from_numpy(GOLD["y_bins"]).to(DEV)
    bin_indices = torch.bucketize(pts[:, 1].contiguous(), y_bins[:-1]) - 1                      # the reference expression (:114)
    assert np.array_equal(bin_indices.cpu().numpy(), GOLD["bins"])
    for S in (128, 256, 512):
        want = _dense(GOLD[f"grid{S}_idx"], GOLD[f"grid{S}_val"], (4, S, S))
        for k in range(4):
            sel = pts[bin_indices == k]
            if len(sel) == 0:
                continue
            img = U.map_points_to_n_imgs(U.transform_points_to_n_pieces(sel, pose, DEV), (S, S), (-40, 40), DEV)
            assert img.shape == (1, S, S) and img.dtype == torch.float32
            assert np.array_equal(img[0].cpu().numpy(), want[k])
            img += 1.0                                                               # callers mutate the result in place (:175)
        cells = U.get_point_position_in_the_img(p2d, (S, S), (-40, 40))
        assert cells.dtype == torch.int64 and np.array_equal(cells.cpu().numpy(), GOLD[f"cells{S}"])
    one = U.get_point_position_in_the_img(p2d[:, :1], (64, 64), (-40, 40))           # single candidate (:203): shape (2,)
    assert one.shape == (2,) and np.array_equal(one.cpu().numpy(), GOLD["cells64"][:, 0])
    with pytest.raises(RuntimeError):
        U.map_points_to_n_imgs(p2d.cpu(), (64, 64), (-40, 40), "cpu")                 # no CPU path


def test_camera_shim_capture_and_partial_point_cloud(tmp_path):
    H, W = 64, 114
    scene = syn.make_scene(21, tri_budget=900)
    poses, _ = syn.random_walk(scene, 2, seed=21)
    mesh = C._Mesh(torch.from_numpy(scene.verts).to(DEV), torch.from_numpy(scene.faces.astype(np.int64)).to(DEV))
    cam = C.Camera(DEV, image_height=H, image_width=W, gathering_factor=0.05, sensor_range=30.0, save_dir_path=str(tmp_path))
    cam.set_pose(torch.tensor(poses[0, :3]), torch.tensor(poses[0, 3:]))
    images, depth = cam.capture_image(mesh)
    assert images.shape == (1, H, W, 3) and depth.shape == (1, H, W, 1) and cam.n_frames_captured == 1
    frame = torch.load(os.path.join(str(tmp_path), "0.pt"))
    assert set(frame) == {"rgb", "zbuf", "mask", "R", "T", "zfar"} and torch.equal(frame["mask"], frame["zbuf"] > -1)
    R, T = O.camera_rt(torch.tensor(poses[:1, :3]), torch.tensor(poses[:1, 3:]))
    zo = O.render_depth(scene.verts, scene.faces, R[0].numpy(), T[0].numpy(), H, W)[0]
    assert np.array_equal(depth[0, :, :, 0].cpu().numpy(), zo)
    assert torch.equal(cam.fov_camera.R.cpu(), R) and torch.allclose(cam.fov_camera.get_camera_center().cpu(), torch.tensor(poses[:1, :3]), atol=1e-4)
    # compute_partial_point_cloud: the reference's own randperm stream (macarons_utils.py:2836-2838)
    mask = depth > -1
    torch.manual_seed(99)
    pc = cam.compute_partial_point_cloud(depth, mask, fov_range=30.0)
    n = int(((zo > -1) & (zo < 30.0)).sum())
    torch.manual_seed(99)
    idx = torch.randperm(n)[: int(n * 0.05)].numpy()
    want = O.partial_point_cloud(zo, R[0].numpy(), T[0].numpy(), 30.0, 0.05, indices=idx)
    assert pc.shape == (int(n * 0.05), 3) and np.array_equal(pc.cpu().numpy(), want)
    pc2, col = cam.compute_partial_point_cloud(depth, mask, images=images, fov_range=30.0, gathering_factor=1.0)
    assert pc2.shape == (n, 3) and col.shape == (n, 3)
    # project_depth_in_3D un-projects every pixel (misses included), row-major
    allp = cam.project_depth_in_3D(depth)
    assert allp.shape == (1, H * W, 3)
    full = O.unproject(zo, R[0].numpy(), T[0].numpy()).reshape(-1, 3)
    hit = (zo > -1).reshape(-1)
    assert np.array_equal(allp[0].cpu().numpy()[hit], full[hit])
    # second pose: history grows, second frame file appears
    cam.set_pose(torch.tensor(poses[1, :3]), torch.tensor(poses[1, 3:]))
    cam.capture_image(mesh)
    assert cam.X_cam_history.shape == (2, 3) and os.path.exists(os.path.join(str(tmp_path), "1.pt"))
