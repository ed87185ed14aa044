"""Host-side mirror of the reference camera for the NBP path (macarons/utility/macarons_utils.py: get_camera_RT
:940-957, Camera.update_camera :2590-2632, capture_image :2743-2786, project_depth_in_3D :2788-2809,
compute_partial_point_cloud :2811-2847), backed by the CUDA rasteriser / back-projection kernels.

Only what the NBP drivers touch is mirrored: pose -> (R, T), depth capture of a mesh,
un-projection, the masked / range-limited / sub-sampled partial point cloud, the camera history.  The pose lattice,
collision bookkeeping and MACARONS depth-network plumbing stay in the reference (out of scope).
"""
from __future__ import annotations

import math
import os

import torch

from .. import ops


def get_camera_RT(X_cam: torch.Tensor, V_cam: torch.Tensor):
    """macarons_utils.py:940-957 + PyTorch3D look_at_view_transform(eye=X, at=X+rays), evaluated with fp32 torch ops
    on the host (a few hundred cameras per step).  Returns R (n,3,3), T (n,3) with x_view = x_world @ R + T."""
    dev = X_cam.device
    X = X_cam.detach().to("cpu", torch.float32).reshape(-1, 3)
    V = V_cam.detach().to("cpu", torch.float32).reshape(-1, 2)
    k = math.pi / 180.0
    elev, azim = -1 * V[:, 0].view(-1, 1), 180.0 + V[:, 1].view(-1, 1)
    cart = torch.stack((torch.cos(k * elev) * torch.sin(k * azim), torch.sin(k * elev),
                        torch.cos(k * elev) * torch.cos(k * azim)), dim=2)
    rays = -(torch.ones(len(V), 1) * cart.view(-1, 3))

    def unit(v):
        return v / v.norm(dim=1, keepdim=True).clamp_min(1e-5)

    z = unit((X + rays) - X)
    up = torch.tensor([[0.0, 1.0, 0.0]]).expand(len(X), 3)
    x = unit(torch.cross(up, z, dim=1))
    y = unit(torch.cross(z, x, dim=1))
    flat = torch.isclose(x, torch.tensor(0.0), atol=5e-3).all(dim=1, keepdim=True)
    if flat.any():
        x = torch.where(flat, unit(torch.cross(y, z, dim=1)), x)
    R = torch.stack((x, y, z), dim=2)
    T = -torch.bmm(R.transpose(1, 2), X[:, :, None])[:, :, 0]
    return R.contiguous().to(dev), T.contiguous().to(dev)


class FoVCamera:
    """The slice of pytorch3d FoVPerspectiveCameras the NBP path reads: .R (1,3,3), .T (1,3), get_camera_center()."""

    def __init__(self, R, T, zfar=750.0, device=None):
        self.R, self.T, self.zfar, self.device = R, T, zfar, device if device is not None else R.device

    def get_camera_center(self):
        return -torch.bmm(self.T[:, None, :], self.R.transpose(1, 2))[:, 0]


class _Mesh:
    """Minimal mesh container with the accessors the reference uses (verts_list()[0], faces_list()[0])."""

    def __init__(self, verts, faces):
        self._v, self._f = verts, faces

    def verts_list(self):
        return [self._v]

    def faces_list(self):
        return [self._f]


class Camera:
    """Depth camera with the reference's method signatures for the NBP path."""

    def __init__(self, device, image_height=256, image_width=456, zfar=750.0, gathering_factor=0.05,
                 sensor_range=70.0, n_interpolation_steps=4, save_dir_path=None):
        self.device = torch.device(device)
        self.image_height, self.image_width = image_height, image_width
        self.zfar, self.gathering_factor, self.sensor_range = zfar, gathering_factor, sensor_range
        self.n_interpolation_steps = n_interpolation_steps
        self.save_dir_path = save_dir_path
        self.n_frames_captured = 0
        self.X_cam = self.V_cam = self.fov_camera = None
        self.X_cam_history = torch.zeros(0, 3, device=self.device)
        self.V_cam_history = torch.zeros(0, 2, device=self.device)
        self._mesh_key, self._mesh_pack = None, None

    # ---- pose bookkeeping (update_camera restricted to explicit poses; the index lattice stays in the reference)
    def set_pose(self, X_cam, V_cam):
        self.X_cam = X_cam.to(self.device, torch.float32).view(1, 3)
        self.V_cam = V_cam.to(self.device, torch.float32).view(1, 2)
        self.X_cam_history = torch.vstack((self.X_cam_history, self.X_cam))
        self.V_cam_history = torch.vstack((self.V_cam_history, self.V_cam))
        R, T = get_camera_RT(self.X_cam, self.V_cam)
        self.fov_camera = FoVCamera(R, T, self.zfar, self.device)

    def get_fov_camera_from_RT(self, R_cam, T_cam):
        return FoVCamera(R_cam, T_cam, self.zfar, self.device)

    # ---- capture_image (macarons_utils.py:2743-2786)
    def _pack(self, mesh):
        v, f = mesh.verts_list()[0], mesh.faces_list()[0]
        # identity + in-place version of the caller's tensors (the key holds references: no address reuse, vertex edits are seen)
        k0 = self._mesh_key
        if k0 is None or k0[0] is not v or k0[1] is not f or k0[2] != v._version or k0[3] != f._version:
            dev = self.device
            self._mesh_pack = (v.to(dev, torch.float32).contiguous(), f.to(dev, torch.int32).contiguous(),
                               torch.tensor([0, v.shape[0]], dtype=torch.int64, device=dev),
                               torch.tensor([0, f.shape[0]], dtype=torch.int64, device=dev),
                               torch.zeros(1, dtype=torch.int32, device=dev), [int(f.shape[0])])
            self._mesh_key = (v, f, v._version, f._version)
        return self._mesh_pack

    def capture_image(self, mesh, fov_camera=None, save_frame=True, dir_path=None):
        """Returns (images (1,H,W,3), depth (1,H,W,1)).  RGB is dead data on the NBP path:
        a constant grey image is returned; depth = view-space z, -1 where nothing is hit."""
        cam = fov_camera if fov_camera is not None else self.fov_camera
        verts, faces, vo, fo, vs, fc = self._pack(mesh)
        H, W = self.image_height, self.image_width
        z = ops.raster_depth(verts, faces, vo, fo, vs, cam.R.reshape(1, 9).contiguous(), cam.T.reshape(1, 3).contiguous(),
                             H, W, fc, [0])
        depth = z.view(1, H, W, 1)
        images = torch.full((1, H, W, 3), 0.5, device=self.device)
        dir_path = dir_path if dir_path is not None else self.save_dir_path
        if save_frame and dir_path is not None:
            frame = {"rgb": images, "zbuf": depth, "mask": depth > -1, "R": cam.R, "T": cam.T, "zfar": self.zfar}
            torch.save(frame, os.path.join(dir_path, str(self.n_frames_captured) + ".pt"))
            self.n_frames_captured += 1
        return images, depth

    # ---- project_depth_in_3D / compute_partial_point_cloud (macarons_utils.py:2788-2847)
    def _backproject(self, depth, mask, cam, fov_range):
        B = depth.shape[0]
        H, W = self.image_height, self.image_width
        z = depth.reshape(B, H, W).to(torch.float32).contiguous()
        cloud = torch.empty((B, (H * W + 3) // 4 * 4, 3), dtype=torch.float32, device=self.device)
        cloud_len = torch.zeros(B, dtype=torch.int32, device=self.device)
        frame_scene = torch.arange(B, dtype=torch.int32, device=self.device)
        m = None if mask is None else mask.reshape(B, H, W).to(torch.uint8).contiguous()
        ops.backproject_append(z, cam.R.reshape(B, 9).contiguous(), cam.T.reshape(B, 3).contiguous(), frame_scene, cloud,
                               cloud_len, mask=m, fov_range=fov_range, gathering_factor=1.0)
        return cloud, cloud_len

    def project_depth_in_3D(self, depth, fov_cameras=None):
        """(B, H*W, 3): every pixel is un-projected, misses (z = -1) included, as the reference does."""
        cam = fov_cameras if fov_cameras is not None else self.fov_camera
        B, HW = depth.shape[0], self.image_height * self.image_width
        ones = torch.ones((B, HW), dtype=torch.uint8, device=self.device)
        cloud, _ = self._backproject(depth, ones, cam, None)
        return cloud[:, :HW]

    def compute_partial_point_cloud(self, depth, mask, images=None, fov_cameras=None, gathering_factor=None, fov_range=None):
        """macarons_utils.py:2811-2847: valid = mask & (depth < fov_range); keep int(n*gathering_factor) of the valid
        world points, chosen by torch.randperm(n) on the default CPU generator exactly as the reference (:2837)."""
        cam = fov_cameras if fov_cameras is not None else self.fov_camera
        cloud, cloud_len = self._backproject(depth, mask, cam, fov_range)
        n = int(cloud_len[0].item())
        world_points = cloud[0, :n]
        gf = self.gathering_factor if gathering_factor is None else gathering_factor
        idx = torch.randperm(n)[: int(n * gf)].to(self.device)
        world_points = world_points[idx]
        if images is None:
            return world_points
        valid = mask.view(-1) != 0
        if fov_range is not None:
            valid &= depth.view(-1) < fov_range
        colors = (0.0 + images.view(-1, 3))[valid][idx]
        return world_points, colors
