"""Torch-tensor front end of the C ABI: validates tensors, passes raw device pointers + the current
CUDA stream to libnbp_b200.so.  PyTorch is plumbing here (device memory, streams), not compute."""
from __future__ import annotations

import math
import os

import torch

from . import _lib

FOV_DEG = 60.0          # FoVPerspectiveCameras default used by the reference (macarons_utils.py:2632)
Z_CLIP = 0.5            # znear / 2 with znear = 1 (PyTorch3D MeshRasterizer default for perspective cameras)


def tan_half_fov(fov_deg: float = FOV_DEG) -> float:
    """tan(fov/2) evaluated the way FoVPerspectiveCameras does (fp32 tensor ops on the host)."""
    fov = (math.pi / 180.0) * torch.tensor([fov_deg], dtype=torch.float32)
    return float(torch.tan(fov / 2).item())


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _chk(t, dtype, name, dev=None):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor (nextbestpath_b200 has no CPU path)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if dev is not None and t.device != dev:
        raise RuntimeError(f"{name} is on {t.device}, expected {dev}")
    return t


class _Workspace:
    """Grow-only per-device scratch buffers owned by the caller side (the library never allocates)."""

    def __init__(self):
        self.bufs = {}

    def get(self, key, nbytes, device):
        b = self.bufs.get((key, device))
        if b is None or b.numel() < nbytes:
            b = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
            self.bufs[(key, device)] = b
        return b


_ws = _Workspace()
RASTER_MAX_VIEW_FACES = 6_000_000      # (view, face) pairs per rasteriser call: bounds its scratch at 4.3 GB


def raster_depth(verts, faces, vert_offsets, face_offsets, view_scene, R, T, H, W, face_counts_host,
                 view_scene_host, zbuf=None, pix_to_face=None, fov_deg=FOV_DEG, z_clip=Z_CLIP, want_faces=False):
    """Batched depth render (a2).  ``face_counts_host``: python list of per-scene face counts;
    ``view_scene_host``: python list of the scene of each view (both only size the workspace)."""
    dev = verts.device
    _chk(verts, torch.float32, "verts"); _chk(faces, torch.int32, "faces", dev)
    _chk(vert_offsets, torch.int64, "vert_offsets", dev); _chk(face_offsets, torch.int64, "face_offsets", dev)
    _chk(view_scene, torch.int32, "view_scene", dev); _chk(R, torch.float32, "R", dev); _chk(T, torch.float32, "T", dev)
    n_views, n_scenes = int(view_scene.numel()), int(face_offsets.numel()) - 1
    total = int(sum(face_counts_host[s] for s in view_scene_host))
    max_f = int(max(face_counts_host)) if len(face_counts_host) else 0
    if zbuf is None:
        zbuf = torch.empty((n_views, H, W), dtype=torch.float32, device=dev)
    else:
        _chk(zbuf, torch.float32, "zbuf", dev)
    if want_faces and pix_to_face is None:
        pix_to_face = torch.empty((n_views, H, W), dtype=torch.int32, device=dev)
    L = _lib.lib()
    # the rasteriser's scratch is 720 bytes per (view, face) pair: views are independent, so large batches go through in groups that
    # keep it under RASTER_MAX_VIEW_FACES pairs (4.3 GB) instead of reserving e.g. 18 GB for 512 views of 50 k-triangle meshes
    v0 = 0
    while v0 < n_views:
        v1, pairs = v0, 0
        while v1 < n_views and (v1 == v0 or pairs + face_counts_host[view_scene_host[v1]] <= RASTER_MAX_VIEW_FACES):
            pairs += face_counts_host[view_scene_host[v1]]
            v1 += 1
        ws = _ws.get("raster", L.nbp_raster_workspace_bytes(v1 - v0, pairs), dev)
        rc = L.nbp_raster_depth_batched(_ptr(verts), _ptr(faces), _ptr(vert_offsets), _ptr(face_offsets), n_scenes,
                                        _ptr(view_scene[v0:v1]), _ptr(R[v0:v1]), _ptr(T[v0:v1]), v1 - v0, pairs, max_f, H, W,
                                        tan_half_fov(fov_deg), z_clip, _ptr(zbuf[v0:v1]),
                                        _ptr(pix_to_face[v0:v1]) if pix_to_face is not None else None,
                                        _ptr(ws), ws.numel(), _stream())
        _lib.check(rc, "nbp_raster_depth_batched")
        v0 = v1
    return (zbuf, pix_to_face) if want_faces else zbuf


def backproject_append(zbuf, R, T, frame_scene, cloud, cloud_len, *, mask=None, frame_uid=None, fov_range=70.0,
                       gathering_factor=0.05, seed=0, fov_deg=FOV_DEG, frame_valid=None, frame_kept=None,
                       overflow=None):
    """Back-project frames and append the kept points to the per-scene clouds (a4+a5).
    zbuf (n_frames,H,W); cloud (n_scenes, cap, 3) fp32; cloud_len (n_scenes,) int32 (updated in place)."""
    dev = zbuf.device
    _chk(zbuf, torch.float32, "zbuf"); _chk(R, torch.float32, "R", dev); _chk(T, torch.float32, "T", dev)
    _chk(frame_scene, torch.int32, "frame_scene", dev); _chk(cloud, torch.float32, "cloud", dev)
    _chk(cloud_len, torch.int32, "cloud_len", dev)
    if mask is not None:
        _chk(mask, torch.uint8, "mask", dev)
    if frame_uid is not None:
        _chk(frame_uid, torch.int32, "frame_uid", dev)
    n_frames, H, W = zbuf.shape[0], zbuf.shape[-2], zbuf.shape[-1]
    n_scenes, cap = cloud.shape[0], cloud.shape[1]
    L = _lib.lib()
    need = L.nbp_backproject_workspace_bytes(n_frames)
    if gathering_factor < 1.0 and os.environ.get("NBP_BP_KEY_CACHE", "1") != "0":
        need += L.nbp_backproject_key_cache_bytes(n_frames, H, W)          # every pixel's selection key is evaluated once, not three times
    ws = _ws.get("backproject", need, dev)
    rc = L.nbp_backproject_append(_ptr(zbuf), _ptr(mask), _ptr(R), _ptr(T), _ptr(frame_scene), _ptr(frame_uid),
                                  n_frames, H, W, tan_half_fov(fov_deg), float(fov_range if fov_range else 0.0),
                                  float(gathering_factor), int(seed) & 0xFFFFFFFFFFFFFFFF, _ptr(cloud), _ptr(cloud_len),
                                  cap, n_scenes, _ptr(frame_valid), _ptr(frame_kept), _ptr(overflow),
                                  _ptr(ws), ws.numel(), _stream())
    _lib.check(rc, "nbp_backproject_append")


def grid_scatter(cloud, cloud_len, pose, slab_bounds, n_bounds, S, *, traj=None, traj_len=None, n_pieces=4,
                 grid_range=(-40.0, 40.0), max_points=-1, out=None):
    """Model-input grid (a6-a10): (n_scenes, n_pieces+1, S, S) fp32 counts."""
    dev = cloud.device
    _chk(cloud, torch.float32, "cloud"); _chk(cloud_len, torch.int32, "cloud_len", dev)
    _chk(pose, torch.float32, "pose", dev); _chk(slab_bounds, torch.float32, "slab_bounds", dev)
    _chk(n_bounds, torch.int32, "n_bounds", dev)
    if traj is not None:
        _chk(traj, torch.float32, "traj", dev); _chk(traj_len, torch.int32, "traj_len", dev)
    n_scenes, cap = cloud.shape[0], cloud.shape[1]
    if out is None:
        out = torch.empty((n_scenes, n_pieces + 1, S, S), dtype=torch.float32, device=dev)
    else:
        _chk(out, torch.float32, "out", dev)
    rc = _lib.lib().nbp_grid_scatter(_ptr(cloud), _ptr(cloud_len), cap, _ptr(traj), _ptr(traj_len),
                                     traj.shape[1] if traj is not None else 0, _ptr(pose), _ptr(slab_bounds),
                                     _ptr(n_bounds), slab_bounds.shape[1], n_scenes, n_pieces, S,
                                     float(grid_range[0]), float(grid_range[1]), int(max_points), _ptr(out), _stream())
    _lib.check(rc, "nbp_grid_scatter")
    return out


def map_points(points_2d, grid_size, grid_range, lens=None):
    """map_points_to_n_imgs semantics on device: (n, m, 2) fp32 -> (n, S0, S1) fp32 counts."""
    _chk(points_2d, torch.float32, "points_2d_batch")
    n, m = points_2d.shape[0], points_2d.shape[1]
    out = torch.empty((n, int(grid_size[0]), int(grid_size[1])), dtype=torch.float32, device=points_2d.device)
    rc = _lib.lib().nbp_map_points(_ptr(points_2d), _ptr(lens), n, m, int(grid_size[0]), int(grid_size[1]),
                                   float(grid_range[0]), float(grid_range[1]), _ptr(out), _stream())
    _lib.check(rc, "nbp_map_points")
    return out


def point_cells(points_2d, grid_size, grid_range):
    """get_point_position_in_the_img semantics on device: (n, 2) fp32 -> (2, n) int64."""
    _chk(points_2d, torch.float32, "points_2d")
    n = points_2d.shape[0]
    out = torch.empty((2, n), dtype=torch.int64, device=points_2d.device)
    rc = _lib.lib().nbp_point_cells(_ptr(points_2d), n, int(grid_size[0]), int(grid_size[1]),
                                    float(grid_range[0]), float(grid_range[1]), _ptr(out), _stream())
    _lib.check(rc, "nbp_point_cells")
    return out


def launch_count() -> int:
    return int(_lib.lib().nbp_launch_count())
