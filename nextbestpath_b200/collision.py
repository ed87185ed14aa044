"""Batched collision tests without Trimesh.

The reference's planner asks "does the straight move between two lattice poses cross the mesh?" one segment at a time through
``trimesh.ray.intersects_location`` on the host (``line_segment_mesh_intersection`` macarons/utility/macarons_utils.py:120-151:
Dijkstra neighbour expansion next_best_path/utility/long_term_utils.py:347, path check next_best_path/testers/nbp_planning.py:142,245)
and "is the camera inside the mesh?" through three axis rays (``check_camera_in_mesh`` long_term_utils.py:158-170).  Here any number
of segments / rays over any number of scenes go through one launch of ``nbp_segments_hit_mesh`` (csrc/collision.cu); because the
mesh is static, a scene's whole lattice-neighbour collision table can be computed once at set-up.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


GT_MAP_HALF_WIDTH_PX = 1.35      # matplotlib's default 1.5 pt line at dpi 100 after the reference's resize to 256 px (oracle/section_oracle.c)


def _stream():
    return torch.cuda.current_stream().cuda_stream


class MeshBatch:
    """Meshes of B scenes packed for the CUDA kernels: verts (sum V,3) fp32, faces (sum F,3) int32 scene-local, int64 offsets."""

    def __init__(self, verts_list, faces_list, device):
        self.dev = torch.device(device)
        vs = [torch.as_tensor(v, dtype=torch.float32).reshape(-1, 3) for v in verts_list]
        fs = [torch.as_tensor(np.asarray(f) if not isinstance(f, torch.Tensor) else f).to(torch.int32).reshape(-1, 3) for f in faces_list]
        self.verts = (torch.cat(vs) if vs else torch.zeros((0, 3))).contiguous().to(self.dev)
        self.faces = (torch.cat(fs) if fs else torch.zeros((0, 3), dtype=torch.int32)).contiguous().to(self.dev)
        self.vert_off = torch.tensor(np.concatenate([[0], np.cumsum([len(v) for v in vs])]), dtype=torch.int64, device=self.dev)
        self.face_off = torch.tensor(np.concatenate([[0], np.cumsum([len(f) for f in fs])]), dtype=torch.int64, device=self.dev)
        self.n_scenes = len(vs)
        self.max_faces = max([len(f) for f in fs] + [0])

    def _run(self, seg, seg_scene, rays, want_count):
        seg = torch.as_tensor(seg, dtype=torch.float32).reshape(-1, 6).contiguous().to(self.dev)
        n = seg.shape[0]
        if seg_scene is None:
            seg_scene = torch.zeros(n, dtype=torch.int32, device=self.dev)
        seg_scene = torch.as_tensor(seg_scene).to(device=self.dev, dtype=torch.int32).contiguous()
        hit = torch.zeros(n, dtype=torch.uint8, device=self.dev)
        count = torch.zeros(n, dtype=torch.int32, device=self.dev) if want_count else None
        if self.faces.numel() == 0:
            return hit.bool(), count
        rc = _lib.lib().nbp_segments_hit_mesh(self.verts.data_ptr(), self.faces.data_ptr(), self.vert_off.data_ptr(), self.face_off.data_ptr(),
                                              self.n_scenes, seg.data_ptr(), seg_scene.data_ptr(), n, 1 if rays else 0, hit.data_ptr(),
                                              count.data_ptr() if count is not None else None, _stream())
        _lib.check(rc, "nbp_segments_hit_mesh")
        return hit.bool(), count

    def segments_hit(self, segments, seg_scene=None, return_counts=False):
        """segments (n,6) = (start, end); seg_scene (n,) scene of each segment.  Returns hit (n,) bool on the device [, counts]."""
        hit, count = self._run(segments, seg_scene, False, return_counts)
        return (hit, count) if return_counts else hit

    def ray_hit_counts(self, origins, directions, ray_scene=None):
        """Number of triangles each forward ray (origin, unit direction) crosses: (n,) int32 on the device."""
        seg = torch.cat((torch.as_tensor(origins, dtype=torch.float32).reshape(-1, 3), torch.as_tensor(directions, dtype=torch.float32).reshape(-1, 3)), dim=1)
        return self._run(seg, ray_scene, True, True)[1]

    def gt_obstacle_maps(self, poses, map_scene=None, S: int = 256, view_size: float = 80.0, half_width_px: float = GT_MAP_HALF_WIDTH_PX):
        """Ground-truth obstacle maps: poses (n,5) camera poses, map_scene (n,) scene of each map (default:
        map b <- scene b).  Returns (n,1,S,S) fp32 of 0/1 on the device = ``current_gt_obs`` of nbp_utils.py:638-639, all maps in one
        launch (the reference: trimesh section + matplotlib + PIL per pose on the host)."""
        pose = torch.as_tensor(poses, dtype=torch.float32).reshape(-1, 5).contiguous().to(self.dev)
        n = pose.shape[0]
        if map_scene is None:
            if n != self.n_scenes:
                raise RuntimeError("gt_obstacle_maps: pass map_scene when the number of poses differs from the number of scenes")
            ms = None
        else:
            ms = torch.as_tensor(map_scene).to(device=self.dev, dtype=torch.int32).contiguous()
            assert ms.numel() == n
        out = torch.empty((n, 1, S, S), dtype=torch.float32, device=self.dev)
        rc = _lib.lib().nbp_gt_obstacle_map(self.verts.data_ptr(), self.faces.data_ptr(), self.vert_off.data_ptr(), self.face_off.data_ptr(),
                                            ms.data_ptr() if ms is not None else None, pose.data_ptr(), n, self.max_faces, S,
                                            float(view_size), float(half_width_px), out.data_ptr(), _stream())
        _lib.check(rc, "nbp_gt_obstacle_map")
        return out

    def neighbour_collision_table(self, positions, neighbours, scene: int = 0):
        """positions (P,3), neighbours (P,K) indices into positions (-1 = none): blocked (P,K) bool = the move crosses the mesh.
        One launch instead of P*K host ray casts inside Dijkstra (long_term_utils.py:334-359)."""
        pos = torch.as_tensor(positions, dtype=torch.float32).to(self.dev)
        nb = torch.as_tensor(neighbours).to(self.dev).long()
        P, K = nb.shape
        valid = nb >= 0
        seg = torch.cat((pos[:, None, :].expand(P, K, 3), pos[nb.clamp_min(0)]), dim=2).reshape(-1, 6)
        hit = self.segments_hit(seg, torch.full((P * K,), scene, dtype=torch.int32, device=self.dev)).view(P, K)
        return hit & valid


def line_segment_mesh_intersection(start_point, end_point, mesh: MeshBatch):
    """Drop-in for macarons_utils.py:120-151 with ``mesh`` a one-scene MeshBatch instead of a trimesh object: Python bool."""
    to_t = lambda p: torch.as_tensor(p, dtype=torch.float32).reshape(3).cpu()
    seg = torch.cat((to_t(start_point), to_t(end_point))).view(1, 6)
    return bool(mesh.segments_hit(seg)[0].item())


def check_camera_in_mesh(mesh_for_check: MeshBatch, camera_position):
    """Drop-in for long_term_utils.py:158-170: odd hit counts along +y, +x and +z."""
    o = torch.as_tensor(camera_position, dtype=torch.float32).reshape(1, 3).cpu().expand(3, 3)
    d = torch.tensor([[0.0, 1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    c = mesh_for_check.ray_hit_counts(o, d).cpu()
    return bool((c % 2 == 1).all())


def get_binary_obstacle_array(mesh: MeshBatch, camera_pose, view_size=80):
    """Drop-in for next_best_path/utility/utils.py:226-262 with ``mesh`` a one-scene MeshBatch: (256, 256) int numpy array of 0/1
    (the reference returns None when the plane misses the mesh; here the map is then all zeros)."""
    m = mesh.gt_obstacle_maps(torch.as_tensor(camera_pose, dtype=torch.float32).reshape(1, 5), map_scene=[0], view_size=float(view_size))
    return m[0, 0].cpu().numpy().astype(int)
