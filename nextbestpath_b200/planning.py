"""Batched re-plan read-out: what the reference does right after ``nbp(...)`` in the re-plan
branch of ``compute_nbp_trajectory`` (next_best_path/testers/nbp_planning.py:166-233) -- obstacle-map fusion and the scoring of
every lattice position -- for all scenes at once, with no host synchronisation inside.  The Dijkstra search and the Trimesh
collision ray that consume these outputs stay on the host (out of scope)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib, ops


def _stream():
    return torch.cuda.current_stream().cuda_stream


def fuse_obstacle_maps(pred_obstacle, cloud, cloud_len, pose, pose_host, traj, traj_len, S=256, grid_range=(-40.0, 40.0),
                       threshold=0.13, max_points=-1):
    """pred_obstacle (B,1,S,S) = NBP's second output.  Returns (fused (B,1,S,S) in {0,1}, full_proj (B,S,S) in {0,1}).
    nbp_planning.py:168-190: two extra histograms of the cloud (all points; the +-0.1 height slice around the camera) through
    the grid-scatter kernel, then one fusion kernel."""
    dev = pred_obstacle.device
    B = pred_obstacle.shape[0]
    ninf = torch.full((B, 1), -float("inf"), dtype=torch.float32, device=dev)
    ones = torch.ones(B, dtype=torch.int32, device=dev)
    allpts = ops.grid_scatter(cloud, cloud_len, pose, ninf, ones, S, traj=traj, traj_len=traj_len, n_pieces=1, grid_range=grid_range,
                              max_points=max_points)                                       # [:,0] all points, [:,1] trajectory
    py = np.asarray(pose_host, dtype=np.float32)[:, 1].astype(np.float64)
    lo = (py - 0.1).astype(np.float32)                                                     # python-float thresholds, compared in fp32
    hi = np.nextafter((py + 0.1).astype(np.float32), np.float32(-np.inf))                  # y < hi  <=>  y <= nextafter(hi, -inf)
    bounds = torch.from_numpy(np.stack((lo, hi), axis=1)).to(dev)
    two = torch.full((B,), 2, dtype=torch.int32, device=dev)
    sl = ops.grid_scatter(cloud, cloud_len, pose, bounds, two, S, n_pieces=1, grid_range=grid_range, max_points=max_points)
    fused = torch.empty((B, 1, S, S), dtype=torch.float32, device=dev)
    full_proj = torch.empty((B, S, S), dtype=torch.float32, device=dev)
    pred = pred_obstacle.contiguous()
    _lib.check(_lib.lib().nbp_obstacle_fuse(pred.data_ptr(), allpts.data_ptr(), 2 * S * S, sl.data_ptr(), 2 * S * S,
                                            allpts[:, 1].data_ptr(), 2 * S * S, B, S, float(np.float32(threshold)),
                                            fused.data_ptr(), full_proj.data_ptr(), _stream()), "nbp_obstacle_fuse")
    return fused, full_proj


def score_candidates(value_map, full_proj, candidates, n_candidates, pose, skip=None, grid_range=(-40.0, 40.0), window=10):
    """value_map (B,8,Sv,Sv), full_proj (B,S,S), candidates (B,M,3) world positions (first n_candidates[b] used), skip (B,M)
    uint8 for known collisions.  Returns dict(valid (B,M) bool, cell (B,M,2) int32, value (B,M), density (B,M), score (B,M)
    float64 = value - 10*density with -inf where invalid), nbp_planning.py:193-231."""
    dev = value_map.device
    B, M = candidates.shape[0], candidates.shape[1]
    value = torch.empty((B, M), dtype=torch.float32, device=dev)
    dens = torch.empty((B, M), dtype=torch.float32, device=dev)
    cell = torch.empty((B, M, 2), dtype=torch.int32, device=dev)
    valid = torch.empty((B, M), dtype=torch.uint8, device=dev)
    vm, fp, cd = value_map.contiguous(), full_proj.contiguous(), candidates.contiguous().float()
    _lib.check(_lib.lib().nbp_candidate_scores(cd.data_ptr(), n_candidates.data_ptr(), M, skip.data_ptr() if skip is not None else None,
                                               pose.data_ptr(), vm.data_ptr(), vm.shape[1], vm.shape[2], fp.data_ptr(), fp.shape[1], B,
                                               float(grid_range[0]), float(grid_range[1]), int(window), value.data_ptr(), dens.data_ptr(),
                                               cell.data_ptr(), valid.data_ptr(), _stream()), "nbp_candidate_scores")
    score = torch.where(valid.bool(), value.double() - 10.0 * dens.double(), torch.full_like(value, -float("inf"), dtype=torch.float64))
    return {"valid": valid.bool(), "cell": cell, "value": value, "density": dens, "score": score}
