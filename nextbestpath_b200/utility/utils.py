"""Drop-in map-builder functions with the reference's signatures
(next_best_path/utility/utils.py:160-223), backed by the CUDA kernels in libnbp_b200.so.

Call sites in the reference: next_best_path/testers/nbp_planning.py:120-121,130-131,172-181 and
next_best_path/utility/nbp_utils.py:628-629,642-643.  Inputs must be CUDA tensors: there is no CPU path.
"""
from __future__ import annotations

import torch

from .. import ops


def get_point_position_in_the_img(points_2d, grid_size, grid_range):
    """utils.py:160-164: rint((p - lo) * S/(hi-lo)) per coordinate -> LongTensor, stacked then squeezed."""
    pts = points_2d.reshape(-1, 2).to(torch.float32).contiguous()
    cells = ops.point_cells(pts, grid_size, grid_range)                  # (2, n)
    return cells.reshape((2,) + tuple(points_2d.shape[:-1])).squeeze()


def transform_points_to_n_pieces(points, camera_pose, device=None, no_rotation=True):
    """utils.py:166-196.  Every call site of the reference uses no_rotation=True, for which the result is
    (-(z - c_z), -(x - c_x)) exactly (the rotation is the identity); shape (1, N, 2)."""
    if not no_rotation:
        raise NotImplementedError("the reference only ever calls this with no_rotation=True")
    cx, _, cz = camera_pose[0], camera_pose[1], camera_pose[2]
    p = points.to(torch.float32)
    out = torch.stack((-(p[:, 2] - cz.to(p.device)), -(p[:, 0] - cx.to(p.device))), dim=1)
    return out.unsqueeze(0)


def map_points_to_n_imgs(points_2d_batch, grid_size, grid_range, device=None):
    """utils.py:198-223: (n, m, 2) points -> (n, *grid_size) fp32 count images (a fresh tensor)."""
    return ops.map_points(points_2d_batch.to(torch.float32).contiguous(), grid_size, grid_range)
