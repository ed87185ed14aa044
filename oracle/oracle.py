"""oracle/oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the geometric half of the NextBestPath exploration hot path
(SURVEY.md section 8 rows a1-a10).  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import this module;
the product package ``nextbestpath_b200`` never does.

Every function cites the reference lines it restates (paths relative to
/root/reference).  Two kinds of pin exist:

* rows a6-a10 (slab split, egocentric transform, grid histogram, model-input
  assembly) restate *reference-owned* torch code; ``tests/golden/make_golden.py``
  ran the reference's own functions in the build container and committed the
  outputs, and ``tests/test_oracle_golden.py`` checks this module against them
  bit-for-bit.  PINNED.
* rows a1-a5 (camera R/T, rasterisation, un-projection) live in PyTorch3D 0.7.4
  (environment.yml:204), which is neither vendored in the reference nor
  installable here, and the reference has no test touching them.  PARITY
  UNPINNED: this module restates the published algorithm and fixes one fp32
  evaluation order, against which the CUDA kernels are compared bit-for-bit.

All arithmetic is fp32 with one rounding per operation (numpy float32 arrays),
which is what ``nvcc`` produces with ``__fmul_rn/__fadd_rn`` (no FMA contraction).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32 = np.float32


# --------------------------------------------------------------------------- C part
def build(verbose: bool = False) -> str:
    """Compile oracle/raster_oracle.c -> oracle/_build/liboracle.so (idempotent)."""
    out = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("raster_oracle.c", "section_oracle.c")]
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(s) for s in srcs):
        r = subprocess.run(["make", "-C", _HERE], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stdout + r.stderr)
        if verbose:
            print(r.stdout)
    return out


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int32)
        lp = ctypes.POINTER(ctypes.c_int64)
        lib.nbp_oracle_render_depth.restype = ctypes.c_int
        lib.nbp_oracle_render_depth.argtypes = [fp, ctypes.c_int64, lp, ctypes.c_int64, fp, fp,
                                                ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                fp, ip, ctypes.c_int]
        lib.nbp_oracle_plane_section_map.restype = ctypes.c_int
        lib.nbp_oracle_plane_section_map.argtypes = [fp, lp, ctypes.c_int64, fp, ctypes.c_int, ctypes.c_float, ctypes.c_float, fp, fp]
        _LIB = lib
    return _LIB


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


# --------------------------------------------------------------------------- a1 camera
def tan_half_fov(fov_deg: float = 60.0) -> np.float32:
    """tan(fov/2) as FoVPerspectiveCameras.compute_projection_matrix evaluates it
    (degrees -> radians as python-float * fp32 tensor, then torch.tan in fp32).
    Reference call site: macarons_utils.py:2632 (FoVPerspectiveCameras(R, T, zfar) -> fov 60, znear 1)."""
    fov = (np.pi / 180.0) * torch.tensor([fov_deg], dtype=torch.float32)
    return f32(torch.tan(fov / 2).item())


def focal_ndc(fov_deg: float = 60.0) -> np.float32:
    """K[0,0] = K[1,1] = 2*znear/(max_x-min_x) with znear 1, aspect 1 = 1/tan(fov/2) in fp32."""
    t = tan_half_fov(fov_deg)
    return f32(f32(2.0) / f32(t + t))


def view_direction(V: torch.Tensor) -> torch.Tensor:
    """Unit view rays for (elev, azim) in degrees.
    Restates get_camera_RT macarons_utils.py:948-951 + get_cartesian_coords CustomGeometry.py:5-24:
    rays = -cart(r=1, elev=-e, azim=180+a)."""
    V = V.to(torch.float32)
    factor = np.pi / 180.0
    elev = -1 * V[:, 0].view(-1, 1)
    azim = 180.0 + V[:, 1].view(-1, 1)
    X = torch.stack((torch.cos(factor * elev) * torch.sin(factor * azim),
                     torch.sin(factor * elev),
                     torch.cos(factor * elev) * torch.cos(factor * azim)), dim=2)
    return -(torch.ones(len(V), 1) * X.view(-1, 3))


def _normalize(v: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    return v / v.norm(dim=1, keepdim=True).clamp_min(eps)


def camera_rt(X_cam: torch.Tensor, V_cam: torch.Tensor):
    """pose -> (R (n,3,3), T (n,3)), world->view is X_view = X_world @ R + T.
    Restates get_camera_RT macarons_utils.py:940-957 and PyTorch3D look_at_view_transform /
    look_at_rotation (eye=X, at=X+rays, up=(0,1,0)): z=norm(at-eye), x=norm(up x z), y=norm(z x x),
    degenerate-x fallback, R has x,y,z as columns, T = -R^T eye."""
    X_cam = X_cam.to(torch.float32).view(-1, 3)
    rays = view_direction(V_cam)
    at = X_cam + rays
    up = torch.tensor([[0.0, 1.0, 0.0]]).expand(len(X_cam), 3)
    z_axis = _normalize(at - X_cam)
    x_axis = _normalize(torch.cross(up, z_axis, dim=1))
    y_axis = _normalize(torch.cross(z_axis, x_axis, dim=1))
    close = torch.isclose(x_axis, torch.tensor(0.0), atol=5e-3).all(dim=1, keepdim=True)
    if close.any():
        x_axis = torch.where(close, _normalize(torch.cross(y_axis, z_axis, dim=1)), x_axis)
    R = torch.stack((x_axis, y_axis, z_axis), dim=2)          # columns
    T = -torch.bmm(R.transpose(1, 2), X_cam[:, :, None])[:, :, 0]
    return R.contiguous(), T.contiguous()


def interpolate_pose(old_pose, new_pose, step: int, n_steps: int = 4, pose_n_azim: int = 8,
                     old_azim_idx: int | None = None, new_azim_idx: int | None = None):
    """Camera.update_camera macarons_utils.py:2590-2632 for interpolation_step < n_steps:
    X = old + (new-old)*step/n ; same for (elev, azim) with the +-360 wrap between azimuth
    index 0 and pose_n_azim-1."""
    old_pose = torch.as_tensor(old_pose, dtype=torch.float32)
    new_pose = torch.as_tensor(new_pose, dtype=torch.float32)
    if step == n_steps:
        return new_pose[:3].clone(), new_pose[3:].clone()
    off = 0.0
    if old_azim_idx is not None and new_azim_idx is not None:
        if old_azim_idx == 0 and new_azim_idx == pose_n_azim - 1:
            off = -360.0
        elif old_azim_idx == pose_n_azim - 1 and new_azim_idx == 0:
            off = 360.0
    X = old_pose[:3] + (new_pose[:3] - old_pose[:3]) * step / n_steps
    V = old_pose[3:] + (new_pose[3:] - old_pose[3:]) * step / n_steps
    V[-1] = V[-1] + off * step / n_steps
    return X, V


# --------------------------------------------------------------------------- a2 raster
def render_depth(verts, faces, R, T, H: int = 256, W: int = 456, fov_deg: float = 60.0,
                 z_clip: float = 0.5, nthreads: int = 1):
    """Depth render of one mesh from one camera -> (zbuf (H,W) fp32 view-z, -1 = miss; pix_to_face int32).
    Reference: Camera.capture_image macarons_utils.py:2743-2786 -> fragments.zbuf; PyTorch3D naive
    rasteriser restated in oracle/raster_oracle.c."""
    v = np.ascontiguousarray(np.asarray(verts, dtype=np.float32))
    f = np.ascontiguousarray(np.asarray(faces, dtype=np.int64))
    Rn = np.ascontiguousarray(np.asarray(R, dtype=np.float32).reshape(9))
    Tn = np.ascontiguousarray(np.asarray(T, dtype=np.float32).reshape(3))
    z = np.empty((H, W), dtype=np.float32)
    p2f = np.empty((H, W), dtype=np.int32)
    _lib().nbp_oracle_render_depth(_fp(v), v.shape[0], f.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)),
                                   f.shape[0], _fp(Rn), _fp(Tn), float(focal_ndc(fov_deg)), float(z_clip),
                                   H, W, _fp(z), p2f.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)),
                                   int(nthreads))
    return z, p2f


# --------------------------------------------------------------------------- a4 unproject
def ndc_tables(H: int = 256, W: int = 456):
    """Camera.__init__ NDC tables macarons_utils.py:2270-2279 (NOT the rasteriser's pixel centres):
    nx[j] = W/m - (j/(m-1))*2 ; ny[i] = H/m - (i/(m-1))*2 ; m = min(H, W); fp32 throughout."""
    m = min(H, W)
    j = np.arange(W, dtype=np.float32)
    i = np.arange(H, dtype=np.float32)
    nx = f32(W / m) - (j / f32(m - 1)) * f32(2)
    ny = f32(H / m) - (i / f32(m - 1)) * f32(2)
    return nx.astype(np.float32), ny.astype(np.float32)


def unproject(zbuf, R, T, fov_deg: float = 60.0):
    """Camera.project_depth_in_3D macarons_utils.py:2788-2809 -> (H*W, 3) world points.
    The reference goes through FoVPerspectiveCameras.unproject_points (4x4 inverse + homogeneous
    divide, not bit-reproducible across devices).  The pin is the algebraically equal closed form
    (SURVEY.md section 7): s = z*tan(fov/2); view = (nx*s, ny*s, z); world = (view - T) @ R^T,
    each op rounded to fp32, sums left to right."""
    z = np.asarray(zbuf, dtype=np.float32)
    H, W = z.shape[-2], z.shape[-1]
    z = z.reshape(H, W)
    nx, ny = ndc_tables(H, W)
    t = tan_half_fov(fov_deg)
    s = z * t
    xv = nx[None, :] * s
    yv = ny[:, None] * s
    Rn = np.asarray(R, dtype=np.float32).reshape(3, 3)
    Tn = np.asarray(T, dtype=np.float32).reshape(3)
    dx, dy, dz = xv - Tn[0], yv - Tn[1], z - Tn[2]
    out = np.empty((H, W, 3), dtype=np.float32)
    for k in range(3):
        out[..., k] = (dx * Rn[k, 0] + dy * Rn[k, 1]) + dz * Rn[k, 2]
    return out.reshape(H * W, 3)


def partial_point_cloud(zbuf, R, T, fov_range: float | None = 70.0, gathering_factor: float = 0.05,
                        indices=None, fov_deg: float = 60.0, mask=None):
    """Camera.compute_partial_point_cloud macarons_utils.py:2811-2847.
    mask = (zbuf > -1) & (zbuf < fov_range); valid world points in row-major pixel order; keep
    k = int(n*gathering_factor) of them: ``indices`` (a permutation prefix, e.g.
    torch.randperm(n)[:k] as :2837) selects which; None with gathering_factor == 1 keeps all."""
    z = np.asarray(zbuf, dtype=np.float32)
    flat = z.reshape(-1)
    mask = (flat > f32(-1)) if mask is None else (np.asarray(mask).reshape(-1) != 0)
    if fov_range is not None:
        mask &= flat < f32(fov_range)
    pts = unproject(z, R, T, fov_deg)[mask]
    n = len(pts)
    k = int(n * gathering_factor)
    if indices is None:
        if k != n:
            raise ValueError("indices are required unless gathering_factor keeps every point")
        return pts
    idx = np.asarray(indices, dtype=np.int64)[:k]
    return pts[idx]


# --------------------------------------------------------------------------- a6-a10 grid
def y_bins_from_verts(verts: torch.Tensor, n_pieces: int = 4) -> torch.Tensor:
    """Slab boundaries, nbp_planning.py:446-451 (min_y+.5, max_y-.5, torch.arange with float step;
    may yield n_pieces+1 or n_pieces+2 elements -- SURVEY.md section 7)."""
    verts = torch.as_tensor(verts, dtype=torch.float32)
    min_y = torch.min(verts, dim=0)[0][1].item() + 0.5
    max_y = torch.max(verts, dim=0)[0][1].item() - 0.5
    bin_width = (max_y - min_y) / n_pieces
    return torch.arange(min_y, max_y + bin_width, bin_width)


def slab_index(y, bounds):
    """nbp_planning.py:114: bucketize(y, y_bins[:-1]) - 1 with right=False
    == (number of boundaries strictly below y) - 1.  ``bounds`` is y_bins[:-1]."""
    y = np.asarray(y, dtype=np.float32)
    b = np.asarray(bounds, dtype=np.float32)
    return (y[:, None] > b[None, :]).sum(axis=1).astype(np.int64) - 1


def transform_points(points, pose):
    """transform_points_to_n_pieces utils.py:166-196 with no_rotation=True (every call site):
    p = (-(z - c_z), -(x - c_x)) -> (N, 2) fp32."""
    p = np.asarray(points, dtype=np.float32).reshape(-1, 3)
    c = np.asarray(pose, dtype=np.float32).reshape(-1)
    return np.stack((-(p[:, 2] - c[2]), -(p[:, 0] - c[0])), axis=1).astype(np.float32)


def cell_index(p2d, S: int, grid_range=(-40, 40)):
    """Rounding shared by map_points_to_n_imgs utils.py:206-207 and get_point_position_in_the_img
    utils.py:160-164: rint((p - lo) * fp32(S/(hi-lo))), half-to-even, each op rounded to fp32."""
    p = np.asarray(p2d, dtype=np.float32)
    scale = f32(S / (grid_range[1] - grid_range[0]))
    lo = f32(grid_range[0])
    return np.rint((p - lo) * scale).astype(np.int64)


def map_points(p2d, S: int, grid_range=(-40, 40)):
    """map_points_to_n_imgs utils.py:198-223 for one image: (S,S) fp32 counts,
    out[r, c] += 1 for every point whose (r, c) lies inside the grid."""
    rc = cell_index(p2d, S, grid_range).reshape(-1, 2)
    ok = (rc[:, 0] >= 0) & (rc[:, 0] < S) & (rc[:, 1] >= 0) & (rc[:, 1] < S)
    out = np.zeros((S, S), dtype=np.float32)
    np.add.at(out, (rc[ok, 0], rc[ok, 1]), f32(1))
    return out


def build_model_input(cloud, pose, bounds, trajectory, S: int = 256, grid_range=(-40, 40), n_pieces: int = 4):
    """Model input assembly nbp_planning.py:114-132,166 -> (n_pieces+1, S, S) fp32:
    n_pieces height-slab count images of the cloud + one count image of the camera trajectory,
    all in the egocentric frame of ``pose`` (x, y, z, elev, azim)."""
    cloud = np.asarray(cloud, dtype=np.float32).reshape(-1, 3)
    out = np.zeros((n_pieces + 1, S, S), dtype=np.float32)
    if len(cloud):
        sl = slab_index(cloud[:, 1], bounds)
        p2d = transform_points(cloud, pose)
        for i in range(n_pieces):
            out[i] = map_points(p2d[sl == i], S, grid_range)
    traj = np.asarray(trajectory, dtype=np.float32).reshape(-1, 3)
    out[n_pieces] = map_points(transform_points(traj, pose), S, grid_range)
    return out


# --------------------------------------------------------------------------- re-plan read-out
def fuse_obstacle_map(pred_obstacle, cloud, pose, trajectory, S: int = 256, grid_range=(-40, 40), threshold: float = 0.13):
    """Obstacle-map fusion of the re-plan branch, nbp_planning.py:168-190.
    pred_obstacle (S,S) probabilities -> (fused (S,S) in {0,1}, full_proj (S,S) in {0,1}).
    PINNED by tests/golden/planner.npz, produced by executing the reference's own lines (make_golden.py)."""
    cloud = np.asarray(cloud, dtype=np.float32).reshape(-1, 3)
    pose = np.asarray(pose, dtype=np.float32).reshape(-1)
    fused = (np.asarray(pred_obstacle, dtype=np.float32) >= f32(threshold)).astype(np.float32)
    p2d = transform_points(cloud, pose)
    full_proj = np.minimum(map_points(p2d, S, grid_range), f32(1))                      # full_pc_projection[... > 1] = 1
    lo, hi = f32(float(pose[1]) - 0.1), f32(float(pose[1]) + 0.1)                       # python float thresholds, compared in fp32
    sel = (cloud[:, 1] < hi) & (cloud[:, 1] > lo)
    slice_img = (map_points(p2d[sel], S, grid_range) > 0).astype(np.float32)
    seen = full_proj > 0
    fused[seen] = slice_img[seen]
    traj_img = map_points(transform_points(trajectory, pose), S, grid_range)
    fused[traj_img > 0] = 0
    return fused, full_proj


def score_candidates(value_map, full_proj, candidates, pose, skip=None, S: int = 256, Sv: int = 64, grid_range=(-40, 40), window: int = 10):
    """Candidate scoring loop nbp_planning.py:193-231 + check_pixel_values macarons_utils.py:86-100.
    value_map (8,Sv,Sv), full_proj (S,S), candidates (M,3).  Returns (valid (M,) bool, cell (M,2) int, score (M,) float64)
    with score = value - 10*density exactly as the Python loop computes it (float64 of two fp32 .item()s)."""
    vm = np.asarray(value_map, dtype=np.float32)
    fp = np.asarray(full_proj, dtype=np.float32)
    cand = np.asarray(candidates, dtype=np.float32).reshape(-1, 3)
    M = len(cand)
    valid = np.zeros(M, dtype=bool); cell = -np.ones((M, 2), dtype=np.int64); score = np.zeros(M, dtype=np.float64)
    max_gain = vm.max(axis=0)
    p2d = transform_points(cand, pose)
    gv = cell_index(p2d, Sv, grid_range)
    gs = cell_index(p2d, S, grid_range)
    for j in range(M):
        if skip is not None and skip[j]:
            continue
        r, c = int(gv[j, 0]), int(gv[j, 1])
        if not (0 <= r < Sv and 0 <= c < Sv):
            continue
        cell[j] = (r, c)
        x, y = int(gs[j, 0]), int(gs[j, 1])
        dens = fp[x, y]                                                     # numpy wraps negative indices like torch
        region = fp[max(x - window, 0): min(x + window + 1, S), max(y - window, 0): min(y + window + 1, S)]
        if not (region == 1).any():
            continue
        valid[j] = True
        score[j] = float(max_gain[r, c]) - 10 * float(dens)
    return valid, cell, score


# --------------------------------------------------------------------------- section 8f row 2: coverage metric
def coverage_percentage(gt, pc, threshold: float = 1.0, weight: int = 2, indices=None, return_flags: bool = False):
    """calculate_coverage_percentage next_best_path/utility/long_term_utils.py:457-468 (random_sample_pc :436-446,
    find_nearest_points_distances :448-455): fraction of ground-truth points ``gt`` (G,3) whose nearest point of the
    (sub-sampled) reconstruction ``pc`` (N,3) is closer than ``threshold``.  If N > weight*G the reference keeps
    ``pc[torch.randperm(N)[:weight*G]]``: pass that index prefix as ``indices``.  Distances are evaluated in strict fp32 as
    d2 = (dx*dx + dy*dy) + dz*dz < threshold*threshold (the arithmetic csrc/coverage.cu is pinned to); torch.cdist itself switches
    to a matmul expansion for large inputs, so the reference is reproducible only to ~1e-3 in distance (tests count the
    ground-truth points that close to the threshold)."""
    g = np.ascontiguousarray(np.asarray(gt, dtype=np.float32).reshape(-1, 3))
    q = np.ascontiguousarray(np.asarray(pc, dtype=np.float32).reshape(-1, 3))
    if len(q) == 0:
        return (0.0, np.zeros(len(g), dtype=bool)) if return_flags else 0.0
    want = int(len(g) * weight)
    if len(q) > want:
        if indices is None:
            raise ValueError("reconstruction longer than weight*G: pass the randperm prefix as `indices`")
        q = q[np.asarray(indices, dtype=np.int64)[:want]]
    thr2 = np.float32(threshold) * np.float32(threshold)
    flags = np.zeros(len(g), dtype=bool)
    step = max(1, (1 << 24) // max(len(q), 1))
    for a in range(0, len(g), step):
        blk = g[a:a + step]
        dx = blk[:, None, 0] - q[None, :, 0]; dy = blk[:, None, 1] - q[None, :, 1]; dz = blk[:, None, 2] - q[None, :, 2]
        d2 = (dx * dx + dy * dy) + dz * dz                                       # float32 throughout, this association
        flags[a:a + step] = (d2 < thr2).any(axis=1)
    cov = float(np.float32(flags.sum()) / np.float32(len(g))) if len(g) else 0.0
    return (cov, flags) if return_flags else cov


# --------------------------------------------------------------------------- section 8f row 3: collision rays
def _dot3(a, b):
    return (a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1]) + a[..., 2] * b[..., 2]


def _cross3(a, b):
    return np.stack((a[..., 1] * b[..., 2] - a[..., 2] * b[..., 1], a[..., 2] * b[..., 0] - a[..., 0] * b[..., 2],
                     a[..., 0] * b[..., 1] - a[..., 1] * b[..., 0]), axis=-1)


def segment_mesh_hits(verts, faces, segments, rays: bool = False):
    """line_segment_mesh_intersection macarons/utility/macarons_utils.py:120-151 (rays=False: segments (n,6) = start, end; returns
    (hit (n,) bool, count (n,) int)) and the ray casts of check_camera_in_mesh next_best_path/utility/long_term_utils.py:158-170
    (rays=True: (origin, unit direction)).  The reference delegates to ``trimesh.ray.intersects_location`` (trimesh 4.1.2, not
    in the tree, not installable here: PARITY UNPINNED).  This restates Trimesh's published ray_triangle formulation in float64:
    plane intersection, barycentric coordinates by Cramer's rule, accept when all barycentrics lie in [-1e-13, 1+1e-13] and the hit
    is forward of the origin (distance > -1e-6); a segment keeps hits with |location - start| < |end - start| (:143-144).
    Every operation is an explicit elementwise float64 op in a fixed order: csrc/collision.cu reproduces it bit for bit."""
    v = np.asarray(verts, dtype=np.float32).astype(np.float64)
    tri = v[np.asarray(faces, dtype=np.int64)]                                   # (F, 3, 3)
    seg = np.asarray(segments, dtype=np.float32).astype(np.float64).reshape(-1, 6)
    v0, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    n = _cross3(e1, e2)
    nn = _dot3(n, n)
    d00, d01, d11 = _dot3(e1, e1), _dot3(e1, e2), _dot3(e2, e2)
    hit = np.zeros(len(seg), dtype=bool); count = np.zeros(len(seg), dtype=np.int64)
    for i, s in enumerate(seg):
        o, d = s[:3], s[3:]
        length = 0.0
        if not rays:
            d = d - o
            length = np.sqrt(_dot3(d, d))
            d = d / length
        with np.errstate(divide="ignore", invalid="ignore"):
            denom = _dot3(n, d[None, :])
            ok = (nn > 0.0) & (np.abs(denom) > 1e-13 * np.sqrt(nn))
            t = _dot3(n, v0 - o[None, :]) / denom
            ok &= t > -1e-6
            p = o[None, :] + t[:, None] * d[None, :]
            w = p - v0
            d20, d21 = _dot3(w, e1), _dot3(w, e2)
            inv = 1.0 / (d00 * d11 - d01 * d01)
            b1 = (d11 * d20 - d01 * d21) * inv
            b2 = (d00 * d21 - d01 * d20) * inv
            b0 = (1.0 - b1) - b2
            lo, hi = -1e-13, 1.0 + 1e-13
            ok &= (b0 > lo) & (b1 > lo) & (b2 > lo) & (b0 < hi) & (b1 < hi) & (b2 < hi)
            if not rays:
                dl = p - o[None, :]
                ok &= np.sqrt(_dot3(dl, dl)) < length
        count[i] = int(ok.sum()); hit[i] = bool(ok.any())
    return hit, count


# --------------------------------------------------------------------------- f3 ground-truth obstacle map
GT_MAP_HALF_WIDTH_PX = 1.35      # matplotlib's 1.5 pt line at 100 dpi, after the resize of the ~198 px axes to 256 px (section_oracle.c)


def gt_obstacle_map(verts, faces, pose, S: int = 256, view: float = 80.0, half_width: float = GT_MAP_HALF_WIDTH_PX, return_segments=False):
    """get_binary_obstacle_array(mesh, camera_pose, view_size) next_best_path/utility/utils.py:226-262 (called nbp_utils.py:638):
    the mesh cut by the horizontal plane through the camera, drawn into an S x S binary image centred on the camera
    (rows towards -z, columns towards -x).  Restatement in oracle/section_oracle.c -- PARITY UNPINNED (trimesh / matplotlib absent).
    Returns (S, S) float32 of 0/1 [, (n_seg, 4) segments x0, z0, x1, z1]."""
    v = np.ascontiguousarray(verts, dtype=f32)
    f = np.ascontiguousarray(faces, dtype=np.int64)
    p = np.ascontiguousarray(np.asarray(pose, dtype=f32).reshape(-1)[:3])
    out = np.zeros((S, S), dtype=f32)
    seg = np.zeros((max(len(f), 1), 4), dtype=f32)
    n = _lib().nbp_oracle_plane_section_map(_fp(v), f.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), f.shape[0], _fp(p), int(S),
                                            float(view), float(half_width), _fp(out), _fp(seg))
    return (out, seg[:n].copy()) if return_segments else out
