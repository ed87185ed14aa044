"""Batched exploration inner loop: B independent scenes advance one pose per step on one GPU.

One step per scene = one ``pose_i`` iteration of ``compute_nbp_trajectory``
(/root/reference/next_best_path/testers/nbp_planning.py:60-355) restricted to the scoped stages
(SURVEY.md section 8d):

  A  back-project the current key frame (5 % sample) and append to the scene's cloud      :96-105
  B  slab split + egocentric histogram of the WHOLE cloud + trajectory image -> (5,S,S)    :114-132
  C  NBP forward (eval) -> value map (8,S/4,S/4), obstacle map (1,S,S); max over headings  :166,194
  D  move: 4 interpolated cameras towards the next key pose, 4 depth renders               :271-274
  E  back-project frames [old key, interp 1, 2, 3] (key frame a second time) and append    :334-352

The planner, collision checks, coverage metric and disk I/O of the reference are out of scope: the
next key pose of every scene is an input of ``step``.  Frames never leave HBM (the reference writes
each one to disk and reloads it, macarons_utils.py:2782,:992).  All state lives in preallocated device
buffers; a step enqueues ~25 geometry kernels plus one CUDA-graph replay of the network (stage C) and performs no host
synchronisation.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import ops
from .utility.camera import get_camera_RT

POINTS_PER_FRAME_MAX = 5836          # int(0.05 * 256 * 456)


@dataclass
class StepOutput:
    value_map: torch.Tensor      # (B, 8, S/4, S/4) fp32
    obstacle_map: torch.Tensor   # (B, 1, S, S) fp32
    value_max: torch.Tensor      # (B, S/4, S/4) fp32, max over the 8 headings (nbp_planning.py:194)
    model_input: torch.Tensor    # (B, 5, S, S) fp32 counts


def slab_bounds_from_verts(verts: torch.Tensor, n_pieces: int = 4) -> torch.Tensor:
    """y_bins[:-1] exactly as the reference builds it (nbp_planning.py:446-451): the torch.arange has
    n_pieces+1 or n_pieces+2 elements depending on float rounding, so it must not be re-derived elsewhere."""
    min_y = torch.min(verts, dim=0)[0][1].item() + 0.5
    max_y = torch.max(verts, dim=0)[0][1].item() - 0.5
    bin_width = (max_y - min_y) / n_pieces
    return torch.arange(min_y, max_y + bin_width, bin_width)[:-1]


def interpolated_poses(old_pose, new_pose, old_az_idx, new_az_idx, n_steps=4, pose_n_azim=8):
    """Camera.update_camera (macarons_utils.py:2590-2632), vectorised over scenes: returns (n_steps, B, 5) poses of
    interpolation steps 1..n_steps; step n_steps lands exactly on ``new_pose``; azimuth wraps between index 0 and 7."""
    old = torch.as_tensor(old_pose, dtype=torch.float32)
    new = torch.as_tensor(new_pose, dtype=torch.float32)
    oa, na = torch.as_tensor(old_az_idx), torch.as_tensor(new_az_idx)
    off = torch.zeros(old.shape[0], dtype=torch.float32)
    off[(oa == 0) & (na == pose_n_azim - 1)] = -360.0
    off[(oa == pose_n_azim - 1) & (na == 0)] = 360.0
    out = torch.empty((n_steps, old.shape[0], 5), dtype=torch.float32)
    for k in range(1, n_steps + 1):
        if k == n_steps:
            out[k - 1] = new
        else:
            p = old + (new - old) * k / n_steps
            p[:, 4] = p[:, 4] + off * k / n_steps
            out[k - 1] = p
    return out


def shard_scenes(n_scenes: int, world_size: int, rank: int):
    """Contiguous block of scene indices owned by `rank` (scenes are independent, so rollouts
    shard with no data-path collective).  Returns (first, count); blocks differ by at most one scene."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank {rank} / world size {world_size}")
    base, extra = divmod(n_scenes, world_size)
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def max_over_ranks(value: float, device=None) -> float:
    """Multi-GPU timings are the max over ranks (one all-reduce of a scalar; the only collective of a rollout)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


class _null_ctx:
    def __enter__(self):
        return None

    def __exit__(self, *a):
        return False


STAGE_NAMES = ("A_backproject_key", "B_grid_scatter", "C_nbp_forward", "D_render_4_views", "E_backproject_4_frames")


class RolloutEngine:
    def __init__(self, scenes, nbp, device, S=256, H=256, W=456, max_steps=100, gathering_factor=0.05,
                 sensor_range=70.0, grid_range=(-40.0, 40.0), n_pieces=4, seed=9):
        self.dev = torch.device(device)
        self.nbp = None
        self.set_network(nbp)
        self.B, self.S, self.H, self.W = len(scenes), S, H, W
        self.gf, self.sensor_range, self.grid_range, self.n_pieces, self.seed = gathering_factor, sensor_range, grid_range, n_pieces, seed
        dev = self.dev
        B = self.B
        # ---- static scene data
        self.face_counts = [int(s.faces.shape[0]) for s in scenes]
        self.verts = torch.from_numpy(np.concatenate([s.verts for s in scenes]).astype(np.float32)).to(dev)
        self.faces = torch.from_numpy(np.concatenate([s.faces for s in scenes]).astype(np.int32)).to(dev)
        self.vert_off = torch.tensor(np.concatenate([[0], np.cumsum([len(s.verts) for s in scenes])]), dtype=torch.int64, device=dev)
        self.face_off = torch.tensor(np.concatenate([[0], np.cumsum(self.face_counts)]), dtype=torch.int64, device=dev)
        bounds = torch.zeros((B, 8), dtype=torch.float32)
        nb = torch.zeros(B, dtype=torch.int32)
        for i, s in enumerate(scenes):
            b = slab_bounds_from_verts(torch.from_numpy(s.verts), n_pieces)
            bounds[i, : len(b)] = b
            nb[i] = len(b)
        self.slab_bounds, self.n_bounds = bounds.to(dev), nb.to(dev)
        # ---- dynamic state
        frames_per_step = 5
        self.cap = ((max_steps + 1) * frames_per_step * int(gathering_factor * H * W + 1) + 3) // 4 * 4
        self.cloud = torch.empty((B, self.cap, 3), dtype=torch.float32, device=dev)
        self.cloud_len = torch.zeros(B, dtype=torch.int32, device=dev)
        self.traj_cap = 8 + 4 * (max_steps + 1)
        self.traj = torch.zeros((B, self.traj_cap, 3), dtype=torch.float32, device=dev)
        self.traj_len = torch.zeros(B, dtype=torch.int32, device=dev)
        self.traj_len_host = 0
        # slot-major frame ring: slot 0 = key frame, 1..3 = interpolated frames, 4 = new key frame
        self.frames = torch.empty((5, B, H, W), dtype=torch.float32, device=dev)
        self.frame_R = torch.empty((5, B, 9), dtype=torch.float32, device=dev)
        self.frame_T = torch.empty((5, B, 3), dtype=torch.float32, device=dev)
        self.pose = torch.zeros((B, 5), dtype=torch.float32, device=dev)
        self.grid = torch.empty((B, n_pieces + 1, S, S), dtype=torch.float32, device=dev)
        self.overflow = torch.zeros(1, dtype=torch.int32, device=dev)
        self.scene_ids = torch.arange(B, dtype=torch.int32, device=dev)
        self.view_scene4 = self.scene_ids.repeat(4).contiguous()             # view = slot * B + scene
        self.view_scene4_host = list(range(B)) * 4
        self.step_idx = 0
        self.max_points_bound = 0
        self.stage_events = None        # set to [] to record CUDA events at the stage boundaries of the next step(s)
        self._uid_base = torch.arange(B, dtype=torch.int32, device=dev) * 8
        self._copy_stream = None        # side stream of the device->host read-back (created on first use)
        self._copy_done = None
        # stages D/E on a side stream under stage C (see step).  Off by default: measured on B200 at 256 scenes the step is
        # throughput-bound (114.4 ms with, 114.8 ms without) -- the geometry CTAs only displace the network's persistent CTAs
        self.overlap_geometry = False
        self._geo_stream = None
        self._after_b = None

    def set_network(self, nbp):
        """The engine consumes (or copies out) the maps of a step before it runs the next forward: NBP.forward may hand out the
        output buffers of its captured graph instead of clones.  StepOutput tensors are therefore valid until the next step()."""
        self.nbp = nbp
        if nbp is not None:
            nbp.static_outputs = True

    # ------------------------------------------------------------------ helpers
    def _uids(self, n_slots, slot0):
        """Unique counter-based RNG stream id per (step, slot, scene), slot-major like the frames."""
        base = (self._uid_base + self.step_idx * 8 * self.B).view(1, self.B)
        return (base + torch.arange(slot0, slot0 + n_slots, dtype=torch.int32, device=self.dev).view(-1, 1)).reshape(-1).contiguous()

    def _render(self, R, T, view_scene, view_scene_host, out):
        ops.raster_depth(self.verts, self.faces, self.vert_off, self.face_off, view_scene, R, T, self.H, self.W,
                         self.face_counts, view_scene_host, zbuf=out)

    def _append_traj(self, X):
        """camera.X_cam_history (macarons_utils.py:2628): X (B, k, 3) device tensor."""
        k = X.shape[1]
        self.traj[:, self.traj_len_host: self.traj_len_host + k] = X
        self.traj_len_host += k
        self.traj_len.fill_(self.traj_len_host)

    # ------------------------------------------------------------------ API
    def reset(self, start_pose, approach_pose=None, approach_az=None, start_az=None):
        """start_pose (B,5) host tensor (x,y,z,elev,azim): renders the first key frame of every scene.
        ``approach_pose`` (B,5) + azimuth indices: the reference's set-up code places the camera on a neighbouring lattice pose
        and walks to the start pose in 4 interpolation steps (macarons/testers/scene.py:466-486), so ``X_cam_history`` -- the
        trajectory image, channel 4 of the model input -- starts with those 1+4 positions; the 4 set-up frames themselves are
        never back-projected by the NBP loop.  Without it the trajectory starts at the start pose."""
        start = torch.as_tensor(start_pose, dtype=torch.float32)
        self.cloud_len.zero_(); self.traj_len_host = 0; self.traj_len.zero_(); self.overflow.zero_()
        self.step_idx = 0; self.max_points_bound = 0
        R, T = get_camera_RT(start[:, :3], start[:, 3:])
        self.pose.copy_(start.to(self.dev))
        self.frame_R[0].copy_(R.reshape(-1, 9).to(self.dev)); self.frame_T[0].copy_(T.to(self.dev))
        self._render(self.frame_R[0], self.frame_T[0], self.scene_ids, list(range(self.B)), self.frames[0])
        if approach_pose is not None:
            appr = torch.as_tensor(approach_pose, dtype=torch.float32)
            steps = interpolated_poses(appr, start, approach_az, start_az)               # (4, B, 5); step 4 == start pose
            hist = torch.cat((appr[None, :, :3], steps[:, :, :3]), dim=0).permute(1, 0, 2).contiguous()
            self._append_traj(hist.to(self.dev))
        else:
            self._append_traj(self.pose[:, None, :3])

    def upload_move(self, cur_pose, next_pose, cur_az, next_az):
        """Host side of stage D: interpolated poses -> (R, T) for the 4 views of every scene, copied to the device
        from pinned memory.  Returns device tensors (poses (4,B,5), R (4*B,9), T (4*B,3)), slot-major."""
        poses = interpolated_poses(cur_pose, next_pose, cur_az, next_az)
        R, T = get_camera_RT(poses.reshape(-1, 5)[:, :3], poses.reshape(-1, 5)[:, 3:])
        pin = lambda t: t.contiguous().pin_memory().to(self.dev, non_blocking=True)
        return pin(poses), pin(R.reshape(-1, 9)), pin(T)

    def coverage(self, index, weight: int = 2, seed=None):
        """calculate_coverage_percentage(gt_scene_pc, full_pc) for every scene (nbp_planning.py:71):
        ``index`` = nextbestpath_b200.coverage.CoverageIndex over the scenes' ground-truth clouds.  Returns (B,) fp32 on the
        device; no host synchronisation."""
        return index.coverage(self.cloud, self.cloud_len, weight=weight, seed=self.seed + self.step_idx if seed is None else seed,
                              max_points=self.max_points_bound)

    def build_model_input(self):
        """Stage B alone: the (B,5,S,S) model input of the current state (used to calibrate / inspect; ``step`` does this itself)."""
        ops.grid_scatter(self.cloud, self.cloud_len, self.pose, self.slab_bounds, self.n_bounds, self.S, traj=self.traj,
                         traj_len=self.traj_len, n_pieces=self.n_pieces, grid_range=self.grid_range,
                         max_points=self.max_points_bound, out=self.grid)
        return self.grid

    def wait_host_outputs(self):
        """Block the host until the read-back started by the last ``step(..., host_out=...)`` has landed.  The device
        keeps running stages D and E of that step meanwhile."""
        if self._copy_done is not None:
            self._copy_done.synchronize()

    def step(self, move, run_network: bool = True, host_out=None):
        """One rollout step for all scenes.  ``move`` = (poses (4,B,5), R (4*B,9), T (4*B,3)) on the device
        (from ``upload_move``): the 4 cameras leading to the next key pose.  ``run_network=False`` advances the
        geometry only (stages A, D, E): used to fast-forward a rollout to a given pose index.
        ``host_out`` = (value_max, value_map, obstacle_map) pinned host tensors: the network outputs are copied to them
        on a side stream as soon as stage C ends, overlapping the copy with stages D and E (which do not depend on the
        maps: the move was decided before the step); ``wait_host_outputs()`` tells when they can be read."""
        B, H, W = self.B, self.H, self.W
        poses4, R4, T4 = move
        n_new = int(self.gf * H * W)
        ev = self.stage_events
        mark = (lambda: ev.append(torch.cuda.Event(enable_timing=True)) or ev[-1].record()) if ev is not None else (lambda: None)
        mark()
        # ---- A: back-project the current key frame
        ops.backproject_append(self.frames[0], self.frame_R[0], self.frame_T[0], self.scene_ids, self.cloud, self.cloud_len,
                               frame_uid=self._uids(1, 0), fov_range=self.sensor_range, gathering_factor=self.gf,
                               seed=self.seed, overflow=self.overflow)
        self.max_points_bound = min(self.cap, self.max_points_bound + n_new)
        mark()
        out1 = out2 = vmax = None
        if run_network:
            # ---- B: model input
            ops.grid_scatter(self.cloud, self.cloud_len, self.pose, self.slab_bounds, self.n_bounds, self.S, traj=self.traj,
                             traj_len=self.traj_len, n_pieces=self.n_pieces, grid_range=self.grid_range,
                             max_points=self.max_points_bound, out=self.grid)
            mark()
            if self.overlap_geometry:
                self._after_b = torch.cuda.Event(); self._after_b.record()
            # ---- C: network
            if self._copy_done is not None:                      # the previous step's read-back still owns the output buffers
                torch.cuda.current_stream(self.dev).wait_event(self._copy_done)
            with torch.no_grad():
                out1, out2 = self.nbp(self.grid)
            vmax = self.nbp.last_value_max                       # max over the 8 headings, fused into the Final1 kernel (row a14)
            mark()
            if host_out is not None:
                if self._copy_stream is None:
                    self._copy_stream = torch.cuda.Stream(device=self.dev)
                    self._copy_done = torch.cuda.Event()
                ready = torch.cuda.Event(); ready.record()
                with torch.cuda.stream(self._copy_stream):
                    self._copy_stream.wait_event(ready)
                    for h, d in zip(host_out, (vmax, out1, out2)):
                        h.copy_(d, non_blocking=True)
                        d.record_stream(self._copy_stream)
                    self._copy_done.record()
        # ---- D + E do not depend on the network (the move was decided before the step): with ``overlap_geometry`` they run on a side
        # stream UNDER stage C and rejoin the main stream at the end of the step (default: one stream, see __init__)
        main = torch.cuda.current_stream(self.dev)
        side = None
        if self.overlap_geometry and run_network:
            if self._geo_stream is None:
                self._geo_stream = torch.cuda.Stream(device=self.dev)
            side = self._geo_stream
            side.wait_event(self._after_b)
        with torch.cuda.stream(side) if side is not None else _null_ctx():
            # ---- D: move + render 4 frames straight into slots 1..4 (slot 4 is the new key frame)
            self._render(R4, T4, self.view_scene4, self.view_scene4_host, self.frames[1:5].view(4 * B, H, W))
            self.frame_R[1:5].copy_(R4.view(4, B, 9)); self.frame_T[1:5].copy_(T4.view(4, B, 3))
            self._append_traj(poses4[:, :, :3].permute(1, 0, 2))
            mark()
            # ---- E: back-project [old key, interp1, interp2, interp3]; frames of a scene append in slot order
            ops.backproject_append(self.frames[0:4].view(4 * B, H, W), self.frame_R[0:4].view(4 * B, 9), self.frame_T[0:4].view(4 * B, 3),
                                   self.view_scene4, self.cloud, self.cloud_len, frame_uid=self._uids(4, 1),
                                   fov_range=self.sensor_range, gathering_factor=self.gf, seed=self.seed, overflow=self.overflow)
            self.max_points_bound = min(self.cap, self.max_points_bound + 4 * n_new)
            # ---- the new key frame becomes slot 0
            self.frames[0].copy_(self.frames[4]); self.frame_R[0].copy_(self.frame_R[4]); self.frame_T[0].copy_(self.frame_T[4])
            self.pose.copy_(poses4[3])
            mark()
            if side is not None:
                done = torch.cuda.Event(); done.record(side)
        if side is not None:
            main.wait_event(done)
        self.step_idx += 1
        return StepOutput(out1, out2, vmax, self.grid) if run_network else None
