"""oracle/driver_lines.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restatement of the two pieces of *driver* code of the reference that call the hot path, so that the drop-in shims
(`nextbestpath_b200.networks.NBP`, `nextbestpath_b200.utility.utils.*`) can be exercised on the GPU box, where
/root/reference does not exist, in exactly the call pattern the reference drivers use:

  pose_map_build_and_forward : next_best_path/testers/nbp_planning.py:112-132 (slab split, per-slab map build, trajectory
                               image) and :166-193 (NBP forward on the 5-channel input, obstacle-map fusion, heading max)
  train_experience_data      : next_best_path/utility/nbp_utils.py:340-391 (micro-batch assembly, forward, sparse gather,
                               NBP.loss, GradScaler-scaled backward, accumulated optimizer step)

Both functions receive every callee (`nbp`, `transform_points_to_n_pieces`, `map_points_to_n_imgs`, `GradScaler`, ...)
as an argument: the test decides whether the reference's CPU functions or the CUDA shims are bound.

PINNED: tests/test_dropin_lines.py executes the reference's own source lines from /root/reference (when present, i.e.
in the build container) and these restatements on the same inputs with the same callees and requires identical results;
tests/golden/dropin.npz holds the outputs of the reference's own lines bound to the reference's own CPU functions
(tests/golden/make_golden.py), which the GPU test compares the shims against.
"""
from __future__ import annotations

import random

import numpy as np
import torch


def pose_map_build_and_forward(full_pc, y_bins, n_pieces, camera_current_pose, X_cam_history, pc2img_size, prediction_range,
                               device, nbp, transform_points_to_n_pieces, map_points_to_n_imgs):
    """One pose of compute_nbp_trajectory between "cloud accumulated" and "candidates scored".
    Returns the tensors the driver holds afterwards, by the driver's own variable names."""
    # ---- nbp_planning.py:114-127: height slabs -> one count image each (an empty slab is a zero image)
    slab_of_point = torch.bucketize(full_pc[:, 1], y_bins[:-1]) - 1
    slab_imgs = []
    for s in range(n_pieces):
        pts = full_pc[slab_of_point == s]
        if len(pts) > 0:
            img = map_points_to_n_imgs(transform_points_to_n_pieces(pts, camera_current_pose, device), pc2img_size,
                                       prediction_range, device)
        else:
            img = torch.zeros(1, pc2img_size[0], pc2img_size[1], device=device)
        slab_imgs.append(img)
    current_pc_imgs = torch.cat(slab_imgs, dim=0).unsqueeze(0)
    # ---- :130-132: the camera trajectory so far as a fifth image
    trajectory_2d = transform_points_to_n_pieces(X_cam_history, camera_current_pose, device)
    current_previous_trajectory_img = map_points_to_n_imgs(trajectory_2d, pc2img_size, prediction_range, device).unsqueeze(0)
    # ---- :166-169: forward, obstacle map binarised at 0.13
    model_input = torch.cat((current_pc_imgs, current_previous_trajectory_img), dim=1).to(device)
    predicted_value_map, raw_obstacle_map = nbp(model_input)
    predicted_obstacle_map = (raw_obstacle_map >= 0.13).float()
    # ---- :172-175: all points -> one binary image
    full_pc_projection = map_points_to_n_imgs(transform_points_to_n_pieces(full_pc, camera_current_pose, device), pc2img_size,
                                              prediction_range, device).unsqueeze(0)
    full_pc_projection[full_pc_projection > 1] = 1
    # ---- :178-183: the +-0.1 slice around the camera height -> binary image
    cam_y = camera_current_pose[1].item()
    near = (full_pc[:, 1] < cam_y + 0.1) & (full_pc[:, 1] > cam_y - 0.1)
    filt_pc_selection_img = map_points_to_n_imgs(transform_points_to_n_pieces(full_pc[near], camera_current_pose, device),
                                                 pc2img_size, prediction_range, device).unsqueeze(0)
    filt_pc_selection_img[filt_pc_selection_img > 0] = 1
    # ---- :186-190: observed cells override the prediction; visited cells are passable
    seen = full_pc_projection > 0
    predicted_obstacle_map[seen] = filt_pc_selection_img[seen]
    predicted_obstacle_map[current_previous_trajectory_img > 0] = 0
    # ---- :193: best heading per value-map cell
    max_gain_map, _ = torch.max(predicted_value_map, dim=1, keepdim=True)
    return {"model_input": model_input, "predicted_value_map": predicted_value_map, "raw_obstacle_map": raw_obstacle_map,
            "predicted_obstacle_map": predicted_obstacle_map, "full_pc_projection": full_pc_projection,
            "max_gain_map": max_gain_map}


def train_experience_data(training_set_db, params, optimizer, nbp, device, current_epoch, GradScaler):
    """nbp_utils.py:340-391.  `training_set_db`: list of replay records (dicts of numpy arrays: current_model_input (1,5,S,S),
    current_gt_2d_layout (1,1,S,S), target_value_map_pixel (K,3) long, actual_coverage_gain (K,), pose_i).
    Gradients of up to 8 micro-batches accumulate before one optimizer step; the running-loss average always divides by 8."""
    random.shuffle(training_set_db)
    scaler = GradScaler()
    accumulate = 8
    epoch_losses, loss_sum, n_pending = [], 0, 0
    bs = params.nbp_batch_size
    for first in range(0, len(training_set_db), bs):
        usable = [rec for rec in training_set_db[first:first + bs] if current_epoch > 1 or rec["pose_i"] > 10]
        tensors = [[torch.from_numpy(np.copy(rec[k])).to(device) for k in
                    ("current_model_input", "current_gt_2d_layout", "target_value_map_pixel", "actual_coverage_gain")] for rec in usable]
        if not tensors:
            continue
        inputs = torch.cat([t[0] for t in tensors])
        layouts = torch.cat([t[1] for t in tensors])
        coords = torch.cat([t[2] for t in tensors]).to(device)
        gains = torch.cat([t[3] for t in tensors]).to(device)
        counts = [len(t[2]) for t in tensors]
        sample_of = torch.repeat_interleave(torch.arange(len(counts), device=device), torch.tensor(counts, device=device))
        value_map, obstacle_map = nbp(inputs)
        picked = value_map[sample_of, coords[:, 0], coords[:, 1], coords[:, 2]]
        batch_loss = nbp.loss(picked, gains, obstacle_map, layouts)
        scaler.scale(batch_loss).backward()
        loss_sum += batch_loss.item()
        n_pending += 1
        if n_pending % accumulate == 0 or first + bs >= len(training_set_db):
            scaler.step(optimizer)
            scaler.update()
            optimizer.zero_grad()
            epoch_losses.append(loss_sum / accumulate)
            loss_sum, n_pending = 0, 0
    return epoch_losses


# ----------------------------------------------------------------------------------------------- seeded demo inputs
def demo_pose_inputs(n_points=20000, n_traj=37, seed=77):
    """A cloud / trajectory / slab boundaries shaped like pose ~10 of a rollout (walls are dense vertical sheets)."""
    g = torch.Generator().manual_seed(seed)
    pose = torch.tensor([12.0, 1.8, -9.0, 0.0, 45.0])
    pc = torch.empty(n_points, 3)
    pc[:, 0] = pose[0] + torch.rand(n_points, generator=g) * 90 - 45
    pc[:, 1] = torch.rand(n_points, generator=g) * 10.5 - 1.6
    pc[:, 2] = pose[2] + torch.rand(n_points, generator=g) * 90 - 45
    k = n_points // 3
    pc[:k, 0] = torch.round(pc[:k, 0] / 7) * 7
    pc[k:2 * k, 2] = torch.round(pc[k:2 * k, 2] / 9) * 9
    pc[2 * k:2 * k + 500, 1] = pose[1] + (torch.rand(500, generator=g) - 0.5) * 0.3
    traj = pose[:3] + torch.cumsum(torch.randn(n_traj, 3, generator=g) * torch.tensor([0.75, 0.0, 0.75]), 0)
    lo, hi = -1.6 + 0.5, 8.9 - 0.5
    width = (hi - lo) / 4
    y_bins = torch.arange(lo, hi + width, width)                     # nbp_planning.py:449-451
    return pc, traj, pose, y_bins


def demo_replay_records(n=4, S=64, K=12, seed=31):
    """Replay records in the reference's msgpack layout (nbp_utils.py:676-683), as numpy arrays."""
    from . import nbp_torch as NT
    g = torch.Generator().manual_seed(seed)
    recs = []
    for i in range(n):
        x = NT.count_like_input(1, S, seed=seed + 1 + i)
        coords = torch.stack((torch.randint(0, 8, (K,), generator=g), torch.randint(0, S // 4, (K,), generator=g),
                              torch.randint(0, S // 4, (K,), generator=g)), dim=-1)
        recs.append({"current_model_input": x.numpy(), "current_gt_2d_layout": (torch.rand(1, 1, S, S, generator=g) < 0.2).float().numpy(),
                     "target_value_map_pixel": coords.numpy(), "actual_coverage_gain": (torch.rand(K, generator=g) * 10).numpy(),
                     "pose_i": 11 + i})
    return recs


class RecordingAdamW(torch.optim.AdamW):
    """AdamW that keeps a copy of the (already unscaled) gradients its step() consumes."""

    def step(self, closure=None):
        self.recorded = [p.grad.detach().clone() for grp in self.param_groups for p in grp["params"]]
        return super().step(closure)
