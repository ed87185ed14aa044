"""Generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE (imported from /root/reference
with sys.modules stubs for its unavailable, unused imports).  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests there read the committed .npz files.
What is pinned:
  mapbuilder.npz : next_best_path/utility/utils.py  transform_points_to_n_pieces (:166-196),
                   map_points_to_n_imgs (:198-223), get_point_position_in_the_img (:160-164) and the
                   slab split expressions of next_best_path/testers/nbp_planning.py:114-115,446-451
  nbp_eval.npz   : next_best_path/networks/nbp_model.py NBP.forward in eval mode (S=128, B=1)
  coverage.npz   : next_best_path/utility/long_term_utils.py calculate_coverage_percentage (:436-468), executed from the reference tree
                   (``--coverage-only`` regenerates just this file)
  dropin.npz     : the driver lines nbp_planning.py:112-132,166-193 and nbp_utils.py:340-391 executed from the reference tree with the
                   reference's own functions bound (``--dropin-only`` regenerates just this file)
  nbp_train.npz  : NBP.forward (train mode) + NBP.loss + backward: loss, per-parameter gradient norms,
                   BN running statistics after the step (S=64, B=2)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def import_reference():
    for name in ["torchsummary", "matplotlib", "matplotlib.pyplot", "trimesh"]:
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.summary = lambda *a, **k: None
            sys.modules[name] = m
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path.insert(0, "/root/reference")
    from next_best_path.networks.nbp_model import NBP
    from next_best_path.utility import utils as ref_utils
    return NBP, ref_utils


def sparse(a):
    a = np.asarray(a)
    idx = np.flatnonzero(a)
    return idx.astype(np.int32), a.reshape(-1)[idx]


def _ref_lines(path, start_marker, end_marker, include_end=True):
    """Source lines of a reference file between two marker substrings (executed from where they lie, never copied)."""
    import textwrap
    lines = open(path).read().split("\n")
    a = next(i for i, l in enumerate(lines) if start_marker in l)
    b = next(i for i, l in enumerate(lines) if end_marker in l and i >= a)
    return textwrap.dedent("\n".join(lines[a: b + 1 if include_end else b]))


def planner_golden(ru):
    """planner.npz: the re-plan read-out.  The reference code is inline in compute_nbp_trajectory, so the
    lines next_best_path/testers/nbp_planning.py:166-233 and macarons_utils.py:86-100 are EXECUTED here from the reference tree
    (exec of the source text with a prepared namespace), with `nbp` stubbed to return fixed maps."""
    import ast
    g = torch.Generator().manual_seed(21)
    N = 5000
    pose = torch.tensor([4.0, 1.8, -6.5, 0.0, 90.0])
    full_pc = torch.empty(N, 3)
    full_pc[:, 0] = pose[0] + torch.rand(N, generator=g) * 70 - 35
    full_pc[:, 1] = torch.rand(N, generator=g) * 10 - 1
    full_pc[:, 2] = pose[2] + torch.rand(N, generator=g) * 70 - 35
    full_pc[:1500, 0] = torch.round(full_pc[:1500, 0] / 6) * 6                      # walls
    full_pc[1500:2500, 1] = pose[1] + (torch.rand(1000, generator=g) - 0.5) * 0.3    # many points near the camera-height slice
    value_map = torch.rand(1, 8, 64, 64, generator=g) * 10
    obstacle = torch.rand(1, 1, 256, 256, generator=g)
    traj = pose[:3] + torch.cumsum(torch.randn(30, 3, generator=g) * torch.tensor([1.5, 0.0, 1.5]), 0)
    ns = {"torch": torch, "ast": ast, "device": "cpu", "pc2img_size": (256, 256), "prediction_range": (-40, 40), "value_map_size": (64, 64),
          "transform_points_to_n_pieces": ru.transform_points_to_n_pieces, "map_points_to_n_imgs": ru.map_points_to_n_imgs,
          "get_point_position_in_the_img": ru.get_point_position_in_the_img, "full_pc": full_pc, "camera_current_pose": pose,
          "nbp": lambda x: (value_map, obstacle.clone()), "Dijkstra_path": [], "path_record": 0}
    exec(_ref_lines("/root/reference/macarons/utility/macarons_utils.py", "def check_pixel_values", "return contains_one"), ns)
    ns["current_pc_imgs"] = torch.zeros(1, 4, 256, 256)
    ns["current_previous_trajectory_img"] = ru.map_points_to_n_imgs(ru.transform_points_to_n_pieces(traj, pose, "cpu"), (256, 256), (-40, 40), "cpu").unsqueeze(0)
    keys, pts = [], []
    for i in range(-14, 15):
        for k in range(-14, 15):
            keys.append(str([i + 20, 0, k + 20])); pts.append(torch.tensor([pose[0] + 3.0 * i, pose[1], pose[2] + 3.0 * k]))
    ns["splited_pose_space"] = dict(zip(keys, pts))
    ns["collision_list"] = [[20, 0, 21], [25, 0, 25], [10, 0, 30]]
    exec(_ref_lines("/root/reference/next_best_path/testers/nbp_planning.py", "predicted_value_map, predicted_obstacle_map = nbp(",
                    "camera_position_value_list.sort("), ns)
    lst = ns["camera_position_value_list"]
    np.savez_compressed(os.path.join(HERE, "planner.npz"), cloud=full_pc.numpy(), pose=pose.numpy(), value_map=value_map[0].numpy(),
                        obstacle=obstacle[0, 0].numpy(), traj=traj.numpy(), cand=torch.stack(pts).numpy(), cand_keys=np.array(keys),
                        collision=np.array(ns["collision_list"]), fused=ns["predicted_obstacle_map"][0, 0].numpy().astype(np.uint8),
                        full_proj=ns["full_pc_projection"][0, 0].numpy().astype(np.uint8),
                        out_keys=np.array([e[0] for e in lst]), out_cell=np.array([[int(e[1][0]), int(e[1][1])] for e in lst]),
                        out_score=np.array([e[2] for e in lst], dtype=np.float64))
    print("planner golden:", len(lst), "valid candidates of", len(keys), "; fused ones", int(ns["predicted_obstacle_map"].sum()))


def coverage_golden():
    """coverage.npz: random_sample_pc / find_nearest_points_distances / calculate_coverage_percentage are
    EXECUTED from next_best_path/utility/long_term_utils.py:436-468 (the module itself cannot be imported: trimesh, pytorch3d)."""
    ns = {"torch": torch}
    exec(_ref_lines("/root/reference/next_best_path/utility/long_term_utils.py", "def random_sample_pc", "return similar_points.item()"), ns)
    ref = ns["calculate_coverage_percentage"]
    g = torch.Generator().manual_seed(33)
    # ground truth: points on the walls / floor of a 60 x 8 x 40 room; reconstruction: noisy samples of part of it
    def surface(n):
        u = torch.rand(n, 3, generator=g) * torch.tensor([60.0, 8.0, 40.0])
        face = torch.randint(0, 5, (n,), generator=g)
        u[face == 0, 0] = 0.0; u[face == 1, 0] = 60.0; u[face == 2, 2] = 0.0; u[face == 3, 2] = 40.0; u[face == 4, 1] = 0.0
        return u
    out = {}
    gt = surface(3000)
    for name, n_pc, part in (("sub", 9000, 0.6), ("all", 2500, 0.35)):
        pc = surface(n_pc)
        pc = pc[pc[:, 0] < 60.0 * part + 1e-3] if name == "all" else pc
        pc = pc[: (n_pc if name == "sub" else len(pc))]
        pc[:, 0] = pc[:, 0] * (part if name == "sub" else 1.0)
        pc = pc + torch.randn(pc.shape, generator=g) * 0.05
        torch.manual_seed(1234)
        val = ref(gt, pc)
        torch.manual_seed(1234)
        idx = torch.randperm(len(pc))[: 2 * len(gt)] if len(pc) > 2 * len(gt) else torch.zeros(0, dtype=torch.long)
        out[f"{name}_pc"] = pc.numpy(); out[f"{name}_idx"] = idx.numpy(); out[f"{name}_cov"] = np.float64(val)
    out["gt"] = gt.numpy()
    out["empty_cov"] = np.float64(ref(gt, torch.zeros(0, 3)))
    np.savez_compressed(os.path.join(HERE, "coverage.npz"), **out)
    print("coverage.npz:", {k: float(v) for k, v in out.items() if k.endswith("_cov")})


def dropin_golden(NBP, ru):
    """dropin.npz: the reference's OWN driver lines, executed from /root/reference and bound to the reference's OWN CPU functions
    (next_best_path/testers/nbp_planning.py:112-132,166-193 and next_best_path/utility/nbp_utils.py:340-391).  The GPU test binds
    the CUDA shims into the same call pattern (oracle/driver_lines.py, pinned to these lines by tests/test_dropin_lines.py)."""
    import random
    import types as _types
    from oracle import driver_lines as DL
    from oracle import nbp_torch as O
    out = {}
    # ---- planning lines
    pc, traj, pose, y_bins = DL.demo_pose_inputs()
    net = NBP(); net.load_state_dict(O.golden_state_dict(seed=9)); net.eval()
    raw = {}

    def nbp(x):
        with torch.no_grad():
            raw["v"], raw["o"] = net(x)
        return raw["v"], raw["o"].clone()

    ns = planning_namespace(ru, pc, traj, pose, y_bins, nbp, "cpu", (256, 256))
    exec_planning_lines(ns)
    out["model_input_idx"], out["model_input_val"] = sparse(torch.cat((ns["current_pc_imgs"], ns["current_previous_trajectory_img"]), 1).numpy())
    out["value_map"] = raw["v"].numpy(); out["raw_obstacle_map"] = raw["o"].numpy()
    out["fused"] = np.packbits(ns["predicted_obstacle_map"].numpy().astype(np.uint8))
    out["full_proj"] = np.packbits(ns["full_pc_projection"].numpy().astype(np.uint8))
    out["max_gain_map"] = ns["max_gain_map"].numpy()
    # ---- training lines
    recs = DL.demo_replay_records()
    net = NBP(); net.load_state_dict(O.golden_state_dict(seed=9)); net.train()
    opt = DL.RecordingAdamW(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    fn = exec_train_lines()
    random.seed(8)
    losses = fn(recs, _types.SimpleNamespace(nbp_batch_size=2), opt, net, "cpu", 2)
    out["train_losses"] = np.array(losses, dtype=np.float64)
    out["train_grad_names"] = np.array([n for n, _ in net.named_parameters()])
    out["train_grad_norms"] = np.array([float(g.double().norm()) for g in opt.recorded])
    out["train_rm_conv1"] = net.state_dict()["Conv1.conv.1.running_mean"].numpy()
    np.savez_compressed(os.path.join(HERE, "dropin.npz"), **out)
    print("dropin.npz: losses", losses, "fused ones", int(ns["predicted_obstacle_map"].sum()), "input points", float(out["model_input_val"].sum()))


def planning_namespace(ru, pc, traj, pose, y_bins, nbp, device, size):
    import types as _types
    return {"torch": torch, "device": device, "pc2img_size": size, "prediction_range": (-40, 40), "n_pieces": 4,
            "transform_points_to_n_pieces": ru.transform_points_to_n_pieces, "map_points_to_n_imgs": ru.map_points_to_n_imgs,
            "full_pc": pc.to(device), "y_bins": y_bins.to(device), "camera_current_pose": pose.to(device),
            "camera": _types.SimpleNamespace(X_cam_history=traj.to(device)), "nbp": nbp}


def exec_planning_lines(ns, root="/root/reference"):
    """nbp_planning.py:112-132 then :166-193, run from where they lie."""
    path = os.path.join(root, "next_best_path/testers/nbp_planning.py")
    exec(_ref_lines(path, "bins = torch.bucketize(full_pc[:, 1]", "current_previous_trajectory_img = previous_trajectory_img.unsqueeze(0)"), ns)
    exec(_ref_lines(path, "predicted_value_map, predicted_obstacle_map = nbp(", "max_gain_map, _ = torch.max("), ns)
    return ns


def exec_train_lines(root="/root/reference", grad_scaler=None):
    """def train_experience_data (nbp_utils.py:340-391), defined from the reference's own source text."""
    import random
    if grad_scaler is None:
        from torch.cuda.amp import GradScaler as grad_scaler                 # the reference's import (nbp_utils.py:14)
    ns = {"torch": torch, "np": np, "random": random, "GradScaler": grad_scaler}
    exec(_ref_lines(os.path.join(root, "next_best_path/utility/nbp_utils.py"), "def train_experience_data(", "return training_loss"), ns)
    return ns["train_experience_data"]


def main():
    NBP, ru = import_reference()
    from oracle import nbp_torch as O

    # ------------------------------------------------------------------ map builder
    g = torch.Generator().manual_seed(8)
    N = 6000
    pts = torch.empty(N, 3)
    pts[:, 0] = torch.rand(N, generator=g) * 120 - 60
    pts[:, 1] = torch.rand(N, generator=g) * 12 - 2
    pts[:, 2] = torch.rand(N, generator=g) * 120 - 60
    # some points exactly on cell boundaries / half-way cases / grid edges
    pts[:64, 0] = 3.0 + torch.arange(64) * 0.15625          # multiples of 1/6.4 -> .5 cases at S=512
    pts[:64, 2] = -7.0 + torch.arange(64) * 0.3125
    pose = torch.tensor([3.0, 1.8, -7.0, 0.0, 135.0])
    verts_y = torch.tensor([[0.0, -1.4853071, 0.0], [0.0, 8.0196469, 0.0]])
    min_y, max_y = torch.min(verts_y, dim=0)[0][1].item() + 0.5, torch.max(verts_y, dim=0)[0][1].item() - 0.5
    bin_width = (max_y - min_y) / 4
    y_bins = torch.arange(min_y, max_y + bin_width, bin_width)
    bins = torch.bucketize(pts[:, 1], y_bins[:-1]) - 1
    out = {"points": pts.numpy(), "pose": pose.numpy(), "y_bins": y_bins.numpy(), "bins": bins.numpy().astype(np.int8)}
    p2d = ru.transform_points_to_n_pieces(pts, pose, "cpu")
    out["p2d"] = p2d.numpy()
    for S in (128, 256, 512):
        imgs = []
        for i in range(4):
            grp = pts[bins == i]
            imgs.append(ru.map_points_to_n_imgs(ru.transform_points_to_n_pieces(grp, pose, "cpu"), (S, S), (-40, 40), "cpu"))
        img = torch.cat(imgs, 0).numpy()
        out[f"grid{S}_idx"], out[f"grid{S}_val"] = sparse(img)
        out[f"cells{S}"] = ru.get_point_position_in_the_img(p2d[0], (S, S), (-40, 40)).numpy()
    out["cells64"] = ru.get_point_position_in_the_img(p2d[0], (64, 64), (-40, 40)).numpy()
    np.savez_compressed(os.path.join(HERE, "mapbuilder.npz"), **out)

    # ------------------------------------------------------------------ NBP eval
    sd = O.golden_state_dict(seed=9)
    net = NBP()
    net.load_state_dict(sd)
    net.eval()
    x = O.count_like_input(1, 128, seed=3)
    with torch.no_grad():
        o1, o2 = net(x)
    xi, xv = sparse(x.numpy())
    np.savez_compressed(os.path.join(HERE, "nbp_eval.npz"), x_idx=xi, x_val=xv, x_shape=np.array(x.shape),
                        out1=o1.numpy(), out2=o2.numpy(),
                        sd_checksum=np.array([float(sum(v.double().abs().sum() for v in sd.values()))]))

    # ------------------------------------------------------------------ NBP train step
    net = NBP()
    net.load_state_dict(O.golden_state_dict(seed=9))
    net.train()
    xb = O.count_like_input(2, 64, seed=4)
    g2 = torch.Generator().manual_seed(5)
    K = 40
    tgt_idx = torch.stack((torch.randint(0, 8, (2, K), generator=g2), torch.randint(0, 16, (2, K), generator=g2),
                           torch.randint(0, 16, (2, K), generator=g2)), dim=-1)          # (B, K, 3) = (ch, gx, gy)
    tgt_val = torch.rand(2, K, generator=g2) * 10
    layout = (torch.rand(2, 1, 64, 64, generator=g2) < 0.2).float()
    p1, p2 = net(xb)
    pred = torch.stack([p1[b, tgt_idx[b, :, 0], tgt_idx[b, :, 1], tgt_idx[b, :, 2]] for b in range(2)])
    loss = net.loss(pred, tgt_val, p2, layout)
    loss.backward()
    names = [n for n, _ in net.named_parameters()]
    gn = np.array([float(p.grad.double().norm()) for _, p in net.named_parameters()])
    probe = {n: p.grad.reshape(-1)[:16].numpy().copy() for n, p in net.named_parameters()
             if n in ("Conv1.conv.0.weight", "Conv5.conv.3.weight", "Up_conv2_2.conv.3.weight", "Final1.weight",
                      "Att4_1.psi.0.weight", "log_vars")}
    st = net.state_dict()
    np.savez_compressed(os.path.join(HERE, "nbp_train.npz"), tgt_idx=tgt_idx.numpy(), tgt_val=tgt_val.numpy(),
                        layout=np.packbits(layout.numpy().astype(np.uint8)), loss=np.array([loss.item()]),
                        out1=p1.detach().numpy(), out2_mean=np.array([float(p2.mean())]),
                        grad_names=np.array(names), grad_norms=gn,
                        rm_conv1=st["Conv1.conv.1.running_mean"].numpy(), rv_conv1=st["Conv1.conv.1.running_var"].numpy(),
                        rm_up22=st["Up_conv2_2.conv.4.running_mean"].numpy(),
                        **{"probe_" + k.replace(".", "_"): v for k, v in probe.items()})
    planner_golden(ru)
    coverage_golden()
    dropin_golden(NBP, ru)
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__" and "--dropin-only" in sys.argv:
    dropin_golden(*import_reference())
    sys.exit(0)

if __name__ == "__main__" and "--coverage-only" in sys.argv:
    coverage_golden()
    sys.exit(0)

if __name__ == "__main__":
    main()
