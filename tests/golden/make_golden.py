"""Generates tests/golden/*.npz by RUNNING THE REFERENCE'S OWN CODE (imported from /root/reference
with sys.modules stubs for its unavailable, unused imports).  Run in the build container only:

    python tests/golden/make_golden.py

The GPU box has no /root/reference; tests there read the committed .npz files.
What is pinned:
  mapbuilder.npz : next_best_path/utility/utils.py  transform_points_to_n_pieces (:166-196),
                   map_points_to_n_imgs (:198-223), get_point_position_in_the_img (:160-164) and the
                   slab split expressions of next_best_path/testers/nbp_planning.py:114-115,446-451
  nbp_eval.npz   : next_best_path/networks/nbp_model.py NBP.forward in eval mode (S=128, B=1)
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
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
