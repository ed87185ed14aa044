"""GPU: the batched rollout engine against a scene-by-scene CPU oracle rollout (same poses, gathering_factor 1 so
that the point sets are deterministic): model-input grids bit-exact, value maps within 1e-3."""
import numpy as np
import pytest
import torch

from nextbestpath_b200 import synthetic as syn
from nextbestpath_b200.networks import NBP
from nextbestpath_b200.rollout import RolloutEngine, interpolated_poses
from nextbestpath_b200.utility.camera import get_camera_RT
from oracle import nbp_torch as NT
from oracle import oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
H, W, S = 48, 86, 64


def _oracle_rollout(scene, poses, az, n_steps, sd, gf, sensor_range, S=S, H=H, W=W, sd64=None):
    """The reference loop (nbp_planning.py:60-355) restricted to the scoped stages, one scene, CPU."""
    verts, faces = scene.verts, scene.faces
    bounds = O.y_bins_from_verts(torch.from_numpy(verts)).numpy()[:-1]
    cloud = np.zeros((0, 3), np.float32)
    traj = [poses[0, :3]]
    R, T = O.camera_rt(torch.tensor(poses[:1, :3]), torch.tensor(poses[:1, 3:]))
    key = (O.render_depth(verts, faces, R[0].numpy(), T[0].numpy(), H, W)[0], R[0].numpy(), T[0].numpy())
    grids, outs = [], []
    for t in range(n_steps):
        cloud = np.concatenate([cloud, O.partial_point_cloud(key[0], key[1], key[2], sensor_range, gf)])
        grid = O.build_model_input(cloud, poses[t], bounds, np.stack(traj), S)
        grids.append(grid)
        with torch.no_grad():
            o1, o2 = NT.forward(sd, torch.from_numpy(grid)[None])
            _, t2 = NT.forward(sd64, torch.from_numpy(grid)[None].double()) if sd64 is not None else (None, o2)
        outs.append((o1[0], o2[0], t2[0]))
        frames = [key]
        for k in range(1, 5):
            X, V = O.interpolate_pose(poses[t], poses[t + 1], k, 4, 8, int(az[t]), int(az[t + 1]))
            Rk, Tk = O.camera_rt(X.view(1, 3), V.view(1, 2))
            frames.append((O.render_depth(verts, faces, Rk[0].numpy(), Tk[0].numpy(), H, W)[0], Rk[0].numpy(), Tk[0].numpy()))
            traj.append(X.numpy())
        for fr in frames[:4]:
            cloud = np.concatenate([cloud, O.partial_point_cloud(fr[0], fr[1], fr[2], sensor_range, gf)])
        key = frames[4]
    return grids, outs, cloud


def test_product_camera_rt_equals_oracle():
    g = torch.Generator().manual_seed(0)
    X = torch.rand(64, 3, generator=g) * 100 - 50
    V = torch.stack((torch.rand(64, generator=g) * 60 - 30, torch.randint(0, 8, (64,), generator=g) * 45.0), 1)
    R0, T0 = O.camera_rt(X, V)
    R1, T1 = get_camera_RT(X, V)
    assert torch.equal(R0, R1) and torch.equal(T0, T1)
    p0, p1 = torch.rand(5, 5, generator=g) * 10, torch.rand(5, 5, generator=g) * 10
    p0[:, 4] = torch.tensor([0.0, 315.0, 45.0, 0.0, 90.0]); p1[:, 4] = torch.tensor([315.0, 0.0, 90.0, 45.0, 90.0])
    a0, a1 = [0, 7, 1, 0, 2], [7, 0, 2, 1, 2]
    mine = interpolated_poses(p0, p1, a0, a1)
    for b in range(5):
        for k in range(1, 5):
            X_, V_ = O.interpolate_pose(p0[b], p1[b], k, 4, 8, a0[b], a1[b])
            assert torch.equal(mine[k - 1, b], torch.cat((X_, V_)))


@pytest.mark.parametrize("n_steps,B,S,H,W,tris,precision", [(3, 3, 64, 48, 86, 600, "mixed"), (3, 3, 64, 48, 86, 600, "fp16x2"),
                                                              (1, 2, 512, 64, 114, 6000, "mixed")])
def test_rollout_matches_oracle_rollout(n_steps, B, S, H, W, tris, precision):
    """Last case: BASELINE configs[3]-shaped step (512x512 grid, larger meshes; image reduced so the CPU oracle finishes).
    The 64x64 cases run with gathering_factor 1 (deterministic point sets): up to ~11 000 points per cell, 20-50x beyond the counts
    the BatchNorm statistics of the test weights were calibrated on -- an out-of-range stress of the number formats, run in both
    parity precisions."""
    scenes = [syn.make_scene(40 + i, tri_budget=tris + 400 * i) for i in range(B)]
    walks = [syn.random_walk(sc, n_steps + 1, seed=70 + i) for i, sc in enumerate(scenes)]
    poses = np.stack([w[0] for w in walks])          # (B, n_steps+1, 5)
    az = np.stack([w[1] for w in walks])
    sd = NT.golden_state_dict(seed=9)
    sd64 = {k: (v.double() if v.is_floating_point() else v) for k, v in sd.items()}
    net = NBP(); net.load_state_dict(sd); net.to(DEV).eval()
    net.precision = precision
    eng = RolloutEngine(scenes, net, DEV, S=S, H=H, W=W, max_steps=n_steps + 1, gathering_factor=1.0, sensor_range=30.0)
    eng.reset(poses[:, 0])
    got = []
    for t in range(n_steps):
        move = eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1])
        out = eng.step(move)
        got.append((out.model_input.cpu().numpy().copy(), out.value_map.cpu(), out.obstacle_map.cpu(), out.value_max.cpu()))
    torch.cuda.synchronize()
    assert eng.overflow.item() == 0
    n_sat = net.e4m3_saturation_count()
    print(f"precision {precision}: {n_sat} e4m3 saturation events")
    assert (n_sat > 0) == (precision == "mixed" and S == 64)          # out-of-range inputs are detected, in-range runs are clean
    lens = eng.cloud_len.cpu().numpy()
    for b in range(B):
        grids, outs, cloud = _oracle_rollout(scenes[b], poses[b], az[b], n_steps, sd, 1.0, 30.0, S, H, W, sd64)
        assert lens[b] == len(cloud)
        assert np.array_equal(eng.cloud[b, : lens[b]].cpu().numpy(), cloud)            # same points, same order
        for t in range(n_steps):
            assert np.array_equal(got[t][0][b], grids[t]), f"scene {b} step {t}: model input differs"
            o1, o2, o2_f64 = outs[t]
            e1 = (got[t][1][b] - o1).abs().max() / o1.abs().max()
            l2 = (got[t][2][b] - o2).norm() / o2.norm()
            # value map: max error relative to the map's maximum, and the obstacle map's l2-relative error: the 1e-3 parity bar.
            # obstacle map, worst single pixel (a sigmoid probability, thresholded at 0.13 by the planner): measured against the
            # float64 evaluation.  With gathering_factor 1 a grid cell here holds up to ~11 000 points (the reference's 5 % sampling:
            # a few hundred), logits overflow to +-inf, and the reference's own fp32 arithmetic is up to 3.3e-3 away from float64 on
            # single pixels -- the bar is 1e-3, or four times fp32's own error where that is larger.
            e2 = (got[t][2][b].double() - o2_f64).abs().max()
            e2_f32 = (o2.double() - o2_f64).abs().max()
            print(f"S={S} scene {b} step {t} (max count {int(grids[t].max())}): value map {float(e1):.2e}, obstacle map l2-rel {float(l2):.2e}, "
                  f"worst pixel vs fp64 {float(e2):.2e} (fp32 oracle vs fp64: {float(e2_f32):.2e})")
            # "mixed" carries ~15-bit operands in 33 of the 38 GEMM layers and its e4m3 planes cover |x| <= 3584: on the 64x64 cases
            # (counts 20-200x beyond the calibrated range) activations leave that window, the module REPORTS it
            # (e4m3_saturation_count() > 0, asserted above) and the obstacle map is only held to 5e-3 (l2) / 1e-1 (single worst pixel:
            # measured 4.8e-3 ... 5.0e-2 on B200 depending on the step, and moving by a few 1e-3 with the summation order of the
            # kernels -- saturated elements carry 11 bits); the value-map bar is unchanged.  In range (the 512-grid case: no saturation) every bar is the 1e-3 one; the fp16x2
            # run of the same out-of-range inputs meets them too.
            bar2, bar_l2 = (max(1e-3, 4.0 * float(e2_f32)), 1e-3) if n_sat == 0 else (1e-1, 5e-3)
            assert e1 <= 1e-3 and l2 <= bar_l2 and e2 <= bar2, (float(e1), float(e2), float(e2_f32), float(l2))
            assert torch.equal(got[t][3][b], got[t][1][b].amax(dim=0))
        assert grids[-1][:4].sum() > 1000 and grids[-1][4].sum() >= (9 if n_steps >= 3 else 1)


def test_rollout_subsampled_statistics():
    """gathering_factor 0.05 (reference default): clouds grow by exactly sum int(n*0.05) per frame and the run is
    reproducible for a fixed seed."""
    B, n_steps = 2, 2
    scenes = [syn.make_scene(50 + i, tri_budget=800) for i in range(B)]
    walks = [syn.random_walk(sc, n_steps + 1, seed=80 + i) for i, sc in enumerate(scenes)]
    poses = np.stack([w[0] for w in walks]); az = np.stack([w[1] for w in walks])
    net = NBP(); net.load_state_dict(NT.golden_state_dict(seed=9)); net.to(DEV).eval()
    runs = []
    for rep in range(2):
        eng = RolloutEngine(scenes, net, DEV, S=S, H=H, W=W, max_steps=n_steps + 1, gathering_factor=0.05, seed=3)
        eng.reset(poses[:, 0])
        for t in range(n_steps):
            out = eng.step(eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1]), run_network=(t == n_steps - 1))
        runs.append((eng.cloud_len.cpu().clone(), eng.cloud.cpu().clone(), out.model_input.cpu().clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][2], runs[1][2])
    n = runs[0][0]
    assert (n > 0).all() and (n <= n_steps * 5 * int(0.05 * H * W)).all()
    assert runs[0][2][:, :4].sum() > 0


def test_rollout_geometry_overlap_is_invisible():
    """Stages D/E on a side stream under the network (RolloutEngine.overlap_geometry) give the same clouds, grids and maps, bit
    for bit, as the single-stream order."""
    B, n_steps = 3, 3
    scenes = [syn.make_scene(65 + i, tri_budget=900) for i in range(B)]
    walks = [syn.random_walk(sc, n_steps + 1, seed=95 + i) for i, sc in enumerate(scenes)]
    poses = np.stack([w[0] for w in walks]); az = np.stack([w[1] for w in walks])
    net = NBP(); net.load_state_dict(NT.golden_state_dict(seed=9)); net.to(DEV).eval()
    runs = []
    for overlap in (True, False):
        eng = RolloutEngine(scenes, net, DEV, S=S, H=H, W=W, max_steps=n_steps + 1, gathering_factor=0.05, seed=4)
        eng.overlap_geometry = overlap
        eng.reset(poses[:, 0])
        outs = []
        for t in range(n_steps):
            o = eng.step(eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1]))
            outs.append((o.model_input.clone(), o.value_map.clone(), o.obstacle_map.clone()))
        torch.cuda.synchronize()
        runs.append((eng.cloud_len.cpu(), eng.cloud.cpu(), eng.frames.cpu(), outs))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][2], runs[1][2])
    for b in range(B):                                              # the cloud buffer is uninitialised beyond each scene's length
        n = int(runs[0][0][b])
        assert n > 0 and torch.equal(runs[0][1][b, :n], runs[1][1][b, :n])
    for a, b in zip(runs[0][3], runs[1][3]):
        assert all(torch.equal(x, y) for x, y in zip(a, b))


def test_rollout_host_outputs_overlap_copy():
    """step(..., host_out=...) reads the maps back on a side stream under stages D/E: the pinned host buffers hold the
    same bits as the device outputs once wait_host_outputs() returns, and the engine state is unaffected."""
    B, n_steps = 2, 2
    scenes = [syn.make_scene(60 + i, tri_budget=700) for i in range(B)]
    walks = [syn.random_walk(sc, n_steps + 1, seed=90 + i) for i, sc in enumerate(scenes)]
    poses = np.stack([w[0] for w in walks]); az = np.stack([w[1] for w in walks])
    net = NBP(); net.load_state_dict(NT.golden_state_dict(seed=9)); net.to(DEV).eval()
    h = (torch.empty((B, S // 4, S // 4)).pin_memory(), torch.empty((B, 8, S // 4, S // 4)).pin_memory(), torch.empty((B, 1, S, S)).pin_memory())
    runs = []
    for use_host in (False, True):
        eng = RolloutEngine(scenes, net, DEV, S=S, H=H, W=W, max_steps=n_steps + 1, gathering_factor=1.0, sensor_range=30.0)
        eng.reset(poses[:, 0])
        for t in range(n_steps):
            mv = eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1])
            out = eng.step(mv, host_out=h if use_host else None)
            if use_host:
                eng.wait_host_outputs()
                assert torch.equal(h[0], out.value_max.cpu()) and torch.equal(h[1], out.value_map.cpu()) and torch.equal(h[2], out.obstacle_map.cpu())
        torch.cuda.synchronize()
        runs.append((eng.cloud_len.cpu().clone(), out.value_map.cpu().clone()))
    assert torch.equal(runs[0][0], runs[1][0]) and torch.equal(runs[0][1], runs[1][1])


def test_reset_with_setup_approach_matches_reference_history():
    """reset(..., approach_pose=...) reproduces the 1+4 set-up positions of the reference's X_cam_history (scene.py:479-486):
    the trajectory channel of the first model input equals the oracle histogram of [neighbour, 3 interpolated, start]."""
    scene = syn.make_scene(77, tri_budget=700)
    poses, az = syn.random_walk(scene, 3, seed=77)
    net = NBP(); net.load_state_dict(NT.golden_state_dict(seed=9)); net.to(DEV).eval()
    eng = RolloutEngine([scene], net, DEV, S=S, H=H, W=W, max_steps=3, gathering_factor=1.0, sensor_range=30.0)
    # walk pose 0 -> pose 1 is the set-up approach; the rollout proper starts at pose 1
    eng.reset(poses[None, 1], approach_pose=poses[None, 0], approach_az=az[None, 0], start_az=az[None, 1])
    out = eng.step(eng.upload_move(poses[None, 1], poses[None, 2], az[None, 1], az[None, 2]))
    hist = [poses[0, :3]] + [O.interpolate_pose(poses[0], poses[1], k, 4, 8, int(az[0]), int(az[1]))[0].numpy() for k in range(1, 5)]
    assert np.array_equal(hist[-1], poses[1, :3])
    bounds = O.y_bins_from_verts(torch.from_numpy(scene.verts)).numpy()[:-1]
    R, T = O.camera_rt(torch.tensor(poses[1:2, :3]), torch.tensor(poses[1:2, 3:]))
    z = O.render_depth(scene.verts, scene.faces, R[0].numpy(), T[0].numpy(), H, W)[0]
    cloud = O.partial_point_cloud(z, R[0].numpy(), T[0].numpy(), 30.0, 1.0)
    grid = O.build_model_input(cloud, poses[1], bounds, np.stack(hist), S)
    assert np.array_equal(out.model_input[0].cpu().numpy(), grid)
    assert grid[4].sum() >= 2
