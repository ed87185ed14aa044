"""SURVEY.md section 8(f) row 1 -- obstacle-map fusion + candidate scoring (nbp_planning.py:166-233).
CPU: the oracle against the fixture produced by EXECUTING the reference's own lines (tests/golden/make_golden.py).
GPU: the CUDA kernels against the oracle and the fixture, bit-exact (0/1 maps, integer cells, fp32 values)."""
import ast
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O


@pytest.fixture(scope="module")
def gold(golden_dir):
    return np.load(os.path.join(golden_dir, "planner.npz"))


def _skip_mask(gold):
    coll = [list(c) for c in gold["collision"].tolist()]
    return np.array([ast.literal_eval(str(k)) in coll for k in gold["cand_keys"]])


def test_oracle_fusion_matches_reference(gold):
    fused, full_proj = O.fuse_obstacle_map(gold["obstacle"], gold["cloud"], gold["pose"], gold["traj"])
    assert np.array_equal(fused.astype(np.uint8), gold["fused"]) and np.array_equal(full_proj.astype(np.uint8), gold["full_proj"])
    assert 0 < gold["fused"].sum() < gold["fused"].size


def test_oracle_scoring_matches_reference(gold):
    valid, cell, score = O.score_candidates(gold["value_map"], gold["full_proj"].astype(np.float32), gold["cand"], gold["pose"], skip=_skip_mask(gold))
    keys = [str(k) for k in gold["cand_keys"]]
    got = {keys[j]: (tuple(cell[j]), score[j]) for j in range(len(keys)) if valid[j]}
    ref = {str(k): (tuple(c), s) for k, c, s in zip(gold["out_keys"], gold["out_cell"].tolist(), gold["out_score"])}
    assert got.keys() == ref.keys() and len(ref) > 100
    for k in ref:
        assert got[k][0] == ref[k][0] and got[k][1] == ref[k][1]            # same cell, bit-identical float64 score
    # the reference sorts by score, descending (stable): same order from our scores
    order = sorted(got, key=lambda k: got[k][1], reverse=True)
    assert [got[k][1] for k in order] == sorted(gold["out_score"], reverse=True)


@pytest.mark.gpu
def test_gpu_readout_matches_oracle_and_fixture(gold):
    from nextbestpath_b200 import planning
    DEV = "cuda:0"
    g = np.random.default_rng(3)
    B, S = 3, 256
    scenes = [{"cloud": gold["cloud"], "pose": gold["pose"], "traj": gold["traj"], "obs": gold["obstacle"], "vm": gold["value_map"],
               "cand": gold["cand"], "skip": _skip_mask(gold)}]
    for b in range(1, B):
        n = int(g.integers(500, 4000))
        pose = np.array([g.uniform(-20, 20), g.uniform(0, 3), g.uniform(-20, 20), 0, 45.0 * g.integers(8)], np.float32)
        cloud = np.stack([pose[0] + g.uniform(-45, 45, n), g.uniform(-1, 9, n), pose[2] + g.uniform(-45, 45, n)], 1).astype(np.float32)
        cloud[: n // 3, 1] = pose[1] + g.uniform(-0.15, 0.15, n // 3)
        cloud[: n // 6, 1] = np.float32(np.float32(float(pose[1]) + 0.1))              # exactly on the upper threshold: excluded
        m = int(g.integers(50, 300))
        cand = np.stack([pose[0] + 3.0 * g.integers(-15, 16, m), np.full(m, pose[1]), pose[2] + 3.0 * g.integers(-15, 16, m)], 1).astype(np.float32)
        scenes.append({"cloud": cloud, "pose": pose, "traj": pose[:3] + g.normal(0, 4, (20, 3)).astype(np.float32), "obs": g.uniform(0, 1, (S, S)).astype(np.float32),
                       "vm": g.uniform(0, 10, (8, 64, 64)).astype(np.float32), "cand": cand, "skip": g.uniform(size=m) < 0.05})
    cap = (max(len(s["cloud"]) for s in scenes) + 3) // 4 * 4
    M = max(len(s["cand"]) for s in scenes)
    cloud = torch.zeros((B, cap, 3)); lens = torch.zeros(B, dtype=torch.int32); traj = torch.zeros((B, 32, 3)); tl = torch.zeros(B, dtype=torch.int32)
    pose = torch.zeros((B, 5)); cand = torch.zeros((B, M, 3)); nc = torch.zeros(B, dtype=torch.int32); skip = torch.zeros((B, M), dtype=torch.uint8)
    obs = torch.zeros((B, 1, S, S)); vm = torch.zeros((B, 8, 64, 64))
    for b, s in enumerate(scenes):
        cloud[b, : len(s["cloud"])] = torch.from_numpy(s["cloud"]); lens[b] = len(s["cloud"])
        traj[b, : len(s["traj"])] = torch.from_numpy(s["traj"]); tl[b] = len(s["traj"])
        pose[b] = torch.from_numpy(s["pose"]); cand[b, : len(s["cand"])] = torch.from_numpy(s["cand"]); nc[b] = len(s["cand"])
        skip[b, : len(s["cand"])] = torch.from_numpy(s["skip"].astype(np.uint8)); obs[b, 0] = torch.from_numpy(s["obs"]); vm[b] = torch.from_numpy(s["vm"])
    d = lambda t: t.to(DEV)
    fused, full_proj = planning.fuse_obstacle_maps(d(obs), d(cloud), d(lens), d(pose), pose.numpy(), d(traj), d(tl), S=S)
    out = planning.score_candidates(d(vm), full_proj, d(cand), d(nc), d(pose), skip=d(skip))
    torch.cuda.synchronize()
    for b, s in enumerate(scenes):
        f_ref, p_ref = O.fuse_obstacle_map(s["obs"], s["cloud"], s["pose"], s["traj"], S)
        assert np.array_equal(fused[b, 0].cpu().numpy(), f_ref) and np.array_equal(full_proj[b].cpu().numpy(), p_ref)
        v_ref, c_ref, s_ref = O.score_candidates(s["vm"], p_ref, s["cand"], s["pose"], skip=s["skip"], S=S)
        m = len(s["cand"])
        assert np.array_equal(out["valid"][b, :m].cpu().numpy(), v_ref)
        assert not out["valid"][b, m:].any()
        assert np.array_equal(out["cell"][b, :m].cpu().numpy()[v_ref], c_ref[v_ref])
        assert np.array_equal(out["score"][b, :m].cpu().numpy()[v_ref], s_ref[v_ref])
    assert np.array_equal(fused[0, 0].cpu().numpy().astype(np.uint8), gold["fused"])          # the reference's own result
    assert int(out["valid"][0].sum()) == len(gold["out_keys"])
