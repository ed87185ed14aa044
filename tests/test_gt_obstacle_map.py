"""SURVEY.md section 8f row 3 (second half): the ground-truth obstacle map, get_binary_obstacle_array
(next_best_path/utility/utils.py:226-262).  Parity is UNPINNED against trimesh / matplotlib (absent from the reference tree and from
this image): the CPU tests pin the restated geometry (oracle/section_oracle.c) with known answers, the GPU tests require the CUDA
kernel to reproduce the oracle bit for bit."""
import numpy as np
import pytest
import torch

from nextbestpath_b200 import synthetic as syn
from oracle import oracle as O


def _box_room(x0, x1, z0, z1, y0=0.0, y1=10.0):
    """Four walls (two triangles each), no floor / ceiling."""
    c = [(x0, z0), (x1, z0), (x1, z1), (x0, z1)]
    verts, faces = [], []
    for i in range(4):
        (ax, az), (bx, bz) = c[i], c[(i + 1) % 4]
        b = len(verts)
        verts += [(ax, y0, az), (bx, y0, bz), (bx, y1, bz), (ax, y1, az)]
        faces += [(b, b + 1, b + 2), (b, b + 2, b + 3)]
    return np.array(verts, np.float32), np.array(faces, np.int64)


def test_oracle_section_known_answers():
    v, f = _box_room(-20.0, 10.0, -5.0, 30.0)
    pose = np.array([0.0, 3.3, 0.0, 0.0, 0.0], np.float32)
    m, seg = O.gt_obstacle_map(v, f, pose, return_segments=True)
    assert m.shape == (256, 256) and set(np.unique(m)) <= {0.0, 1.0}
    assert len(seg) == 8                                            # every wall triangle is cut once
    # continuous image coordinates: column u = (x_cam + 40 - x) * 3.2, row v = (z_cam + 40 - z) * 3.2; pixel k has its centre at k + 0.5.
    # walls x = 10, -20 -> u = 96, 192; z = 30, -5 -> v = 32, 144: the two pixels whose centres are 0.5 away (1.5 > half width 1.35)
    cols = np.flatnonzero(m[80].astype(bool))                       # a row that crosses only the two x-walls
    rows = np.flatnonzero(m[:, 150].astype(bool))
    assert set(cols) == {95, 96, 191, 192} and set(rows) == {31, 32, 143, 144}, (cols, rows)
    # the walls end at the corners (the round cap adds less than one pixel centre)
    assert m[31, 95:193].all() and not m[31, :94].any() and not m[31, 194:].any()
    # moving the camera moves the picture the other way (egocentric), by 3.2 px per unit
    m2 = O.gt_obstacle_map(v, f, np.array([5.0, 3.3, -2.5, 0.0, 0.0], np.float32))
    assert np.array_equal(np.roll(np.roll(m, 16, axis=1), -8, axis=0)[8:-8, 16:-16], m2[8:-8, 16:-16])
    # out of the 80 x 80 window -> empty; plane above the walls -> empty
    assert O.gt_obstacle_map(v, f, np.array([200.0, 3.3, 0.0, 0, 0], np.float32)).sum() == 0
    assert O.gt_obstacle_map(v, f, np.array([0.0, 11.0, 0.0, 0, 0], np.float32)).sum() == 0


def test_oracle_section_degenerate_faces():
    y = 2.0
    pose = np.array([0.0, y, 0.0, 0, 0], np.float32)
    verts = np.array([[-5, y, -5], [5, y, -5], [0, 7, -5],          # 0: an edge in the plane -> that edge
                      [-5, y, 5], [0, 6, 5], [3, 8, 5],             # 1: touches the plane at one vertex -> nothing
                      [-5, y, 10], [5, y, 10], [0, y, 15],          # 2: lies in the plane -> nothing
                      [10, y, 0], [10, 0, -4], [10, 4, 4]], np.float32)   # 3: one vertex on the plane, the others on opposite sides
    faces = np.arange(12, dtype=np.int64).reshape(4, 3)
    m, seg = O.gt_obstacle_map(verts, faces, pose, return_segments=True)
    assert len(seg) == 2
    assert np.allclose(sorted(seg[0].tolist()), sorted([-5.0, -5.0, 5.0, -5.0]))
    assert np.allclose(seg[1], [10.0, 0.0, 10.0, 0.0])             # vertex (10, y, 0) to the crossing of the opposite edge at z = 0
    assert m.sum() > 0


@pytest.mark.gpu
def test_cuda_gt_obstacle_maps_bit_exact_vs_oracle():
    from nextbestpath_b200.collision import MeshBatch, get_binary_obstacle_array
    dev = "cuda:0"
    scenes = [syn.make_scene(300 + i, tri_budget=1500 + 2500 * i) for i in range(3)]
    mesh = MeshBatch([s.verts for s in scenes], [s.faces for s in scenes], dev)
    poses, which = [], []
    for i, s in enumerate(scenes):
        p, _ = syn.random_walk(s, 4, seed=300 + i)
        for k in range(4):
            poses.append(p[k]); which.append(i)
    v, f = _box_room(-20.0, 10.0, -5.0, 30.0)
    room = MeshBatch([v], [f], dev)
    got = mesh.gt_obstacle_maps(np.stack(poses), map_scene=which)
    assert got.shape == (12, 1, 256, 256) and got.dtype == torch.float32
    tot = 0
    for j, (p, i) in enumerate(zip(poses, which)):
        want = O.gt_obstacle_map(scenes[i].verts, scenes[i].faces, p)
        assert np.array_equal(got[j, 0].cpu().numpy(), want), f"map {j} (scene {i}) differs from the oracle"
        tot += want.sum()
    assert tot > 5000
    for S, view in ((128, 80.0), (512, 80.0), (256, 40.0)):
        g = room.gt_obstacle_maps(np.array([[1.0, 3.3, 2.0, 0, 0]], np.float32), S=S, view_size=view)
        assert np.array_equal(g[0, 0].cpu().numpy(), O.gt_obstacle_map(v, f, [1.0, 3.3, 2.0], S=S, view=view))
    arr = get_binary_obstacle_array(room, torch.tensor([0.0, 3.3, 0.0, 0.0, 0.0]), view_size=80)      # the reference's call (nbp_utils.py:638)
    assert arr.shape == (256, 256) and arr.dtype.kind == "i" and np.array_equal(arr, O.gt_obstacle_map(v, f, [0.0, 3.3, 0.0]).astype(int))
