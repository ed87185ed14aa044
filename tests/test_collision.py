"""segment / ray vs mesh tests (line_segment_mesh_intersection macarons_utils.py:120-151,
check_camera_in_mesh long_term_utils.py:158-170).  Trimesh is absent, so the pin is the oracle restatement (PARITY UNPINNED
against Trimesh itself): CPU known answers for the oracle, bit-exact agreement of the CUDA kernel with it."""
import numpy as np
import pytest
import torch

from nextbestpath_b200 import synthetic as syn
from oracle import oracle as O


def _box(lo, hi):
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    v = np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])], np.float32)
    f = np.array([[0, 1, 3], [0, 3, 2], [4, 6, 7], [4, 7, 5], [0, 4, 5], [0, 5, 1], [2, 3, 7], [2, 7, 6], [0, 2, 6], [0, 6, 4], [1, 5, 7], [1, 7, 3]], np.int32)
    return v, f


def test_oracle_known_answers():
    v, f = _box([0, 0, 0], [10, 4, 6])
    # (a generic interior point: rays through the centre of a wall would graze the diagonal its two triangles share and count twice,
    #  which Trimesh's multiple-hit query does as well)
    seg = np.array([[5.3, 1.1, 2.2, 15, 1.1, 2.2],       # from inside through the x = 10 wall
                    [5.3, 1.1, 2.2, 9, 1.1, 2.2],        # stays inside
                    [5.3, 1.1, 2.2, 9.999, 1.1, 2.2],    # stops just short of the wall
                    [-5, 1.1, 2.2, -1, 1.1, 2.2],        # outside, pointing at the box but too short
                    [-5, 1.1, 2.2, 20, 1.1, 2.2],        # through both walls
                    [5.3, 10, 2.2, 15, 10, 2.2]],        # above the box, parallel to its top
                   np.float32)
    hit, cnt = O.segment_mesh_hits(v, f, seg)
    assert hit.tolist() == [True, False, False, False, True, False]
    assert cnt[0] == 1 and cnt[4] == 2
    # inside test: three axis rays from an interior point cross the surface once each; from outside 0 or 2 times
    org = np.array([[5.3, 1.1, 2.2]] * 3 + [[20, 1.1, 2.2]] * 3, np.float32)
    d = np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1]] * 2, np.float32)
    _, c = O.segment_mesh_hits(v, f, np.concatenate([org, d], 1), rays=True)
    assert (c[:3] % 2 == 1).all() and (c[3:] % 2 == 0).all()
    # hits behind the origin do not count
    _, c = O.segment_mesh_hits(v, f, np.array([[20, 1.1, 2.2, 1, 0, 0]], np.float32), rays=True)
    assert c[0] == 0


@pytest.mark.gpu
def test_cuda_segments_bit_exact_vs_oracle():
    from nextbestpath_b200.collision import MeshBatch, check_camera_in_mesh, line_segment_mesh_intersection
    dev = "cuda:0"
    scenes = [syn.make_scene(90, tri_budget=900), syn.make_scene(91, tri_budget=2500)]
    box_v, box_f = _box([0, 0, 0], [10, 4, 6])
    vl = [s.verts for s in scenes] + [box_v]; fl = [s.faces for s in scenes] + [box_f]
    mb = MeshBatch(vl, fl, dev)
    rng = np.random.default_rng(3)
    segs, sid = [], []
    for si in range(3):
        lo, hi = vl[si].min(0), vl[si].max(0)
        a = rng.uniform(lo, hi, (150, 3)); b = a + rng.normal(0, 1, (150, 3)) * rng.choice([1.0, 6.0, 30.0], (150, 1))
        segs.append(np.concatenate([a, b], 1).astype(np.float32)); sid += [si] * 150
    # lattice-like moves: 3-unit axis steps at camera height, as the planner's neighbour expansion produces
    for si in range(2):
        pos = syn.lattice_positions(scenes[si])[:120]
        step = np.array([[3.0, 0, 0], [0, 0, 3.0], [-3.0, 0, 0], [0, 0, -3.0]])[rng.integers(0, 4, len(pos))]
        segs.append(np.concatenate([pos, pos + step], 1).astype(np.float32)); sid += [si] * len(pos)
    seg = np.concatenate(segs)
    hit, cnt = mb.segments_hit(seg, torch.tensor(sid, dtype=torch.int32), return_counts=True)
    hit, cnt = hit.cpu().numpy(), cnt.cpu().numpy()
    o = 0
    for s_arr, si in zip(segs, [0, 1, 2, 0, 1]):
        h, c = O.segment_mesh_hits(vl[si], fl[si], s_arr)
        assert np.array_equal(hit[o:o + len(s_arr)], h) and np.array_equal(cnt[o:o + len(s_arr)], c)
        o += len(s_arr)
    assert hit.any() and not hit.all()
    # any-hit only (no counts) agrees with the counting run
    assert np.array_equal(mb.segments_hit(seg, torch.tensor(sid, dtype=torch.int32)).cpu().numpy(), hit)
    # drop-ins
    box = MeshBatch([box_v], [box_f], dev)
    assert line_segment_mesh_intersection(torch.tensor([5.3, 1.1, 2.2]), torch.tensor([15.0, 1.1, 2.2]), box) is True
    assert line_segment_mesh_intersection([5.3, 1.1, 2.2], [9.0, 1.1, 2.2], box) is False
    assert check_camera_in_mesh(box, torch.tensor([5.3, 1.1, 2.2])) is True and check_camera_in_mesh(box, torch.tensor([50.0, 1.1, 2.2])) is False
    # neighbour table of a small lattice
    pos = torch.tensor([[5.3, 1.1, 2.2], [9.0, 1.1, 2.2], [15.0, 1.1, 2.2]])
    nb = torch.tensor([[1, 2], [0, 2], [0, -1]])
    tab = box.neighbour_collision_table(pos, nb).cpu()
    assert tab.tolist() == [[False, True], [False, True], [True, False]]
