"""CPU: host-side logic of the section-8f modules that needs no GPU -- the ground-truth grid the coverage kernel walks
(coverage.CoverageIndex), mesh packing for the collision kernel (collision.MeshBatch), the surface sampler, and the
"no CPU fallback" contract of the product entry points."""
import numpy as np
import pytest
import torch

from nextbestpath_b200 import synthetic as syn
from nextbestpath_b200.collision import MeshBatch
from nextbestpath_b200.coverage import CELL_MARGIN, CoverageIndex, calculate_coverage_percentage


def test_coverage_index_grid_is_consistent():
    g = torch.Generator().manual_seed(4)
    gts = [torch.rand(3000, 3, generator=g) * torch.tensor([40.0, 8.0, 25.0]) - 12.0, torch.rand(17, 3, generator=g), torch.zeros(0, 3)]
    idx = CoverageIndex(gts, "cpu", threshold=1.0)
    assert idx.B == 3 and idx.gt_counts == [3000, 17, 0] and idx.total_gt == 3017
    assert idx.cell >= 1.0 and abs(idx.cell - CELL_MARGIN) < 1e-6
    inv = np.float32(1.0) / np.float32(idx.cell)
    for b in range(2):
        g0, g1 = int(idx.gt_off[b]), int(idx.gt_off[b + 1])
        pts = idx.gt_sorted[g0:g1]
        # same multiset of points as the input
        assert torch.equal(torch.sort(pts.flatten()).values, torch.sort(gts[b].flatten()).values)
        nx, ny, nz = [int(v) for v in idx.dims[b]]
        cs = idx.cell_start[int(idx.cell_off[b]): int(idx.cell_off[b + 1])]
        assert len(cs) == nx * ny * nz + 1 and int(cs[0]) == 0 and int(cs[-1]) == g1 - g0 and bool((cs[1:] >= cs[:-1]).all())
        # every point sits in the cell range its own coordinates select, with the kernel's fp32 expression
        c = torch.floor((pts - idx.origin[b]) * torch.tensor(inv)).long()
        assert bool((c >= 0).all()) and bool((c[:, 0] < nx).all()) and bool((c[:, 1] < ny).all()) and bool((c[:, 2] < nz).all())
        key = (c[:, 2] * ny + c[:, 1]) * nx + c[:, 0]
        assert bool((key[1:] >= key[:-1]).all())                       # cell-sorted
        pos = torch.arange(g1 - g0)
        assert bool((cs[key].long() <= pos).all()) and bool((pos < cs[key + 1].long()).all())
    # the empty scene owns a 1-cell grid with no points
    assert int(idx.cell_off[3] - idx.cell_off[2]) == 2


def test_mesh_batch_packing_and_surface_sampler():
    scenes = [syn.make_scene(5, tri_budget=600), syn.make_scene(6, tri_budget=900)]
    mb = MeshBatch([s.verts for s in scenes], [s.faces for s in scenes], "cpu")
    assert mb.n_scenes == 2 and mb.verts.dtype == torch.float32 and mb.faces.dtype == torch.int32
    assert mb.vert_off.tolist() == [0, len(scenes[0].verts), len(scenes[0].verts) + len(scenes[1].verts)]
    assert mb.face_off.tolist() == [0, len(scenes[0].faces), len(scenes[0].faces) + len(scenes[1].faces)]
    assert int(mb.faces[: len(scenes[0].faces)].max()) < len(scenes[0].verts)            # scene-local indices
    p = syn.sample_surface(scenes[0], 2000, seed=1)
    assert p.shape == (2000, 3) and p.dtype == np.float32
    lo, hi = scenes[0].verts.min(0) - 1e-4, scenes[0].verts.max(0) + 1e-4
    assert ((p >= lo) & (p <= hi)).all()
    assert np.array_equal(p, syn.sample_surface(scenes[0], 2000, seed=1)) and not np.array_equal(p, syn.sample_surface(scenes[0], 2000, seed=2))


def test_product_paths_refuse_cpu_tensors():
    gt = torch.rand(50, 3)
    with pytest.raises(RuntimeError):
        calculate_coverage_percentage(gt, gt)                          # no CPU fallback (the oracle is test infrastructure)
    assert calculate_coverage_percentage(gt, torch.zeros(0, 3)) == 0.   # the reference's early return needs no device
    idx = CoverageIndex([gt], "cpu")
    with pytest.raises(RuntimeError):
        idx.coverage(gt.view(1, -1, 3), torch.tensor([50], dtype=torch.int32))
    from nextbestpath_b200.networks import NBP
    with pytest.raises(RuntimeError):
        NBP()(torch.zeros(1, 5, 32, 32))


def test_package_seeded_weights_and_inputs_equal_the_oracles():
    """bench.py's GPU arm builds its weights from the package (no oracle import on the product path); the recipe is the
    oracle's, so the CPU baseline (oracle weights) and the GPU arm run the same network."""
    import torch
    from nextbestpath_b200 import synthetic as syn
    from nextbestpath_b200.networks import NBP
    from oracle import nbp_torch as NT
    net = NBP()
    a, b = syn.seeded_nbp_state_dict(net, 9), NT.seeded_state_dict(9)
    assert list(a) == list(b) == list(net.state_dict())
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.equal(syn.count_like_input(2, 32, seed=5), NT.count_like_input(2, 32, seed=5))
