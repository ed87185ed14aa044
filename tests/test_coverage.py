"""the coverage metric calculate_coverage_percentage
(/root/reference/next_best_path/utility/long_term_utils.py:436-468).

CPU: the oracle against the fixture produced by executing the reference's own function (tests/golden/make_golden.py).
GPU: csrc/coverage.cu (grid-bucketed ground truth, all scenes in one launch) against the oracle: covered-point counts bit-exact
for explicit samples, the drop-in shim against the fixture under the reference's own random stream, the keyed-permutation
sample statistically."""
import os

import numpy as np
import pytest
import torch

from oracle import oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "coverage.npz"))


def _borderline(gt, pc, thr=1.0, tol=2e-3):
    """ground-truth points whose nearest-neighbour distance is within tol of the threshold (fp64): the only points on which
    torch.cdist's matmul expansion and an exact distance may disagree"""
    g, q = np.asarray(gt, np.float64), np.asarray(pc, np.float64)
    dmin = np.full(len(g), np.inf)
    for a in range(0, len(g), 512):
        d = np.sqrt(((g[a:a + 512, None, :] - q[None, :, :]) ** 2).sum(-1))
        dmin[a:a + 512] = d.min(1)
    return int((np.abs(dmin - thr) < tol).sum())


@pytest.mark.parametrize("name", ["sub", "all"])
def test_oracle_matches_reference_fixture(name):
    gt, pc, idx = GOLD["gt"], GOLD[f"{name}_pc"], GOLD[f"{name}_idx"]
    cov = O.coverage_percentage(gt, pc, 1.0, 2, indices=idx if len(idx) else None)
    sampled = pc[idx] if len(idx) else pc
    assert abs(cov - float(GOLD[f"{name}_cov"])) * len(gt) <= _borderline(gt, sampled) + 1e-6
    assert (len(idx) > 0) == (name == "sub")                       # "sub" exercises random_sample_pc, "all" the pass-through


def test_oracle_edge_cases():
    gt = GOLD["gt"]
    assert O.coverage_percentage(gt, np.zeros((0, 3), np.float32)) == 0.0 == float(GOLD["empty_cov"])
    assert O.coverage_percentage(gt[:10], gt[:10] + np.float32(0.25)) == 1.0
    assert O.coverage_percentage(gt[:10], gt[:10] + np.float32(100.0)) == 0.0
    # strict inequality at the threshold (long_term_utils.py:466 uses <)
    assert O.coverage_percentage(np.zeros((1, 3), np.float32), np.array([[1.0, 0.0, 0.0]], np.float32)) == 0.0


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="live reference only exists in the build container")
def test_oracle_against_live_reference_code():
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import make_golden as MG
    ns = {"torch": torch}
    exec(MG._ref_lines("/root/reference/next_best_path/utility/long_term_utils.py", "def random_sample_pc", "return similar_points.item()"), ns)
    g = torch.Generator().manual_seed(5)
    gt = torch.rand(1500, 3, generator=g) * torch.tensor([30.0, 6.0, 30.0])
    pc = torch.rand(4000, 3, generator=g) * torch.tensor([20.0, 6.0, 30.0])
    torch.manual_seed(77); ref = ns["calculate_coverage_percentage"](gt, pc)
    torch.manual_seed(77); idx = torch.randperm(len(pc))[: 2 * len(gt)]
    mine = O.coverage_percentage(gt.numpy(), pc.numpy(), 1.0, 2, indices=idx.numpy())
    assert abs(mine - ref) * len(gt) <= _borderline(gt.numpy(), pc.numpy()[idx.numpy()]) + 1e-6


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_coverage_counts_bit_exact_batched():
    from nextbestpath_b200.coverage import CoverageIndex
    dev = "cuda:0"
    g = torch.Generator().manual_seed(11)
    gts = [torch.from_numpy(GOLD["gt"]), torch.rand(777, 3, generator=g) * torch.tensor([25.0, 5.0, 18.0]) - 7.0,
           torch.rand(2000, 3, generator=g) * 40.0, torch.rand(50, 3, generator=g)]
    pcs = [torch.from_numpy(GOLD["sub_pc"]), torch.rand(5000, 3, generator=g) * torch.tensor([30.0, 7.0, 20.0]) - 9.0,
           torch.zeros(0, 3), torch.rand(60, 3, generator=g) * 3.0 - 1.0]                     # scene 2: empty cloud; scene 3: len <= 2G
    idxs = [torch.from_numpy(GOLD["sub_idx"]), torch.randperm(5000, generator=g)[: 2 * 777], None, None]
    B = len(gts)
    cap = max(len(p) for p in pcs) + 5
    cloud = torch.full((B, cap, 3), float("nan"))
    kmax = max(len(i) for i in idxs if i is not None)
    sidx = torch.zeros((B, kmax), dtype=torch.int64)
    for b in range(B):
        cloud[b, : len(pcs[b])] = pcs[b]
        if idxs[b] is not None:
            sidx[b, : len(idxs[b])] = idxs[b]
    ln = torch.tensor([len(p) for p in pcs], dtype=torch.int32)
    index = CoverageIndex(gts, dev, threshold=1.0)
    cov, cnt = index.coverage(cloud.to(dev), ln.to(dev), weight=2, sample_idx=sidx, return_counts=True)
    torch.cuda.synchronize()
    for b in range(B):
        ref, flags = O.coverage_percentage(gts[b].numpy(), pcs[b].numpy(), 1.0, 2, indices=idxs[b].numpy() if idxs[b] is not None else None, return_flags=True)
        assert int(cnt[b]) == int(flags.sum()), (b, int(cnt[b]), int(flags.sum()))
        assert float(cov[b]) == np.float32(ref)
    assert float(cov[0]) == np.float32(GOLD["sub_cov"]) and float(cov[2]) == 0.0


@pytest.mark.gpu
def test_cuda_coverage_dropin_follows_reference_random_stream():
    from nextbestpath_b200.coverage import calculate_coverage_percentage
    dev = "cuda:0"
    gt = torch.from_numpy(GOLD["gt"]).to(dev)
    for name in ("sub", "all"):
        torch.manual_seed(1234)                                    # the seed the fixture was generated under
        got = calculate_coverage_percentage(gt, torch.from_numpy(GOLD[f"{name}_pc"]).to(dev))
        assert isinstance(got, float) and got == float(np.float32(GOLD[f"{name}_cov"]))
    assert calculate_coverage_percentage(gt, torch.zeros(0, 3, device=dev)) == 0.0
    with pytest.raises(RuntimeError):
        calculate_coverage_percentage(gt.cpu(), gt.cpu())


@pytest.mark.gpu
def test_cuda_coverage_dropin_cache_is_not_keyed_on_the_address():
    """A new ground-truth cloud of the same shape -- possibly at the address the allocator just recycled -- or an in-place
    edit of the cached one must rebuild the index (the cache used to be keyed on data_ptr + shape)."""
    from nextbestpath_b200.coverage import calculate_coverage_percentage
    dev = "cuda:0"
    pc = torch.from_numpy(GOLD["all_pc"]).to(dev)
    gt_a = torch.from_numpy(GOLD["gt"]).to(dev)
    a = calculate_coverage_percentage(gt_a, pc)
    ptr = gt_a.data_ptr()
    del gt_a
    gt_b = torch.from_numpy(GOLD["gt"] + np.array([500.0, 0.0, 0.0], np.float32)).to(dev)      # far away: nothing is covered
    b = calculate_coverage_percentage(gt_b, pc)
    assert a > 0.05 and b == 0.0, (a, b, gt_b.data_ptr() == ptr)
    gt_b -= torch.tensor([500.0, 0.0, 0.0], device=dev)                                          # in-place edit of the cached tensor
    assert calculate_coverage_percentage(gt_b, pc) == a


@pytest.mark.gpu
def test_cuda_coverage_keyed_sample_is_a_uniform_subset():
    from nextbestpath_b200.coverage import CoverageIndex
    dev = "cuda:0"
    gt, pc = torch.from_numpy(GOLD["gt"]), torch.from_numpy(GOLD["sub_pc"])
    index = CoverageIndex([gt], dev)
    cloud = pc.view(1, -1, 3).contiguous().to(dev)
    ln = torch.tensor([len(pc)], dtype=torch.int32, device=dev)
    a = float(index.coverage(cloud, ln, seed=1)); b = float(index.coverage(cloud, ln, seed=1)); c = float(index.coverage(cloud, ln, seed=2))
    assert a == b                                                  # counter-based: reproducible
    full = O.coverage_percentage(gt.numpy(), pc.numpy(), 1.0, weight=10 ** 6)          # no sub-sampling: upper bound
    assert abs(a - float(GOLD["sub_cov"])) < 0.03 and abs(c - float(GOLD["sub_cov"])) < 0.03 and a <= full + 1e-6
