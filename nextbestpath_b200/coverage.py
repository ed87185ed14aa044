"""Batched surface-coverage metric.

``calculate_coverage_percentage(gt_scene_pc, full_pc)`` runs at the top of every pose iteration of both reference drivers
(next_best_path/testers/nbp_planning.py:71, next_best_path/utility/nbp_utils.py:572; definition
next_best_path/utility/long_term_utils.py:437-468) and builds a |GT| x |sample| ``torch.cdist`` matrix each time.  Here the ground
truth of every scene is bucketed once into a uniform grid (``CoverageIndex``, set-up) and one CUDA launch per step marks the
ground-truth points that have a reconstruction point within the threshold, for all scenes (``csrc/coverage.cu``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _lib

CELL_MARGIN = 1.0 + 2.0 ** -10        # cell edge slightly above the threshold: the +-1 cell neighbourhood then covers the radius
                                      # with room for the fp32 rounding of (p - origin) * (1/cell)


def _stream():
    return torch.cuda.current_stream().cuda_stream


class CoverageIndex:
    """Ground-truth surface clouds of B scenes bucketed into per-scene uniform grids (built once per scene set).

    Set-up only (not on the per-step path): cell keys and the cell-sorted order are computed with torch ops on the device."""

    def __init__(self, gt_clouds, device, threshold: float = 1.0):
        self.dev = torch.device(device)
        self.threshold = float(threshold)
        self.cell = float(np.float32(self.threshold * CELL_MARGIN))
        inv = torch.tensor(np.float32(1.0) / np.float32(self.cell), dtype=torch.float32, device=self.dev)
        pts, starts, origins, dims, goff, coff = [], [], [], [], [0], [0]
        for g in gt_clouds:
            g = torch.as_tensor(g, dtype=torch.float32).to(self.dev).reshape(-1, 3).contiguous()
            if g.shape[0] == 0:
                origins.append(torch.zeros(3, device=self.dev)); dims.append(torch.ones(3, dtype=torch.int32, device=self.dev))
                starts.append(torch.zeros(2, dtype=torch.int32, device=self.dev))
                goff.append(goff[-1]); coff.append(coff[-1] + 2)
                continue
            o = g.min(dim=0).values
            c = torch.floor((g - o) * inv).to(torch.int64)                        # same fp32 expression as the kernel
            d = c.max(dim=0).values + 1
            key = (c[:, 2] * d[1] + c[:, 1]) * d[0] + c[:, 0]
            order = torch.argsort(key, stable=True)
            ncell = int(d[0] * d[1] * d[2])
            st = torch.searchsorted(key[order].contiguous(), torch.arange(ncell + 1, device=self.dev, dtype=torch.int64)).to(torch.int32)
            pts.append(g[order]); starts.append(st); origins.append(o); dims.append(d.to(torch.int32))
            goff.append(goff[-1] + g.shape[0]); coff.append(coff[-1] + ncell + 1)
        self.B = len(gt_clouds)
        self.gt_sorted = torch.cat(pts).contiguous() if pts else torch.zeros((0, 3), device=self.dev)
        self.cell_start = torch.cat(starts).contiguous()
        self.origin = torch.stack(origins).contiguous()
        self.dims = torch.stack(dims).contiguous()
        self.gt_off = torch.tensor(goff, dtype=torch.int64, device=self.dev)
        self.cell_off = torch.tensor(coff, dtype=torch.int64, device=self.dev)
        self.gt_counts = [goff[i + 1] - goff[i] for i in range(self.B)]
        self.total_gt = goff[-1]
        self.covered = torch.empty(max(self.total_gt, 1), dtype=torch.uint8, device=self.dev)

    def coverage(self, cloud, cloud_len, weight: int = 2, seed: int = 0, sample_idx=None, max_points: int = -1, return_counts=False):
        """cloud (B, cap, 3) fp32, cloud_len (B,) int32 on the device.  Returns coverage (B,) fp32 on the device (no host sync).
        ``sample_idx`` (B, K >= weight*G_b) int64: explicit subset for the scenes whose cloud is longer than weight*G_b -- pass
        ``torch.randperm(len)[:k]`` to follow the reference's own random stream; default: a keyed permutation (``seed``)."""
        if not cloud.is_cuda:
            raise RuntimeError("coverage runs on CUDA tensors only (no CPU fallback; the oracle is oracle.oracle.coverage_percentage)")
        B = self.B
        assert cloud.shape[0] == B and cloud.shape[2] == 3 and cloud.dtype == torch.float32 and cloud.is_contiguous()
        cap = cloud.shape[1]
        kmax = min(cap if max_points < 0 else min(cap, max_points), weight * max(self.gt_counts + [0]))
        out = torch.empty(B, dtype=torch.float32, device=self.dev)
        counts = torch.empty(B, dtype=torch.int32, device=self.dev) if return_counts else None
        sidx = None
        if sample_idx is not None:                                                # used only for scenes with len > weight*G
            sidx = sample_idx.to(device=self.dev, dtype=torch.int64).contiguous()
            assert sidx.dim() == 2 and sidx.shape[0] == B
        rc = _lib.lib().nbp_coverage_percentage(cloud.data_ptr(), cap, cloud_len.data_ptr(), sidx.data_ptr() if sidx is not None else None,
                                                sidx.shape[1] if sidx is not None else 0, self.gt_sorted.data_ptr(), self.gt_off.data_ptr(),
                                                self.cell_start.data_ptr(), self.cell_off.data_ptr(), self.origin.data_ptr(), self.dims.data_ptr(),
                                                B, kmax, self.total_gt, self.cell, self.threshold, int(weight), int(seed) & (2 ** 64 - 1),
                                                self.covered.data_ptr(), out.data_ptr(), counts.data_ptr() if counts is not None else None, _stream())
        _lib.check(rc, "nbp_coverage_percentage")
        return (out, counts) if return_counts else out


_index_cache = {}


def calculate_coverage_percentage(pc1, pc2, threshold=1, weight=2):
    """Drop-in for next_best_path/utility/long_term_utils.py:457-468 (same arguments, returns a Python float): pc1 = ground-truth
    points (N,3), pc2 = reconstruction (M,3), CUDA tensors.  The sub-sampling draws ``torch.randperm(len(pc2))`` from the default
    CPU generator exactly like ``random_sample_pc`` (:436-446), so a seeded run follows the reference's random stream.
    The grid over pc1 is cached while the drivers pass the SAME tensor object, unmodified (identity + in-place version counter;
    the cache holds a reference, so the allocator cannot hand its address to another scene's cloud)."""
    if len(pc2) == 0:
        return 0.
    if not (isinstance(pc1, torch.Tensor) and pc1.is_cuda and isinstance(pc2, torch.Tensor) and pc2.is_cuda):
        raise RuntimeError("calculate_coverage_percentage: CUDA tensors only (no CPU fallback)")
    ent = _index_cache.get("gt")
    if ent is not None and ent[0] is pc1 and ent[1] == pc1._version and ent[2] == float(threshold):
        idx = ent[3]
    else:
        idx = CoverageIndex([pc1.detach().float()], pc1.device, threshold=float(threshold))
        _index_cache["gt"] = (pc1, pc1._version, float(threshold), idx)
    n2, want = pc2.shape[0], int(len(pc1) * weight)
    sample = None
    if n2 > want:
        sample = torch.randperm(n2)[:want].view(1, -1)                      # long_term_utils.py:445
    cloud = pc2.detach().float().contiguous().view(1, n2, 3)
    ln = torch.tensor([n2], dtype=torch.int32, device=pc2.device)
    return float(idx.coverage(cloud, ln, weight=int(weight), sample_idx=sample).item())
