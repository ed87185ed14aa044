#!/usr/bin/env python
"""BASELINE.json configs[4]: raster + occupancy-scatter microbench, 1k-100k triangles x 128-1024 envs, achieved HBM GB/s
against the measured peak (MEASURED_PEAKS.json).  Algorithmic bytes:
  raster       per view : F*36 B (3 verts x 3 fp32 per triangle) + H*W*4 B zbuf
  back-project per frame: H*W*4 B zbuf + k*12 B points (k = int(0.05*n_valid))
  grid scatter per env  : N*12 B points + 5*S*S*4 B grid
Timing: CUDA events, 3 warm-up + 5 timed launches, working sets larger than L2 except where noted."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nextbestpath_b200 import ops, synthetic as syn
from nextbestpath_b200.utility.camera import get_camera_RT

DEV = "cuda:0"
PEAK = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else 6650.0


def timed(fn, warm=3, it=5):
    for _ in range(warm): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(it): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / it


def main():
    ap = argparse.ArgumentParser(); ap.add_argument("--quick", action="store_true"); ap.add_argument("--only", default=""); a = ap.parse_args()
    H, W, S = 256, 456, 256
    rows = []
    tris_list = [1000, 3000, 10000, 30000, 100000] if not a.quick else [3000, 30000]
    env_list = [128, 256, 512, 1024] if not a.quick else [128, 512]
    # ---- raster + back-projection
    for F in (tris_list if a.only in ("", "raster") else []):
        base = [syn.make_scene(500 + i, tri_budget=F) for i in range(8)]          # 8 distinct meshes, reused round-robin
        for E in env_list:
            if F * E > 30_000_000:                                               # bound the workspace (2*F*E records of 96 B)
                continue
            scenes = [base[i % 8] for i in range(E)]
            verts = torch.from_numpy(np.concatenate([s.verts for s in scenes])).to(DEV)
            faces = torch.from_numpy(np.concatenate([s.faces for s in scenes]).astype(np.int32)).to(DEV)
            fc = [len(s.faces) for s in scenes]
            vo = torch.tensor(np.concatenate([[0], np.cumsum([len(s.verts) for s in scenes])]), dtype=torch.int64, device=DEV)
            fo = torch.tensor(np.concatenate([[0], np.cumsum(fc)]), dtype=torch.int64, device=DEV)
            poses = np.stack([syn.random_walk(s, 1, seed=i)[0][0] for i, s in enumerate(scenes)])
            R, T = get_camera_RT(torch.tensor(poses[:, :3]), torch.tensor(poses[:, 3:]))
            R, T = R.reshape(-1, 9).contiguous().to(DEV), T.contiguous().to(DEV)
            vs = torch.arange(E, dtype=torch.int32, device=DEV); vsh = list(range(E))
            z = torch.empty((E, H, W), device=DEV)
            ms = timed(lambda: ops.raster_depth(verts, faces, vo, fo, vs, R, T, H, W, fc, vsh, zbuf=z))
            nbytes = sum(fc) * 36 + E * H * W * 4
            rows.append({"kernel": "raster", "tris": int(np.mean(fc)), "envs": E, "ms": ms, "GBps": nbytes / ms / 1e6, "frac_of_peak": nbytes / ms / 1e6 / PEAK,
                         "views_per_s": E / ms * 1e3, "hit_fraction": float((z > -1).float().mean())})
            cloud = torch.empty((E, 8192, 3), device=DEV); cl = torch.zeros(E, dtype=torch.int32, device=DEV)
            fk = torch.zeros(E, dtype=torch.int32, device=DEV)

            def bp():
                cl.zero_()
                ops.backproject_append(z, R, T, vs, cloud, cl, fov_range=70.0, gathering_factor=0.05, seed=1, frame_kept=fk)
            ms = timed(bp)
            nbytes = E * H * W * 4 + int(fk.sum().item()) * 12
            rows.append({"kernel": "backproject+select", "tris": int(np.mean(fc)), "envs": E, "ms": ms, "GBps": nbytes / ms / 1e6, "frac_of_peak": nbytes / ms / 1e6 / PEAK})
            del verts, faces, z, cloud
            torch.cuda.empty_cache()
    # ---- grid scatter
    for N in (([30_000, 300_000, 1_500_000, 3_000_000] if not a.quick else [300_000]) if a.only in ("", "scatter") else []):
        for E in env_list:
            if N * E * 12 > 40e9:
                continue
            g = torch.Generator(device=DEV).manual_seed(N + E)
            cloud = torch.rand((E, N, 3), device=DEV, generator=g)
            cloud[..., 0] = cloud[..., 0] * 160 - 80; cloud[..., 2] = cloud[..., 2] * 160 - 80; cloud[..., 1] = cloud[..., 1] * 10 - 1
            cloud[:, : N // 2, 0] = torch.round(cloud[:, : N // 2, 0] / 12) * 12          # wall-like clustering -> atomic contention
            lens = torch.full((E,), N, dtype=torch.int32, device=DEV)
            pose = torch.zeros((E, 5), device=DEV)
            bounds = torch.tensor([[0.0, 2.0, 4.0, 6.0, 0, 0, 0, 0]], device=DEV).repeat(E, 1).contiguous()
            nb = torch.full((E,), 4, dtype=torch.int32, device=DEV)
            out = torch.empty((E, 5, S, S), device=DEV)
            ms = timed(lambda: ops.grid_scatter(cloud, lens, pose, bounds, nb, S, max_points=N, out=out))
            nbytes = E * (N * 12 + 5 * S * S * 4)
            rows.append({"kernel": "grid_scatter", "points": N, "envs": E, "ms": ms, "GBps": nbytes / ms / 1e6, "frac_of_peak": nbytes / ms / 1e6 / PEAK,
                         "binned_fraction": float(out[:, :4].sum() / (E * N))})
            del cloud, out
            torch.cuda.empty_cache()
    # ---- coverage metric: G ground-truth points per scene, clouds of N points (sample = 2*G)
    from nextbestpath_b200.coverage import CoverageIndex
    for (G, N) in (([(20000, 300_000), (20000, 1_500_000), (50000, 1_500_000)] if not a.quick else [(20000, 300_000)]) if a.only in ("", "coverage") else []):
        for E in ([64, 256] if not a.quick else [64]):
            base = [syn.make_scene(700 + i, "simple") for i in range(8)]
            gts = [syn.sample_surface(base[i % 8], G, seed=i) for i in range(E)]
            index = CoverageIndex(gts, DEV, threshold=1.0)
            cloud = torch.empty((E, N, 3), device=DEV)
            for i in range(E):                                   # reconstruction = noisy surface samples of a part of the scene
                sp = torch.from_numpy(syn.sample_surface(base[i % 8], 20000, seed=1000 + i)).to(DEV)
                sp = sp[sp[:, 0] < sp[:, 0].median()]
                rep = sp[torch.randint(0, len(sp), (N,), device=DEV)]
                cloud[i] = rep + torch.randn((N, 3), device=DEV) * 0.05
            lens = torch.full((E,), N, dtype=torch.int32, device=DEV)
            out = [None]
            ms = timed(lambda: out.__setitem__(0, index.coverage(cloud, lens, seed=3)))
            nbytes = E * (2 * G * 12 + G * 12 + G)                # sampled points + ground truth + flags
            tic = __import__("time").perf_counter()
            gt0, pc0 = torch.from_numpy(gts[0]), cloud[0].cpu()
            smp = pc0[torch.randperm(N)[: 2 * G]]
            cpu_cov = float((torch.cdist(gt0, smp).min(dim=1).values < 1.0).float().mean())      # the reference's expression, one scene, host cores
            cpu_ms = (__import__("time").perf_counter() - tic) * 1e3
            rows.append({"kernel": "coverage", "gt_points": G, "cloud_points": N, "envs": E, "ms": ms, "GBps": nbytes / ms / 1e6,
                         "frac_of_peak": nbytes / ms / 1e6 / PEAK, "coverage_scene0": float(out[0][0]), "cpu_cdist_scene0": cpu_cov,
                         "cpu_cdist_ms_per_scene": cpu_ms})
            del cloud
            torch.cuda.empty_cache()
    # ---- collision table: every lattice position x 8 horizontal neighbours, all scenes in one launch
    if a.only in ("", "collision"):
        from nextbestpath_b200.collision import MeshBatch
        for E in ([64, 256] if not a.quick else [32]):
            base = [syn.make_scene(800 + i, "simple") for i in range(8)]
            scenes = [base[i % 8] for i in range(E)]
            mb = MeshBatch([s.verts for s in scenes], [s.faces for s in scenes], DEV)
            segs, sid = [], []
            offs = np.array([[3, 0, 0], [-3, 0, 0], [0, 0, 3], [0, 0, -3], [3, 0, 3], [3, 0, -3], [-3, 0, 3], [-3, 0, -3]], np.float32)
            for i, s in enumerate(scenes):
                pos = syn.lattice_positions(s).astype(np.float32)
                sg = np.concatenate([np.repeat(pos, 8, 0), np.repeat(pos, 8, 0) + np.tile(offs, (len(pos), 1))], 1)
                segs.append(sg); sid.append(np.full(len(sg), i, np.int32))
            seg = torch.from_numpy(np.concatenate(segs)).to(DEV); sc = torch.from_numpy(np.concatenate(sid)).to(DEV)
            out = [None]
            ms = timed(lambda: out.__setitem__(0, mb.segments_hit(seg, sc)))
            tests = float(sum(len(sg) * len(s.faces) for sg, s in zip(segs, scenes)))
            rows.append({"kernel": "segments_hit_mesh", "envs": E, "segments": int(seg.shape[0]), "mean_faces": float(np.mean([len(s.faces) for s in scenes])),
                         "ms": ms, "ray_triangle_tests_per_s": tests / ms * 1e3, "blocked_fraction": float(out[0].float().mean()),
                         "parity": "tests/test_collision.py (bit-exact against the CPU restatement)"})
    print(json.dumps({"hbm_peak_GBps": PEAK, "rows": rows}))
    for r in rows:
        print("  ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}" for k, v in r.items()), file=sys.stderr)


if __name__ == "__main__":
    main()
