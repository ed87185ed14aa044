#!/usr/bin/env python
"""bench.py -- rollout steps/s of the B200-native NextBestPath exploration inner loop.

    python bench.py --gpus N --steps K --warmup W            # this framework (N>1: launched under torchrun)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path on the host cores

Metric (BASELINE.json): rollout scene-steps per second at 256 parallel AiMDoom-simple-shaped scenes, 256x256 grid
(configs[1]).  One "step" advances ALL scenes by one pose: per scene 1+4 back-projections, a re-binning of the whole
accumulated cloud into the 5x256x256 model input, one NBP forward (eval), 4 depth renders.
The 256 scenes are sharded across the N ranks with no data-path collective ("scaling": "strong": total work fixed).

The timed rollout starts at pose `--prefill` (default 50 = mid-rollout, the mean cloud size of a 100-pose rollout,
~1.46 M points per scene): the first `prefill` poses are advanced geometry-only before timing starts.

value  : inputs (meshes, trajectory cameras) resident in HBM, CUDA-event timed, max over ranks
e2e    : same steps driven from HOST buffers: per step the camera poses are interpolated on the host, (R,T) copied
         from pinned memory, and the value maps / obstacle maps are read back to the host inside the timed region
roofline: the dominant kernel (conv_gemm_f16, tcgen05) timed live with CUDA events around every launch
cpu_baseline: the oracle (CPU restatement of the reference path; the reference's own files cannot run here, see
         DESIGN.md) timed on this box's host cores on a bounded sample: whole steps of ONE scene at the same state
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# stdout carries exactly ONE line (the JSON): the real stdout is kept aside and fd 1 is pointed at stderr, so that library
# banners (e.g. "NCCL version ..." printed on communicator creation) cannot end up in front of the result line
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


METRIC = "rollout_scene_steps_per_sec"
UNIT = "scene-steps/s"
FLOP_PER_SCENE_STEP = {128: 45.603e9, 256: 182.411e9, 512: 729.645e9}     # conv 2*MAC, SURVEY.md section 6


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "bf16": d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1590.0)),
                "which": "measured (MEASURED_PEAKS.json, bf16 sustained)"}
    return {"hbm_gbs": 6650.0, "bf16": 1590.0, "which": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown," \
            "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 9 for i in range(4) if r[5 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ workload
def make_workload(n_scenes, level, n_poses, rank0_scene_index=0, seed0=1000):
    from nextbestpath_b200 import synthetic as syn
    scenes = [syn.make_scene(seed0 + rank0_scene_index + i, level) for i in range(n_scenes)]
    walks = [syn.random_walk(sc, n_poses, seed=seed0 + rank0_scene_index + i) for i, sc in enumerate(scenes)]
    poses = np.stack([w[0] for w in walks])       # (B, n_poses, 5)
    az = np.stack([w[1] for w in walks])
    return scenes, poses, az


# ------------------------------------------------------------------------------------------------ CPU leg
def cpu_steps(scene, poses, az, cloud0, traj0, t0, n_steps, S, sd, threads, H=256, W=456, gf=0.05, sensor_range=70.0):
    """Whole rollout steps of ONE scene on the host cores with the oracle (the reference path restated):
    naive PyTorch3D-style rasteriser (rows split over `threads` POSIX threads), numpy un-projection and histogram,
    torch fp32 NBP forward with `threads` intra-op threads.  Returns seconds per step (list)."""
    from oracle import nbp_torch as NT
    from oracle import oracle as O
    torch.set_num_threads(threads)
    bounds = O.y_bins_from_verts(torch.from_numpy(scene.verts)).numpy()[:-1]
    cloud = cloud0.copy()
    traj = [p for p in traj0]
    g = torch.Generator().manual_seed(9)
    cam = lambda X, V: [a[0].numpy() for a in O.camera_rt(torch.as_tensor(X).view(1, 3), torch.as_tensor(V).view(1, 2))]
    R, T = cam(poses[t0, :3], poses[t0, 3:])
    key = (O.render_depth(scene.verts, scene.faces, R, T, H, W, nthreads=threads)[0], R, T)

    def part(fr):
        z = fr[0]
        n = int(((z > -1) & (z < sensor_range)).sum())
        idx = torch.randperm(n, generator=g)[: int(n * gf)].numpy()          # macarons_utils.py:2836-2838
        return O.partial_point_cloud(z, fr[1], fr[2], sensor_range, gf, indices=idx)

    times = []
    for t in range(t0, t0 + n_steps):
        tic = time.perf_counter()
        cloud = np.concatenate([cloud, part(key)])
        grid = O.build_model_input(cloud, poses[t], bounds, np.stack(traj), S)
        with torch.no_grad():
            o1, o2 = NT.forward(sd, torch.from_numpy(grid)[None])
        o1.amax(dim=1)
        frames = [key]
        for k in range(1, 5):
            X, V = O.interpolate_pose(poses[t], poses[t + 1], k, 4, 8, int(az[t]), int(az[t + 1]))
            R, T = cam(X, V)
            frames.append((O.render_depth(scene.verts, scene.faces, R, T, H, W, nthreads=threads)[0], R, T))
            traj.append(X.numpy())
        for fr in frames[:4]:
            cloud = np.concatenate([cloud, part(fr)])
        key = frames[4]
        times.append(time.perf_counter() - tic)
    return times


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the reference's PyTorch3D/Trimesh
    code cannot run in this image) on all host threads.  Each step = one scene-step of the same workload/config."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import nbp_torch as NT
    from oracle import oracle as O
    O.build()
    threads = os.cpu_count() or 1
    S = args.grid
    scenes, poses, az = make_workload(1, args.level, args.prefill + args.warmup + args.steps + 2)
    sd = NT.golden_state_dict(seed=9)
    # state at pose `prefill`: a cloud of the size the GPU arm starts from (uniform in the footprint)
    rng = np.random.default_rng(0)
    n0 = args.prefill * 5 * int(0.05 * 256 * 456)
    sc = scenes[0]
    lo, hi = sc.verts.min(0), sc.verts.max(0)
    cloud0 = (rng.uniform(0, 1, (n0, 3)) * (hi - lo) + lo).astype(np.float32)
    traj0 = [poses[0, min(i, args.prefill), :3] for i in range(1 + 4 * args.prefill)]
    times = cpu_steps(sc, poses[0], az[0], cloud0, traj0, args.prefill, args.warmup + args.steps, S, sd, threads)
    timed = times[args.warmup:]
    ms = 1e3 * float(np.mean(timed))
    value = 1e3 / ms
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.scenes} parallel AiMDoom-{args.level}-shaped scenes, {S}x{S} grid, inference rollout "
                                   f"(BASELINE configs[1]); reference arm = one scene per step, sequentially, as nbp_planning.py:395",
                       "image": "256x456", "prefill_pose": args.prefill, "mesh_level": args.level},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.steps} whole scene-steps of 1 scene at pose {args.prefill} (cloud {n0} pts), "
                                       f"after {args.warmup} warm-up"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "oracle port of the reference CPU path: PyTorch3D 0.7.4 naive rasteriser restated in C, numpy un-projection "
                    "and histogram, torch fp32 NBP; pytorch3d/trimesh are not installable here (DESIGN.md)"}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU leg
def run_ours(args):
    import torch.distributed as dist
    from nextbestpath_b200 import _lib, ops
    from nextbestpath_b200.networks import NBP
    from nextbestpath_b200.rollout import RolloutEngine, shard_scenes
    from oracle import nbp_torch as NT            # only: golden weights + the cpu_baseline leg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()

    S = args.grid
    B_total = args.scenes
    per = [shard_scenes(B_total, world, r)[1] for r in range(world)]
    first, B = shard_scenes(B_total, world, rank)
    n_total = args.prefill + 2 * (args.warmup + args.steps) + 3
    scenes, poses, az = make_workload(B, args.level, n_total + 8, rank0_scene_index=first)
    sd = NT.golden_state_dict(seed=9)
    net = NBP(); net.load_state_dict(sd); net.to(dev).eval()
    net.precision = args.precision
    net.max_chunk = args.chunk
    eng = RolloutEngine(scenes, net, dev, S=S, max_steps=n_total + 1, seed=9)
    eng.reset(poses[:, 0])
    t = 0
    for _ in range(args.prefill):                                   # geometry-only fast-forward to pose `prefill`
        eng.step(eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1]), run_network=False)
        t += 1
    torch.cuda.synchronize()
    cloud_pts = eng.cloud_len.float().mean().item()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from nextbestpath_b200.rollout import max_over_ranks as _mor
    max_over_ranks = lambda x: _mor(x, dev)

    # ================= value: trajectory cameras resident in HBM
    n_run = args.warmup + args.steps
    moves = [eng.upload_move(poses[:, t + i], poses[:, t + i + 1], az[:, t + i], az[:, t + i + 1]) for i in range(n_run)]
    torch.cuda.synchronize()
    for i in range(args.warmup):
        eng.step(moves[i])
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    launches0 = ops.launch_count()
    _lib.check(L.nbp_conv_profile_begin(60000), "nbp_conv_profile_begin")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, n_run):
        eng.step(moves[i])
    e1.record()
    barrier()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count() - launches0
    cms, cfl, cn, cdrop = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint64(), ctypes.c_uint64()
    _lib.check(L.nbp_conv_profile_end(ctypes.byref(cms), ctypes.byref(cfl), ctypes.byref(cn), ctypes.byref(cdrop)), "nbp_conv_profile_end")
    clocks = sampler.stop() if rank == 0 else None
    t += n_run
    ms_per_step = ms_total / args.steps
    value = B_total * args.steps / (ms_total / 1e3)

    # ================= per-stage device time of one extra step (diagnostic, outside both timed regions)
    from nextbestpath_b200.rollout import STAGE_NAMES
    eng.stage_events = []
    eng.step(eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1]))
    torch.cuda.synchronize()
    evs = eng.stage_events
    eng.stage_events = None
    stage_ms = {nm: evs[i].elapsed_time(evs[i + 1]) for i, nm in enumerate(STAGE_NAMES)}
    t += 1

    # ================= e2e: host-driven steps (host pose interpolation, pinned H2D, D2H of the maps)
    h_val = torch.empty((B, S // 4, S // 4), dtype=torch.float32).pin_memory()
    h_map8 = torch.empty((B, 8, S // 4, S // 4), dtype=torch.float32).pin_memory()
    h_obs = torch.empty((B, 1, S, S), dtype=torch.float32).pin_memory()

    def e2e_step(i):
        mv = eng.upload_move(poses[:, t + i], poses[:, t + i + 1], az[:, t + i], az[:, t + i + 1])
        eng.step(mv, host_out=(h_val, h_map8, h_obs))              # D2H on a side stream right after the forward, under stages D/E
        eng.wait_host_outputs()                                    # the caller consumes the maps before planning the next pose
        return mv

    for i in range(args.warmup):
        e2e_step(i)
    barrier()
    tic = time.perf_counter()
    e0.record()
    for i in range(args.warmup, n_run):
        mv = e2e_step(i)
    e1.record()
    barrier()                                                      # includes torch.cuda.synchronize(): stages D/E of the last step
    wall = time.perf_counter() - tic
    e2e_ms = max_over_ranks(max(e0.elapsed_time(e1), wall * 1e3))
    e2e_value = B_total * args.steps / (e2e_ms / 1e3)
    h2d = sum(x.numel() * x.element_size() for x in mv)
    d2h = sum(x.numel() * x.element_size() for x in (h_val, h_map8, h_obs))

    # ================= roofline of the dominant kernel
    peaks = load_peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")      # from the committed `ncu --set full` capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch_mean")
    conv_ms, conv_flops = cms.value, cfl.value
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "conv_gemm_f16 (tcgen05.mma kind::f16, TMEM accumulators, TMA operands)",
                "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"],
                "traffic": traffic, "traffic_source": "profiles/conv_traffic.json: mean dram read+write bytes of the launches in the committed ncu --set full capture",
                "algorithmic_flops_per_launch_mean": conv_flops / max(int(cn.value), 1), "peak_source": peaks["which"],
                "launches_timed": int(cn.value), "launches_dropped": int(cdrop.value), "kernel_ms_per_step": conv_ms / args.steps,
                "share_of_step": conv_ms / ms_total if ms_total > 0 else None,
                "flops_counted": "algorithmic 2*M*N*K of the reference's convolutions (182.4 GFLOP per scene-step at 256x256; the fused "
                                 "upsample+conv layers are counted as the 3x3 conv on the upsampled image they replace); "
                                 + ("precision fp16x2 executes 3 tensor-core passes per executed flop, and the fused up-sampling "
                                    "executes 2.25x fewer MACs on 6 layers: executed tensor flops = 2.47 x algorithmic"
                                    if args.precision == "fp16x2" else "precision fp16: 1 pass"),
                "mma_passes": 3 if args.precision == "fp16x2" else 1}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ================= cpu_baseline (rank 0, N=1 only): bounded sample on the host cores
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n0 = int(eng.cloud_len[0].item())
        cloud0 = eng.cloud[0, :n0].cpu().numpy()
        traj0 = list(eng.traj[0, : eng.traj_len_host].cpu().numpy())
        tcur = t + n_run                                             # engine state = pose tcur
        times = cpu_steps(scenes[0], poses[0], az[0], cloud0, traj0, min(tcur, poses.shape[1] - 5), 1 + args.cpu_steps, S, sd, threads)
        v = 1.0 / float(np.mean(times[1:]))
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_steps} whole scene-steps of scene 0 at its current state ({n0} cloud points) after 1 warm-up; "
                         f"oracle port of the reference CPU path (naive rasteriser in C over {threads} threads, numpy histogram, torch fp32 NBP)"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f16x2 (split-fp16 operands, fp32 accumulate; fp32-grade results)" if args.precision == "fp16x2" else "f16 (fp32 accumulate)",
            "data": "synthetic",
            "config": {"workload": f"{B_total} parallel AiMDoom-{args.level}-shaped scenes, {S}x{S} grid, inference rollout "
                                   f"(BASELINE configs[{3 if (args.level == 'insane' and S == 512) else 1}])",
                       "scenes_total": B_total, "scenes_per_gpu": per, "image": "256x456", "mesh_level": args.level,
                       "mean_faces_per_scene": float(np.mean([s.n_faces for s in scenes])), "prefill_pose": args.prefill,
                       "mean_cloud_points_per_scene_at_start": cloud_pts, "nbp_chunk": args.chunk, "precision": args.precision,
                       "l2": "inputs larger than L2 (per-step working set > 5 GB: clouds, frames, activations)",
                       "parallelism": f"scenes sharded over {world} rank(s), no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "stage_ms": stage_ms,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scenes", type=int, default=256)
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--level", default="simple")
    ap.add_argument("--prefill", type=int, default=50)
    ap.add_argument("--chunk", type=int, default=32)
    ap.add_argument("--precision", default="fp16x2", choices=["fp16x2", "fp16"])
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
