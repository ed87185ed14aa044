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
roofline: the dominant kernel (conv_gemm_f16, tcgen05) timed live with CUDA events around every launch, over a second pass
         of the same K steps with the network launched kernel by kernel (the `value` pass replays a captured CUDA graph of
         the network, and events inside a captured graph cannot be timed)
cpu_baseline: the oracle (CPU restatement of the reference path; the reference's own files cannot run here, see
         DESIGN.md) timed on this box's host cores on a bounded sample: whole steps of ONE scene from the SAME state the
         GPU arm is timed from (pose `prefill`)
extra blocks (reported baselines / secondary configs, each outside the two timed regions above):
  latency_b1     : NBP.forward at batch 1 (BASELINE configs[0]: the reference calls the net with one scene), 128^2 and 256^2
  cudnn_baseline : the same network under torch + cuDNN on this GPU (fp32 with TF32 off, and TF32 on), "the kernel to beat"
  configs3_reduced: BASELINE configs[3] shape (AiMDoom-insane-shaped meshes, 512x512 grid) at 32 rollouts per GPU: one GPU's share of the
                   configuration; at N = 4 ranks the block IS configs[3] (128 rollouts across 4 GPUs)
  train          : BASELINE configs[2] shape -- NBP fwd + loss + bwd + AdamW on 256^2 tiles, 64 tiles per GPU per optimizer
                   step, NCCL all-reduce of the flat gradient across the N ranks
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
    from nextbestpath_b200 import synthetic as syn

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
    n_total = args.prefill + 2 * (args.warmup + args.steps) + (args.steps + 1) + 4
    scenes, poses, az = make_workload(B, args.level, n_total + 8, rank0_scene_index=first)
    eng = RolloutEngine(scenes, None, dev, S=S, max_steps=n_total + 1, seed=9)
    eng.overlap_geometry = args.overlap
    eng.reset(poses[:, 0])
    t = 0
    for _ in range(args.prefill):                                   # geometry-only fast-forward to pose `prefill`
        eng.step(eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1]), run_network=False)
        t += 1
    torch.cuda.synchronize()
    # seeded weights; BatchNorm running statistics = batch statistics of 8 model inputs of THIS workload (one train-mode pass on the
    # CUDA train path): what training on such grids would leave there.  Statistics calibrated on unrelated inputs would put every
    # activation far outside the range a trained network works in.
    net = syn.calibrated_nbp(dev, seed=9, calib_x=eng.build_model_input()[: min(B, 8)].clone())
    net.precision = args.precision
    net.max_chunk = args.chunk
    eng.set_network(net)

    def all_mean(local_sum, local_n):
        v = torch.tensor([float(local_sum), float(local_n)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(v)
        return float(v[0] / v[1])

    cloud_pts = all_mean(eng.cloud_len.double().sum().item(), B)                      # over ALL ranks' scenes
    mean_faces = all_mean(sum(s.n_faces for s in scenes), B)
    # state of scene 0 at pose `prefill`: what the cpu_baseline leg starts from (the same state the GPU arm is timed from)
    cpu_state = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n0 = int(eng.cloud_len[0].item())
        cpu_state = (eng.cloud[0, :n0].cpu().numpy(), list(eng.traj[0, : eng.traj_len_host].cpu().numpy()), t)

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
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    prof = os.environ.get("NBP_BENCH_CUDA_PROFILER") == "1"         # ncu/nsys --profile-from-start off: capture exactly the timed steps
    if prof:
        torch.cuda.profiler.start()
    e0.record()
    for i in range(args.warmup, n_run):
        eng.step(moves[i])
    e1.record()
    barrier()
    if prof:
        torch.cuda.profiler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    t += n_run
    ms_per_step = ms_total / args.steps
    value = B_total * args.steps / (ms_total / 1e3)

    # ================= per-stage device time of one extra step (diagnostic, outside both timed regions)
    from nextbestpath_b200.rollout import STAGE_NAMES
    eng.stage_events = []
    eng.overlap_geometry = False                                   # the diagnostic wants the stages back to back on one stream
    eng.step(eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1]))
    torch.cuda.synchronize()
    eng.overlap_geometry = args.overlap
    evs = eng.stage_events
    eng.stage_events = None
    stage_ms = {nm: evs[i].elapsed_time(evs[i + 1]) for i, nm in enumerate(STAGE_NAMES)}
    t += 1
    sb0, sb1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sb0.record()
    for _ in range(5):                                             # stage B again on the same state: the single sample above is noisy
        eng.build_model_input()
    sb1.record()
    torch.cuda.synchronize()
    stage_ms["B_grid_scatter_mean_of_5"] = sb0.elapsed_time(sb1) / 5

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

    t += n_run
    # ================= roofline pass: the same K steps, network launched kernel by kernel, CUDA events around every conv launch
    net.use_cuda_graph = False
    moves = [eng.upload_move(poses[:, t + i], poses[:, t + i + 1], az[:, t + i], az[:, t + i + 1]) for i in range(args.steps + 1)]
    eng.step(moves[0])
    barrier()
    _lib.check(L.nbp_conv_profile_begin(60000), "nbp_conv_profile_begin")
    e0.record()
    for i in range(1, args.steps + 1):
        eng.step(moves[i])
    e1.record()
    barrier()
    prof_ms_total = e0.elapsed_time(e1)
    cms, cfl, cn, cdrop = ctypes.c_double(), ctypes.c_double(), ctypes.c_uint64(), ctypes.c_uint64()
    _lib.check(L.nbp_conv_profile_end(ctypes.byref(cms), ctypes.byref(cfl), ctypes.byref(cn), ctypes.byref(cdrop)), "nbp_conv_profile_end")
    net.use_cuda_graph = True
    t += args.steps + 1

    # ================= roofline of the dominant kernel
    peaks = load_peaks()
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "conv_traffic.json")      # from the committed `ncu --set full` capture
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch_mean")
    conv_ms, conv_flops = cms.value, cfl.value
    achieved = conv_flops / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "conv_gemm_f16 (tcgen05.mma kind::f16 + kind::f8f6f4, TMEM accumulators, TMA operands)",
                "achieved": achieved, "peak": peaks["bf16"], "unit": "TFLOP/s", "frac": achieved / peaks["bf16"],
                "traffic": traffic, "traffic_source": "profiles/conv_traffic.json: mean dram read+write bytes of the launches in the committed ncu --set full capture",
                "algorithmic_flops_per_launch_mean": conv_flops / max(int(cn.value), 1), "peak_source": peaks["which"],
                "launches_timed": int(cn.value), "launches_dropped": int(cdrop.value), "kernel_ms_per_step": conv_ms / args.steps,
                "share_of_step": conv_ms / prof_ms_total if prof_ms_total > 0 else None,
                "timed_in": f"a second pass of the same {args.steps} steps with per-launch CUDA events and the network launched kernel by "
                            f"kernel ({prof_ms_total / args.steps:.2f} ms/step there; the value pass replays the network as a CUDA graph)",
                "step_ms_in_profiled_pass": prof_ms_total / args.steps,
                "flops_counted": "algorithmic 2*M*N*K of the reference's convolutions (182.4 GFLOP per scene-step at 256x256; the fused "
                                 "upsample+conv layers are counted as the 3x3 conv on the upsampled image they replace); "
                                 + {"mixed": "precision mixed: the fp16 hi product + one e4m3 reduction of twice the K at twice the rate = 2 fp16-pass "
                                             "equivalents per flop (3 in the five fp16x2 encoder layers); the fused up-sampling executes 2.25x fewer "
                                             "MACs on 6 layers",
                                    "fp16x2": "precision fp16x2 executes 3 tensor-core passes per executed flop, and the fused up-sampling "
                                              "executes 2.25x fewer MACs on 6 layers: executed tensor flops = 2.47 x algorithmic",
                                    "fp16": "precision fp16: 1 pass"}[args.precision],
                "mma_pass_equivalents": {"mixed": 2.13, "fp16x2": 3, "fp16": 1}[args.precision],
                "e4m3_saturation_events": net.e4m3_saturation_count()}

    # ================= extra blocks (secondary configs and reported baselines; all outside the timed regions above)
    del eng, moves
    torch.cuda.empty_cache()
    latency = None if args.no_extras else latency_b1(net, dev)
    cfg3 = None if args.no_extras else cfg3_block(args, dev, world, rank, net, max_over_ranks, barrier)
    train = None if args.no_extras else train_block(args, dev, world, rank, max_over_ranks, barrier)
    cudnn = None
    if rank == 0 and world == 1 and not args.no_extras:
        cudnn = cudnn_baseline(net, dev, B_total, S, args.chunk, value)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ================= cpu_baseline (rank 0, N=1 only): bounded sample on the host cores, from the state the GPU arm was timed from
    cpu = None
    if cpu_state is not None:
        from oracle import nbp_torch as NT                          # the checker doubles as the CPU baseline (kind: "port")
        threads = os.cpu_count() or 1
        cloud0, traj0, t_state = cpu_state
        times = cpu_steps(scenes[0], poses[0], az[0], cloud0, traj0, t_state, 1 + args.cpu_steps, S, NT.golden_state_dict(seed=9), threads)
        v = 1.0 / float(np.mean(times[1:]))
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{args.cpu_steps} whole scene-steps of scene 0 from pose {t_state} ({len(cloud0)} cloud points: the state the GPU arm's "
                         f"timed region starts from) after 1 warm-up; oracle port of the reference CPU path (naive rasteriser in C over "
                         f"{threads} threads, numpy histogram, torch fp32 NBP)"}

    full_cfg1 = (B_total == 256 and args.level == "simple" and S == 256)
    full_cfg3 = (B_total == 128 and args.level == "insane" and S == 512 and world == 4)
    label = "BASELINE configs[1]" if full_cfg1 else "BASELINE configs[3]" if full_cfg3 else \
        f"a reduced / non-BASELINE configuration ({B_total} scenes, {args.level}, {S}x{S}, {world} GPU)"
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"mixed": "f16 + e4m3 corrections (5 encoder layers split-fp16 x2), fp32 accumulate; value maps 1e-4 of fp32",
                      "fp16x2": "f16x2 (split-fp16 operands, fp32 accumulate; fp32-grade results)", "fp16": "f16 (fp32 accumulate)"}[args.precision],
            "data": "synthetic",
            "config": {"workload": f"{B_total} parallel AiMDoom-{args.level}-shaped scenes, {S}x{S} grid, inference rollout ({label})",
                       "scenes_total": B_total, "scenes_per_gpu": per, "image": "256x456", "mesh_level": args.level,
                       "mean_faces_per_scene": mean_faces, "prefill_pose": args.prefill,
                       "mean_cloud_points_per_scene_at_start": cloud_pts, "nbp_chunk": args.chunk, "precision": args.precision,
                       "network_launch": "CUDA graph replay (captured once per shape)",
                       "geometry_overlap": "stages D/E on a side stream under stage C" if args.overlap else "off (single stream)",
                       "weights": "seeded (no checkpoints offline); BatchNorm statistics calibrated on 8 model inputs of this workload",
                       "l2": "inputs larger than L2 (per-step working set > 5 GB: clouds, frames, activations)",
                       "parallelism": f"scenes sharded over {world} rank(s), no collective"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "stage_ms": stage_ms,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "latency_b1": latency, "cudnn_baseline": cudnn, "train": train, "configs3_reduced": cfg3}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


# ------------------------------------------------------------------------------------------------ extra blocks
def _time_ms(fn, n, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def latency_b1(net, dev):
    """BASELINE configs[0]: the reference drivers call the network with ONE scene (nbp_planning.py:166,395).  Device time per
    NBP.forward call at batch 1 through the drop-in module (graph replay + the copy of the input into the graph's buffer +
    the clones the drop-in API returns), and the host wall time of the blocking call."""
    from nextbestpath_b200 import synthetic as syn
    out = {}
    static = net.static_outputs
    net.static_outputs = False
    for S in (128, 256):
        x = syn.count_like_input(1, S, seed=5).to(dev)
        with torch.no_grad():
            dev_ms = _time_ms(lambda: net(x), 30, warm=3)
            torch.cuda.synchronize()
            tic = time.perf_counter()
            for _ in range(30):
                o1, _ = net(x)
                o1[0, 0, 0, 0].item()
            wall_ms = (time.perf_counter() - tic) / 30 * 1e3
        out[f"S{S}"] = {"device_ms": dev_ms, "blocking_call_ms": wall_ms, "gflop": FLOP_PER_SCENE_STEP[S] / 1e9}
    net.static_outputs = static
    out["note"] = "B = 1: one 128-pixel tile row per SM at the deep levels; latency is set by ~60 dependent kernels, not by tensor throughput"
    return out


def cudnn_baseline(net, dev, B_total, S, chunk, our_value):
    """torch + cuDNN on the same GPU: the reference's nbp_model.py architecture (functional restatement, same weights) in eval
    mode under no_grad, in chunks of `chunk` scenes -- fp32 with TF32 disabled (the arithmetic the parity bar is stated in) and
    with TF32 enabled (cuDNN's tensor-core path; its value-map error against the fp32 run is reported).  Network only."""
    from nextbestpath_b200 import synthetic as syn
    from oracle import nbp_torch as NT                               # the functional torch restatement of nbp_model.py (a baseline, not the product)
    sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
    x = syn.count_like_input(chunk, S, seed=6).to(dev)
    n_chunks = (B_total + chunk - 1) // chunk
    res = {}
    outs = {}
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.benchmark = True
        for name, tf32 in (("fp32", False), ("tf32", True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            with torch.no_grad():
                ms = _time_ms(lambda: NT.forward(sd, x), 3, warm=2)
                outs[name] = NT.forward(sd, x)[0]
            res[name] = {"ms_per_chunk": ms, "network_only_scene_steps_per_s": chunk / (ms / 1e3),
                         "ms_per_256_scene_forward": ms * n_chunks,
                         "algorithmic_tflops": chunk * FLOP_PER_SCENE_STEP[S] / (ms * 1e-3) / 1e12}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
    with torch.no_grad():
        net_static = net.static_outputs
        net.static_outputs = False
        ours = net(x)[0]
        ours_ms = _time_ms(lambda: net(x), 5, warm=2)
        net.static_outputs = net_static
    ref = outs["fp32"]
    rel = lambda a: float((a - ref).abs().max() / ref.abs().max())
    res["tf32"]["value_map_max_rel_err_vs_cudnn_fp32"] = rel(outs["tf32"])
    res["ours"] = {"ms_per_chunk": ours_ms, "network_only_scene_steps_per_s": chunk / (ours_ms / 1e3),
                   "algorithmic_tflops": chunk * FLOP_PER_SCENE_STEP[S] / (ours_ms * 1e-3) / 1e12,
                   "value_map_max_rel_err_vs_cudnn_fp32": rel(ours)}
    res["speedup_vs_cudnn_fp32"] = res["fp32"]["ms_per_chunk"] / ours_ms
    res["speedup_vs_cudnn_tf32"] = res["tf32"]["ms_per_chunk"] / ours_ms
    res["note"] = (f"eval forward of {chunk} scenes at {S}x{S}, torch {torch.__version__} / cuDNN {torch.backends.cudnn.version()}, "
                   "cudnn.benchmark on, same weights and input; the parity bar (1e-3) is met by ours and by cuDNN fp32 only")
    return res


def cfg3_block(args, dev, world, rank, net, max_over_ranks, barrier):
    """BASELINE configs[3] per-GPU share: AiMDoom-insane-shaped meshes (~50 k triangles), 512x512 grid, 32 rollouts per GPU (the full
    configuration is 128 rollouts across 4 GPUs = 32 per GPU; the network runs them in chunks of 8 scenes: 8 scenes at 512^2 have the
    tensor sizes of 32 scenes at 256^2).  Same engine, same network module (its graph for the new shape is captured in the warm-up)."""
    from nextbestpath_b200.rollout import RolloutEngine
    n_sc, prefill, warm, steps, S = 32, 20, 2, 3, 512     # 32 rollouts per GPU: at 4 ranks exactly configs[3] (128 rollouts across 4 GPUs)
    old_chunk, net.max_chunk = net.max_chunk, 8                                   # 8 scenes at 512^2 = the tensor sizes of 32 scenes at 256^2
    scenes, poses, az = make_workload(n_sc, "insane", prefill + warm + steps + 4, rank0_scene_index=rank * n_sc, seed0=5000)
    eng = RolloutEngine(scenes, net, dev, S=S, max_steps=prefill + warm + steps + 3, seed=9)
    eng.reset(poses[:, 0])
    t = 0
    for _ in range(prefill):
        eng.step(eng.upload_move(poses[:, t], poses[:, t + 1], az[:, t], az[:, t + 1]), run_network=False)
        t += 1
    moves = [eng.upload_move(poses[:, t + i], poses[:, t + i + 1], az[:, t + i], az[:, t + i + 1]) for i in range(warm + steps)]
    for i in range(warm):
        eng.step(moves[i])
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(warm, warm + steps):
        eng.step(moves[i])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    out = {"metric": METRIC, "value": world * n_sc * steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "ms_per_step": ms / steps,
           "scenes_per_gpu": n_sc, "grid": f"{S}x{S}", "mesh_level": "insane", "mean_faces_per_scene": float(np.mean([s.n_faces for s in scenes])),
           "prefill_pose": prefill, "mean_cloud_points_per_scene": float(eng.cloud_len.float().mean().item()), "steps": steps,
           "algorithmic_tflops_network": world * n_sc * steps * FLOP_PER_SCENE_STEP[S] / (ms * 1e-3) / 1e12,
           "workload": ("BASELINE configs[3]: 128 insane rollouts on a 512x512 grid across 4 GPUs" if world == 4 else
                        f"BASELINE configs[3] shape, reduced to {world * n_sc} rollouts on {world} GPU (full: 128 rollouts across 4 GPUs)")}
    net.max_chunk = old_chunk
    del eng, moves
    torch.cuda.empty_cache()
    return out


def train_block(args, dev, world, rank, max_over_ranks, barrier):
    """BASELINE configs[2] shape: train_nbp.py's optimizer step (nbp_utils.py:340-391) on 256x256 map tiles, 64 tiles per GPU
    per optimizer step in micro-batches of 32 (gradients accumulated), ONE NCCL all-reduce of the flat 199.9 MB gradient per
    step, AdamW.  tiles/s over all ranks; algorithmic FLOPs = 3 x forward."""
    from nextbestpath_b200 import ops
    from nextbestpath_b200 import synthetic as syn
    from nextbestpath_b200.networks import NBP
    from nextbestpath_b200.train import FlatGradAllReduce, train_step
    S, K, tiles, micro = 256, 64, args.train_tiles, args.train_micro
    net = NBP()
    net.load_state_dict(syn.seeded_nbp_state_dict(net, 9))
    net.to(dev)
    opt = torch.optim.AdamW(net.parameters(), lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)       # nbp_utils.py:228
    red = FlatGradAllReduce(net.parameters())
    g = torch.Generator().manual_seed(rank)
    mbs = []
    for i in range(0, tiles, micro):
        b = min(micro, tiles - i)
        x = syn.count_like_input(b, S, seed=100 * rank + i).to(dev)
        tp = torch.stack((torch.randint(0, 8, (b, K), generator=g), torch.randint(0, S // 4, (b, K), generator=g),
                          torch.randint(0, S // 4, (b, K), generator=g)), -1).to(dev)
        mbs.append((x, tp, (torch.rand(b, K, generator=g) * 10).to(dev), (torch.rand(b, 1, S, S, generator=g) < 0.2).float().to(dev)))
    loss = train_step(net, opt, mbs, red)
    barrier()
    l0 = ops.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    n_steps = 2
    ar_ms = 0.0
    for _ in range(n_steps):
        loss = train_step(net, opt, mbs, red)
        ar_ms += red.last_ms()
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1) / n_steps)
    tiles_s = world * tiles / (ms / 1e3)
    out = {"metric": "nbp_train_tiles_per_sec", "value": tiles_s, "unit": "tiles/s", "n_gpus": world, "ms_per_optimizer_step": ms,
           "tiles_per_gpu_per_step": tiles, "global_tiles_per_step": world * tiles, "micro_batch": micro, "grid": S, "loss": float(loss),
           "algorithmic_tflops": tiles_s * 3 * FLOP_PER_SCENE_STEP[S] / 1e12,
           "frac_of_bf16_peak_per_gpu": tiles_s / world * 3 * FLOP_PER_SCENE_STEP[S] / 1e12 / load_peaks()["bf16"],
           "grad_allreduce_bytes": red.flat.numel() * 4 if world > 1 else 0, "grad_allreduce_ms": ar_ms / n_steps if world > 1 else 0.0,
           "gpu_launches_per_step": int((ops.launch_count() - l0) // n_steps), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
           "workload": f"BASELINE configs[2] shape ({'the full 512-tile, 8-GPU configuration' if world == 8 and tiles == 64 else f'{world * tiles} tiles per step on {world} GPU'}): "
                       "fp16x2 tcgen05 forward / dgrad / wgrad, train-mode BatchNorm, AdamW, synthetic count tiles"}
    del net, opt, red, mbs
    torch.cuda.empty_cache()
    return out


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
    ap.add_argument("--precision", default="mixed", choices=["mixed", "fp16x2", "fp16"])
    ap.add_argument("--cpu-steps", type=int, default=8)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--overlap", action="store_true", help="run stages D/E on a side stream under the network (measured: no gain)")
    ap.add_argument("--no-extras", action="store_true", help="skip the latency_b1 / cudnn_baseline / train blocks")
    ap.add_argument("--train-tiles", type=int, default=64, help="tiles per GPU per optimizer step of the train block")
    ap.add_argument("--train-micro", type=int, default=32)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
