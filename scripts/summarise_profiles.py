#!/usr/bin/env python
"""Turn the raw ncu exports of scripts/capture_profiles.sh (gpurun_out/, scratch) into the committed summaries under profiles/."""
import csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
F = {"ns": 1e-3, "nsecond": 1e-3, "us": 1, "usecond": 1, "ms": 1e3, "msecond": 1e3, "byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
HBM = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6459.3


def load_raw(path):
    rows = list(csv.reader(open(path, errors="ignore")))
    h, units, data = rows[0], rows[1], rows[2:]
    col = {c: i for i, c in enumerate(h)}
    def val(r, c):
        try:
            return float(r[col[c]].replace(",", "")) * F.get(units[col[c]], 1)
        except Exception:
            return float("nan")
    return col, data, val


ORDER = ["Conv1.b"] + [f"Conv{l}.{ab}" for l in range(2, 6) for ab in "ab"]
for dec, lvls in ((1, (5, 4)), (2, (5, 4, 3, 2))):
    for l in lvls:
        ORDER += [f"Up{l}_{dec}", f"Att{l}_{dec}", f"Up_conv{l}_{dec}.a", f"Up_conv{l}_{dec}.b"]

# ---- conv, full set
col, data, val = load_raw(os.path.join(G, "r02_conv_full_raw.csv"))
out = ["layer            <N,MODE,HALO,PAIR>      us  tensor-pipe%  L2-hit%  dram MB  dram GB/s   regs"]
tot_d = tot_us = 0.0
for name, r in zip(ORDER, data[-33:]):
    kn = r[col["Kernel Name"]]
    us = val(r, "gpu__time_duration.sum"); d = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    tot_d += d; tot_us += us
    out.append(f"{name:16s} {kn[kn.find('<'):kn.find('>') + 1]:16s} {us:8.1f} {val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'):9.1f} "
               f"{val(r, 'lts__t_sector_hit_rate.pct'):10.1f} {d / 1e6:8.1f} {d / us / 1e3:9.0f} {val(r, 'launch__registers_per_thread'):6.0f}")
out.append(f"total {tot_us:.1f} us; mean DRAM bytes per launch {tot_d / 33 / 1e6:.1f} MB")
open(os.path.join(P, "r02_conv_gemm_ncu_full.txt"), "w").write(
    "ncu --profile-from-start off --set full --clock-control none -k regex:conv_gemm -c 33 python scripts/profile_forward.py mixed 32 256 1   (scripts/capture_profiles.sh)\n"
    "(the 33 conv launches of ONE 32-scene 256x256 eval forward, default mixed precision, CTA-pair kernels; MODE 1 = fp16x2, 2 = fp16+e4m3, 3 = fp16+e4m3 with\n"
    " split stages; HALO = taps per halo stage (0: plain stages); PAIR 1 = tcgen05 cta_group::2 tiles; the attention GEMMs and Up_conv2_2.b run the dot epilogue; bracketed by cudaProfilerStart/Stop; raw page read in-session, the 85 MB .ncu-rep is scratch)\n" + "\n".join(out) + "\n")
json.dump({"dram_bytes_per_launch_mean": tot_d / 33, "launches": 33,
           "source": "profiles/r02_conv_gemm_ncu_full.txt (ncu --set full, the 33 conv launches of one 32-scene forward, mixed precision, CTA pairs)"},
          open(os.path.join(P, "conv_traffic.json"), "w"), indent=1)
print("\n".join(out[-6:]))

# ---- geometry, full set
col, data, val = load_raw(os.path.join(G, "r02_geometry_full_raw.csv"))
lines = []
for r in data:
    name = r[col["Kernel Name"]].split("(")[0]
    us = val(r, "gpu__time_duration.sum"); d = val(r, "dram__bytes_read.sum") + val(r, "dram__bytes_write.sum")
    lines.append(f"{name:22s} grid {int(val(r, 'launch__grid_size')):6d} x {int(val(r, 'launch__block_size')):4d} {us:7.1f} us | dram {d / 1e6:7.1f} MB = {d / us / 1e3:7.1f} GB/s = "
                 f"{d / us / 1e3 / HBM:5.3f} of the measured HBM peak | IPC {val(r, 'sm__inst_executed.avg.per_cycle_elapsed'):.2f} | warps active "
                 f"{val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active'):.0f} % | L2 hit {val(r, 'lts__t_sector_hit_rate.pct'):.0f} % | L2 RED sectors {int(val(r, 'lts__t_sectors_srcunit_tex_op_red.sum'))}")
BEFORE = """bp_select              grid     32 x 1024   155.2 us | dram    15.0 MB =    96.4 GB/s = 0.015 of the measured HBM peak | IPC 0.51 | warps active 50 % | L2 hit 40 % | L2 RED sectors 0
bp_write               grid     32 x 1024   152.0 us | dram    15.4 MB =   101.2 GB/s = 0.016 of the measured HBM peak | IPC 0.53 | warps active 47 % | L2 hit 37 % | L2 RED sectors 0
grid_scatter           grid  12128 x  256   352.4 us | dram   598.8 MB =  1699.2 GB/s = 0.263 of the measured HBM peak | IPC 1.23 | warps active 66 % | L2 hit 37 % | L2 RED sectors 19215445
raster_setup           grid   2816 x  256    37.4 us | dram     9.1 MB =   242.1 GB/s = 0.037 of the measured HBM peak | IPC 0.77 | warps active 44 % | L2 hit 54 % | L2 RED sectors 13
raster_tiles           grid  59392 x  256   488.5 us | dram    26.8 MB =    55.0 GB/s = 0.009 of the measured HBM peak | IPC 3.04 | warps active 60 % | L2 hit 63 % | L2 RED sectors 0
bp_select              grid    128 x 1024   158.4 us | dram    59.9 MB =   377.9 GB/s = 0.059 of the measured HBM peak | IPC 2.06 | warps active 50 % | L2 hit 40 % | L2 RED sectors 0
bp_write               grid    128 x 1024   168.5 us | dram    61.4 MB =   364.3 GB/s = 0.056 of the measured HBM peak | IPC 1.91 | warps active 45 % | L2 hit 35 % | L2 RED sectors 0
bp_select              grid     32 x 1024   153.7 us | dram    15.0 MB =    97.3 GB/s = 0.015 of the measured HBM peak | IPC 0.52 | warps active 50 % | L2 hit 40 % | L2 RED sectors 0"""
READING = """reading (algorithmic bytes per SURVEY 8d in brackets):
  grid_scatter   first half of round 2: 599 MB of DRAM traffic for 32 x 1.5 M points x 12 B = 576 MB [+ 42 MB of grid], no wasted traffic, 0.26 of the HBM peak; the
  (_hash)        limiter was the 19.2 M fp32 RED sectors (2.5 points per RED): wall cells are hit by every frame that sees the wall.  grid_scatter_hash counts every
                 16384-point chunk in a shared-memory hash table first (one RED per distinct cell and chunk): 7.7 M RED sectors, same DRAM bytes, 0.35 of the HBM
                 peak; the kernel is now issue-bound (IPC 2.4, top stall: the barriers between its clear / insert / flush phases).
  bp_select /    one CTA of 1024 threads per frame; with the key scratch (every pixel's 6-round Feistel key is evaluated once, in the first select pass, and read
  bp_write       back by the second pass and by bp_write) and a warp-parallel histogram scan: ~80 us per kernel and 32-frame call instead of ~155.  Still latency
                 of the per-frame passes, not bandwidth (IPC 0.4-1.7): at the bench's 256 / 1024 frames per call the CTAs fill the GPU (stage A 0.66 ms, E 2.0 ms).
  raster_tiles   IPC 3.0 of 4: instruction-issue bound by the op-for-op pinned fp32 arithmetic; the 8 x 4 warp footprint and the z-prune are worth 2-6 %.
"""
open(os.path.join(P, "r02_geometry_ncu_full.txt"), "w").write(
    "NBP_BENCH_CUDA_PROFILER=1 ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:'grid_scatter|bp_select|bp_write|raster_tiles|raster_setup' -c 8 \\\n"
    "    python bench.py --scenes 32 --steps 2 --warmup 3 --no-extras --no-cpu-baseline      (scripts/capture_profiles.sh; rollout data: 32 scenes x ~1.5 M cloud points)\n\n"
    "---- first half of round 2 (direct RED scatter, three key evaluations per pixel, 16 x 2 raster warps)\n" + BEFORE +
    "\n\n---- end of round 2 (grid_scatter_hash, key scratch + warp scan, 8 x 4 raster warps + z-prune)\n" + "\n".join(lines) + "\n\n" + READING)
print("\n".join(lines))

# ---- launch list
txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "launch_list.py"), os.path.join(G, "r02_launches_32scenes.csv")], capture_output=True, text=True).stdout
open(os.path.join(P, "r02_launches_two_steps_32scenes.txt"), "w").write(
    "NBP_BENCH_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python bench.py --scenes 32 --steps 2 --warmup 3 --no-extras --no-cpu-baseline\n"
    "(the two timed rollout steps of a 32-scene run = one network chunk per step; mixed precision, CTA-pair kernels; CUDA-graph kernel nodes are profiled individually; cold-cache, serialised times)\n" + txt)
import shutil
shutil.copy(os.path.join(G, "r02_launches_32scenes.csv"), os.path.join(P, "r02_launches_32scenes.csv"))
print(txt)
