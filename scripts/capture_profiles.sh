#!/bin/bash
# ncu captures behind profiles/r02_*: run on a B200 box from the repo root (gpurun -- bash scripts/capture_profiles.sh).
# Never a bench value: everything printed under ncu is discarded.
set -x
OUT=gpurun_out
if [ "$1" != "conv-only" ]; then
# (a) launch list of whole rollout steps at 32 scenes (one network chunk per step), single metric
# (NBP_BENCH_CUDA_PROFILER=1: bench.py brackets its timed `value` steps with cudaProfilerStart/Stop)
NBP_BENCH_CUDA_PROFILER=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/r02_launches_32scenes.csv \
    python bench.py --scenes 32 --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $OUT/r02_launches_32scenes.log 2>&1
# (b) geometry kernels on rollout data (clouds of ~1.45 M points per scene, 32 scenes), full set
NBP_BENCH_CUDA_PROFILER=1 ncu --profile-from-start off --set full --import-source on --clock-control none \
    -k regex:'grid_scatter|bp_select|bp_write|raster_tiles|raster_setup' -c 8 \
    -o $OUT/r02_geometry_full -f python bench.py --scenes 32 --steps 2 --warmup 3 --no-extras --no-cpu-baseline > $OUT/r02_geometry_full.log 2>&1
ncu -i $OUT/r02_geometry_full.ncu-rep --page raw --csv > $OUT/r02_geometry_full_raw.csv 2>/dev/null
rm -f $OUT/r02_geometry_full.ncu-rep
if [ "$1" = "geometry-only" ]; then ls -la $OUT | grep r02_ | tail; exit 0; fi
fi
if true; then
# (c) the conv kernel, all 33 layers of one 32-scene forward in the mixed precision, full set
ncu --profile-from-start off --set full --clock-control none -k regex:conv_gemm -c 33 -o $OUT/r02_conv_full -f \
    python scripts/profile_forward.py mixed 32 256 1 > $OUT/r02_conv_full.log 2>&1
ncu -i $OUT/r02_conv_full.ncu-rep --page raw --csv > $OUT/r02_conv_full_raw.csv 2>/dev/null
rm -f $OUT/r02_conv_full.ncu-rep            # 80+ MB: gpurun only copies back 64 MiB; the raw page above is what the summaries are made from
fi
ls -la $OUT | grep r02_ | tail -12
