#!/bin/bash
# ncu --set full captures of the stage kernels inside one bench.py solve (run under gpurun; ONE GPU).
#   tools/profile_kernels.sh <tag>      -> gpurun_out/prof_<tag>_main.ncu-rep (+ optional variants)
# The captured launches are iteration 7 of the 4th solve (3 warm-up solves x 10 iterations x 3 stage kernels skipped:
# backward_fused (K1 + K2), forward_first, forward_fanout).
set -u
TAG=${1:-r1}
mkdir -p gpurun_out
COMMON="--set full --clock-control none --import-source on"
timeout 600 ncu $COMMON -k regex:'backward_fused_kernel|forward_first_kernel|forward_fanout_kernel' \
    --launch-skip 108 --launch-count 3 -f -o gpurun_out/prof_${TAG}_main \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_${TAG}_main.log 2>&1
if [ "${PROFILE_COOP:-0}" = "1" ]; then
  NMPC_B200_BWD_GS=4 timeout 600 ncu $COMMON -k regex:'backward_coop_kernel' --launch-skip 36 --launch-count 1 -f \
      -o gpurun_out/prof_${TAG}_coop python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_${TAG}_coop.log 2>&1
fi
if [ "${PROFILE_WIDE:-0}" = "1" ]; then
  # K2 for many inputs (centroidal motion 9 x 16, B = 1024): the 2nd sweep of the 2nd solve
  timeout 300 ncu $COMMON -k regex:'backward_wide_kernel' --launch-skip 5 --launch-count 1 -f \
      -o gpurun_out/prof_${TAG}_wide python tools/time_configs.py centroidal_profile > gpurun_out/prof_${TAG}_wide.log 2>&1
fi
ls -la gpurun_out/*.ncu-rep
