#!/bin/bash
# ncu --set full captures of the stage kernels inside one bench.py solve (run under gpurun; ONE GPU), and the DRAM
# traffic file bench.py's roofline.traffic reads.
#   GIT_COMMIT=$(git rev-parse --short HEAD) tools/profile_kernels.sh <tag>
#     -> gpurun_out/prof_<tag>_main.ncu-rep, profiles/<tag>_stage_kernels.md, profiles/dram_traffic.json
# The captured launches are iteration 7 of the 4th solve (3 warm-up solves x 10 iterations x 3 stage kernels skipped).
set -u
TAG=${1:-r2}
mkdir -p gpurun_out profiles
COMMON="--set full --clock-control none --import-source on"
timeout 600 ncu $COMMON -k regex:'backward_lanes_kernel|forward_first_split_kernel|forward_fanout_split_kernel' \
    --launch-skip 108 --launch-count 3 -f -o gpurun_out/prof_${TAG}_main \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_${TAG}_main.log 2>&1
python tools/summarize_ncu.py rep gpurun_out/prof_${TAG}_main.ncu-rep gpurun_out/${TAG}_stage_kernels.md gpurun_out/dram_traffic_${TAG}.json
python - <<PY
import json
raw = json.load(open("gpurun_out/dram_traffic_${TAG}.json"))
names = {"backward": "backward_lanes_kernel", "forward_first": "forward_first_split_kernel", "forward_rest": "forward_fanout_split_kernel"}
out = {"commit": "${GIT_COMMIT:-unknown}", "file": "profiles/${TAG}_stage_kernels.md",
       "command": "ncu --set full --clock-control none ... python bench.py --steps 1 --warmup 3 (iteration 7 of the 4th solve)",
       "kernels": {k: raw[v] for k, v in names.items() if v in raw}}
json.dump(out, open("gpurun_out/dram_traffic.json", "w"), indent=1)
print(json.dumps(out)[:400])
PY
if [ "${PROFILE_WIDE:-0}" = "1" ]; then
  # K2 for many inputs (centroidal motion 9 x 16, B = 1024): the 2nd sweep of the 2nd solve
  timeout 300 ncu $COMMON -k regex:'backward_wide_kernel' --launch-skip 5 --launch-count 1 -f \
      -o gpurun_out/prof_${TAG}_wide python tools/time_configs.py centroidal_profile > gpurun_out/prof_${TAG}_wide.log 2>&1
  python tools/summarize_ncu.py rep gpurun_out/prof_${TAG}_wide.ncu-rep gpurun_out/${TAG}_wide_kernel.md /dev/null
fi
if [ "${PROFILE_BIG:-0}" = "1" ]; then
  # the large-batch kernels: B = 131072 (thread-per-instance fused K2, in-warp fan-out K3)
  timeout 600 ncu $COMMON -k regex:'backward_fused_kernel|forward_kernel' --launch-skip 32 --launch-count 2 -f \
      -o gpurun_out/prof_${TAG}_big python bench.py --batch 131072 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/prof_${TAG}_big.log 2>&1
  python tools/summarize_ncu.py rep gpurun_out/prof_${TAG}_big.ncu-rep gpurun_out/${TAG}_large_batch.md /dev/null
fi
ls -la gpurun_out/*.ncu-rep | tail -5
