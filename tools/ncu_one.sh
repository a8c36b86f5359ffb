#!/bin/bash
# ncu --set full capture of selected launches of one command (run under gpurun; ONE GPU).
#   tools/ncu_one.sh <tag> <kernel regex> <launch-skip> <launch-count> <command...>
# -> gpurun_out/prof_<tag>.ncu-rep ; read here with tools/summarize_ncu.py / tools/ncu_stalls.py
set -u
TAG=$1; REGEX=$2; SKIP=$3; COUNT=$4; shift 4
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$REGEX" \
    --launch-skip $SKIP --launch-count $COUNT -f -o gpurun_out/prof_${TAG} "$@" > gpurun_out/prof_${TAG}.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/prof_${TAG}.log
