#!/bin/bash
# usage: tools/run_exp.sh <tag> <batches> variant...   -- each variant in its own process under a timeout
TAG=$1; shift
BATCHES=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  name=${v%%=*}
  timeout 240 python tools/exp_variants.py --batch $BATCHES --out gpurun_out/exp_${TAG}_${name}.json "$v" > gpurun_out/exp_${TAG}_${name}.log 2>&1
  echo "$name rc=$?"
  tail -2 gpurun_out/exp_${TAG}_${name}.log | cut -c1-1500
done
