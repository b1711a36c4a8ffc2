#!/bin/bash
# compute-sanitizer passes over small shapes of every kernel family.  Run on the GPU box:
#   bash tools/sanitize.sh <tag>      -> gpurun_out/sanitize_<tag>_{memcheck,racecheck,synccheck}.log
# racecheck sees shared-memory hazards only; the kernels synchronise through mbarriers (arrive/try_wait), which the
# tool models since CUDA 12.x -- a report there is a real finding or a modelling gap, both worth reading.
set -u
tag=${1:-r02}
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck; do
  cases="k1 k2s k2l k2l_checked literal traceback all_triples"
  [ "$tool" = racecheck ] && cases="k1 k2s k2l"
  timeout 900 $CS --tool $tool --print-limit 20 --error-exitcode 9 python tools/sanitize_cases.py $cases \
      > gpurun_out/sanitize_${tag}_${tool}.log 2>&1
  echo "$tool rc=$?" | tee -a gpurun_out/sanitize_${tag}_${tool}.log
  tail -3 gpurun_out/sanitize_${tag}_${tool}.log
done
