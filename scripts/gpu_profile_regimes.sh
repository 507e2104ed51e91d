#!/bin/bash
# ncu evidence for the two regimes of the scan on one configs[3] shard (12.5M x 1280): Q = 4096 (tensor-bound) and Q = 64 / 16
# (bandwidth-bound).  Plain run first, then one --set full capture of the full-shard scan launch of each.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for wl in cfg3shard cfg3shardq64 cfg3shardq16; do
  A="--steps 2 --warmup 1 --no-cpu-baseline --no-north-star --workload $wl"
  python bench.py $A > gpurun_out/reg_${wl}_plain.log 2>&1 || { echo "$wl plain failed"; continue; }
  # the full-shard FILTER scan is the LAST scan launch of a step: skip the warm-up steps' launches, capture one
  ncu --set full --clock-control none --import-source on -k regex:"scan_tc" -s 8 -c 3 -o gpurun_out/reg_${wl}_scan \
      python bench.py $A > gpurun_out/reg_${wl}_ncu.log 2>&1
  echo "$wl rc=$?"
done
