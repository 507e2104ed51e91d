#!/bin/bash
# ncu passes for the bench command (B200_PROFILING.md recipe): plain run first, then the launch list, then one
# --set full capture of the scan kernel.  Usage: gpu_profile.sh <tag> [bench args...]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
TAG=${1:-prof}; shift
KREGEX=${KREGEX:-scan_tc}; KSKIP=${KSKIP:-3}; KCOUNT=${KCOUNT:-3}
ARGS="--steps 2 --warmup 1 --no-cpu-baseline $*"
python bench.py $ARGS > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
python bench.py $ARGS > gpurun_out/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:$KREGEX -s $KSKIP -c $KCOUNT -o gpurun_out/${TAG}_scan \
    python bench.py $ARGS > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
tail -n 3 gpurun_out/${TAG}_plain.log
