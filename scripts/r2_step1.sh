#!/bin/bash
# round 2, step 1: hot-list select — parity, phase trace, short bench with and without the hot lists, launch list
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log; }
run a1_search 1500 python -m pytest tests/test_gpu_search.py -q -m gpu -x --timeout 900
run a1_trace 300 python scripts/dev/trace_select.py
PLANT=0 run a1_trace_unplanted 300 python scripts/dev/trace_select.py
run a1_bench_hot 600 python bench.py --steps 20 --warmup 5 --no-north-star --no-cpu-baseline
RVO_OPTS=hot=0 run a1_bench_nohot 600 python bench.py --steps 20 --warmup 5 --no-north-star --no-cpu-baseline
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/a1_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-north-star > gpurun_out/a1_ncu1.log 2>&1
echo "launch list rc=$?"
