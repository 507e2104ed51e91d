#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-600; }
run f1_pool 900 python -m pytest tests/test_gpu_maskpool.py -q -m gpu -x --timeout 300
run f1_dropin 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_selfjoin.py -q -m gpu -x --timeout 600
run f1_bench 900 python bench.py --steps 20 --warmup 5 --no-north-star --no-cpu-baseline
run f1_q1big 600 python bench.py --workload cfg3shardq1 --steps 10 --warmup 3 --no-cpu-baseline --no-north-star
run f1_cfg1q1 600 python bench.py --workload cfg1q1 --steps 50 --warmup 5 --no-cpu-baseline --no-north-star
run f1_poolbench 300 python scripts/bench_maskpool.py
