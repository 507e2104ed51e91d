#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-600; }
run d1_selfjoin 900 python -m pytest tests/test_gpu_selfjoin.py -q -m gpu -x --timeout 600
run d1_bench 1500 python bench.py --steps 20 --warmup 5
run d1_ref 600 python bench.py --impl reference --steps 20 --warmup 5
