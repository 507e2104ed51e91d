#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run c1_ref 900 python -m pytest tests/test_reference_source.py tests/test_gpu_multidevice.py -q -m gpu --timeout 600 -rs
run c1_search 1500 python -m pytest tests/test_gpu_search.py tests/test_gpu_dropin.py -q -m gpu -x --timeout 900
run c1_cfg0 300 python bench.py --workload cfg0 --steps 200 --warmup 20 --no-cpu-baseline --no-north-star
run c1_q1big 600 python bench.py --workload cfg3shardq1 --steps 10 --warmup 3 --no-cpu-baseline --no-north-star
run c1_cfg1q1 600 python bench.py --workload cfg1q1 --steps 50 --warmup 5 --no-cpu-baseline --no-north-star
