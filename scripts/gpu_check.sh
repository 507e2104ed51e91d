#!/bin/bash
# What the driver runs at round end, on one GPU box: the whole GPU suite, smoke(), the reference arm and the default bench.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-6} gpurun_out/$name.log | cut -c1-400; }
run z_tests 2400 python -m pytest tests -x -q -m gpu --timeout 900 -rs
run z_smoke 300 python __graft_entry__.py --smoke
run z_ref 900 python bench.py --impl reference --steps 20 --warmup 5
run z_bench 1800 python bench.py --steps 20 --warmup 5
