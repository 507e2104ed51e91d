#!/bin/bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-25} gpurun_out/$name.log; }
run c1_ref 900 python -m pytest tests/test_reference_source.py tests/test_gpu_multidevice.py -q -m gpu --timeout 600 -rs
