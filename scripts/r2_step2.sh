#!/bin/bash
# round 2, step 2: boundary — multi-device collections, persistence, tables, fp16 K1, NaN guards; then everything
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-15} gpurun_out/$name.log; }
run b1_multi 900 python -m pytest tests/test_gpu_multidevice.py -q -m gpu -x --timeout 600
run b1_pool 900 python -m pytest tests/test_gpu_maskpool.py -q -m gpu --timeout 300
run b1_dropin 900 python -m pytest tests/test_gpu_dropin.py tests/test_gpu_selfjoin.py -q -m gpu --timeout 300
run b1_search 1500 python -m pytest tests/test_gpu_search.py -q -m gpu --timeout 900
run b1_smoke 300 python __graft_entry__.py --smoke
