#!/bin/bash
# One GPU-box pass: staged parity tests (each file in its own process so that a trapped kernel cannot poison
# the rest), then a short bench.  Logs land in gpurun_out/.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n 12 gpurun_out/$name.log; }
run t_dense 600 python -m pytest tests/test_gpu_search.py -q -m gpu -k dense --timeout 200
run t_pool 900 python -m pytest tests/test_gpu_maskpool.py -q -m gpu --timeout 300
run t_search 1800 python -m pytest tests/test_gpu_search.py -q -m gpu -k "not dense" --timeout 900
run t_dropin 600 python -m pytest tests/test_gpu_dropin.py -q -m gpu --timeout 300
run smoke 300 python __graft_entry__.py --smoke
run bench 1200 python bench.py --steps 20 --warmup 5
