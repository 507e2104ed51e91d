#!/bin/bash
# 2-GPU pass: real-device multi-shard collection, NCCL/peer-push parity, bench at N=2 (strong-scaled 100M x 1280 included)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-12} gpurun_out/$name.log | cut -c1-700; }
run e2_multi 900 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_sharded_nccl.py tests/test_reference_source.py -q -m gpu --timeout 600 -rs
run e2_bench 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5
