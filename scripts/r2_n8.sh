#!/bin/bash
# N-GPU pass (N = GPUs of the box): NCCL / peer-push parity at world N, one-process multi-GPU collection on real devices, bench at N
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-8} gpurun_out/$name.log | cut -c1-500; }
run j${N}_multi 900 python -m pytest tests/test_gpu_multidevice.py tests/test_gpu_sharded_nccl.py -q -m gpu --timeout 600 -rs
run j${N}_check 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 scripts/check_sharded_nccl.py
run j${N}_bench 1800 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 20 --warmup 5
run j${N}_bench_nccl 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29523 bench.py --gpus $N --steps 20 --warmup 5 --exchange nccl --no-north-star
