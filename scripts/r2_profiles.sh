#!/bin/bash
# round-2 evidence: launch lists (ncu --metrics gpu__time_duration.sum) of the search steps + --set full captures of the kernels that changed
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-300; }
A="--steps 2 --warmup 1 --no-cpu-baseline --no-north-star"
run k1_dropin 900 python -m pytest tests/test_gpu_dropin.py -q -m gpu -x --timeout 600
run k1_load 1200 python scripts/bench_load.py
# plain runs first (numbers), then the profiler passes of the same commands
run k1_cfg1_plain 600 python bench.py $A
run k1_cfg0_plain 600 python bench.py $A --workload cfg0
run k1_q1big_plain 600 python bench.py $A --workload cfg3shardq1
LL="ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv"
timeout 600 $LL --log-file gpurun_out/k1_cfg1_launches.csv python bench.py $A > gpurun_out/k1_cfg1_ncu1.log 2>&1; echo "ll cfg1 $?"
timeout 600 $LL --log-file gpurun_out/k1_cfg0_launches.csv python bench.py $A --workload cfg0 > gpurun_out/k1_cfg0_ncu1.log 2>&1; echo "ll cfg0 $?"
timeout 600 $LL --log-file gpurun_out/k1_q1big_launches.csv python bench.py $A --workload cfg3shardq1 > gpurun_out/k1_q1big_ncu1.log 2>&1; echo "ll q1big $?"
FULL="ncu --set full --clock-control none --import-source on"
timeout 600 $FULL -k regex:"select_kernel|seed_tau|normalize_rows" -s 8 -c 4 -o gpurun_out/k1_chain python bench.py $A > gpurun_out/k1_chain_ncu.log 2>&1; echo "full chain $?"
timeout 600 $FULL -k regex:"scan_small|dense_topk" -s 4 -c 2 -o gpurun_out/k1_small python bench.py $A --workload cfg0 > gpurun_out/k1_small_ncu.log 2>&1; echo "full small $?"
timeout 600 $FULL -k regex:"scan_small" -s 6 -c 2 -o gpurun_out/k1_q1big python bench.py $A --workload cfg3shardq1 > gpurun_out/k1_q1big_ncu.log 2>&1; echo "full q1big $?"
run k1_pool1280_plain 300 python scripts/bench_maskpool.py
timeout 600 $FULL -k regex:mask_pool_tc -s 30 -c 2 -o gpurun_out/k1_pool1280 python scripts/bench_maskpool.py > gpurun_out/k1_pool1280_ncu.log 2>&1; echo "full pool $?"
timeout 900 $FULL -k regex:"scan_tc2|pairs_kernel" -s 40 -c 2 -o gpurun_out/k1_selfjoin python scripts/bench_selfjoin.py 400000 > gpurun_out/k1_selfjoin_ncu.log 2>&1; echo "full selfjoin $?"
