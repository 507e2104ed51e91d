#!/bin/bash
# final round-2 evidence for the two search workloads the driver's line carries (configs[1], configs[0]): plain runs first, then the
# launch lists and --set full captures of the same commands (scripts/make_r02_profiles.py turns them into profiles/r02_*.md)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
run() { name=$1; shift; timeout "$1" "${@:2}" > gpurun_out/$name.log 2>&1; echo "$name rc=$?"; tail -n 2 gpurun_out/$name.log | cut -c1-200; }
A="--steps 2 --warmup 1 --no-cpu-baseline --no-north-star --no-selfjoin"
run k1_cfg1_plain 300 python bench.py $A
run k1_cfg0_plain 300 python bench.py $A --workload cfg0
run k1_chain_timeline 120 python scripts/dev/trace_chain.py
LL="ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv"
timeout 400 $LL --log-file gpurun_out/k1_cfg1_launches.csv python bench.py $A > gpurun_out/k1_cfg1_ncu1.log 2>&1; echo "ll cfg1 $?"
timeout 400 $LL --log-file gpurun_out/k1_cfg0_launches.csv python bench.py $A --workload cfg0 > gpurun_out/k1_cfg0_ncu1.log 2>&1; echo "ll cfg0 $?"
FULL="ncu --set full --clock-control none --import-source on"
timeout 400 $FULL -k regex:"select_kernel|seed_tau|normalize_rows" -s 8 -c 4 -o gpurun_out/k1_chain -f python bench.py $A > gpurun_out/k1_chain_ncu.log 2>&1; echo "full chain $?"
timeout 400 $FULL -k regex:"scan_tc2_kernel" -s 6 -c 2 -o gpurun_out/h1_cur -f python bench.py $A > gpurun_out/h1_cur_ncu.log 2>&1; echo "full scan $?"
timeout 400 $FULL -k regex:"scan_small|dense_topk" -s 4 -c 2 -o gpurun_out/k1_small -f python bench.py $A --workload cfg0 > gpurun_out/k1_small_ncu.log 2>&1; echo "full small $?"
