#!/bin/bash
# Breadth measurements: other workloads through bench.py and the K1 mask-pool timing.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for wl in ${WORKLOADS:-cfg0 cfg3shardq16 cfg3shard}; do
  echo "=== $wl"
  timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -n 1 | python -c "
import sys,json
l=json.loads(sys.stdin.read())
print('value %.1f q/s  step %.3f ms  scan %.3f ms  hbm_frac %.3f tensor_frac %s e2e %.1f (%.3f ms) launches %d ok=%s' % (l['value'], l['ms_per_step'], l['roofline']['kernel_ms'], l['roofline']['frac'], (l['roofline_tensor'] or {}).get('frac'), l['e2e']['value'], l['e2e']['ms_per_step'], l['gpu_launches'], l['results_ok']))"
done
echo "=== mask pool cfg2"
timeout 600 python scripts/bench_maskpool.py
