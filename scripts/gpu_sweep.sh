#!/bin/bash
# Option sweep for the scan kernel on the bench workload: RVO_OPTS="name=value,..." is applied by bench.py.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
for opts in "$@"; do
  echo "=== $opts"
  RVO_OPTS="$opts" timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>&1 | tail -n 1 | python -c "
import sys,json
l=json.loads(sys.stdin.read())
print('value %.0f q/s  step %.3f ms  scan %.3f ms  hbm_frac %.3f  e2e %.0f  ok=%s' % (l['value'], l['ms_per_step'], l['roofline']['kernel_ms'], l['roofline']['frac'], l['e2e']['value'], l['results_ok']))"
done
