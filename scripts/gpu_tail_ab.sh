#!/bin/bash
# scripts/gpu_tail_ab.sh -- the one-kernel V-cycle tail: parity tests, then A/B of the headline bench.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
: > gpurun_out/tail_ab.txt
run() {
  echo "== $*" >> gpurun_out/tail_ab.txt
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2> gpurun_out/tail_err.log | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('it/s %.1f ms %.3f e2e %.1f cg %.2f parity %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['cg']['iterations_per_arap_iteration'], d.get('parity')))
k=d['kernels']
print('   ', {n: round(v['avg_us'],1) for n,v in k.items() if n.startswith('mg_')})
" >> gpurun_out/tail_ab.txt 2>&1
  tail -2 gpurun_out/tail_err.log >> gpurun_out/tail_ab.txt
}
run ARAP_TAIL_ROWS=0
run ARAP_TAIL_ROWS=4096
run ARAP_TAIL_ROWS=4096 ARAP_TAIL_CLUSTER=16
run ARAP_TAIL_ROWS=16384 ARAP_TAIL_PARENT_ROWS=200000
run ARAP_TAIL_ROWS=16384 ARAP_TAIL_PARENT_ROWS=200000 ARAP_TAIL_CLUSTER=16
run ARAP_TAIL_ROWS=4096 ARAP_TAIL_CLUSTER=4
cat gpurun_out/tail_ab.txt
