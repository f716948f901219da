#!/bin/bash
# GPU-box helper: A/B of the fill kernels' share of the block slots x workers in flight on the default bench
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 6 --warmup 3 --no-cpu $EXTRA 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value'],3), round(d['e2e']['value'],3), round(d['ms_per_step'],1), d['host_cores_busy'])"; }
for f in 1.0 0.7 0.5 0.35; do
  EXTRA="" run share${f}_default VM_FILL_SM_FRAC=$f
  EXTRA="--workers 6 --ahead 3" run share${f}_w6a3 VM_FILL_SM_FRAC=$f
  EXTRA="--workers 8 --chunk 2500 --ahead 3" run share${f}_w8c2500a3 VM_FILL_SM_FRAC=$f
done
