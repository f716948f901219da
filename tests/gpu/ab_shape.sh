#!/bin/bash
# GPU-box helper: job shapes (workers x chunk x jobs ahead) on the default bench
mkdir -p gpurun_out
run() { tag=$1; shift; python bench.py --steps 8 --warmup 4 --no-cpu "$@" 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value'],3), round(d['e2e']['value'],3), round(d['ms_per_step'],1), d['host_cores_busy'], d['gpu_launches'], d['config'].get('hbm_used_gb'))"; }
run default
run w4_c10000_a3 --workers 4 --chunk 10000 --ahead 3
run w6_c10000_a5 --workers 6 --chunk 10000 --ahead 5
run w3_c10000_a2 --workers 3 --chunk 10000 --ahead 2
run w8_c5000_a4 --workers 8 --chunk 5000 --ahead 4
