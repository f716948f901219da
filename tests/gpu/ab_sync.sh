#!/bin/bash
# GPU-box helper: A/B of the stream-wait mode and the number of hardware queues on the default bench
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python bench.py --steps 6 --warmup 3 --no-cpu $EXTRA 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value'],3), round(d['e2e']['value'],3), round(d['ms_per_step'],1), d['host_cores_busy'])"; }
run base X=1
run conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
run spin300 VM_SYNC_SPIN_US=300
run spinall VM_SYNC_SPIN_US=-1
run conn32_spin300 CUDA_DEVICE_MAX_CONNECTIONS=32 VM_SYNC_SPIN_US=300
EXTRA="--workers 8 --chunk 2500 --ahead 3" run w8c2500a3 X=1
EXTRA="--workers 8 --chunk 2500 --ahead 3" run w8c2500a3_conn32 CUDA_DEVICE_MAX_CONNECTIONS=32
EXTRA="--workers 6 --ahead 3" run w6a3 X=1
EXTRA="--workers 6 --ahead 3" run w6a3_conn32_spin CUDA_DEVICE_MAX_CONNECTIONS=32 VM_SYNC_SPIN_US=300
