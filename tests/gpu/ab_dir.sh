#!/bin/bash
# GPU-box helper: position-directory sizing (minimum run length, bucket count) on the default and the 250 Mb workloads
mkdir -p gpurun_out
run() { tag=$1; wl=$2; shift; shift; env "$@" python bench.py --workload $wl --steps 6 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', round(d['value'],3), round(d['ms_per_step'],1), 'hits', d['roofline']['all_kernels_ms'].get('k_reseed_hits'), 'hbm', d['config'].get('hbm_used_gb'))"; }
run cfg1_default cfg1 X=1
run cfg1_run8_b16 cfg1 VM_KB_MIN_RUN=8 VM_KB_MAX_BUCKETS=16
run cfg1_run8_b64 cfg1 VM_KB_MIN_RUN=8 VM_KB_MAX_BUCKETS=64
run cfg1_run4_b256 cfg1 VM_KB_MIN_RUN=4 VM_KB_MAX_BUCKETS=256
run cfg3_default cfg3 X=1
run cfg3_b512 cfg3 VM_KB_MAX_BUCKETS=512
run cfg3_run16_b4096 cfg3 VM_KB_MIN_RUN=16 VM_KB_MAX_BUCKETS=4096
