#!/bin/bash
# GPU-box helper: the round's evidence run -- default bench, reference arm, ncu launch list of the bench command,
# one `ncu --set full` capture per hot kernel.  Outputs under gpurun_out/.
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -2 gpurun_out/bench_default.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_bench.csv \
    python bench.py --reads 2000 --steps 2 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1
bash tests/gpu/ncu_full.sh fill "vm_fillb_kernel|vm_fill_kernel" 14
bash tests/gpu/ncu_full.sh edupper vm_ed_upper_kernel 1
bash tests/gpu/ncu_full.sh reseed "vm_reseed" 2
bash tests/gpu/ncu_full.sh chain "vm_chain_exact" 3
bash tests/gpu/ncu_full.sh seed "vm_sketch|vm_seed" 5
ls -la gpurun_out | head -30
