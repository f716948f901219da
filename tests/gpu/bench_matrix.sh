#!/bin/bash
# GPU-box helper: GPU parity tests, then bench.py over a matrix of "reads workers chunk" configurations.
# usage: bash tests/gpu/bench_matrix.sh "2000 1 0" "10000 4 0" ...   (outputs under gpurun_out/)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5; fi
for cfg in "$@"; do
  set -- $cfg
  timeout 600 python bench.py --reads $1 --steps ${STEPS:-3} --warmup 3 --no-cpu --workers $2 --chunk $3 --ahead ${AHEAD:-2} \
      > gpurun_out/b$1_w$2_c$3.json 2> gpurun_out/b$1_w$2_c$3.err
  tail -2 gpurun_out/b$1_w$2_c$3.err
done
