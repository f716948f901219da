#!/bin/bash
# GPU-box helper: the default bench under different environments.  usage: bash tests/gpu/ab_env.sh "tag VAR=val ... [-- bench args]" ...
mkdir -p gpurun_out
for spec in "$@"; do
  set -- $spec
  tag=$1; shift
  envs=(); args=()
  while [ $# -gt 0 ] && [ "$1" != "--" ]; do envs+=("$1"); shift; done
  [ "$1" == "--" ] && shift
  args=("$@")
  env "${envs[@]}" timeout 280 python bench.py --steps 5 --warmup 3 --no-cpu "${args[@]}" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  tail -1 gpurun_out/ab_$tag.err
done
