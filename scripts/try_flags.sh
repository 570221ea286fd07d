#!/bin/bash
# experiment: rebuild the library on the GPU box with extra nvcc flags and time the headline kernel
for flags in "$@"; do
  PN_EXTRA_NVCC_FLAGS="$flags" python code-adaptive-prob-ode-solvers_b200/build.py --force > /dev/null 2>&1
  echo "=== flags: $flags"
  python scripts/gpu_quick.py 65536 1 4 2>&1 | grep -E "kernel info|iter 2|stats" | tail -3
done
python code-adaptive-prob-ode-solvers_b200/build.py --force > /dev/null 2>&1
