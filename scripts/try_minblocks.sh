#!/bin/bash
# experiment: solver-kernel occupancy vs register cap (rebuilds on the GPU box)
for mb in 1 2 3 4; do
  PN_EXTRA_NVCC_FLAGS="-DPN_MINBLOCKS=$mb" python code-adaptive-prob-ode-solvers_b200/build.py --force > /dev/null 2>&1
  echo "=== PN_MINBLOCKS=$mb"
  python scripts/gpu_quick.py 65536 2>&1 | grep -E "kernel info|iter 1"
done
python code-adaptive-prob-ode-solvers_b200/build.py --force > /dev/null 2>&1
